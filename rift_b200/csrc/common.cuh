// Shared device/host helpers for the rift_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace rift {

// ---- error plumbing: every C-ABI entry returns 0 / negative and leaves a message here
void set_last_error(const std::string& msg);
const char* get_last_error();

#define RIFT_CUDA_OK(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::rift::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @ " +  \
                                   __FILE__ + ":" + std::to_string(__LINE__));                   \
            return -2;                                                                            \
        }                                                                                         \
    } while (0)

#define RIFT_REQUIRE(cond, msg)                                                  \
    do {                                                                         \
        if (!(cond)) {                                                           \
            ::rift::set_last_error(std::string("invalid argument: ") + (msg));   \
            return -1;                                                           \
        }                                                                        \
    } while (0)

extern long long g_kernel_launches;    // every kernel launch of this library passes through RIFT_LAUNCH_OK

#define RIFT_LAUNCH_OK()                                                                          \
    do {                                                                                          \
        ++::rift::g_kernel_launches;                                                              \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            ::rift::set_last_error(std::string("kernel launch: ") + cudaGetErrorString(_e) +      \
                                   " @ " + __FILE__ + ":" + std::to_string(__LINE__));           \
            return -2;                                                                            \
        }                                                                                         \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts
// with pdl_grid_sync(): `launch_dependents` lets the NEXT kernel of the stream be scheduled while this one
// runs (its CTAs then park in `griddepcontrol.wait`), `wait` blocks until the PREVIOUS kernel has completed and
// its memory is visible.  Memory semantics are therefore those of plain stream order; what is gained is the
// launch latency / prologue of ~1.5k small kernels per policy update.  RIFT_B200_PDL=0 launches without the
// attribute (both instructions are then no-ops).
bool pdl_enabled();

#ifdef __CUDACC__
#ifdef RIFT_PDL_EARLY_TRIGGER
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() {}      // dependents are released when this grid's CTAs exit
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() { pdl_trigger(); pdl_wait(); }
// Selective early trigger: small-footprint kernels that sit between two GEMMs of the dependent chains (LayerNorm, attention,
// activation backward) release their dependent at START, so the following GEMM's CTAs become resident beside them and run
// their set-up (barriers, TMEM allocation, descriptor prefetch) before they wait.  Measured on one box, same run:
// 8.23 vs 8.32 ms per step.  (Doing this in EVERY kernel is slower - 8.58 ms: dependents of multi-wave kernels and of the
// GEMMs park on SMs that the grid's own later waves and the other streams need.)  -DRIFT_PDL_NO_SELECTIVE compiles it out.
#ifndef RIFT_PDL_NO_SELECTIVE
__device__ __forceinline__ void pdl_grid_sync_sel() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); pdl_wait(); }
#else
__device__ __forceinline__ void pdl_grid_sync_sel() { pdl_trigger(); pdl_wait(); }
#endif

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float gelu_erf(float x) {       // nn.GELU() default: exact erf form
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
#endif

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

}  // namespace rift
