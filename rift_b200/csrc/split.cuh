// Split-bf16 operand planes: x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits).
// Producers of GEMM operands write these planes directly so the tcgen05 GEMM is a pure TMA -> UMMA pipe.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace rift {

struct Planes {              // [rows, Kp] bf16 each; columns [C, Kp) are zero
    uint16_t* hi = nullptr; uint16_t* lo = nullptr; int Kp = 0;
    __host__ __device__ bool on() const { return hi != nullptr; }
};

#ifdef __CUDACC__
__device__ __forceinline__ void split_store(const Planes& p, long long row, int col, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    const long long i = row * p.Kp + col;
    p.hi[i] = *reinterpret_cast<const uint16_t*>(&h);
    p.lo[i] = *reinterpret_cast<const uint16_t*>(&l);
}
// zero the pad columns [C, Kp) of one row; call with a few threads (lane / nlanes)
__device__ __forceinline__ void split_zero_pad(const Planes& p, long long row, int C, int lane, int nlanes) {
    for (int c = C + lane; c < p.Kp; c += nlanes) { p.hi[row * p.Kp + c] = 0; p.lo[row * p.Kp + c] = 0; }
}
__device__ __forceinline__ void split4_store(const Planes& p, long long row, int col, float a, float b, float c, float d) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b), h2 = __float2bfloat16_rn(c),
                        h3 = __float2bfloat16_rn(d);
    __nv_bfloat162 x0 = __halves2bfloat162(h0, h1), x1 = __halves2bfloat162(h2, h3);
    __nv_bfloat162 y0 = __floats2bfloat162_rn(a - __bfloat162float(h0), b - __bfloat162float(h1));
    __nv_bfloat162 y1 = __floats2bfloat162_rn(c - __bfloat162float(h2), d - __bfloat162float(h3));
    uint2 uh, ul;
    uh.x = *reinterpret_cast<uint32_t*>(&x0); uh.y = *reinterpret_cast<uint32_t*>(&x1);
    ul.x = *reinterpret_cast<uint32_t*>(&y0); ul.y = *reinterpret_cast<uint32_t*>(&y1);
    const long long i = row * p.Kp + col;            // col % 4 == 0 and Kp % 64 == 0 -> 8-byte aligned
    *reinterpret_cast<uint2*>(p.hi + i) = uh;
    *reinterpret_cast<uint2*>(p.lo + i) = ul;
}
#endif

}  // namespace rift
