// Exact-fp32 SIMT GEMM with generic operand strides and the fused epilogue described in ops.h.
//
// Role: (1) the numerically exact device path every tensor-core kernel is validated against,
// (2) the carrier for shapes the tcgen05 path does not take (K < 16, N == 1, strided views,
// transposed operands of the weight-gradient products).  64x64x16 tiles, 256 threads, 4x4 per thread.
#include <stdlib.h>

#include "common.cuh"
#include "ops.h"

namespace rift {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

struct Epilogue {
    const float* bias; const float* colscale;
    const float* pre; long long ldpre; int pre_div;
    const float* res; long long ldres; int res_div; int res_mod;
    int act; float beta; float alpha;
    float* preact; long long ldc;
};

__device__ __forceinline__ float apply_epilogue(const Epilogue& e, float acc, int m, int n, float cold) {
    float v = acc * e.alpha;
    if (e.pre) v += e.pre[(long long)(m / e.pre_div) * e.ldpre + n];
    if (e.colscale) v *= e.colscale[n];
    if (e.bias) v += e.bias[n];
    if (e.preact) e.preact[(long long)m * e.ldc + n] = v;
    if (e.act == ACT_RELU) v = fmaxf(v, 0.f);
    else if (e.act == ACT_GELU) v = gelu_erf(v);
    if (e.res) {
        const int rr = e.res_mod > 0 ? (m % e.res_mod) : (m / e.res_div);
        v += e.res[(long long)rr * e.ldres + n];
    }
    if (e.beta != 0.f) v += e.beta * cold;
    return v;
}

// precision emulation for numerics studies (env RIFT_B200_EMULATE): 0 exact, 1 tf32 round-to-nearest,
// 2 tf32 truncation (what kind::tf32 does to raw fp32 bits), 3 bf16 round-to-nearest
__device__ __forceinline__ float emulate_round(float x, int mode) {
    if (mode == 1) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return __uint_as_float(u); }
    if (mode == 2) return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    if (mode == 3) return __uint_as_float(((__float_as_uint(x) + 0x7fffu + ((__float_as_uint(x) >> 16) & 1u)) & 0xffff0000u));
    return x;
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbn,
                 long long sbk, float* __restrict__ C, long long ldc, int M, int N, int K, int k_chunk, Epilogue ep,
                 float* __restrict__ split_ws, int emu) {
    pdl_grid_sync();
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Bs[BK][BN + PADM];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_chunk;
    const int kend = min(K, kbeg + k_chunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int m, k;
            if (A_KMAJOR) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < kend) ? emulate_round(__ldg(A + (long long)gm * sam + (long long)gk * sak), emu) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int n, k;
            if (B_KMAJOR) { k = e & 15; n = e >> 4; } else { n = e & 63; k = e >> 6; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < kend) ? emulate_round(__ldg(B + (long long)gn * sbn + (long long)gk * sbk), emu) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            if (split_ws) {
                split_ws[((long long)blockIdx.z * M + m) * N + n] = acc[i][j];
            } else {
                float* c = C + (long long)m * ldc + n;
                *c = apply_epilogue(ep, acc[i][j], m, n, ep.beta != 0.f ? *c : 0.f);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
gemm_splitk_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, long long ldc, int M, int N,
                          Epilogue ep, int Nw) {
    pdl_grid_sync();
    const long long total = (long long)M * N, slab = (long long)M * Nw;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int m = (int)(i / N), n = (int)(i - (long long)m * N);
        const float* w = ws + (long long)m * Nw + n;
        float a = 0.f;
        for (int z = 0; z < splits; ++z) a += w[(long long)z * slab];     // fixed order: deterministic
        float* c = C + (long long)m * ldc + n;
        *c = apply_epilogue(ep, a, m, n, ep.beta != 0.f ? *c : 0.f);
    }
}

// fixed-order reduction of split-K partials [splits, M, N] + the epilogue of `a` into a.C (shared with gemm_tc.cu)
int launch_splitk_reduce(const float* ws, int splits, const GemmArgs& a, cudaStream_t st, int ws_pitch) {
    Epilogue ep{a.bias, a.colscale, a.pre, a.ldpre, a.pre_div, a.res, a.ldres, a.res_div, a.res_mod, a.act, a.beta, a.alpha, a.preact, a.ldc};
    const long long total = (long long)a.M * a.N;
    launch_k(gemm_splitk_reduce_kernel, (int)min((long long)148 * 8, (total + 255) / 256), 256, 0, st, ws, splits, a.C, a.ldc, a.M, a.N, ep,
             ws_pitch > 0 ? ws_pitch : a.N);
    RIFT_LAUNCH_OK();
    return 0;
}

// N == 1 (the probability head's last Linear): one warp per row, lanes stride the reduction index
__global__ void __launch_bounds__(256)
gemv_rowdot_kernel(const float* __restrict__ A, long long sam, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ C, long long ldc, int M, int K) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* a = A + (long long)row * sam;
    float s0 = 0.f, s1 = 0.f;
    int k = lane;
    for (; k + 32 < K; k += 64) { s0 = fmaf(a[k], __ldg(w + k), s0); s1 = fmaf(a[k + 32], __ldg(w + k + 32), s1); }
    if (k < K) s0 = fmaf(a[k], __ldg(w + k), s0);
    const float s = warp_sum(s0 + s1);
    if (lane == 0) C[(long long)row * ldc] = s + (bias ? bias[0] : 0.f);
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
    RIFT_REQUIRE(a.A && a.B && a.C, "gemm: null operand");
    if (a.M <= 0 || a.N <= 0) return 0;
    RIFT_REQUIRE(a.K > 0, "gemm: K must be positive");
    if (a.N == 1 && a.sak == 1 && a.sbk == 1 && a.M >= 256 && !a.colscale && !a.pre && !a.res && !a.preact && a.act == ACT_NONE &&
        a.beta == 0.f && a.alpha == 1.f && a.split_k <= 1) {
        launch_k(gemv_rowdot_kernel, cdiv(a.M, 8), 256, 0, st, a.A, a.sam, a.B, a.bias, a.C, a.ldc, a.M, a.K);
        RIFT_LAUNCH_OK();
        return 0;
    }
    RIFT_REQUIRE(a.pre_div > 0 && a.res_div > 0, "gemm: broadcast divisors must be positive");
    Epilogue ep{a.bias, a.colscale, a.pre, a.ldpre, a.pre_div, a.res, a.ldres, a.res_div, a.res_mod, a.act, a.beta, a.alpha, a.preact, a.ldc};
    int splits = a.split_k > 1 ? a.split_k : 1;
    RIFT_REQUIRE(splits == 1 || a.split_ws != nullptr, "gemm: split_k needs a workspace");
    int k_chunk = a.K;
    if (splits > 1) {
        k_chunk = ((a.K + splits - 1) / splits + BK - 1) / BK * BK;
        splits = (a.K + k_chunk - 1) / k_chunk;
    }
    dim3 grid(cdiv(a.N, BN), cdiv(a.M, BM), splits);
    float* ws = splits > 1 ? a.split_ws : nullptr;
    static int emu = -1;
    if (emu < 0) { const char* s = getenv("RIFT_B200_EMULATE"); emu = s ? atoi(s) : 0; }
    const bool ak = (a.sak == 1), bk = (a.sbk == 1);
#define RIFT_GEMM_LAUNCH(AK, BK_)                                                                                     \
    launch_k(gemm_simt_kernel<AK, BK_>, grid, 256, 0, st, a.A, a.sam, a.sak, a.B, a.sbn, a.sbk, a.C, a.ldc, a.M, a.N, a.K, \
                                                    k_chunk, ep, ws, emu)
    if (ak && bk) RIFT_GEMM_LAUNCH(true, true);
    else if (ak && !bk) RIFT_GEMM_LAUNCH(true, false);
    else if (!ak && bk) RIFT_GEMM_LAUNCH(false, true);
    else RIFT_GEMM_LAUNCH(false, false);
#undef RIFT_GEMM_LAUNCH
    RIFT_LAUNCH_OK();
    if (splits > 1) {
        const long long total = (long long)a.M * a.N;
        launch_k(gemm_splitk_reduce_kernel, (int)min((long long)148 * 8, (total + 255) / 256), 256, 0, st, ws, splits, a.C, a.ldc,
                                                                                                    a.M, a.N, ep, a.N);
        RIFT_LAUNCH_OK();
    }
    return 0;
}

}  // namespace rift
