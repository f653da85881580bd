// Row-wise / gather / attention kernels of the Pluto policy (everything that is not a GEMM).
// Reference semantics are cited per kernel (paths relative to
// /root/reference/rift/cbv/planning/pluto/model/).
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"
#include "ops.h"

namespace rift {

#define GRID1D(n, t) (int)min((long long)148 * 16, ((long long)(n) + (t) - 1) / (t))
#define FOR_GRID(i, n) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)(n); i += (long long)gridDim.x * blockDim.x)

// =====================================================================================
// LayerNorm (eps 1e-5, affine), optional fused ReLU (MLPLayer / FourierEmbedding: Linear->LN->ReLU)
// and optional second output y2 = y + add[row % rowmod] (decoder m2m attention: q = k = LN(x) + m_pos).
// =====================================================================================
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ y, long long ldy, int relu,
                 const float* __restrict__ add_rowmod, int rowmod, float* __restrict__ y2, float* __restrict__ mean_out,
                 float* __restrict__ rstd_out, Planes yp, Planes y2p) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        const float* xr = x + (long long)row * ldx;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += xr[c];
        const float mean = warp_sum(s) / C;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q += d * d; }
        const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
        for (int c = lane; c < C; c += 32) {
            float v = (xr[c] - mean) * rstd * gamma[c] + beta[c];
            if (relu) v = fmaxf(v, 0.f);
            if (y) y[(long long)row * ldy + c] = v;
            if (yp.on()) split_store(yp, row, c, v);
            if (add_rowmod) {
                const float v2 = v + add_rowmod[(long long)(row % rowmod) * C + c];
                if (y2) y2[(long long)row * ldy + c] = v2;
                if (y2p.on()) split_store(y2p, row, c, v2);
            }
        }
        if (yp.on()) split_zero_pad(yp, row, C, lane, 32);
        if (y2p.on()) split_zero_pad(y2p, row, C, lane, 32);
    }
}

// 16-byte vectorised variant (C % 4 == 0, 16-byte aligned rows): every lane owns float4 column groups, so the
// fp32 and the split-bf16 plane stores are 16 / 8 byte vector stores.  The row is cached in registers (C <= 1024).
// LPR = lanes per row (32, or 16 / 8 for narrow rows so that a warp works on 2 / 4 rows at once).
template <int LPR>
__global__ void __launch_bounds__(256)
layernorm4_kernel(const float* __restrict__ x, long long ldx, int rows, int C, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ y, long long ldy, int relu,
                  const float* __restrict__ add_rowmod, int rowmod, float* __restrict__ y2, float* __restrict__ mean_out,
                  float* __restrict__ rstd_out, Planes yp, Planes y2p) {
    pdl_grid_sync_sel();
    constexpr int RPW = 32 / LPR;                     // rows per warp
    constexpr int NV = LPR == 32 ? 8 : 1;             // float4 groups per lane (narrow rows: C <= 4 * LPR)
    const int lane = threadIdx.x & 31;
    const int sl = lane % LPR, sub = lane / LPR;
    const int wpb = blockDim.x >> 5;
    const int C4 = C >> 2;
    const float invC = 1.f / C;
    for (int row0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW; row0 < rows; row0 += gridDim.x * wpb * RPW) {
        const int row = row0 + sub;
        const bool rok = row < rows;
        const float4* xr = reinterpret_cast<const float4*>(x + (long long)(rok ? row : 0) * ldx);
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = sl + LPR * i;
            if (c4 < C4) { v[i] = xr[c4]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = sl + LPR * i;
            if (c4 < C4) {
                const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / C + 1e-5f);
        (void)invC;
        if (!rok) continue;
        if (sl == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = sl + LPR * i;
            if (c4 < C4) {
                const float4 g = reinterpret_cast<const float4*>(gamma)[c4], b = reinterpret_cast<const float4*>(beta)[c4];
                float4 o;
                o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
                o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                if (y) *reinterpret_cast<float4*>(y + (long long)row * ldy + 4 * c4) = o;
                if (yp.on()) split4_store(yp, row, 4 * c4, o.x, o.y, o.z, o.w);
                if (add_rowmod) {
                    const float4 a = *reinterpret_cast<const float4*>(add_rowmod + (long long)(row % rowmod) * C + 4 * c4);
                    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                    if (y2) *reinterpret_cast<float4*>(y2 + (long long)row * ldy + 4 * c4) = o;
                    if (y2p.on()) split4_store(y2p, row, 4 * c4, o.x, o.y, o.z, o.w);
                }
            }
        }
        if (yp.on()) split_zero_pad(yp, row, C, sl, LPR);
        if (y2p.on()) split_zero_pad(y2p, row, C, sl, LPR);
    }
}

int launch_layernorm(const float* x, long long ldx, int rows, int C, const float* gamma, const float* beta, float* y,
                     long long ldy, int relu, const float* add_rowmod, int rowmod, float* y2, float* mean, float* rstd,
                     cudaStream_t st, Planes yp, Planes y2p) {
    if (rows <= 0) return 0;
    RIFT_REQUIRE((y2 == nullptr && !y2p.on()) || (add_rowmod != nullptr && rowmod > 0), "layernorm: y2 needs add_rowmod");
    if (y2 == nullptr && !y2p.on()) add_rowmod = nullptr;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((C & 3) == 0 && C <= 1024 && (ldx & 3) == 0 && (ldy & 3) == 0 && al16(x) && al16(y) && al16(y2) && al16(gamma) &&
        al16(beta) && al16(add_rowmod)) {
        // narrow rows: 16 / 8 lanes per row, several rows per warp; one pass per warp (no grid-stride cap) keeps
        // enough loads in flight to cover the HBM latency
        if (C <= 32)
            launch_k(layernorm4_kernel<8>, min(cdiv(rows, 32), 148 * 64), 256, 0, st, x, ldx, rows, C, gamma, beta, y, ldy, relu,
                     add_rowmod, rowmod, y2, mean, rstd, yp, y2p);
        else if (C <= 64)
            launch_k(layernorm4_kernel<16>, min(cdiv(rows, 16), 148 * 64), 256, 0, st, x, ldx, rows, C, gamma, beta, y, ldy, relu,
                     add_rowmod, rowmod, y2, mean, rstd, yp, y2p);
        else
            launch_k(layernorm4_kernel<32>, min(cdiv(rows, 8), 148 * 64), 256, 0, st, x, ldx, rows, C, gamma, beta, y, ldy, relu,
                     add_rowmod, rowmod, y2, mean, rstd, yp, y2p);
        RIFT_LAUNCH_OK();
        return 0;
    }
    launch_k(layernorm_kernel, min(cdiv(rows, 8), 148 * 8), 256, 0, st, x, ldx, rows, C, gamma, beta, y, ldy, relu, add_rowmod,
                                                                  rowmod, y2, mean, rstd, yp, y2p);
    RIFT_LAUNCH_OK();
    return 0;
}

// ---- LayerNorm backward.  dy_eff = dy * (y > 0) when the ReLU was fused.
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy_eff * gamma
// dgamma += sum_rows dy_eff * xhat ; dbeta += sum_rows dy_eff   (two-stage, fixed order -> deterministic)
constexpr int LNB_MAXC = 1024;
constexpr int LNB_BLOCKS = 592;       // partial rows of the parameter-gradient reduction (4 CTAs per SM)

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int rows,
                     int C, const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ y_relu, long long ldy, float* __restrict__ dx, long long lddx,
                     int dx_accumulate, float* __restrict__ partial /*[grid][2][C]*/, float* __restrict__ dgamma_atomic,
                     float* __restrict__ dbeta_atomic) {
    pdl_grid_sync();
    extern __shared__ float sm[];      // [8 warps][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dg[LNB_MAXC / 32], db[LNB_MAXC / 32];
#pragma unroll
    for (int i = 0; i < LNB_MAXC / 32; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
        const float* xr = x + (long long)row * ldx;
        const float* dyr = dy + (long long)row * lddy;
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_MAXC / 32; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
                float d = dyr[c];
                if (y_relu && !(y_relu[(long long)row * ldy + c] > 0.f)) d = 0.f;
                const float xh = (xr[c] - mu) * rs;
                const float g = d * gamma[c];
                s1 += g; s2 += g * xh;
                dg[i] += d * xh; db[i] += d;
            }
        }
        s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
        if (dx) {
#pragma unroll
            for (int i = 0; i < LNB_MAXC / 32; ++i) {
                const int c = lane + 32 * i;
                if (c < C) {
                    float d = dyr[c];
                    if (y_relu && !(y_relu[(long long)row * ldy + c] > 0.f)) d = 0.f;
                    const float xh = (xr[c] - mu) * rs;
                    const float v = rs * (d * gamma[c] - s1 - xh * s2);
                    float* o = dx + (long long)row * lddx + c;
                    *o = dx_accumulate ? *o + v : v;
                }
            }
        }
    }
    if (partial == nullptr && dgamma_atomic == nullptr) return;
#pragma unroll
    for (int i = 0; i < LNB_MAXC / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < C) { sm[(warp * 2 + 0) * C + c] = dg[i]; sm[(warp * 2 + 1) * C + c] = db[i]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
        const int which = c / C, cc = c - which * C;
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += sm[(w * 2 + which) * C + cc];
        if (dgamma_atomic) atomicAdd((which ? dbeta_atomic : dgamma_atomic) + cc, a);   // straight into the gradient arena
        else partial[((long long)blockIdx.x * 2 + which) * C + cc] = a;
    }
}

__global__ void __launch_bounds__(256)
partial_reduce_accumulate_kernel(const float* __restrict__ partial, int nb, long long stride, int C, float* __restrict__ out,
                                 int accumulate) {
    pdl_grid_sync();
    FOR_GRID(c, C) {
        float a = 0.f;
        for (int b = 0; b < nb; ++b) a += partial[(long long)b * stride + c];
        out[c] = accumulate ? out[c] + a : a;
    }
}

int layernorm_bwd_scratch_floats(int C) { return LNB_BLOCKS * 2 * C; }

__global__ void colsum_final_kernel(const float* __restrict__ partial, int nb, long long stride, int C, float* __restrict__ out,
                                    int accumulate);
// dgamma (blockIdx.y = 0) and dbeta (1) of one LayerNorm in a single launch
__global__ void colsum_final2_kernel(const float* __restrict__ partial, int nb, int C, float* __restrict__ out0,
                                     float* __restrict__ out1);
// stream the second reduction stage runs on (see SideStream)
static inline int second_stage_stream(cudaStream_t st, const SideStream& fin, cudaStream_t* out) {
    *out = st;
    if (fin.st && fin.st != st) {
        RIFT_CUDA_OK(cudaEventRecord(fin.ev, st));
        RIFT_CUDA_OK(cudaStreamWaitEvent(fin.st, fin.ev, 0));
        *out = fin.st;
    }
    return 0;
}

// 16-byte vectorised backward (C = 128 * NV, rows 16-byte aligned): each lane owns NV float4 column groups, the
// whole row lives in registers between the statistics pass and the dx pass, and the per-lane dgamma / dbeta
// accumulators have exactly the width the row needs.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd4_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ dy, long long lddy, int rows, int C,
                      const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ y_relu, long long ldy, float* __restrict__ dx, long long lddx,
                      int dx_accumulate, float* __restrict__ partial /*[grid][2][C]*/, float* __restrict__ dgamma_atomic,
                      float* __restrict__ dbeta_atomic, Planes dxp, const uint8_t* __restrict__ zero_flag, int zero_div) {
    pdl_grid_sync_sel();
    extern __shared__ float sm[];      // [8 warps][2][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C4 = C >> 2;
    float4 dg[NV], db[NV], gm[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c4 = lane + 32 * i;
        gm[i] = c4 < C4 ? reinterpret_cast<const float4*>(gamma)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
        const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
        const float4* dyr = reinterpret_cast<const float4*>(dy + (long long)row * lddy);
        const float4* yr = y_relu ? reinterpret_cast<const float4*>(y_relu + (long long)row * ldy) : nullptr;
        const float mu = mean[row], rs = rstd[row];
        float4 xh[NV], d[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < C4) {
                const float4 xv = xr[c4];
                float4 dv = dyr[c4];
                if (yr) {
                    const float4 yv = yr[c4];
                    if (!(yv.x > 0.f)) dv.x = 0.f; if (!(yv.y > 0.f)) dv.y = 0.f;
                    if (!(yv.z > 0.f)) dv.z = 0.f; if (!(yv.w > 0.f)) dv.w = 0.f;
                }
                xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                d[i] = dv;
                const float4 g = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
                s1 += (g.x + g.y) + (g.z + g.w);
                s2 += (g.x * xh[i].x + g.y * xh[i].y) + (g.z * xh[i].z + g.w * xh[i].w);
                dg[i].x += dv.x * xh[i].x; dg[i].y += dv.y * xh[i].y; dg[i].z += dv.z * xh[i].z; dg[i].w += dv.w * xh[i].w;
                db[i].x += dv.x; db[i].y += dv.y; db[i].z += dv.z; db[i].w += dv.w;
            }
        }
        s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
        if (dx) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c4 = lane + 32 * i;
                if (c4 < C4) {
                    float4 v;
                    v.x = rs * (d[i].x * gm[i].x - s1 - xh[i].x * s2); v.y = rs * (d[i].y * gm[i].y - s1 - xh[i].y * s2);
                    v.z = rs * (d[i].z * gm[i].z - s1 - xh[i].z * s2); v.w = rs * (d[i].w * gm[i].w - s1 - xh[i].w * s2);
                    float4* o = reinterpret_cast<float4*>(dx + (long long)row * lddx) + c4;
                    if (dx_accumulate) { const float4 t = *o; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    // rows whose forward output was overwritten with zero pass no gradient on (decoder: padded reference lines)
                    if (zero_flag && zero_flag[row / zero_div]) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    *o = v;
                    if (dxp.on()) {              // the updated residual-stream gradient also leaves as split-bf16 planes (next GEMM's dY)
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                                            h3 = __float2bfloat16_rn(v.w);
                        const __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
                        const __nv_bfloat162 cc = __floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
                        const __nv_bfloat162 dd = __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
                        uint2 uh, ul;
                        uh.x = *reinterpret_cast<const uint32_t*>(&a); uh.y = *reinterpret_cast<const uint32_t*>(&b);
                        ul.x = *reinterpret_cast<const uint32_t*>(&cc); ul.y = *reinterpret_cast<const uint32_t*>(&dd);
                        *reinterpret_cast<uint2*>(dxp.hi + (long long)row * dxp.Kp + 4 * c4) = uh;
                        *reinterpret_cast<uint2*>(dxp.lo + (long long)row * dxp.Kp + 4 * c4) = ul;
                    }
                }
            }
        }
    }
    if (partial == nullptr && dgamma_atomic == nullptr) return;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c4 = lane + 32 * i;
        if (c4 < C4) {
            *reinterpret_cast<float4*>(&sm[(warp * 2 + 0) * C + 4 * c4]) = dg[i];
            *reinterpret_cast<float4*>(&sm[(warp * 2 + 1) * C + 4 * c4]) = db[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
        const int which = c / C, cc = c - which * C;
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += sm[(w * 2 + which) * C + cc];
        if (dgamma_atomic) atomicAdd((which ? dbeta_atomic : dgamma_atomic) + cc, a);   // straight into the gradient arena
        else partial[((long long)blockIdx.x * 2 + which) * C + cc] = a;
    }
}

// Parameter gradients: by default every block adds its partial dgamma / dbeta straight into the gradient arena with
// atomicAdd (they ACCUMULATE, like every parameter gradient of the backward; summation order not fixed).
// RIFT_B200_LN_ATOMIC=0 restores the two-stage fixed-order reduction through `scratch` (+ colsum_final2 on `fin`).
int launch_layernorm_bwd(const float* x, long long ldx, const float* dy, long long lddy, int rows, int C,
                         const float* gamma, const float* mean, const float* rstd, const float* y_for_relu, long long ldy,
                         float* dx, long long lddx, int dx_accumulate, float* dgamma, float* dbeta, float* scratch,
                         cudaStream_t st, SideStream fin, Planes dxp, const uint8_t* zero_flag, int zero_div) {
    if (rows <= 0) return 0;
    {
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool vec = (C & 63) == 0 && (ldx & 3) == 0 && (lddy & 3) == 0 && (ldy & 3) == 0 && (lddx & 3) == 0 && al16(x) && al16(dy) &&
                         al16(y_for_relu) && al16(dx) && al16(gamma);
        RIFT_REQUIRE(!(dxp.on() || zero_flag) || vec,
                     "layernorm_bwd: plane output / row zeroing need the vector form (C % 64 == 0, 16-byte aligned rows)");
    }
    RIFT_REQUIRE(!dxp.on() || (dxp.Kp == C && dx != nullptr), "layernorm_bwd: plane pitch must equal C");
    if (zero_div < 1) zero_div = 1;
    RIFT_REQUIRE(C <= LNB_MAXC, "layernorm_bwd: C too large");
    RIFT_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma/dbeta go together");
    RIFT_REQUIRE(dgamma == nullptr || scratch != nullptr, "layernorm_bwd: scratch required for parameter gradients");
    static const bool atomic = [] { const char* e = getenv("RIFT_B200_LN_ATOMIC"); return !(e && atoi(e) == 0); }();
    float* part = (dgamma && !atomic) ? scratch : nullptr;
    float* dga = (dgamma && atomic) ? dgamma : nullptr;
    float* dba = (dgamma && atomic) ? dbeta : nullptr;
    int nb_used = 0;
    {
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool vec = (C & 3) == 0 && (ldx & 3) == 0 && (lddy & 3) == 0 && (ldy & 3) == 0 && (lddx & 3) == 0 && al16(x) &&
                         al16(dy) && al16(y_for_relu) && al16(dx) && al16(gamma);
        if (vec) {
            // more, smaller blocks than the partial buffer has slots is not possible: LNB_BLOCKS partial rows
            // rows per warp: fewer blocks mean fewer atomics on the 2 C parameter-gradient addresses (RIFT_B200_LNB_RPW)
            static const int rpw = [] { const char* e = getenv("RIFT_B200_LNB_RPW"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 1; }();
            const int nb4 = min(cdiv(rows, 8 * rpw), LNB_BLOCKS);
            const size_t smem4 = (size_t)8 * 2 * C * sizeof(float);
            const int nv = cdiv(C, 128);
#define RIFT_LNB4(NV)                                                                                                       \
    launch_k(layernorm_bwd4_kernel<NV>, nb4, 256, smem4, st, x, ldx, dy, lddy, rows, C, gamma, mean, rstd, y_for_relu, ldy, dx, \
                                                       lddx, dx_accumulate, part, dga, dba, dxp, zero_flag, zero_div)
            static bool attr4 = false;
            if (!attr4) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * LNB_MAXC * 4));
                attr4 = true;
            }
            if (nv <= 1) RIFT_LNB4(1); else if (nv == 2) RIFT_LNB4(2); else if (nv <= 4) RIFT_LNB4(4); else RIFT_LNB4(8);
#undef RIFT_LNB4
            RIFT_LAUNCH_OK();
            nb_used = nb4;
        }
    }
    if (!nb_used) {
        const int nb = min(cdiv(rows, 8), LNB_BLOCKS);
        const size_t smem = (size_t)8 * 2 * C * sizeof(float);
        static bool attr = false;
        if (!attr) {
            RIFT_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * LNB_MAXC * 4));
            attr = true;
        }
        launch_k(layernorm_bwd_kernel, nb, 256, smem, st, x, ldx, dy, lddy, rows, C, gamma, mean, rstd, y_for_relu, ldy, dx, lddx,
                                                    dx_accumulate, part, dga, dba);
        RIFT_LAUNCH_OK();
        nb_used = nb;
    }
    if (part) {
        cudaStream_t s2;
        int r2 = second_stage_stream(st, fin, &s2);
        if (r2) return r2;
        launch_k(colsum_final2_kernel, dim3(cdiv(C, 32), 2), 1024, 0, s2, scratch, nb_used, C, dgamma, dbeta);
        RIFT_LAUNCH_OK();
    }
    return 0;
}

// ---- column sum (bias gradients): out[c] (+)= sum_r x[r, c], two-stage deterministic
// block (32 columns x 8 row lanes): blockIdx.x = 32-column chunk, blockIdx.y = row slab; each warp reads
// 128 contiguous bytes of a row, four independent accumulators per thread keep loads in flight.
// scratch: [slabs][C] with slabs <= 148 (callers size it 148 * C).
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, long long ldx, int rows, int C, float* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int stride = gridDim.y * 8;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < C) {
        int r = blockIdx.y * 8 + ty;
        for (; r + 3 * stride < rows; r += 4 * stride) {
            a0 += x[(long long)r * ldx + c];
            a1 += x[(long long)(r + stride) * ldx + c];
            a2 += x[(long long)(r + 2 * stride) * ldx + c];
            a3 += x[(long long)(r + 3 * stride) * ldx + c];
        }
        for (; r < rows; r += stride) a0 += x[(long long)r * ldx + c];
    }
    sm[ty][tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i][tx];
        partial[(long long)blockIdx.y * C + c] = s;
    }
}
// out[c] (+)= sum over nb partial rows, 8 row lanes per column, fixed order
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int nb, long long stride, int C, float* __restrict__ out, int accumulate) {
    pdl_grid_sync();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < C) for (int b = ty; b < nb; b += 8) a += partial[(long long)b * stride + c];
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i][tx];
        out[c] = accumulate ? out[c] + s : s;
    }
}
__global__ void __launch_bounds__(1024)
colsum_final2_kernel(const float* __restrict__ partial, int nb, int C, float* __restrict__ out0, float* __restrict__ out1) {
    pdl_grid_sync();
    __shared__ float sm[32][33];                       // 32 row lanes x 32 columns, combined in a fixed order
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const float* p = partial + (long long)blockIdx.y * C;          // layout [nb][2][C]
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
        int b = ty;
        for (; b + 32 < nb; b += 64) { a0 += p[(long long)b * 2 * C + c]; a1 += p[(long long)(b + 32) * 2 * C + c]; }
        if (b < nb) a0 += p[(long long)b * 2 * C + c];
    }
    sm[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += sm[i][tx];
        float* out = blockIdx.y ? out1 : out0;
        out[c] += s;
    }
}
// first stage over a split-bf16 plane pair: value = hi + lo (16 mantissa bits, ample for a bias gradient)
__global__ void __launch_bounds__(256)
colsum_planes_partial_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, int Kp, int rows, int C,
                             float* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ float sm[8][65];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + 2 * tx;              // two columns (one bf16x2 word per plane) per thread
    const int stride = gridDim.y * 8;
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
        for (int r = blockIdx.y * 8 + ty; r < rows; r += stride) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(hi + (long long)r * Kp + c);
            const __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(lo + (long long)r * Kp + c);
            a0 += __bfloat162float(h.x) + __bfloat162float(l.x);
            a1 += __bfloat162float(h.y) + __bfloat162float(l.y);
        }
    }
    sm[ty][2 * tx] = a0; sm[ty][2 * tx + 1] = a1;
    __syncthreads();
    if (ty == 0) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cc = c + u;
            if (cc < C) {
                float s2 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s2 += sm[i][2 * tx + u];
                partial[(long long)blockIdx.y * C + cc] = s2;
            }
        }
    }
}
int launch_colsum_planes(const Planes& p, int rows, int C, float* out, int accumulate, float* scratch, cudaStream_t st,
                         SideStream fin) {
    if (C <= 0 || rows <= 0) return 0;
    RIFT_REQUIRE(p.on() && (C % 2) == 0, "colsum_planes: planes required, C must be even");
    const int chunks = cdiv(C, 64);
    const int slabs = max(1, min(148, min(cdiv(rows, 32), cdiv(148 * 4, chunks))));
    launch_k(colsum_planes_partial_kernel, dim3(chunks, slabs), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(p.hi),
             reinterpret_cast<const __nv_bfloat16*>(p.lo), p.Kp, rows, C, scratch);
    RIFT_LAUNCH_OK();
    return launch_colsum_final(scratch, slabs, C, out, accumulate, st, fin);
}
int launch_colsum_final(const float* partial, int slabs, int C, float* out, int accumulate, cudaStream_t st, SideStream fin) {
    if (C <= 0 || slabs <= 0) return 0;
    cudaStream_t s2;
    int r2 = second_stage_stream(st, fin, &s2);
    if (r2) return r2;
    launch_k(colsum_final_kernel, cdiv(C, 32), 256, 0, s2, partial, slabs, C, C, out, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}
int launch_colsum(const float* x, long long ldx, int rows, int C, float* out, int accumulate, float* scratch,
                  cudaStream_t st, SideStream fin) {
    if (C <= 0) return 0;
    const int chunks = cdiv(C, 32);
    int slabs = max(1, min(148, min(cdiv(rows, 32), cdiv(148 * 4, chunks))));
    launch_k(colsum_partial_kernel, dim3(chunks, slabs), 256, 0, st, x, ldx, rows, C, scratch);
    RIFT_LAUNCH_OK();
    cudaStream_t s2;
    int r2 = second_stage_stream(st, fin, &s2);
    if (r2) return r2;
    launch_k(colsum_final_kernel, chunks, 256, 0, s2, scratch, slabs, C, C, out, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// Multi-head attention for short sequences (S <= ~256, head_dim 32/64): one CTA per (batch, head),
// K and V of that head staged in shared memory, one query per thread with an online softmax.
// nn.MultiheadAttention semantics (layers/transformer.py:73-94, modules/planning_decoder.py:42-86):
// scores = (q * hd^-0.5) k^T, key_padding_mask -> -inf, softmax, @ v.  A fully masked row yields 0.
// =====================================================================================
__device__ __forceinline__ long long attn_row(int b, int inner_n, long long outer, long long inner) {
    return (long long)(b / inner_n) * outer + (long long)(b % inner_n) * inner;
}

template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(AttnArgs a) {
    pdl_grid_sync();
    extern __shared__ float sm[];
    float* Ks = sm;
    float* Vs = sm + (size_t)a.Sk * HD;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const long long krow0 = attn_row(b, a.k_inner_n, a.k_outer, a.k_inner);
    for (int e = threadIdx.x; e < a.Sk * HD; e += blockDim.x) {
        const int j = e / HD, d = e - j * HD;
        const long long r = krow0 + (long long)j * a.k_seq;
        Ks[e] = a.k[r * a.ldk + h * HD + d];
        Vs[e] = a.v[r * a.ldv + h * HD + d];
    }
    __syncthreads();
    const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * a.Sk : nullptr;
    const long long qrow0 = attn_row(b, a.q_inner_n, a.q_outer, a.q_inner);
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < a.Sq; i += gridDim.y * blockDim.x) {
        const long long qr = qrow0 + (long long)i * a.q_seq;
        float q[HD], acc[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = a.q[qr * a.ldq + h * HD + d] * a.scale; acc[d] = 0.f; }
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < a.Sk; ++j) {
            if (kpm && kpm[j]) continue;
            float s = 0.f;
            const float4* k4 = reinterpret_cast<const float4*>(Ks + j * HD);     // broadcast 16-byte shared loads
            const float4* v4 = reinterpret_cast<const float4*>(Vs + j * HD);
#pragma unroll
            for (int d = 0; d < HD / 4; ++d) {
                const float4 kk = k4[d];
                s = fmaf(q[4 * d], kk.x, s); s = fmaf(q[4 * d + 1], kk.y, s);
                s = fmaf(q[4 * d + 2], kk.z, s); s = fmaf(q[4 * d + 3], kk.w, s);
            }
            const float mn = fmaxf(m, s);
            const float corr = __expf(m - mn);      // m = -inf on the first key: exp(-inf) = 0
            const float p = __expf(s - mn);
            l = l * corr + p;
#pragma unroll
            for (int d = 0; d < HD / 4; ++d) {
                const float4 vv = v4[d];
                acc[4 * d] = acc[4 * d] * corr + p * vv.x; acc[4 * d + 1] = acc[4 * d + 1] * corr + p * vv.y;
                acc[4 * d + 2] = acc[4 * d + 2] * corr + p * vv.z; acc[4 * d + 3] = acc[4 * d + 3] * corr + p * vv.w;
            }
            m = mn;
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
        const long long orow = a.o_custom ? attn_row(b, a.o_inner_n, a.o_outer, a.o_inner) + (long long)i * a.o_seq : qr;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 o4 = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
            if (a.o) *reinterpret_cast<float4*>(a.o + orow * a.ldo + h * HD + d) = o4;
            if (a.o_planes.on()) split4_store(a.o_planes, orow, h * HD + d, o4.x, o4.y, o4.z, o4.w);
        }
        if (a.lse) a.lse[((long long)b * a.H + h) * a.Sq + i] = l > 0.f ? m + logf(l) : INFINITY;
    }
}

// ---- head_dim 32, cooperative form: PB (batch, head) problems per CTA, G lanes per query.  Q, K, V are staged with
// coalesced 16-byte loads into rows padded to 36 floats.  The G lanes of a query split the CHANNELS (32 / G each):
// every key costs a partial dot product plus log2(G) shuffles inside the lane group, the online-softmax state is
// replicated, and each lane ends up owning its slice of the output row - few registers, several CTAs per SM.
constexpr int AF_PITCH = 36;
__host__ __device__ inline size_t attn_fwd2_smem_floats(int Sq, int Sk) {
    const size_t n = (size_t)Sq * AF_PITCH + (size_t)2 * Sk * AF_PITCH + ((Sk + 3) / 4);
    return (n + 3) & ~(size_t)3;
}

template <int G>
__global__ void __launch_bounds__(512)
attention2_kernel(AttnArgs a, int PB, int tpp) {
    pdl_grid_sync_sel();
    constexpr int HD = 32, CPL = HD / G;
    extern __shared__ __align__(16) float sm[];
    const int Sq = a.Sq, Sk = a.Sk;
    const int pl = threadIdx.x / tpp, t = threadIdx.x - pl * tpp;
    const long long prob = (long long)blockIdx.x * PB + pl;
    const bool live = prob < (long long)a.B * a.H;
    float* base = sm + (size_t)pl * attn_fwd2_smem_floats(Sq, Sk);
    float* Qs = base;
    float* Ks = Qs + (size_t)Sq * AF_PITCH;
    float* Vs = Ks + (size_t)Sk * AF_PITCH;
    uint8_t* msk = reinterpret_cast<uint8_t*>(Vs + (size_t)Sk * AF_PITCH);
    const int b = live ? (int)(prob / a.H) : 0, h = live ? (int)(prob % a.H) : 0;
    const long long krow0 = attn_row(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row(b, a.q_inner_n, a.q_outer, a.q_inner);
    if (live) {
        for (int e = t; e < Sk * 8; e += tpp) {
            const int j = e >> 3, c = (e & 7) * 4;
            const long long r = krow0 + (long long)j * a.k_seq;
            *reinterpret_cast<float4*>(Ks + j * AF_PITCH + c) = *reinterpret_cast<const float4*>(a.k + r * a.ldk + h * HD + c);
            *reinterpret_cast<float4*>(Vs + j * AF_PITCH + c) = *reinterpret_cast<const float4*>(a.v + r * a.ldv + h * HD + c);
        }
        for (int e = t; e < Sq * 8; e += tpp) {
            const int i = e >> 3, c = (e & 7) * 4;
            float4 q = *reinterpret_cast<const float4*>(a.q + (qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + c);
            q.x *= a.scale; q.y *= a.scale; q.z *= a.scale; q.w *= a.scale;
            *reinterpret_cast<float4*>(Qs + i * AF_PITCH + c) = q;
        }
        const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * Sk : nullptr;
        for (int j = t; j < Sk; j += tpp) msk[j] = kpm ? kpm[j] : 0;
    }
    __syncthreads();
    const int row = t / G, g = t % G;
    if (!(live && row < Sq)) return;                 // whole lane groups leave together; shuffles below name the group only
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) & ~(G - 1));
    const int i = row, c0 = g * CPL;
    float q[CPL], acc[CPL];
#pragma unroll
    for (int d = 0; d < CPL; d += 4) {
        const float4 x = *reinterpret_cast<const float4*>(Qs + i * AF_PITCH + c0 + d);
        q[d] = x.x; q[d + 1] = x.y; q[d + 2] = x.z; q[d + 3] = x.w;
        acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < Sk; ++j) {
        if (msk[j]) continue;
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < CPL; d += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AF_PITCH + c0 + d);
            sc = fmaf(q[d], kk.x, sc); sc = fmaf(q[d + 1], kk.y, sc); sc = fmaf(q[d + 2], kk.z, sc); sc = fmaf(q[d + 3], kk.w, sc);
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1) sc += __shfl_xor_sync(gmask, sc, o);
        const float mn = fmaxf(m, sc);
        const float corr = __expf(m - mn);          // m = -inf on the first key: exp(-inf) = 0
        const float p = __expf(sc - mn);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < CPL; d += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(Vs + j * AF_PITCH + c0 + d);
            acc[d] = acc[d] * corr + p * vv.x; acc[d + 1] = acc[d + 1] * corr + p * vv.y;
            acc[d + 2] = acc[d + 2] * corr + p * vv.z; acc[d + 3] = acc[d + 3] * corr + p * vv.w;
        }
        m = mn;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    const long long qr = qrow0 + (long long)i * a.q_seq;
    const long long orow = a.o_custom ? attn_row(b, a.o_inner_n, a.o_outer, a.o_inner) + (long long)i * a.o_seq : qr;
#pragma unroll
    for (int d = 0; d < CPL; d += 4) {
        const float4 o4 = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
        if (a.o) *reinterpret_cast<float4*>(a.o + orow * a.ldo + h * HD + c0 + d) = o4;
        if (a.o_planes.on()) split4_store(a.o_planes, orow, h * HD + c0 + d, o4.x, o4.y, o4.z, o4.w);
    }
    if (a.lse && g == 0) a.lse[((long long)b * a.H + h) * Sq + i] = l > 0.f ? m + logf(l) : INFINITY;
}

// ---- head_dim 32, register-tiled form (sequences of 24 .. 128 rows; see attention_bwd3_kernel for the tiling) ----------------
//   A  S tile = 4 consecutive queries x 4 keys strided by Sk4 / 4 (conflict-free K rows) over 32 channels -> Ss (masked: -inf)
//   B  row softmax, 4 lanes per row: m, l, P = exp(S - m) left unnormalised in Ss, 1 / l and the log-sum-exp per row
//   C  O tile = 4 queries x 4 channels, loop over keys; scaled by 1 / l on the way out
__host__ __device__ inline int attn_fwd3_skp(int Sk) { const int s4 = (Sk + 3) & ~3; return (s4 & 7) ? s4 : s4 + 4; }
__host__ __device__ inline size_t attn_fwd3_smem_floats(int Sq, int Sk) {
    const int Sq4 = (Sq + 3) & ~3, Sk4 = (Sk + 3) & ~3;
    return (size_t)Sq4 * AF_PITCH + (size_t)2 * Sk4 * AF_PITCH + (size_t)Sq4 * attn_fwd3_skp(Sk) + Sq4 + Sk4 / 4;
}
__global__ void __launch_bounds__(256, 3)
attention3_kernel(AttnArgs a) {
    pdl_grid_sync_sel();
    constexpr int HD = 32, T = 256;
    extern __shared__ __align__(16) float sm[];
    const int Sq = a.Sq, Sk = a.Sk, Sq4 = (Sq + 3) & ~3, Sk4 = (Sk + 3) & ~3, SkP = attn_fwd3_skp(Sk);
    float* Qs = sm;
    float* Ks = Qs + (size_t)Sq4 * AF_PITCH;
    float* Vs = Ks + (size_t)Sk4 * AF_PITCH;
    float* Ss = Vs + (size_t)Sk4 * AF_PITCH;             // [Sq4][SkP]
    float* inv = Ss + (size_t)Sq4 * SkP;                 // [Sq4]
    uint8_t* msk = reinterpret_cast<uint8_t*>(inv + Sq4);
    const int t = threadIdx.x;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const long long krow0 = attn_row(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row(b, a.q_inner_n, a.q_outer, a.q_inner);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = t; e < Sk4 * 8; e += T) {
        const int j = e >> 3, c = (e & 7) * 4;
        const long long r = krow0 + (long long)j * a.k_seq;
        *reinterpret_cast<float4*>(Ks + j * AF_PITCH + c) = j < Sk ? *reinterpret_cast<const float4*>(a.k + r * a.ldk + h * HD + c) : z4;
        *reinterpret_cast<float4*>(Vs + j * AF_PITCH + c) = j < Sk ? *reinterpret_cast<const float4*>(a.v + r * a.ldv + h * HD + c) : z4;
    }
    for (int e = t; e < Sq4 * 8; e += T) {
        const int i = e >> 3, c = (e & 7) * 4;
        float4 q = z4;
        if (i < Sq) {
            q = *reinterpret_cast<const float4*>(a.q + (qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + c);
            q.x *= a.scale; q.y *= a.scale; q.z *= a.scale; q.w *= a.scale;
        }
        *reinterpret_cast<float4*>(Qs + i * AF_PITCH + c) = q;
    }
    {
        const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * Sk : nullptr;
        for (int j = t; j < Sk4; j += T) msk[j] = j < Sk ? (kpm ? kpm[j] : 0) : 1;
    }
    __syncthreads();
    const int tj = Sk4 >> 2, ntile = (Sq4 >> 2) * tj;
    for (int u = t; u < ntile; u += T) {
        const int i0 = (u / tj) << 2, jb = u % tj;
        float s[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) s[r][x] = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
            float4 q[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) q[r] = *reinterpret_cast<const float4*>(Qs + (i0 + r) * AF_PITCH + c);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float4 kk = *reinterpret_cast<const float4*>(Ks + (jb + x * tj) * AF_PITCH + c);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    s[r][x] = fmaf(q[r].x, kk.x, s[r][x]); s[r][x] = fmaf(q[r].y, kk.y, s[r][x]);
                    s[r][x] = fmaf(q[r].z, kk.z, s[r][x]); s[r][x] = fmaf(q[r].w, kk.w, s[r][x]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) Ss[(i0 + r) * SkP + jb + x * tj] = msk[jb + x * tj] ? -INFINITY : s[r][x];
    }
    __syncthreads();
    {
        const int g = t & 3;
        const unsigned gmask = 0xfu << ((t & 31) & ~3);
        for (int row = t >> 2; row < Sq4; row += T / 4) {
            float* sr = Ss + row * SkP;
            float m = -INFINITY;
            for (int j = g; j < Sk; j += 4) m = fmaxf(m, sr[j]);
            m = fmaxf(m, __shfl_xor_sync(gmask, m, 1));
            m = fmaxf(m, __shfl_xor_sync(gmask, m, 2));
            float l = 0.f;
            if (m > -INFINITY) {
                for (int j = g; j < Sk; j += 4) { const float p = __expf(sr[j] - m); sr[j] = p; l += p; }
            } else {
                for (int j = g; j < Sk; j += 4) sr[j] = 0.f;
            }
            l += __shfl_xor_sync(gmask, l, 1);
            l += __shfl_xor_sync(gmask, l, 2);
            if (g == 0) {
                inv[row] = l > 0.f ? 1.f / l : 0.f;
                if (a.lse && row < Sq) a.lse[((long long)b * a.H + h) * Sq + row] = l > 0.f ? m + logf(l) : INFINITY;
            }
        }
    }
    __syncthreads();
    const int nq = (Sq4 >> 2) * 8;
    for (int u = t; u < nq; u += T) {
        const int i0 = (u >> 3) << 2, c = (u & 7) << 2;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[r][x] = 0.f;
        for (int j = 0; j < Sk; ++j) {
            const float4 vv = *reinterpret_cast<const float4*>(Vs + j * AF_PITCH + c);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float p = Ss[(i0 + r) * SkP + j];
                acc[r][0] = fmaf(p, vv.x, acc[r][0]); acc[r][1] = fmaf(p, vv.y, acc[r][1]);
                acc[r][2] = fmaf(p, vv.z, acc[r][2]); acc[r][3] = fmaf(p, vv.w, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + r;
            if (i >= Sq) continue;
            const float iv = inv[i];
            const long long qr = qrow0 + (long long)i * a.q_seq;
            const long long orow = a.o_custom ? attn_row(b, a.o_inner_n, a.o_outer, a.o_inner) + (long long)i * a.o_seq : qr;
            const float4 o4 = make_float4(acc[r][0] * iv, acc[r][1] * iv, acc[r][2] * iv, acc[r][3] * iv);
            if (a.o) *reinterpret_cast<float4*>(a.o + orow * a.ldo + h * HD + c) = o4;
            if (a.o_planes.on()) split4_store(a.o_planes, orow, h * HD + c, o4.x, o4.y, o4.z, o4.w);
        }
    }
}

// ---- head_dim 32, tiny sequences (decoder r2r: 6 reference lines, m2m: 12 modes): ONE WARP per (batch, head) problem, lane =
// channel.  q, k, v of the whole problem live in registers (3 S values per lane, every index a compile-time constant), a score
// is one warp reduction, soft-max and the P V product are replicated / per-lane: no shared memory, no block barrier.
// S = compile-time sequence length (Sq == Sk == S).  Same conventions as the kernels above: masked keys are skipped, a row
// without keys yields 0 and lse = +inf.  EXPERIMENT, off by default (RIFT_B200_ATTN_SMALL=1): the S^2 five-step warp reductions
// make it slower than the cooperative kernels - forward + backward 89 vs 37 us at 3072 problems of 12 x 12, 36 vs 27 us at 6144
// of 6 x 6, 8.65 vs 8.23 ms per step.
template <int S>
__global__ void __launch_bounds__(128)
attention_small_kernel(AttnArgs a) {
    pdl_grid_sync_sel();
    constexpr int HD = 32;
    const int lane = threadIdx.x & 31;
    const long long prob = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (prob >= (long long)a.B * a.H) return;
    const int b = (int)(prob / a.H), h = (int)(prob % a.H);
    const long long krow0 = attn_row(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row(b, a.q_inner_n, a.q_outer, a.q_inner);
    const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * S : nullptr;
    float q[S], k[S], v[S];
    unsigned masked = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        q[i] = a.q[(qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + lane] * a.scale;
        const long long r = krow0 + (long long)i * a.k_seq;
        k[i] = a.k[r * a.ldk + h * HD + lane];
        v[i] = a.v[r * a.ldv + h * HD + lane];
        if (kpm && kpm[i]) masked |= 1u << i;
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        float sc[S];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            sc[j] = warp_sum(q[i] * k[j]);
            if (!((masked >> j) & 1u)) m = fmaxf(m, sc[j]);
        }
        float l = 0.f, o = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const float p = ((masked >> j) & 1u) ? 0.f : __expf(sc[j] - m);
            l += p;
            o = fmaf(p, v[j], o);
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const long long qr = qrow0 + (long long)i * a.q_seq;
        const long long orow = a.o_custom ? attn_row(b, a.o_inner_n, a.o_outer, a.o_inner) + (long long)i * a.o_seq : qr;
        if (a.o) a.o[orow * a.ldo + h * HD + lane] = o * inv;
        if (a.o_planes.on()) split_store(a.o_planes, orow, h * HD + lane, o * inv);
        if (a.lse && lane == 0) a.lse[((long long)b * a.H + h) * S + i] = l > 0.f ? m + logf(l) : INFINITY;
    }
}
static bool attention_small_on() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_ATTN_SMALL"); return e && atoi(e) != 0; }();   // measured slower: off
    return on;
}

int launch_attention(const AttnArgs& a, cudaStream_t st) {
    if (a.B <= 0 || a.Sq <= 0) return 0;
    RIFT_REQUIRE(a.hd == 32 || a.hd == 64, "attention: head_dim must be 32 or 64");
    RIFT_REQUIRE(a.Sk > 0, "attention: empty key sequence");
    {
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool vec = a.hd == 32 && al16(a.q) && al16(a.k) && al16(a.v) && al16(a.o) && (a.ldq & 3) == 0 && (a.ldk & 3) == 0 &&
                         (a.ldv & 3) == 0 && (a.ldo & 3) == 0;
        const int smax = max(a.Sq, a.Sk);
        const size_t per = attn_fwd2_smem_floats(a.Sq, a.Sk) * sizeof(float);
        static const bool legacy = getenv("RIFT_B200_ATTN_FWD_LEGACY") != nullptr;
        if (a.hd == 32 && !legacy && attention_small_on() && a.Sq == a.Sk && (a.Sq == 6 || a.Sq == 12)) {
            const unsigned nb = (unsigned)(((long long)a.B * a.H + 3) / 4);
            if (a.Sq == 6) launch_k(attention_small_kernel<6>, nb, 128, 0, st, a);
            else launch_k(attention_small_kernel<12>, nb, 128, 0, st, a);
            RIFT_LAUNCH_OK();
            return 0;
        }
        static const bool tiled = [] { const char* e = getenv("RIFT_B200_ATTN_FWD_TILED"); return !(e && atoi(e) == 0); }();
        if (vec && !legacy && tiled && min(a.Sq, a.Sk) >= 24 && smax <= 128 && attn_fwd3_smem_floats(a.Sq, a.Sk) * sizeof(float) <= 160 * 1024) {
            static bool attr3 = false;
            if (!attr3) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr3 = true;
            }
            launch_k(attention3_kernel, (unsigned)(a.B * a.H), 256, attn_fwd3_smem_floats(a.Sq, a.Sk) * sizeof(float), st, a);
            RIFT_LAUNCH_OK();
            return 0;
        }
        if (vec && !legacy && 4 * a.Sq <= 512 && per <= 200 * 1024) {
            constexpr int G = 4;
            const int tpp = (G * max(a.Sq, min(smax, 128 / G)) + 31) / 32 * 32;     // enough threads to stage K / V quickly too
            int PB = max(1, 256 / tpp);
            while (PB > 1 && PB * per > 96 * 1024) --PB;
            const long long nprob = (long long)a.B * a.H;
            static bool attr2 = false;
            if (!attr2) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(attention2_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr2 = true;
            }
            launch_k(attention2_kernel<G>, (unsigned)((nprob + PB - 1) / PB), PB * tpp, PB * per, st, a, PB, tpp);
            RIFT_LAUNCH_OK();
            return 0;
        }
    }
    const size_t smem = (size_t)2 * a.Sk * a.hd * sizeof(float);
    RIFT_REQUIRE(smem <= 200 * 1024, "attention: key sequence too long for the shared-memory kernel");
    const int threads = min(128, (a.Sq + 31) / 32 * 32);
    dim3 grid(a.B * a.H, cdiv(a.Sq, 128));
    static bool attr = false;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RIFT_CUDA_OK(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (a.hd == 32) launch_k(attention_kernel<32>, grid, threads, smem, st, a);
    else launch_k(attention_kernel<64>, grid, threads, smem, st, a);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// 1-D neighbourhood attention (NATTEN 0.14 NeighborhoodAttention1D, dilation 1; call site
// layers/embedding.py:169-178).  One warp per (sequence, head); lane = channel.
// window(i) = [clamp(i - k/2, 0, L - k), +k) ; logit = (q_i * hd^-0.5) . k_j + rpb[h, j - i + k - 1]
// qkv row layout: [3][heads][hd]  (qkv.reshape(B, L, 3, heads, hd))
// =====================================================================================
constexpr int NAT_MAXL = 32, NAT_MAXK = 7;

__global__ void __launch_bounds__(128)
nat_attention_kernel(const float* __restrict__ qkv, int n_seq, int L, int heads, int hd, int ksize,
                     const float* __restrict__ rpb, float* __restrict__ out, Planes op) {
    pdl_grid_sync();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_seq * heads) return;
    const int n = warp / heads, h = warp % heads;
    const int dim = heads * hd;
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)n * L * 3 * dim + h * hd;
    const bool on = lane < hd;
    for (int i = 0; i < L; ++i) {
        const int start = min(max(i - ksize / 2, 0), L - ksize);
        const float q = on ? base[(long long)i * 3 * dim + lane] * scale : 0.f;
        float logit[NAT_MAXK];
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < NAT_MAXK; ++kk) {
            if (kk < ksize) {
                const int j = start + kk;
                const float kv = on ? base[(long long)j * 3 * dim + dim + lane] : 0.f;
                logit[kk] = warp_sum(q * kv) + rpb[h * (2 * ksize - 1) + (j - i + ksize - 1)];
                mx = fmaxf(mx, logit[kk]);
            }
        }
        float den = 0.f;
#pragma unroll
        for (int kk = 0; kk < NAT_MAXK; ++kk)
            if (kk < ksize) { logit[kk] = __expf(logit[kk] - mx); den += logit[kk]; }
        float o = 0.f;
#pragma unroll
        for (int kk = 0; kk < NAT_MAXK; ++kk)
            if (kk < ksize && on) o += logit[kk] * base[(long long)(start + kk) * 3 * dim + 2 * dim + lane];
        if (on) {
            if (out) out[((long long)n * L + i) * dim + h * hd + lane] = o / den;
            if (op.on()) split_store(op, (long long)n * L + i, h * hd + lane, o / den);
        }
        if (op.on() && h == 0) split_zero_pad(op, (long long)n * L + i, dim, lane, 32);
    }
}

// Register-resident form for the shapes the history encoder uses (L, ksize compile-time): the whole (sequence,
// head) slice - q, k, v of every position for this lane's channel - is loaded up front (3 L independent 128-byte
// warp loads in flight instead of a dependent load chain per position) and the loops are fully unrolled.
template <int L, int KS>
__global__ void __launch_bounds__(128)
nat_attention_reg_kernel(const float* __restrict__ qkv, int n_seq, int heads, int hd, const float* __restrict__ rpb,
                         float* __restrict__ out, Planes op) {
    pdl_grid_sync();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_seq * heads) return;
    const int n = warp / heads, h = warp % heads;
    const int dim = heads * hd;
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)n * L * 3 * dim + h * hd;
    const bool on = lane < hd;
    float q[L], k[L], v[L], rb[2 * KS - 1];
#pragma unroll
    for (int i = 0; i < L; ++i) {
        q[i] = on ? base[(long long)i * 3 * dim + lane] : 0.f;
        k[i] = on ? base[(long long)i * 3 * dim + dim + lane] : 0.f;
        v[i] = on ? base[(long long)i * 3 * dim + 2 * dim + lane] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 2 * KS - 1; ++t) rb[t] = rpb[h * (2 * KS - 1) + t];
#pragma unroll
    for (int i = 0; i < L; ++i) {
        constexpr int half = KS / 2;
        const int start = (i - half < 0) ? 0 : ((i - half > L - KS) ? L - KS : i - half);
        const float qi = q[i] * scale;
        float logit[KS];
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            logit[kk] = warp_sum(qi * k[start + kk]) + rb[start + kk - i + KS - 1];
            mx = fmaxf(mx, logit[kk]);
        }
        float den = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) { logit[kk] = __expf(logit[kk] - mx); den += logit[kk]; }
        float o = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) o += logit[kk] * v[start + kk];
        if (on) {
            if (out) out[((long long)n * L + i) * dim + h * hd + lane] = o / den;
            if (op.on()) split_store(op, (long long)n * L + i, h * hd + lane, o / den);
        }
        if (op.on() && h == 0) split_zero_pad(op, (long long)n * L + i, dim, lane, 32);
    }
}

int launch_nat_attention(const float* qkv, int n_seq, int L, int heads, int hd, int ksize, const float* rpb, float* out,
                         cudaStream_t st, Planes op) {
    if (n_seq <= 0) return 0;
    RIFT_REQUIRE(hd <= 32 && ksize <= NAT_MAXK && L >= ksize && L <= NAT_MAXL, "nat_attention: unsupported shape");
    const int nblk = cdiv((long long)n_seq * heads * 32, 128);
#define RIFT_NAT_REG(LL, KK)                                                                                            \
    if (L == LL && ksize == KK) {                                                                                       \
        launch_k(nat_attention_reg_kernel<LL, KK>, nblk, 128, 0, st, qkv, n_seq, heads, hd, rpb, out, op);              \
        RIFT_LAUNCH_OK();                                                                                               \
        return 0;                                                                                                       \
    }
    RIFT_NAT_REG(20, 3) RIFT_NAT_REG(10, 3) RIFT_NAT_REG(5, 5)
#undef RIFT_NAT_REG
    launch_k(nat_attention_kernel, cdiv((long long)n_seq * heads * 32, 128), 128, 0, st, qkv, n_seq, L, heads, hd, ksize, rpb, out, op);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// Conv1d(k=3, pad=1) as GEMM: im2col in the weight's own (C_in, 3) flattening so that
// weight.view(C_out, C_in*3) is the GEMM B operand unchanged (layers/embedding.py:90-106,109-120,33-60).
// x is channel-last (n_seq, L, C).
// =====================================================================================
__global__ void im2col_k3_kernel(const float* __restrict__ x, int n_seq, int L, int Lout, int C, int stride,
                                 float* __restrict__ out, Planes op) {
    pdl_grid_sync();
    const int W = op.on() ? op.Kp : C * 3;          // with planes the pad columns are written (as zeros) too
    const long long total = (long long)n_seq * Lout * W;
    FOR_GRID(e, total) {
        const int j = (int)(e % W);
        const long long row = e / W;               // n * Lout + t
        float v = 0.f;
        if (j < C * 3) {
            const int k = j % 3, c = j / 3;
            const int t = (int)(row % Lout);
            const long long n = row / Lout;
            const int ts = t * stride - 1 + k;
            v = (ts >= 0 && ts < L) ? x[(n * L + ts) * C + c] : 0.f;
            if (out) out[row * (C * 3) + j] = v;
        }
        if (op.on()) split_store(op, row, j, v);
    }
}
// planes-only form: one thread = four consecutive columns j (8-byte plane stores instead of four 2-byte ones)
__global__ void __launch_bounds__(256)
im2col_k3_planes4_kernel(const float* __restrict__ x, int n_seq, int L, int Lout, int C, int stride, Planes op) {
    pdl_grid_sync();
    const int W4 = op.Kp >> 2;
    const long long total = (long long)n_seq * Lout * W4;
    FOR_GRID(e, total) {
        const int j0 = (int)(e % W4) << 2;
        const long long row = e / W4;              // n * Lout + t
        const int t = (int)(row % Lout);
        const long long n = row / Lout;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + q;
            const int k = j % 3, c = j / 3;
            const int ts = t * stride - 1 + k;
            v[q] = (j < C * 3 && ts >= 0 && ts < L) ? __ldg(x + (n * L + ts) * C + c) : 0.f;
        }
        split4_store(op, row, j0, v[0], v[1], v[2], v[3]);
    }
}
int launch_im2col_k3(const float* x, int n_seq, int L, int C, int stride, float* out, cudaStream_t st, Planes op) {
    const int Lout = (L + 2 - 3) / stride + 1;
    const long long total = (long long)n_seq * Lout * (op.on() ? op.Kp : C * 3);
    if (total <= 0) return 0;
    if (op.on() && out == nullptr) {
        launch_k(im2col_k3_planes4_kernel, GRID1D(total / 4, 256), 256, 0, st, x, n_seq, L, Lout, C, stride, op);
        RIFT_LAUNCH_OK();
        return 0;
    }
    launch_k(im2col_k3_kernel, GRID1D(total, 256), 256, 0, st, x, n_seq, L, Lout, C, stride, out, op);
    RIFT_LAUNCH_OK();
    return 0;
}
// only the last output position (the encoder keeps out[:, :, -1], layers/embedding.py:87)
__global__ void im2col_k3_last_kernel(const float* __restrict__ x, int n_seq, int L, int C, float* __restrict__ out,
                                      Planes op) {
    pdl_grid_sync();
    const int W = op.on() ? op.Kp : C * 3;
    const long long total = (long long)n_seq * W;
    FOR_GRID(e, total) {
        const int j = (int)(e % W);
        const long long n = e / W;
        float v = 0.f;
        if (j < C * 3) {
            const int k = j % 3, c = j / 3;
            const int ts = L - 2 + k;
            v = (ts >= 0 && ts < L) ? x[(n * L + ts) * C + c] : 0.f;
            if (out) out[n * (C * 3) + j] = v;
        }
        if (op.on()) split_store(op, n, j, v);
    }
}
int launch_im2col_k3_last(const float* x, int n_seq, int L, int C, float* out, cudaStream_t st, Planes op) {
    const long long total = (long long)n_seq * (op.on() ? op.Kp : C * 3);
    if (total <= 0) return 0;
    launch_k(im2col_k3_last_kernel, GRID1D(total, 256), 256, 0, st, x, n_seq, L, C, out, op);
    RIFT_LAUNCH_OK();
    return 0;
}

// FPN top-down: dst += F.interpolate(src, scale_factor=Ld/Ls, mode='linear', align_corners=False)
__global__ void fpn_upsample_add_kernel(float* __restrict__ dst, const float* __restrict__ src, int n_seq, int Ld, int Ls,
                                        int C) {
    pdl_grid_sync();
    const long long total = (long long)n_seq * Ld * C;
    const float rscale = (float)Ls / (float)Ld;
    FOR_GRID(e, total) {
        const int c = (int)(e % C);
        const long long r = e / C;
        const int j = (int)(r % Ld);
        const long long n = r / Ld;
        float sp = ((float)j + 0.5f) * rscale - 0.5f;
        sp = fmaxf(sp, 0.f);
        const int i0 = min((int)sp, Ls - 1);
        const int i1 = min(i0 + 1, Ls - 1);
        const float w1 = sp - (float)i0, w0 = 1.f - w1;
        dst[e] += w0 * src[(n * Ls + i0) * C + c] + w1 * src[(n * Ls + i1) * C + c];
    }
}
__global__ void __launch_bounds__(256)
fpn_upsample_add4_kernel(float* __restrict__ dst, const float* __restrict__ src, int n_seq, int Ld, int Ls, int C) {
    pdl_grid_sync();
    const int C4 = C >> 2;
    const long long total = (long long)n_seq * Ld * C4;
    const float rscale = (float)Ls / (float)Ld;
    FOR_GRID(e, total) {
        const int c = (int)(e % C4) << 2;
        const long long r = e / C4;
        const int j = (int)(r % Ld);
        const long long n = r / Ld;
        float sp = ((float)j + 0.5f) * rscale - 0.5f;
        sp = fmaxf(sp, 0.f);
        const int i0 = min((int)sp, Ls - 1);
        const int i1 = min(i0 + 1, Ls - 1);
        const float w1 = sp - (float)i0, w0 = 1.f - w1;
        const float4 a = *reinterpret_cast<const float4*>(src + (n * Ls + i0) * C + c);
        const float4 b = *reinterpret_cast<const float4*>(src + (n * Ls + i1) * C + c);
        float4* o = reinterpret_cast<float4*>(dst + r * C + c);
        float4 d = *o;
        d.x += w0 * a.x + w1 * b.x; d.y += w0 * a.y + w1 * b.y; d.z += w0 * a.z + w1 * b.z; d.w += w0 * a.w + w1 * b.w;
        *o = d;
    }
}
int launch_fpn_upsample_add(float* dst, const float* src, int n_seq, int Ld, int Ls, int C, cudaStream_t st) {
    const long long total = (long long)n_seq * Ld * C;
    if (total <= 0) return 0;
    if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        launch_k(fpn_upsample_add4_kernel, GRID1D(total / 4, 256), 256, 0, st, dst, src, n_seq, Ld, Ls, C);
        RIFT_LAUNCH_OK();
        return 0;
    }
    launch_k(fpn_upsample_add_kernel, GRID1D(total, 256), 256, 0, st, dst, src, n_seq, Ld, Ls, C);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// PointsEncoder pooling (layers/embedding.py:271-296): features of invalid points are 0 and take
// part in the max; argmax = -1 when the winning value is such a zero (gradient is dropped there).
// =====================================================================================
__global__ void masked_maxpool_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, int groups, int n, int C,
                                      float* __restrict__ out, int* __restrict__ argmax, Planes op) {
    pdl_grid_sync();
    const long long total = (long long)groups * C;
    FOR_GRID(e, total) {
        const int c = (int)(e % C);
        const long long g = e / C;
        float best = -INFINITY;
        int arg = -1;
        for (int p = 0; p < n; ++p) {
            const bool v = mask[g * n + p] != 0;
            const float val = v ? x[(g * n + p) * C + c] : 0.f;
            if (val > best) { best = val; arg = v ? p : -1; }
        }
        if (out) out[e] = best;
        if (op.on()) {
            split_store(op, g, c, best);
            if (c == 0) split_zero_pad(op, g, C, 0, 1);
        }
        if (argmax) argmax[e] = arg;
    }
}
int launch_masked_maxpool(const float* x, const uint8_t* mask, int groups, int n, int C, float* out, int* argmax,
                          cudaStream_t st, Planes op) {
    const long long total = (long long)groups * C;
    if (total <= 0) return 0;
    launch_k(masked_maxpool_kernel, GRID1D(total, 256), 256, 0, st, x, mask, groups, n, C, out, argmax, op);
    RIFT_LAUNCH_OK();
    return 0;
}

__global__ void mask_any_kernel(const uint8_t* __restrict__ mask, int rows, int n, uint8_t* __restrict__ any_out,
                                uint8_t* __restrict__ none_out) {
    pdl_grid_sync();
    FOR_GRID(r, rows) {
        bool a = false;
        for (int i = 0; i < n; ++i) a = a || (mask[r * n + i] != 0);
        if (any_out) any_out[r] = a ? 1 : 0;
        if (none_out) none_out[r] = a ? 0 : 1;
    }
}
int launch_mask_any(const uint8_t* mask, int rows, int n, uint8_t* any_out, uint8_t* none_out, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(mask_any_kernel, GRID1D(rows, 128), 128, 0, st, mask, rows, n, any_out, none_out);
    RIFT_LAUNCH_OK();
    return 0;
}

// key padding (pluto_model.py:136-138): token s of sample b is padded when it has no valid step / point
__global__ void token_masks_kernel(const uint8_t* __restrict__ agent_valid, int agent_T, int Th,
                                   const uint8_t* __restrict__ map_valid, int P, int bs, int A, int Mp,
                                   uint8_t* __restrict__ agent_any, uint8_t* __restrict__ key_pad) {
    pdl_grid_sync();
    const int S = A + Mp;
    FOR_GRID(e, (long long)bs * S) {
        const int s = (int)(e % S);
        const long long b = e / S;
        bool a = false;
        if (s < A) {
            const uint8_t* v = agent_valid + (b * A + s) * agent_T;
            for (int t = 0; t < Th; ++t) a = a || (v[t] != 0);
            agent_any[b * A + s] = a ? 1 : 0;
        } else {
            const uint8_t* v = map_valid + (b * Mp + (s - A)) * P;
            for (int t = 0; t < P; ++t) a = a || (v[t] != 0);
        }
        key_pad[e] = a ? 0 : 1;
    }
}
int launch_token_masks(const uint8_t* agent_valid, int agent_T, int Th, const uint8_t* map_valid, int P, int bs, int A,
                       int Mp, uint8_t* agent_any, uint8_t* key_pad, cudaStream_t st) {
    if (bs <= 0) return 0;
    launch_k(token_masks_kernel, GRID1D((long long)bs * (A + Mp), 128), 128, 0, st, agent_valid, agent_T, Th, map_valid, P, bs, A,
                                                                             Mp, agent_any, key_pad);
    RIFT_LAUNCH_OK();
    return 0;
}

// eval-mode BatchNorm1d folded into the preceding Linear: y = (x W^T) * scale + shift
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, const float* __restrict__ lin_bias, int n,
                               float* __restrict__ scale, float* __restrict__ shift) {
    pdl_grid_sync();
    FOR_GRID(i, n) {
        const float s = gamma[i] / sqrtf(var[i] + 1e-5f);
        scale[i] = s;
        shift[i] = ((lin_bias ? lin_bias[i] : 0.f) - mean[i]) * s + beta[i];
    }
}
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* lin_bias,
                   int n, float* scale, float* shift, cudaStream_t st) {
    launch_k(bn_fold_kernel, cdiv(n, 128), 128, 0, st, gamma, beta, mean, var, lin_bias, n, scale, shift);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// Input feature builders
// =====================================================================================
// AgentEncoder.forward / to_vector (modules/agent_encoder.py:41-76): feat (n_agents, Th-1, 9) channel-last
__global__ void agent_features_kernel(const float* __restrict__ pos, const float* __restrict__ heading,
                                      const float* __restrict__ vel, const float* __restrict__ shape,
                                      const uint8_t* __restrict__ valid, int n_agents, int Th, int Ts,
                                      float* __restrict__ feat) {
    pdl_grid_sync();
    const int L = Th - 1;
    const long long total = (long long)n_agents * L;
    FOR_GRID(e, total) {
        const int t = (int)(e % L);
        const long long n = e / L;
        const long long i0 = n * Ts + t, i1 = i0 + 1;
        const bool vm = valid[i0] && valid[i1];
        float* f = feat + e * 9;
        const float dh = vm ? heading[i1] - heading[i0] : 0.f;
        f[0] = vm ? pos[i1 * 2] - pos[i0 * 2] : 0.f;
        f[1] = vm ? pos[i1 * 2 + 1] - pos[i0 * 2 + 1] : 0.f;
        f[2] = vm ? vel[i1 * 2] - vel[i0 * 2] : 0.f;
        f[3] = vm ? vel[i1 * 2 + 1] - vel[i0 * 2 + 1] : 0.f;
        f[4] = cosf(dh);
        f[5] = sinf(dh);
        f[6] = shape[i1 * 2];
        f[7] = shape[i1 * 2 + 1];
        f[8] = vm ? 1.f : 0.f;
    }
}
int launch_agent_features(const float* pos, const float* heading, const float* vel, const float* shape,
                          const uint8_t* valid, int n_agents, int Th, int Tstride, float* feat, cudaStream_t st) {
    const long long total = (long long)n_agents * (Th - 1);
    if (total <= 0) return 0;
    launch_k(agent_features_kernel, GRID1D(total, 128), 128, 0, st, pos, heading, vel, shape, valid, n_agents, Th, Tstride, feat);
    RIFT_LAUNCH_OK();
    return 0;
}

// MapEncoder polygon feature (modules/map_encoder.py:43-60): (n_poly, P, 10)
__global__ void map_features_kernel(const float* __restrict__ pp, const float* __restrict__ pv, const float* __restrict__ po,
                                    const float* __restrict__ pc, int n_poly, int P, float* __restrict__ feat) {
    pdl_grid_sync();
    const long long total = (long long)n_poly * P;
    FOR_GRID(e, total) {
        const int p = (int)(e % P);
        const long long m = e / P;
        const long long s0 = (m * 3 + 0) * P + p, s1 = (m * 3 + 1) * P + p, s2 = (m * 3 + 2) * P + p;
        float* f = feat + e * 10;
        const float x0 = pp[s0 * 2], y0 = pp[s0 * 2 + 1];
        f[0] = x0 - pc[m * 3];
        f[1] = y0 - pc[m * 3 + 1];
        f[2] = pv[s0 * 2];
        f[3] = pv[s0 * 2 + 1];
        f[4] = cosf(po[s0]);
        f[5] = sinf(po[s0]);
        f[6] = pp[s1 * 2] - x0;
        f[7] = pp[s1 * 2 + 1] - y0;
        f[8] = pp[s2 * 2] - x0;
        f[9] = pp[s2 * 2 + 1] - y0;
    }
}
int launch_map_features(const float* point_position, const float* point_vector, const float* point_orientation,
                        const float* polygon_center, int n_poly, int P, float* feat, cudaStream_t st) {
    const long long total = (long long)n_poly * P;
    if (total <= 0) return 0;
    launch_k(map_features_kernel, GRID1D(total, 128), 128, 0, st, point_position, point_vector, point_orientation, polygon_center,
                                                            n_poly, P, feat);
    RIFT_LAUNCH_OK();
    return 0;
}

// PlanningDecoder reference-line feature (modules/planning_decoder.py:137-160): (n_ref, Pr, 6) and
// r_pos (n_ref, 3) = [position[0], orientation[0]]
__global__ void ref_features_kernel(const float* __restrict__ pos, const float* __restrict__ vec, const float* __restrict__ ori,
                                    int n_ref, int Pr, float* __restrict__ feat, float* __restrict__ rpos) {
    pdl_grid_sync();
    const long long total = (long long)n_ref * Pr;
    FOR_GRID(e, total) {
        const int p = (int)(e % Pr);
        const long long r = e / Pr;
        float* f = feat + e * 6;
        const float x0 = pos[(r * Pr) * 2], y0 = pos[(r * Pr) * 2 + 1];
        f[0] = pos[e * 2] - x0;
        f[1] = pos[e * 2 + 1] - y0;
        f[2] = vec[e * 2];
        f[3] = vec[e * 2 + 1];
        f[4] = cosf(ori[e]);
        f[5] = sinf(ori[e]);
        if (p == 0) { rpos[r * 3] = x0; rpos[r * 3 + 1] = y0; rpos[r * 3 + 2] = ori[e]; }
    }
}
int launch_ref_features(const float* position, const float* vector, const float* orientation, int n_ref, int Pr,
                        float* feat, float* rpos, cudaStream_t st) {
    const long long total = (long long)n_ref * Pr;
    if (total <= 0) return 0;
    launch_k(ref_features_kernel, GRID1D(total, 128), 128, 0, st, position, vector, orientation, n_ref, Pr, feat, rpos);
    RIFT_LAUNCH_OK();
    return 0;
}

// token positions for pos_emb (pluto_model.py:123-135): [x, y, wrap(angle)], angle wrapped like
// (angle + pi) % (2 pi) - pi with torch's floor-mod semantics
__device__ __forceinline__ float wrap_angle(float a) {
    const float two_pi = 6.283185307179586f;
    const float x = a + 3.141592653589793f;
    float m = fmodf(x, two_pi);
    if (m != 0.f && (m < 0.f)) m += two_pi;
    return m - 3.141592653589793f;
}
__global__ void token_pos_kernel(const float* __restrict__ apos, const float* __restrict__ ahead, const float* __restrict__ pc,
                                 int bs, int A, int Th, int Ts, int Mp, float* __restrict__ pos) {
    pdl_grid_sync();
    const int S = A + Mp;
    FOR_GRID(e, (long long)bs * S) {
        const int s = (int)(e % S);
        const long long b = e / S;
        float x, y, ang;
        if (s < A) {
            const long long i = (b * A + s) * Ts + (Th - 1);
            x = apos[i * 2]; y = apos[i * 2 + 1]; ang = ahead[i];
        } else {
            const long long i = b * Mp + (s - A);
            x = pc[i * 3]; y = pc[i * 3 + 1]; ang = pc[i * 3 + 2];
        }
        pos[e * 3] = x; pos[e * 3 + 1] = y; pos[e * 3 + 2] = wrap_angle(ang);
    }
}
int launch_token_pos(const float* agent_pos, const float* agent_heading, const float* polygon_center, int bs, int A,
                     int Th, int Tstride, int Mp, float* pos, cudaStream_t st) {
    if (bs <= 0) return 0;
    launch_k(token_pos_kernel, GRID1D((long long)bs * (A + Mp), 128), 128, 0, st, agent_pos, agent_heading, polygon_center, bs, A,
                                                                           Th, Tstride, Mp, pos);
    RIFT_LAUNCH_OK();
    return 0;
}

// FourierEmbedding input features (layers/fourier_embedding.py:45-50) for input dimension `dsel`:
// [cos(x f 2 pi) (nfreq), sin(x f 2 pi) (nfreq), x], zero padded to ldf columns
__global__ void fourier_features_kernel(const float* __restrict__ x, int rows, int d, int dsel, const float* __restrict__ freqs,
                                        int nfreq, float* __restrict__ feat, int ldfeat, Planes op) {
    pdl_grid_sync();
    const int ldf = op.on() ? op.Kp : ldfeat;          // loop width: the planes' zero pad is written too
    FOR_GRID(e, (long long)rows * ldf) {
        const int c = (int)(e % ldf);
        const long long r = e / ldf;
        const float xv = x[r * d + dsel];
        float v = 0.f;
        if (c < 2 * nfreq) {
            const int j = c < nfreq ? c : c - nfreq;
            const float ang = ((xv * freqs[dsel * nfreq + j]) * 2.f) * 3.141592653589793f;
            v = c < nfreq ? cosf(ang) : sinf(ang);
        } else if (c == 2 * nfreq) {
            v = xv;
        }
        if (op.on()) split_store(op, r, c, v);
        if (feat && c < ldfeat) feat[r * ldfeat + c] = v;
    }
}
int launch_fourier_features(const float* x, int rows, int d, int dsel, const float* freqs, int nfreq, float* feat,
                            int ldf, cudaStream_t st, Planes op) {
    if (rows <= 0) return 0;
    RIFT_REQUIRE(ldf >= 2 * nfreq + 1, "fourier_features: ldf too small");
    const int width = op.on() ? op.Kp : ldf;
    launch_k(fourier_features_kernel, GRID1D((long long)rows * width, 256), 256, 0, st, x, rows, d, dsel, freqs, nfreq, feat, ldf, op);
    RIFT_LAUNCH_OK();
    return 0;
}

// StateAttentionEncoder tokens (modules/agent_encoder.py:116-122): tok[b,i,:] = x[b,i] * W_i[:,0] + b_i + pos_embed[i]
struct StateTokParams { const float* w[8]; const float* b[8]; };
__global__ void state_tokens_kernel(const float* __restrict__ cur, int cs_stride, int bs, int n_tok, int D, StateTokParams p,
                                    const float* __restrict__ pos_embed, float* __restrict__ toks) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)bs * n_tok * D) {
        const int c = (int)(e % D);
        const long long r = e / D;
        const int i = (int)(r % n_tok);
        const long long b = r / n_tok;
        toks[e] = cur[b * cs_stride + i] * p.w[i][c] + p.b[i][c] + pos_embed[i * D + c];
    }
}
int launch_state_tokens(const float* cur_state, int cs_stride, int bs, int n_tok, int D, const float* const* w,
                        const float* const* b, const float* pos_embed, float* toks, cudaStream_t st) {
    RIFT_REQUIRE(n_tok <= 8, "state_tokens: at most 8 state channels");
    if (bs <= 0) return 0;
    StateTokParams p;
    for (int i = 0; i < 8; ++i) { p.w[i] = i < n_tok ? w[i] : nullptr; p.b[i] = i < n_tok ? b[i] : nullptr; }
    launch_k(state_tokens_kernel, GRID1D((long long)bs * n_tok * D, 256), 256, 0, st, cur_state, cs_stride, bs, n_tok, D, p,
                                                                               pos_embed, toks);
    RIFT_LAUNCH_OK();
    return 0;
}

// agent tokens (modules/agent_encoder.py:78-94): history embedding for valid agents, 0 otherwise,
// token 0 <- ego state embedding, + type_emb[category]; written into tokens[b, a, :] of (bs, S, D)
__global__ void agent_assemble_kernel(const float* __restrict__ x_hist, const float* __restrict__ x_ego,
                                      const uint8_t* __restrict__ agent_any, const int8_t* __restrict__ category,
                                      const float* __restrict__ type_emb, int bs, int A, int S, int D,
                                      float* __restrict__ tokens) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)bs * A * D) {
        const int c = (int)(e % D);
        const long long r = e / D;
        const int a = (int)(r % A);
        const long long b = r / A;
        float v;
        if (a == 0) v = x_ego[b * D + c];
        else v = agent_any[r] ? x_hist[r * D + c] : 0.f;
        v += type_emb[(int)category[r] * D + c];
        tokens[(b * S + a) * D + c] = v;
    }
}
int launch_agent_assemble(const float* x_hist, const float* x_ego, const uint8_t* agent_any, const int8_t* category,
                          const float* type_emb, int bs, int A, int S, int D, float* tokens, cudaStream_t st) {
    if (bs <= 0) return 0;
    launch_k(agent_assemble_kernel, GRID1D((long long)bs * A * D, 256), 256, 0, st, x_hist, x_ego, agent_any, category, type_emb,
                                                                             bs, A, S, D, tokens);
    RIFT_LAUNCH_OK();
    return 0;
}

// map tokens (modules/map_encoder.py:79-93)
__global__ void map_assemble_kernel(const float* __restrict__ x_poly, const float* __restrict__ x_speed,
                                    const int8_t* __restrict__ ptype, const uint8_t* __restrict__ on_route,
                                    const int8_t* __restrict__ tl, const uint8_t* __restrict__ has_speed,
                                    const float* __restrict__ type_emb, const float* __restrict__ route_emb,
                                    const float* __restrict__ tl_emb, const float* __restrict__ unknown_emb, int bs, int Mp,
                                    int A, int S, int D, float* __restrict__ tokens) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)bs * Mp * D) {
        const int c = (int)(e % D);
        const long long r = e / D;
        const int m = (int)(r % Mp);
        const long long b = r / Mp;
        float v = x_poly[e];
        v += type_emb[(int)ptype[r] * D + c] + route_emb[(on_route[r] ? 1 : 0) * D + c] + tl_emb[(int)tl[r] * D + c] +
             (has_speed[r] ? x_speed[e] : unknown_emb[c]);
        tokens[(b * S + A + m) * D + c] = v;
    }
}
int launch_map_assemble(const float* x_poly, const float* x_speed, const int8_t* ptype, const uint8_t* on_route,
                        const int8_t* tl, const uint8_t* has_speed, const float* type_emb, const float* route_emb,
                        const float* tl_emb, const float* unknown_emb, int bs, int Mp, int A, int S, int D, float* tokens,
                        cudaStream_t st) {
    if (bs <= 0 || Mp <= 0) return 0;
    launch_k(map_assemble_kernel, GRID1D((long long)bs * Mp * D, 256), 256, 0, st, x_poly, x_speed, ptype, on_route, tl, has_speed,
                                                                            type_emb, route_emb, tl_emb, unknown_emb, bs,
                                                                            Mp, A, S, D, tokens);
    RIFT_LAUNCH_OK();
    return 0;
}

// decoder query init: q[row, :] = u[row / Mo, :] + v[row % Mo, :]
// (q_proj(cat[r_emb, m_emb]) split into its two column blocks, modules/planning_decoder.py:162-167)
__global__ void query_init_kernel(const float* __restrict__ u, const float* __restrict__ v, long long rows, int Mo, int D,
                                  float* __restrict__ q) {
    pdl_grid_sync();
    FOR_GRID(e, rows * D) {
        const int c = (int)(e % D);
        const long long r = e / D;
        q[e] = u[(r / Mo) * D + c] + v[(r % Mo) * D + c];
    }
}
int launch_query_init(const float* u, const float* v, int rows, int Mo, int D, float* q, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(query_init_kernel, GRID1D((long long)rows * D, 256), 256, 0, st, u, v, rows, Mo, D, q);
    RIFT_LAUNCH_OK();
    return 0;
}

__global__ void zero_rows_kernel(float* __restrict__ x, const uint8_t* __restrict__ flag, int div, long long rows, int C) {
    pdl_grid_sync();
    FOR_GRID(e, rows * C) {
        if (flag[(e / C) / div]) x[e] = 0.f;
    }
}
int launch_zero_rows(float* x, const uint8_t* rowflag, int flag_div, int rows, int C, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(zero_rows_kernel, GRID1D((long long)rows * C, 256), 256, 0, st, x, rowflag, flag_div, rows, C);
    RIFT_LAUNCH_OK();
    return 0;
}

// trajectory = cat([loc, yaw, vel], -1) with each head viewed (.., T, 2)  (modules/planning_decoder.py:180-186)
__global__ void interleave_heads_kernel(const float* __restrict__ loc, const float* __restrict__ yaw, const float* __restrict__ vel,
                                        long long rows, int T, float* __restrict__ out) {
    pdl_grid_sync();
    FOR_GRID(e, rows * T * 6) {
        const int c = (int)(e % 6);
        const long long rt = e / 6;                  // row * T + t
        const float* src = c < 2 ? loc : (c < 4 ? yaw : vel);
        out[e] = src[rt * 2 + (c & 1)];
    }
}
int launch_interleave_heads(const float* loc, const float* yaw, const float* vel, long long rows, int T, float* out,
                            cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(interleave_heads_kernel, GRID1D(rows * T * 6, 256), 256, 0, st, loc, yaw, vel, rows, T, out);
    RIFT_LAUNCH_OK();
    return 0;
}

// probability.masked_fill_(r_padding_mask, -1e6)  (pluto_model.py:203); pi is (rows=bs*R, Mo)
__global__ void mask_logits_kernel(float* __restrict__ pi, const uint8_t* __restrict__ r_pad, long long rows, int Mo, float fill) {
    pdl_grid_sync();
    FOR_GRID(e, rows * Mo) {
        if (r_pad[e / Mo]) pi[e] = fill;
    }
}
int launch_mask_logits(float* pi, const uint8_t* r_pad, long long rows, int Mo, float fill, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(mask_logits_kernel, GRID1D(rows * Mo, 256), 256, 0, st, pi, r_pad, rows, Mo, fill);
    RIFT_LAUNCH_OK();
    return 0;
}

// out[b*nrows + i, :] = x[b*ldx_batch + (row0+i)*C + :]
__global__ void gather_rows_kernel(const float* __restrict__ x, long long ldb, int bs, int row0, int nrows, int C,
                                   float* __restrict__ out) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)bs * nrows * C) {
        const int c = (int)(e % C);
        const long long r = e / C;
        const int i = (int)(r % nrows);
        const long long b = r / nrows;
        out[e] = x[b * ldb + (long long)(row0 + i) * C + c];
    }
}
int launch_gather_rows(const float* x, long long ldx_batch, int bs, int row0, int nrows, int C, float* out, cudaStream_t st) {
    if (bs <= 0 || nrows <= 0) return 0;
    launch_k(gather_rows_kernel, GRID1D((long long)bs * nrows * C, 256), 256, 0, st, x, ldx_batch, bs, row0, nrows, C, out);
    RIFT_LAUNCH_OK();
    return 0;
}

// dy *= act'(.) in place.  ReLU: ref = post-activation (y > 0) ; GELU: ref = pre-activation
__global__ void act_bwd_kernel(const float* __restrict__ ref, float* __restrict__ dy, long long n, int act) {
    pdl_grid_sync();
    FOR_GRID(e, n) {
        if (act == ACT_RELU) { if (!(ref[e] > 0.f)) dy[e] = 0.f; }
        else if (act == ACT_GELU) dy[e] *= gelu_erf_grad(ref[e]);
    }
}
// four elements per thread; with `planes` the product leaves as split-bf16 planes (flat: pitch == row length) INSTEAD of
// being written back to dy - the consumer is a GEMM that reads planes only
__device__ __forceinline__ float4 act_grad4(float4 d, const float4 r, int act) {
    if (act == ACT_RELU) {
        if (!(r.x > 0.f)) d.x = 0.f; if (!(r.y > 0.f)) d.y = 0.f; if (!(r.z > 0.f)) d.z = 0.f; if (!(r.w > 0.f)) d.w = 0.f;
    } else {
        d.x *= gelu_erf_grad(r.x); d.y *= gelu_erf_grad(r.y); d.z *= gelu_erf_grad(r.z); d.w *= gelu_erf_grad(r.w);
    }
    return d;
}
__device__ __forceinline__ void split4_flat(uint2* __restrict__ hi, uint2* __restrict__ lo, long long e, const float4 d) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(d.x), h1 = __float2bfloat16_rn(d.y), h2 = __float2bfloat16_rn(d.z), h3 = __float2bfloat16_rn(d.w);
    const __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    const __nv_bfloat162 cc = __floats2bfloat162_rn(d.x - __bfloat162float(h0), d.y - __bfloat162float(h1));
    const __nv_bfloat162 dd = __floats2bfloat162_rn(d.z - __bfloat162float(h2), d.w - __bfloat162float(h3));
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&a); uh.y = *reinterpret_cast<const uint32_t*>(&b);
    ul.x = *reinterpret_cast<const uint32_t*>(&cc); ul.y = *reinterpret_cast<const uint32_t*>(&dd);
    hi[e] = uh; lo[e] = ul;
}
__global__ void __launch_bounds__(256)
act_bwd4_kernel(const float4* __restrict__ ref, float4* __restrict__ dy, long long n4, int act, uint2* __restrict__ hi, uint2* __restrict__ lo) {
    pdl_grid_sync_sel();
    const long long stride = (long long)gridDim.x * 256;
    long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    for (; e + stride < n4; e += 2 * stride) {            // two independent 16-byte streams per thread in flight
        const float4 r0 = __ldg(ref + e), r1 = __ldg(ref + e + stride);
        const float4 d0 = act_grad4(dy[e], r0, act), d1 = act_grad4(dy[e + stride], r1, act);
        if (hi) { split4_flat(hi, lo, e, d0); split4_flat(hi, lo, e + stride, d1); }
        else { dy[e] = d0; dy[e + stride] = d1; }
    }
    if (e < n4) {
        const float4 d0 = act_grad4(dy[e], __ldg(ref + e), act);
        if (hi) split4_flat(hi, lo, e, d0); else dy[e] = d0;
    }
}
int launch_act_bwd(const float* pre_or_post, float* dy, long long n, int act, cudaStream_t st, Planes planes) {
    if (n <= 0) return 0;
    RIFT_REQUIRE(act == ACT_RELU || act == ACT_GELU || (act == ACT_NONE && !planes.on()), "act_bwd: unknown activation");
    const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(pre_or_post) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
    RIFT_REQUIRE(!planes.on() || vec, "act_bwd: plane output needs 16-byte aligned rows of a multiple of 4 elements");
    if (act == ACT_NONE) return 0;
    if (vec) {
        launch_k(act_bwd4_kernel, (int)min((long long)148 * 8, (n / 4 + 255) / 256), 256, 0, st, reinterpret_cast<const float4*>(pre_or_post),
                 reinterpret_cast<float4*>(dy), n / 4, act, reinterpret_cast<uint2*>(planes.hi), reinterpret_cast<uint2*>(planes.lo));
    } else {
        launch_k(act_bwd_kernel, GRID1D(n, 256), 256, 0, st, pre_or_post, dy, n, act);
    }
    RIFT_LAUNCH_OK();
    return 0;
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    pdl_grid_sync();
    FOR_GRID(e, n) dst[e] += src[e];
}
int launch_add_inplace(float* dst, const float* src, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    launch_k(add_inplace_kernel, GRID1D(n, 256), 256, 0, st, dst, src, n);
    RIFT_LAUNCH_OK();
    return 0;
}

// candidate_trajectories = cat([traj[..., :2], atan2(traj[..., 3], traj[..., 2])])  (pluto_model.py:205-212)
__global__ void traj_outputs_kernel(const float* __restrict__ traj, long long n, float* __restrict__ cand) {
    pdl_grid_sync();
    FOR_GRID(e, n) {
        const float* t = traj + e * 6;
        cand[e * 3] = t[0];
        cand[e * 3 + 1] = t[1];
        cand[e * 3 + 2] = atan2f(t[3], t[2]);
    }
}
int launch_traj_outputs(const float* trajectory, long long n_traj, int T, float* cand, cudaStream_t st) {
    const long long n = n_traj * T;
    if (n <= 0) return 0;
    launch_k(traj_outputs_kernel, GRID1D(n, 256), 256, 0, st, trajectory, n, cand);
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift
