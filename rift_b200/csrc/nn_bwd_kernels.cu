// Backward kernels of the non-GEMM operators (see nn_kernels.cu for the forward definitions).
// All reductions are two-stage / fixed-order, so gradients are run-to-run deterministic.
#include <stdlib.h>

#include "common.cuh"
#include "ops.h"

namespace rift {

#define GRID1D(n, t) (int)min((long long)148 * 16, ((long long)(n) + (t) - 1) / (t))
#define FOR_GRID(i, n) \
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)(n); i += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ long long attn_row_b(int b, int inner_n, long long outer, long long inner) {
    return (long long)(b / inner_n) * outer + (long long)(b % inner_n) * inner;
}

// =====================================================================================
// Multi-head attention backward (recompute form).  One CTA per (batch, head); Q, K, V, dO of the head are
// staged in shared memory.  p_ij = exp(s_ij - lse_i), delta_i = sum_j p_ij dP_ij, dS = p (dP - delta).
//   phase A (thread = query i): delta_i and dQ_i = scale * sum_j dS_ij K_j
//   phase B (thread = key j)  : dV_j = sum_i p_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
// Gradients are written with the operand's own row mapping (each element exactly once).  With a shared
// query (q_outer = 0, o_custom) dQ is written per batch row of `dq` using the OUTPUT mapping.
// =====================================================================================
template <int HD>
__device__ __forceinline__ float dot_bcast(const float (&r)[HD], const float* __restrict__ s) {
    const float4* s4 = reinterpret_cast<const float4*>(s);          // broadcast 16-byte shared loads
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
        const float4 t = s4[d];
        a = fmaf(r[4 * d], t.x, a); a = fmaf(r[4 * d + 1], t.y, a); a = fmaf(r[4 * d + 2], t.z, a); a = fmaf(r[4 * d + 3], t.w, a);
    }
    return a;
}
template <int HD>
__device__ __forceinline__ void axpy_bcast(float (&acc)[HD], float w, const float* __restrict__ s) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
        const float4 t = s4[d];
        acc[4 * d] = fmaf(w, t.x, acc[4 * d]); acc[4 * d + 1] = fmaf(w, t.y, acc[4 * d + 1]);
        acc[4 * d + 2] = fmaf(w, t.z, acc[4 * d + 2]); acc[4 * d + 3] = fmaf(w, t.w, acc[4 * d + 3]);
    }
}

// The thread's own row (query in phase A, key in phase B) lives in registers; the other operand is read from
// shared memory with warp-wide broadcast loads, so there are no bank conflicts.
template <int HD>
__global__ void __launch_bounds__(128)
attention_bwd_kernel(AttnArgs a, const float* __restrict__ d_o, long long lddo, float* __restrict__ dq, long long lddq,
                     float* __restrict__ dk, float* __restrict__ dv, long long lddk, long long lddv) {
    pdl_grid_sync();
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                                  // [Sq][HD]  (pre-scaled)
    float* dOs = Qs + (size_t)a.Sq * HD;             // [Sq][HD]
    float* Ks = dOs + (size_t)a.Sq * HD;             // [Sk][HD]
    float* Vs = Ks + (size_t)a.Sk * HD;              // [Sk][HD]
    float* delta = Vs + (size_t)a.Sk * HD;           // [Sq]
    float* lse_s = delta + a.Sq;                     // [Sq]
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const long long krow0 = attn_row_b(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row_b(b, a.q_inner_n, a.q_outer, a.q_inner);
    const long long orow0 = a.o_custom ? attn_row_b(b, a.o_inner_n, a.o_outer, a.o_inner) : qrow0;
    const long long oseq = a.o_custom ? a.o_seq : a.q_seq;
    for (int e = threadIdx.x; e < a.Sk * HD; e += blockDim.x) {
        const int j = e / HD, d = e - j * HD;
        const long long r = krow0 + (long long)j * a.k_seq;
        Ks[e] = a.k[r * a.ldk + h * HD + d];
        Vs[e] = a.v[r * a.ldv + h * HD + d];
    }
    for (int e = threadIdx.x; e < a.Sq * HD; e += blockDim.x) {
        const int i = e / HD, d = e - i * HD;
        Qs[e] = a.q[(qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + d] * a.scale;
        dOs[e] = d_o[(orow0 + (long long)i * oseq) * lddo + h * HD + d];
    }
    for (int i = threadIdx.x; i < a.Sq; i += blockDim.x) lse_s[i] = a.lse[((long long)b * a.H + h) * a.Sq + i];
    __syncthreads();
    const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * a.Sk : nullptr;
    // ---- phase A: thread = query i
    for (int i = threadIdx.x; i < a.Sq; i += blockDim.x) {
        const float lse = lse_s[i];
        float q[HD], go[HD], acc[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = Qs[i * HD + d]; go[d] = dOs[i * HD + d]; acc[d] = 0.f; }
        float dl = 0.f;
        if (lse != INFINITY) {
            for (int j = 0; j < a.Sk; ++j) {
                if (kpm && kpm[j]) continue;
                dl += __expf(dot_bcast<HD>(q, Ks + j * HD) - lse) * dot_bcast<HD>(go, Vs + j * HD);
            }
            for (int j = 0; j < a.Sk; ++j) {
                if (kpm && kpm[j]) continue;
                const float ds = __expf(dot_bcast<HD>(q, Ks + j * HD) - lse) * (dot_bcast<HD>(go, Vs + j * HD) - dl);
                axpy_bcast<HD>(acc, ds, Ks + j * HD);
            }
        }
        delta[i] = dl;
        const long long dr = a.o_custom ? (orow0 + (long long)i * oseq) : (qrow0 + (long long)i * a.q_seq);
#pragma unroll
        for (int d = 0; d < HD; ++d) dq[dr * lddq + h * HD + d] = acc[d] * a.scale;
    }
    __syncthreads();
    // ---- phase B: thread = key j
    for (int j = threadIdx.x; j < a.Sk; j += blockDim.x) {
        float kr[HD], vr[HD], ak[HD], av[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) { kr[d] = Ks[j * HD + d]; vr[d] = Vs[j * HD + d]; ak[d] = 0.f; av[d] = 0.f; }
        if (!(kpm && kpm[j])) {
            for (int i = 0; i < a.Sq; ++i) {
                const float lse = lse_s[i];
                if (lse == INFINITY) continue;
                const float p = __expf(dot_bcast<HD>(kr, Qs + i * HD) - lse);
                const float ds = p * (dot_bcast<HD>(vr, dOs + i * HD) - delta[i]);
                axpy_bcast<HD>(av, p, dOs + i * HD);
                axpy_bcast<HD>(ak, ds, Qs + i * HD);              // Qs is pre-scaled, so dK already carries `scale`
            }
        }
        const long long r = krow0 + (long long)j * a.k_seq;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            dk[r * lddk + h * HD + d] = ak[d];
            dv[r * lddv + h * HD + d] = av[d];
        }
    }
}

// ---- head_dim 32, cooperative form ------------------------------------------------------------------
// Several (batch, head) problems per CTA (PB) and G lanes per row, each lane owning 32 / G CHANNELS.
//   phase A (lane group = query i): per key a partial dot product + log2(G) shuffles give s_ij and dP_ij;
//       P = exp(S - lse) and dP are written to shared memory once, delta_i = sum_j P dP, and in a second sweep
//       dQ_i[channels of the lane] += dS_ij K_j with dS = P (dP - delta).
//   phase B (lane group = key j): dV_j += P_ij dO_i, dK_j += dS_ij Q_i on the lane's channels - no reduction at all.
// Rows are padded to 36 floats (16-byte aligned); few registers per thread, several CTAs per SM.
constexpr int AB_PITCH = 36;

__host__ __device__ inline size_t attn_bwd2_smem_floats(int Sq, int Sk) {
    const int Skp = Sk | 1;
    const size_t n = (size_t)2 * Sq * AB_PITCH + (size_t)2 * Sk * AB_PITCH + (size_t)2 * Sq * Skp + (size_t)2 * Sq + ((Sk + 3) / 4);
    return (n + 3) & ~(size_t)3;                         // problems stay 16-byte aligned
}

template <int G>
__global__ void __launch_bounds__(512)
attention_bwd2_kernel(AttnArgs a, const float* __restrict__ d_o, long long lddo, float* __restrict__ dq, long long lddq,
                      float* __restrict__ dk, float* __restrict__ dv, long long lddk, long long lddv, int PB, int tpp, Planes pq, Planes pk,
                      Planes pv) {
    pdl_grid_sync_sel();
    constexpr int HD = 32, CPL = HD / G;
    extern __shared__ __align__(16) float sm[];
    const int Sq = a.Sq, Sk = a.Sk, Skp = Sk | 1;
    const int pl = threadIdx.x / tpp;                    // problem slot inside the CTA
    const int t = threadIdx.x - pl * tpp;                // thread inside the problem
    const long long prob = (long long)blockIdx.x * PB + pl;
    const bool live = prob < (long long)a.B * a.H;
    float* base = sm + (size_t)pl * attn_bwd2_smem_floats(Sq, Sk);
    float* Qs = base;                                    // [Sq][36]  (pre-scaled)
    float* dOs = Qs + (size_t)Sq * AB_PITCH;             // [Sq][36]
    float* Ks = dOs + (size_t)Sq * AB_PITCH;             // [Sk][36]
    float* Vs = Ks + (size_t)Sk * AB_PITCH;              // [Sk][36]
    float* Ps = Vs + (size_t)Sk * AB_PITCH;              // [Sq][Skp]
    float* dPs = Ps + (size_t)Sq * Skp;                  // [Sq][Skp]
    float* lse_s = dPs + (size_t)Sq * Skp;               // [Sq]
    float* delta = lse_s + Sq;                           // [Sq]
    uint8_t* msk = reinterpret_cast<uint8_t*>(delta + Sq);   // [Sk]
    const int b = live ? (int)(prob / a.H) : 0, h = live ? (int)(prob % a.H) : 0;
    const long long krow0 = attn_row_b(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row_b(b, a.q_inner_n, a.q_outer, a.q_inner);
    const long long orow0 = a.o_custom ? attn_row_b(b, a.o_inner_n, a.o_outer, a.o_inner) : qrow0;
    const long long oseq = a.o_custom ? a.o_seq : a.q_seq;
    if (live) {
        // 16-byte loads: 8 lanes per row
        for (int e = t; e < Sk * 8; e += tpp) {
            const int j = e >> 3, c = (e & 7) * 4;
            const long long r = krow0 + (long long)j * a.k_seq;
            *reinterpret_cast<float4*>(Ks + j * AB_PITCH + c) = *reinterpret_cast<const float4*>(a.k + r * a.ldk + h * HD + c);
            *reinterpret_cast<float4*>(Vs + j * AB_PITCH + c) = *reinterpret_cast<const float4*>(a.v + r * a.ldv + h * HD + c);
        }
        for (int e = t; e < Sq * 8; e += tpp) {
            const int i = e >> 3, c = (e & 7) * 4;
            float4 q = *reinterpret_cast<const float4*>(a.q + (qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + c);
            q.x *= a.scale; q.y *= a.scale; q.z *= a.scale; q.w *= a.scale;
            *reinterpret_cast<float4*>(Qs + i * AB_PITCH + c) = q;
            *reinterpret_cast<float4*>(dOs + i * AB_PITCH + c) =
                *reinterpret_cast<const float4*>(d_o + (orow0 + (long long)i * oseq) * lddo + h * HD + c);
        }
        for (int i = t; i < Sq; i += tpp) lse_s[i] = a.lse[((long long)b * a.H + h) * Sq + i];
        const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * Sk : nullptr;
        for (int j = t; j < Sk; j += tpp) msk[j] = kpm ? kpm[j] : 0;
    }
    __syncthreads();
    const int row = t / G, g = t % G, c0 = g * CPL;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((threadIdx.x & 31) & ~(G - 1));
    // ---- phase A: lane group = query i (whole groups skip together; the shuffles name the group only)
    if (live && row < Sq) {
        const int i = row;
        const float lse = lse_s[i];
        const bool dead = (lse == INFINITY);
        float q[CPL], go[CPL], acc[CPL];
#pragma unroll
        for (int d = 0; d < CPL; d += 4) {
            const float4 x = *reinterpret_cast<const float4*>(Qs + i * AB_PITCH + c0 + d);
            const float4 y = *reinterpret_cast<const float4*>(dOs + i * AB_PITCH + c0 + d);
            q[d] = x.x; q[d + 1] = x.y; q[d + 2] = x.z; q[d + 3] = x.w;
            go[d] = y.x; go[d + 1] = y.y; go[d + 2] = y.z; go[d + 3] = y.w;
            acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
        }
        float dl = 0.f;
        for (int j = 0; j < Sk; ++j) {
            float sc = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < CPL; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AB_PITCH + c0 + d);
                const float4 vv = *reinterpret_cast<const float4*>(Vs + j * AB_PITCH + c0 + d);
                sc = fmaf(q[d], kk.x, sc); sc = fmaf(q[d + 1], kk.y, sc); sc = fmaf(q[d + 2], kk.z, sc); sc = fmaf(q[d + 3], kk.w, sc);
                dp = fmaf(go[d], vv.x, dp); dp = fmaf(go[d + 1], vv.y, dp); dp = fmaf(go[d + 2], vv.z, dp); dp = fmaf(go[d + 3], vv.w, dp);
            }
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                sc += __shfl_xor_sync(gmask, sc, o);
                dp += __shfl_xor_sync(gmask, dp, o);
            }
            const float p = (dead || msk[j]) ? 0.f : __expf(sc - lse);
            if ((j & (G - 1)) == g) { Ps[i * Skp + j] = p; dPs[i * Skp + j] = dp; }
            dl = fmaf(p, dp, dl);
        }
        if (g == 0) delta[i] = dl;
        __syncwarp(gmask);                               // P / dP written by the other lanes of the group
        for (int j = 0; j < Sk; ++j) {
            const float ds = Ps[i * Skp + j] * (dPs[i * Skp + j] - dl);
#pragma unroll
            for (int d = 0; d < CPL; d += 4) {
                const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AB_PITCH + c0 + d);
                acc[d] = fmaf(ds, kk.x, acc[d]); acc[d + 1] = fmaf(ds, kk.y, acc[d + 1]);
                acc[d + 2] = fmaf(ds, kk.z, acc[d + 2]); acc[d + 3] = fmaf(ds, kk.w, acc[d + 3]);
            }
        }
        const long long dr = a.o_custom ? (orow0 + (long long)i * oseq) : (qrow0 + (long long)i * a.q_seq);
        // the gradients leave as fp32 and / or directly as the split-bf16 planes the in-projection's backward GEMMs read
        if (dq) {
            float* dst = dq + dr * lddq + h * HD + c0;
#pragma unroll
            for (int d = 0; d < CPL; d += 4)
                *reinterpret_cast<float4*>(dst + d) =
                    make_float4(acc[d] * a.scale, acc[d + 1] * a.scale, acc[d + 2] * a.scale, acc[d + 3] * a.scale);
        }
        if (pq.on()) {
#pragma unroll
            for (int d = 0; d < CPL; d += 4)
                split4_store(pq, dr, h * HD + c0 + d, acc[d] * a.scale, acc[d + 1] * a.scale, acc[d + 2] * a.scale, acc[d + 3] * a.scale);
        }
    }
    __syncthreads();
    // ---- phase B: lane group = key j, lane = channel slice; no reductions
    if (live && row < Sk) {
        const int j = row;
        float ak[CPL], av[CPL];
#pragma unroll
        for (int d = 0; d < CPL; ++d) { ak[d] = 0.f; av[d] = 0.f; }
        for (int i = 0; i < Sq; ++i) {
            const float p = Ps[i * Skp + j];
            const float ds = p * (dPs[i * Skp + j] - delta[i]);
#pragma unroll
            for (int d = 0; d < CPL; d += 4) {
                const float4 oo = *reinterpret_cast<const float4*>(dOs + i * AB_PITCH + c0 + d);
                const float4 qq = *reinterpret_cast<const float4*>(Qs + i * AB_PITCH + c0 + d);
                av[d] = fmaf(p, oo.x, av[d]); av[d + 1] = fmaf(p, oo.y, av[d + 1]);
                av[d + 2] = fmaf(p, oo.z, av[d + 2]); av[d + 3] = fmaf(p, oo.w, av[d + 3]);
                ak[d] = fmaf(ds, qq.x, ak[d]); ak[d + 1] = fmaf(ds, qq.y, ak[d + 1]);       // Qs is pre-scaled: dK carries `scale`
                ak[d + 2] = fmaf(ds, qq.z, ak[d + 2]); ak[d + 3] = fmaf(ds, qq.w, ak[d + 3]);
            }
        }
        const long long r = krow0 + (long long)j * a.k_seq;
        if (dk) {
            float* dkp = dk + r * lddk + h * HD + c0;
            float* dvp = dv + r * lddv + h * HD + c0;
#pragma unroll
            for (int d = 0; d < CPL; d += 4) {
                *reinterpret_cast<float4*>(dkp + d) = make_float4(ak[d], ak[d + 1], ak[d + 2], ak[d + 3]);
                *reinterpret_cast<float4*>(dvp + d) = make_float4(av[d], av[d + 1], av[d + 2], av[d + 3]);
            }
        }
        if (pk.on()) {
#pragma unroll
            for (int d = 0; d < CPL; d += 4) {
                split4_store(pk, r, h * HD + c0 + d, ak[d], ak[d + 1], ak[d + 2], ak[d + 3]);
                split4_store(pv, r, h * HD + c0 + d, av[d], av[d + 1], av[d + 2], av[d + 3]);
            }
        }
    }
}

// ---- head_dim 32, register-tiled form (sequences of 24 .. 128 rows: encoder self-attention, decoder cross-attention) -------
// One (batch, head) problem per CTA of 256 threads.  Every product of the backward is computed in 4 x 4 register tiles, so a
// 16-byte shared-memory load feeds 16 FMAs instead of 4 and no shuffle is needed (the cooperative form above spends three
// quarters of its instructions on loads, shuffles and address arithmetic; profiles/r2_attention_ncu.md):
//   A1  S tile (4 queries x 4 keys over 32 channels) -> P = exp(S - lse), masked keys / dead rows -> 0      -> Ps
//   A2  dP tile = dO V^T                                                                                     -> dSs
//   A3  delta_i = sum_j P_ij dP_ij (one thread per query row, ascending j), then dS = P (dP - delta) in place
//   B   dK / dV tiles (4 keys x 4 channels, loop over queries) and dQ tiles (4 queries x 4 channels, loop over keys)
// Rows are padded to multiples of 4 (pad rows are zero; pad keys are masked, pad queries are dead rows).
__host__ __device__ inline int attn_bwd3_skp(int Sk) { const int s4 = (Sk + 3) & ~3; return (s4 & 7) ? s4 : s4 + 4; }   // = 4 (mod 8)
__host__ __device__ inline size_t attn_bwd3_smem_floats(int Sq, int Sk) {
    const int Sq4 = (Sq + 3) & ~3, Sk4 = (Sk + 3) & ~3;
    return (size_t)2 * Sq4 * AB_PITCH + (size_t)2 * Sk4 * AB_PITCH + (size_t)2 * Sq4 * attn_bwd3_skp(Sk) + (size_t)2 * Sq4 + Sk4 / 4;
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}

__global__ void __launch_bounds__(256, 3)
attention_bwd3_kernel(AttnArgs a, const float* __restrict__ d_o, long long lddo, float* __restrict__ dq, long long lddq,
                      float* __restrict__ dk, float* __restrict__ dv, long long lddk, long long lddv, Planes pq, Planes pk, Planes pv) {
    pdl_grid_sync_sel();
    constexpr int HD = 32, T = 256;
    extern __shared__ __align__(16) float sm[];
    const int Sq = a.Sq, Sk = a.Sk, Sq4 = (Sq + 3) & ~3, Sk4 = (Sk + 3) & ~3, SkP = attn_bwd3_skp(Sk);
    float* Qs = sm;                                      // [Sq4][36]  (pre-scaled)
    float* dOs = Qs + (size_t)Sq4 * AB_PITCH;            // [Sq4][36]
    float* Ks = dOs + (size_t)Sq4 * AB_PITCH;            // [Sk4][36]
    float* Vs = Ks + (size_t)Sk4 * AB_PITCH;             // [Sk4][36]
    float* Ps = Vs + (size_t)Sk4 * AB_PITCH;             // [Sq4][SkP]
    float* dSs = Ps + (size_t)Sq4 * SkP;                 // [Sq4][SkP]  dP, then dS
    float* lse_s = dSs + (size_t)Sq4 * SkP;              // [Sq4]
    uint8_t* msk = reinterpret_cast<uint8_t*>(lse_s + 2 * Sq4);   // [Sk4]  (one spare row of Sq4 floats in between)
    const int t = threadIdx.x;
    const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
    const long long krow0 = attn_row_b(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row_b(b, a.q_inner_n, a.q_outer, a.q_inner);
    const long long orow0 = a.o_custom ? attn_row_b(b, a.o_inner_n, a.o_outer, a.o_inner) : qrow0;
    const long long oseq = a.o_custom ? a.o_seq : a.q_seq;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = t; e < Sk4 * 8; e += T) {
        const int j = e >> 3, c = (e & 7) * 4;
        const long long r = krow0 + (long long)j * a.k_seq;
        *reinterpret_cast<float4*>(Ks + j * AB_PITCH + c) = j < Sk ? *reinterpret_cast<const float4*>(a.k + r * a.ldk + h * HD + c) : z4;
        *reinterpret_cast<float4*>(Vs + j * AB_PITCH + c) = j < Sk ? *reinterpret_cast<const float4*>(a.v + r * a.ldv + h * HD + c) : z4;
    }
    for (int e = t; e < Sq4 * 8; e += T) {
        const int i = e >> 3, c = (e & 7) * 4;
        float4 q = z4, g = z4;
        if (i < Sq) {
            q = *reinterpret_cast<const float4*>(a.q + (qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + c);
            q.x *= a.scale; q.y *= a.scale; q.z *= a.scale; q.w *= a.scale;
            g = *reinterpret_cast<const float4*>(d_o + (orow0 + (long long)i * oseq) * lddo + h * HD + c);
        }
        *reinterpret_cast<float4*>(Qs + i * AB_PITCH + c) = q;
        *reinterpret_cast<float4*>(dOs + i * AB_PITCH + c) = g;
    }
    for (int i = t; i < Sq4; i += T) lse_s[i] = i < Sq ? a.lse[((long long)b * a.H + h) * Sq + i] : INFINITY;
    {
        const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * Sk : nullptr;
        for (int j = t; j < Sk4; j += T) msk[j] = j < Sk ? (kpm ? kpm[j] : 0) : 1;
    }
    __syncthreads();
    // ---- A1 / A2.  A tile = 4 consecutive queries x the 4 keys {jb, jb + tj, jb + 2 tj, jb + 3 tj}: neighbouring threads then read
    // neighbouring K / V rows (pitch 36 floats = 4 banks apart, conflict-free), where consecutive keys per thread would put
    // the rows of a warp 16 banks apart (7-way conflicts on every K / V load)
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
            for (int r = 0; r < 4; ++r) q[r] = *reinterpret_cast<const float4*>(Qs + (i0 + r) * AB_PITCH + c);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float4 kk = *reinterpret_cast<const float4*>(Ks + (jb + x * tj) * AB_PITCH + c);
#pragma unroll
                for (int r = 0; r < 4; ++r) s[r][x] = dot4(q[r], kk, s[r][x]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float lse = lse_s[i0 + r];
            const bool dead = (lse == INFINITY);
#pragma unroll
            for (int x = 0; x < 4; ++x)
                Ps[(i0 + r) * SkP + jb + x * tj] = (dead || msk[jb + x * tj]) ? 0.f : __expf(s[r][x] - lse);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) s[r][x] = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
            float4 g[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) g[r] = *reinterpret_cast<const float4*>(dOs + (i0 + r) * AB_PITCH + c);
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float4 vv = *reinterpret_cast<const float4*>(Vs + (jb + x * tj) * AB_PITCH + c);
#pragma unroll
                for (int r = 0; r < 4; ++r) s[r][x] = dot4(g[r], vv, s[r][x]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x) dSs[(i0 + r) * SkP + jb + x * tj] = s[r][x];
    }
    __syncthreads();
    // ---- A3: four lanes per query row: delta_i = sum_j P_ij dP_ij (lane partials combined in a fixed order), then dS = P (dP - delta)
    // in place on the lane's own columns
    {
        const int g = t & 3;
        const unsigned gmask = 0xfu << ((t & 31) & ~3);
        for (int i = t >> 2; i < Sq4; i += T / 4) {
            const float* pr = Ps + i * SkP;
            float* dr = dSs + i * SkP;
            float dl = 0.f;
            for (int j = g; j < Sk; j += 4) dl = fmaf(pr[j], dr[j], dl);
            dl += __shfl_xor_sync(gmask, dl, 1);
            dl += __shfl_xor_sync(gmask, dl, 2);
            for (int j = g; j < Sk4; j += 4) dr[j] = pr[j] * (dr[j] - dl);
        }
    }
    __syncthreads();
    // ---- B: (key block, channel quad) items produce dK and dV, (query block, channel quad) items produce dQ
    const int nkv = tj * 8, nq = (Sq4 >> 2) * 8;
    for (int u = t; u < nkv + nq; u += T) {
        if (u < nkv) {
            const int j0 = (u >> 3) << 2, c = (u & 7) << 2;
            float ak[4][4], av[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int x = 0; x < 4; ++x) { ak[r][x] = 0.f; av[r][x] = 0.f; }
            for (int i = 0; i < Sq; ++i) {
                const float4 pp = *reinterpret_cast<const float4*>(Ps + i * SkP + j0);
                const float4 ds = *reinterpret_cast<const float4*>(dSs + i * SkP + j0);
                const float4 qq = *reinterpret_cast<const float4*>(Qs + i * AB_PITCH + c);
                const float4 oo = *reinterpret_cast<const float4*>(dOs + i * AB_PITCH + c);
#define RIFT_AKV(r, P, D)                                                                                                   \
    av[r][0] = fmaf(P, oo.x, av[r][0]); av[r][1] = fmaf(P, oo.y, av[r][1]); av[r][2] = fmaf(P, oo.z, av[r][2]);                 \
    av[r][3] = fmaf(P, oo.w, av[r][3]); ak[r][0] = fmaf(D, qq.x, ak[r][0]); ak[r][1] = fmaf(D, qq.y, ak[r][1]);                 \
    ak[r][2] = fmaf(D, qq.z, ak[r][2]); ak[r][3] = fmaf(D, qq.w, ak[r][3]);
                RIFT_AKV(0, pp.x, ds.x) RIFT_AKV(1, pp.y, ds.y) RIFT_AKV(2, pp.z, ds.z) RIFT_AKV(3, pp.w, ds.w)     // Qs is pre-scaled: dK carries `scale`
#undef RIFT_AKV
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (j0 + r >= Sk) continue;
                const long long row = krow0 + (long long)(j0 + r) * a.k_seq;
                if (dk) {
                    *reinterpret_cast<float4*>(dk + row * lddk + h * HD + c) = make_float4(ak[r][0], ak[r][1], ak[r][2], ak[r][3]);
                    *reinterpret_cast<float4*>(dv + row * lddv + h * HD + c) = make_float4(av[r][0], av[r][1], av[r][2], av[r][3]);
                }
                if (pk.on()) {
                    split4_store(pk, row, h * HD + c, ak[r][0], ak[r][1], ak[r][2], ak[r][3]);
                    split4_store(pv, row, h * HD + c, av[r][0], av[r][1], av[r][2], av[r][3]);
                }
            }
        } else {
            const int v = u - nkv;
            const int i0 = (v >> 3) << 2, c = (v & 7) << 2;
            float acc[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int x = 0; x < 4; ++x) acc[r][x] = 0.f;
            for (int j = 0; j < Sk; ++j) {
                const float4 kk = *reinterpret_cast<const float4*>(Ks + j * AB_PITCH + c);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float ds = dSs[(i0 + r) * SkP + j];
                    acc[r][0] = fmaf(ds, kk.x, acc[r][0]); acc[r][1] = fmaf(ds, kk.y, acc[r][1]);
                    acc[r][2] = fmaf(ds, kk.z, acc[r][2]); acc[r][3] = fmaf(ds, kk.w, acc[r][3]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (i0 + r >= Sq) continue;
                const long long row = a.o_custom ? (orow0 + (long long)(i0 + r) * oseq) : (qrow0 + (long long)(i0 + r) * a.q_seq);
                const float x0 = acc[r][0] * a.scale, x1 = acc[r][1] * a.scale, x2 = acc[r][2] * a.scale, x3 = acc[r][3] * a.scale;
                if (dq) *reinterpret_cast<float4*>(dq + row * lddq + h * HD + c) = make_float4(x0, x1, x2, x3);
                if (pq.on()) split4_store(pq, row, h * HD + c, x0, x1, x2, x3);
            }
        }
    }
}

// ---- head_dim 32, tiny sequences (S = 6 / 12): one warp per (batch, head) problem, lane = channel; q, k, v, dO and the dK / dV
// accumulators of the whole problem live in registers (see attention_small_kernel)
template <int S>
__global__ void __launch_bounds__(128, 3)
attention_small_bwd_kernel(AttnArgs a, const float* __restrict__ d_o, long long lddo, float* __restrict__ dq, long long lddq,
                           float* __restrict__ dk, float* __restrict__ dv, long long lddk, long long lddv, Planes pq, Planes pk, Planes pv) {
    pdl_grid_sync_sel();
    constexpr int HD = 32;
    const int lane = threadIdx.x & 31;
    const long long prob = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (prob >= (long long)a.B * a.H) return;
    const int b = (int)(prob / a.H), h = (int)(prob % a.H);
    const long long krow0 = attn_row_b(b, a.k_inner_n, a.k_outer, a.k_inner);
    const long long qrow0 = attn_row_b(b, a.q_inner_n, a.q_outer, a.q_inner);
    const long long orow0 = a.o_custom ? attn_row_b(b, a.o_inner_n, a.o_outer, a.o_inner) : qrow0;
    const long long oseq = a.o_custom ? a.o_seq : a.q_seq;
    const uint8_t* kpm = a.kpm ? a.kpm + (long long)(a.kpm_mod > 0 ? (b + a.kpm_off) % a.kpm_mod : b / a.kpm_div) * S : nullptr;
    float q[S], k[S], v[S], go[S], ak[S], av[S], lse[S];
    unsigned masked = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        q[i] = a.q[(qrow0 + (long long)i * a.q_seq) * a.ldq + h * HD + lane] * a.scale;
        const long long r = krow0 + (long long)i * a.k_seq;
        k[i] = a.k[r * a.ldk + h * HD + lane];
        v[i] = a.v[r * a.ldv + h * HD + lane];
        go[i] = d_o[(orow0 + (long long)i * oseq) * lddo + h * HD + lane];
        lse[i] = a.lse[((long long)b * a.H + h) * S + i];
        ak[i] = 0.f; av[i] = 0.f;
        if (kpm && kpm[i]) masked |= 1u << i;
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const bool dead = (lse[i] == INFINITY);
        float p[S], dp[S];
        float dl = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const float sc = warp_sum(q[i] * k[j]);
            dp[j] = warp_sum(go[i] * v[j]);
            p[j] = (dead || ((masked >> j) & 1u)) ? 0.f : __expf(sc - lse[i]);
            dl = fmaf(p[j], dp[j], dl);
        }
        float dqv = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            const float ds = p[j] * (dp[j] - dl);
            dqv = fmaf(ds, k[j], dqv);
            ak[j] = fmaf(ds, q[i], ak[j]);                  // q is pre-scaled: dK carries `scale`
            av[j] = fmaf(p[j], go[i], av[j]);
        }
        const long long dr = a.o_custom ? (orow0 + (long long)i * oseq) : (qrow0 + (long long)i * a.q_seq);
        if (dq) dq[dr * lddq + h * HD + lane] = dqv * a.scale;
        if (pq.on()) split_store(pq, dr, h * HD + lane, dqv * a.scale);
    }
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const long long r = krow0 + (long long)j * a.k_seq;
        if (dk) { dk[r * lddk + h * HD + lane] = ak[j]; dv[r * lddv + h * HD + lane] = av[j]; }
        if (pk.on()) { split_store(pk, r, h * HD + lane, ak[j]); split_store(pv, r, h * HD + lane, av[j]); }
    }
}
static bool attention_small_bwd_on() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_ATTN_SMALL"); return e && atoi(e) != 0; }();   // measured slower: off
    return on;
}

// shape-only: will launch_attention_bwd take the kernel above (the one that can write planes)?
bool attention_bwd_planes_ok(int Sq, int Sk, int hd) {
    static const bool legacy = getenv("RIFT_B200_ATTN_BWD_LEGACY") != nullptr;
    const int smax = max(Sq, Sk);
    return hd == 32 && !legacy && 4 * smax <= 512 && attn_bwd2_smem_floats(Sq, Sk) * sizeof(float) <= 200 * 1024;
}
// register-tiled kernel: sequences long enough to fill a CTA with 4 x 4 tiles (RIFT_B200_ATTN_BWD_TILED=0 switches it off)
static bool attention_bwd_tiled(int Sq, int Sk, int hd) {
    static const bool on = [] { const char* e = getenv("RIFT_B200_ATTN_BWD_TILED"); return !(e && atoi(e) == 0); }();
    return on && hd == 32 && min(Sq, Sk) >= 24 && max(Sq, Sk) <= 128 && attn_bwd3_smem_floats(Sq, Sk) * sizeof(float) <= 160 * 1024;
}

int launch_attention_bwd(const AttnArgs& a, const float* d_o, long long lddo, float* dq, long long lddq, float* dk,
                         float* dv, long long lddk, long long lddv, cudaStream_t st, const AttnBwdPlanes& pl) {
    if (a.B <= 0 || a.Sq <= 0) return 0;
    const bool planes = pl.q.on() || pl.k.on() || pl.v.on();
    RIFT_REQUIRE(pl.k.on() == pl.v.on(), "attention_bwd: dK and dV planes go together");
    RIFT_REQUIRE((dq != nullptr || pl.q.on()) && (dk != nullptr || pl.k.on()) && ((dk != nullptr) == (dv != nullptr)),
                 "attention_bwd: every gradient needs an fp32 or a plane destination");
    RIFT_REQUIRE(a.hd == 32 || a.hd == 64, "attention_bwd: head_dim must be 32 or 64");
    RIFT_REQUIRE(a.lse != nullptr, "attention_bwd: forward must have saved the log-sum-exp");
    {
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool vec = a.hd == 32 && al16(a.q) && al16(a.k) && al16(a.v) && al16(d_o) && al16(dq) && al16(dk) && al16(dv) &&
                         (a.ldq & 3) == 0 && (a.ldk & 3) == 0 && (a.ldv & 3) == 0 && (lddo & 3) == 0 && (lddq & 3) == 0 &&
                         (lddk & 3) == 0 && (lddv & 3) == 0;
        static const bool legacy = getenv("RIFT_B200_ATTN_BWD_LEGACY") != nullptr;
        if (a.hd == 32 && !legacy && attention_small_bwd_on() && a.Sq == a.Sk && (a.Sq == 6 || a.Sq == 12)) {
            const unsigned nb = (unsigned)(((long long)a.B * a.H + 3) / 4);
            if (a.Sq == 6) launch_k(attention_small_bwd_kernel<6>, nb, 128, 0, st, a, d_o, lddo, dq, lddq, dk, dv, lddk, lddv, pl.q, pl.k, pl.v);
            else launch_k(attention_small_bwd_kernel<12>, nb, 128, 0, st, a, d_o, lddo, dq, lddq, dk, dv, lddk, lddv, pl.q, pl.k, pl.v);
            RIFT_LAUNCH_OK();
            return 0;
        }
        if (vec && !legacy && attention_bwd_tiled(a.Sq, a.Sk, a.hd)) {
            const size_t smem3 = attn_bwd3_smem_floats(a.Sq, a.Sk) * sizeof(float);
            static bool attr3 = false;
            if (!attr3) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(attention_bwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr3 = true;
            }
            launch_k(attention_bwd3_kernel, (unsigned)(a.B * a.H), 256, smem3, st, a, d_o, lddo, dq, lddq, dk, dv, lddk, lddv, pl.q, pl.k, pl.v);
            RIFT_LAUNCH_OK();
            return 0;
        }
        const int smax = max(a.Sq, a.Sk);
        const size_t per = attn_bwd2_smem_floats(a.Sq, a.Sk) * sizeof(float);
        if (vec && !legacy && 4 * smax <= 512 && per <= 200 * 1024) {
            constexpr int G = 4;
            const int tpp = (G * smax + 31) / 32 * 32;
            int PB = max(1, 256 / tpp);
            while (PB > 1 && PB * per > 96 * 1024) --PB;
            const long long nprob = (long long)a.B * a.H;
            static bool attr2 = false;
            if (!attr2) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(attention_bwd2_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr2 = true;
            }
            launch_k(attention_bwd2_kernel<G>, (unsigned)((nprob + PB - 1) / PB), PB * tpp, PB * per, st, a, d_o, lddo, dq, lddq, dk, dv, lddk,
                                                                                                 lddv, PB, tpp, pl.q, pl.k, pl.v);
            RIFT_LAUNCH_OK();
            return 0;
        }
    }
    RIFT_REQUIRE(!planes, "attention_bwd: plane outputs need the vector kernel (see attention_bwd_planes_ok)");
    const size_t smem = ((size_t)2 * a.Sq * a.hd + (size_t)2 * a.Sk * a.hd + 2 * a.Sq) * sizeof(float);
    RIFT_REQUIRE(smem <= 200 * 1024, "attention_bwd: sequence too long for the shared-memory kernel");
    static bool attr = false;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RIFT_CUDA_OK(cudaFuncSetAttribute(attention_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const int threads = min(128, (max(a.Sq, a.Sk) + 31) / 32 * 32);
    if (a.hd == 32) launch_k(attention_bwd_kernel<32>, a.B * a.H, threads, smem, st, a, d_o, lddo, dq, lddq, dk, dv, lddk, lddv);
    else launch_k(attention_bwd_kernel<64>, a.B * a.H, threads, smem, st, a, d_o, lddo, dq, lddq, dk, dv, lddk, lddv);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// Neighbourhood attention backward: one warp per (sequence, head), lane = channel; dk / dv accumulate in
// shared memory (a key is seen by up to `ksize` queries).  drpb partials per warp: [n_seq*heads][2k-1].
// =====================================================================================
constexpr int NATB_MAXL = 32, NATB_MAXK = 7;

__global__ void __launch_bounds__(128)
nat_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, int n_seq, int L, int heads, int hd,
                         int ksize, const float* __restrict__ rpb, float* __restrict__ dqkv, float* __restrict__ drpb_partial) {
    pdl_grid_sync();
    __shared__ float s_dk[4][NATB_MAXL][32];
    __shared__ float s_dv[4][NATB_MAXL][32];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * 4 + wib;
    if (warp >= n_seq * heads) return;
    const int n = warp / heads, h = warp % heads;
    const int dim = heads * hd;
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)n * L * 3 * dim + h * hd;
    float* dbase = dqkv + (long long)n * L * 3 * dim + h * hd;
    const bool on = lane < hd;
    for (int j = 0; j < L; ++j) { s_dk[wib][j][lane] = 0.f; s_dv[wib][j][lane] = 0.f; }
    float drpb[2 * NATB_MAXK - 1];
#pragma unroll
    for (int t = 0; t < 2 * NATB_MAXK - 1; ++t) drpb[t] = 0.f;
    __syncwarp();
    for (int i = 0; i < L; ++i) {
        const int start = min(max(i - ksize / 2, 0), L - ksize);
        const float q = on ? base[(long long)i * 3 * dim + lane] * scale : 0.f;
        const float go = on ? d_out[((long long)n * L + i) * dim + h * hd + lane] : 0.f;
        float p[NATB_MAXK], dp[NATB_MAXK];
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < NATB_MAXK; ++kk) {
            p[kk] = 0.f; dp[kk] = 0.f;
            if (kk < ksize) {
                const int j = start + kk;
                const float kv = on ? base[(long long)j * 3 * dim + dim + lane] : 0.f;
                const float vv = on ? base[(long long)j * 3 * dim + 2 * dim + lane] : 0.f;
                p[kk] = warp_sum(q * kv) + rpb[h * (2 * ksize - 1) + (j - i + ksize - 1)];
                dp[kk] = warp_sum(go * vv);
                mx = fmaxf(mx, p[kk]);
            }
        }
        float den = 0.f;
#pragma unroll
        for (int kk = 0; kk < NATB_MAXK; ++kk)
            if (kk < ksize) { p[kk] = __expf(p[kk] - mx); den += p[kk]; }
        float dl = 0.f;
#pragma unroll
        for (int kk = 0; kk < NATB_MAXK; ++kk)
            if (kk < ksize) { p[kk] /= den; dl += p[kk] * dp[kk]; }
        float dqv = 0.f;
#pragma unroll
        for (int kk = 0; kk < NATB_MAXK; ++kk) {
            if (kk < ksize) {
                const int j = start + kk;
                const float ds = p[kk] * (dp[kk] - dl);
                if (on) {
                    const float kv = base[(long long)j * 3 * dim + dim + lane];
                    dqv += ds * kv;
                    s_dk[wib][j][lane] += ds * q;                 // q already carries the scale
                    s_dv[wib][j][lane] += p[kk] * go;
                }
                // relative index (j - i + k - 1) depends on i only through `start`: accumulate per slot
#pragma unroll
                for (int t = 0; t < 2 * NATB_MAXK - 1; ++t)
                    if (t == j - i + ksize - 1) drpb[t] += ds;
            }
        }
        if (on) dbase[(long long)i * 3 * dim + lane] = dqv * scale;
    }
    __syncwarp();
    if (on) {
        for (int j = 0; j < L; ++j) {
            dbase[(long long)j * 3 * dim + dim + lane] = s_dk[wib][j][lane];
            dbase[(long long)j * 3 * dim + 2 * dim + lane] = s_dv[wib][j][lane];
        }
    }
    if (drpb_partial && lane == 0) {
        for (int t = 0; t < 2 * ksize - 1; ++t) drpb_partial[(long long)warp * (2 * ksize - 1) + t] = drpb[t];
    }
}

// Register-resident form (see nat_attention_reg_kernel): q, k, v, dO of the whole (sequence, head) slice and the
// dk / dv accumulators live in registers, every index is a compile-time constant after unrolling.
template <int L, int KS>
__global__ void __launch_bounds__(128)
nat_attention_bwd_reg_kernel(const float* __restrict__ qkv, const float* __restrict__ d_out, int n_seq, int heads, int hd,
                             const float* __restrict__ rpb, float* __restrict__ dqkv, float* __restrict__ drpb_partial, Planes dp_out) {
    pdl_grid_sync();
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (warp >= n_seq * heads) return;
    const int n = warp / heads, h = warp % heads;
    const int dim = heads * hd;
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)n * L * 3 * dim + h * hd;
    float* dbase = dqkv ? dqkv + (long long)n * L * 3 * dim + h * hd : nullptr;
    const bool on = lane < hd;
    float q[L], k[L], v[L], go[L], dk[L], dv[L], rb[2 * KS - 1], drpb[2 * KS - 1];
#pragma unroll
    for (int i = 0; i < L; ++i) {
        q[i] = on ? base[(long long)i * 3 * dim + lane] * scale : 0.f;
        k[i] = on ? base[(long long)i * 3 * dim + dim + lane] : 0.f;
        v[i] = on ? base[(long long)i * 3 * dim + 2 * dim + lane] : 0.f;
        go[i] = on ? d_out[((long long)n * L + i) * dim + h * hd + lane] : 0.f;
        dk[i] = 0.f; dv[i] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < 2 * KS - 1; ++t) { rb[t] = rpb[h * (2 * KS - 1) + t]; drpb[t] = 0.f; }
#pragma unroll
    for (int i = 0; i < L; ++i) {
        constexpr int half = KS / 2;
        const int start = (i - half < 0) ? 0 : ((i - half > L - KS) ? L - KS : i - half);
        float p[KS], dp[KS];
        float mx = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            p[kk] = warp_sum(q[i] * k[start + kk]) + rb[start + kk - i + KS - 1];
            dp[kk] = warp_sum(go[i] * v[start + kk]);
            mx = fmaxf(mx, p[kk]);
        }
        float den = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) { p[kk] = __expf(p[kk] - mx); den += p[kk]; }
        float dl = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) { p[kk] /= den; dl += p[kk] * dp[kk]; }
        float dqv = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            const float ds = p[kk] * (dp[kk] - dl);
            dqv += ds * k[start + kk];
            dk[start + kk] += ds * q[i];                  // q already carries the scale
            dv[start + kk] += p[kk] * go[i];
            drpb[start + kk - i + KS - 1] += ds;
        }
        if (on && dbase) dbase[(long long)i * 3 * dim + lane] = dqv * scale;
        if (on && dp_out.on()) split_store(dp_out, (long long)n * L + i, h * hd + lane, dqv * scale);
    }
    if (on && dbase) {
#pragma unroll
        for (int j = 0; j < L; ++j) {
            dbase[(long long)j * 3 * dim + dim + lane] = dk[j];
            dbase[(long long)j * 3 * dim + 2 * dim + lane] = dv[j];
        }
    }
    if (on && dp_out.on()) {               // the qkv projection's dY directly as split-bf16 planes [n_seq * L, 3 dim]
#pragma unroll
        for (int j = 0; j < L; ++j) {
            split_store(dp_out, (long long)n * L + j, dim + h * hd + lane, dk[j]);
            split_store(dp_out, (long long)n * L + j, 2 * dim + h * hd + lane, dv[j]);
        }
    }
    if (drpb_partial && lane == 0) {
#pragma unroll
        for (int t = 0; t < 2 * KS - 1; ++t) drpb_partial[(long long)warp * (2 * KS - 1) + t] = drpb[t];
    }
}

bool nat_attention_bwd_planes_ok(int L, int ksize) { return (L == 20 && ksize == 3) || (L == 10 && ksize == 3) || (L == 5 && ksize == 5); }

int launch_nat_attention_bwd(const float* qkv, const float* d_out, int n_seq, int L, int heads, int hd, int ksize,
                             const float* rpb, float* dqkv, float* drpb_partial, cudaStream_t st, Planes dqkv_planes) {
    if (n_seq <= 0) return 0;
    RIFT_REQUIRE(dqkv != nullptr || dqkv_planes.on(), "nat_attention_bwd: no destination");
    RIFT_REQUIRE(!dqkv_planes.on() || (nat_attention_bwd_planes_ok(L, ksize) && dqkv_planes.Kp == 3 * heads * hd),
                 "nat_attention_bwd: plane output needs the register kernel and pitch 3 * dim");
    RIFT_REQUIRE(hd <= 32 && ksize <= NATB_MAXK && L >= ksize && L <= NATB_MAXL, "nat_attention_bwd: unsupported shape");
#define RIFT_NATB_REG(LL, KK)                                                                                               \
    if (L == LL && ksize == KK) {                                                                                           \
        launch_k(nat_attention_bwd_reg_kernel<LL, KK>, cdiv((long long)n_seq * heads, 4), 128, 0, st, qkv, d_out, n_seq,    \
                 heads, hd, rpb, dqkv, drpb_partial, dqkv_planes);                                                          \
        RIFT_LAUNCH_OK();                                                                                                   \
        return 0;                                                                                                           \
    }
    RIFT_NATB_REG(20, 3) RIFT_NATB_REG(10, 3) RIFT_NATB_REG(5, 5)
#undef RIFT_NATB_REG
    launch_k(nat_attention_bwd_kernel, cdiv((long long)n_seq * heads, 4), 128, 0, st, qkv, d_out, n_seq, L, heads, hd, ksize, rpb, dqkv,
                                                                             drpb_partial);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// gather / scatter style backward kernels
// =====================================================================================
// max-pool: the gradient of a pooled value goes to the arg-max point (nowhere when the winner was a padded zero)
template <int VEC>
__global__ void __launch_bounds__(256)
masked_maxpool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ argmax, int groups, int n, int C,
                          float* __restrict__ dx, int accumulate) {
    pdl_grid_sync();
    const int CV = C / VEC;
    FOR_GRID(e, (long long)groups * n * CV) {
        const int c = (int)(e % CV) * VEC;
        const long long r = e / CV;
        const int p = (int)(r % n);
        const long long g = r / n;
        if (VEC == 4) {
            const int4 am = *reinterpret_cast<const int4*>(argmax + g * C + c);
            const float4 d = *reinterpret_cast<const float4*>(dout + g * C + c);
            float4 v = make_float4(am.x == p ? d.x : 0.f, am.y == p ? d.y : 0.f, am.z == p ? d.z : 0.f, am.w == p ? d.w : 0.f);
            float4* o = reinterpret_cast<float4*>(dx + r * C + c);
            if (accumulate) { const float4 t = *o; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
            *o = v;
        } else {
            const float v = (argmax[g * C + c] == p) ? dout[g * C + c] : 0.f;
            dx[r * C + c] = accumulate ? dx[r * C + c] + v : v;
        }
    }
}
int launch_masked_maxpool_bwd(const float* dout, const int* argmax, int groups, int n, int C, float* dx, int accumulate,
                              cudaStream_t st) {
    const long long total = (long long)groups * n * C;
    if (total <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((C & 3) == 0 && al16(dout) && al16(argmax) && al16(dx))
        launch_k(masked_maxpool_bwd_kernel<4>, GRID1D(total / 4, 256), 256, 0, st, dout, argmax, groups, n, C, dx, accumulate);
    else
        launch_k(masked_maxpool_bwd_kernel<1>, GRID1D(total, 256), 256, 0, st, dout, argmax, groups, n, C, dx, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}

// col2im for the k=3, pad=1 im2col (gather form: every dx element sums its <= 3 contributions)
__global__ void col2im_k3_kernel(const float* __restrict__ dcols, int n_seq, int L, int Lout, int C, int stride,
                                 float* __restrict__ dx, int accumulate) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)n_seq * L * C) {
        const int c = (int)(e % C);
        const long long r = e / C;
        const int ts = (int)(r % L);
        const long long n = r / L;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int num = ts + 1 - k;                   // t * stride
            if (num >= 0 && num % stride == 0) {
                const int t = num / stride;
                if (t < Lout) a += dcols[((n * Lout + t) * C + c) * 3 + k];
            }
        }
        dx[e] = accumulate ? dx[e] + a : a;
    }
}
int launch_col2im_k3(const float* dcols, int n_seq, int L, int C, int stride, float* dx, int accumulate, cudaStream_t st) {
    const int Lout = (L + 2 - 3) / stride + 1;
    const long long total = (long long)n_seq * L * C;
    if (total <= 0) return 0;
    launch_k(col2im_k3_kernel, GRID1D(total, 256), 256, 0, st, dcols, n_seq, L, Lout, C, stride, dx, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}
// im2col_k3_last: dx[n, L-2+k, c] += dcols[n, c*3+k] for k = 0, 1 ; everything else 0
__global__ void col2im_k3_last_kernel(const float* __restrict__ dcols, int n_seq, int L, int C, float* __restrict__ dx) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)n_seq * L * C) {
        const int c = (int)(e % C);
        const long long r = e / C;
        const int ts = (int)(r % L);
        const long long n = r / L;
        const int k = ts - (L - 2);
        dx[e] = (k == 0 || k == 1) ? dcols[(n * C + c) * 3 + k] : 0.f;
    }
}
int launch_col2im_k3_last(const float* dcols, int n_seq, int L, int C, float* dx, cudaStream_t st) {
    const long long total = (long long)n_seq * L * C;
    if (total <= 0) return 0;
    launch_k(col2im_k3_last_kernel, GRID1D(total, 256), 256, 0, st, dcols, n_seq, L, C, dx);
    RIFT_LAUNCH_OK();
    return 0;
}

// transpose of fpn_upsample_add: dsrc[n, i, c] += sum_j w(j -> i) ddst[n, j, c]
// Only destination positions whose source coordinate sp(j) lies within (i - 1, i + 1) can carry weight: the j loop covers that
// window (plus one position of slack on both sides; the weight test inside is the exact one, so skipped terms are exact zeros
// and the sum has the same terms in the same order as the full loop).  VEC = 4 channels per thread.
template <int VEC>
__global__ void __launch_bounds__(256)
fpn_upsample_add_bwd_kernel(const float* __restrict__ ddst, float* __restrict__ dsrc, int n_seq, int Ld, int Ls, int C) {
    pdl_grid_sync();
    const float rscale = (float)Ls / (float)Ld;
    const int CV = C / VEC;
    FOR_GRID(e, (long long)n_seq * Ls * CV) {
        const int c = (int)(e % CV) * VEC;
        const long long r = e / CV;
        const int i = (int)(r % Ls);
        const long long n = r / Ls;
        const int j_lo = (i == 0) ? 0 : max(0, (int)floorf(((float)i - 0.5f) / rscale - 0.5f) - 1);
        const int j_hi = (i == Ls - 1) ? Ld - 1 : min(Ld - 1, (int)ceilf(((float)i + 1.5f) / rscale - 0.5f) + 1);
        float a[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) a[v] = 0.f;
        for (int j = j_lo; j <= j_hi; ++j) {
            float sp = ((float)j + 0.5f) * rscale - 0.5f;
            sp = fmaxf(sp, 0.f);
            const int i0 = min((int)sp, Ls - 1);
            const int i1 = min(i0 + 1, Ls - 1);
            const float w1 = sp - (float)i0, w0 = 1.f - w1;
            float w = 0.f;
            if (i0 == i) w += w0;
            if (i1 == i) w += w1;
            if (w != 0.f) {
                const float* p = ddst + (n * Ld + j) * C + c;
                if (VEC == 4) {
                    const float4 x = *reinterpret_cast<const float4*>(p);
                    a[0] += w * x.x; a[1 % VEC] += w * x.y; a[2 % VEC] += w * x.z; a[3 % VEC] += w * x.w;
                } else {
                    a[0] += w * p[0];
                }
            }
        }
        float* o = dsrc + (n * Ls + i) * C + c;
        if (VEC == 4) {
            float4 t = *reinterpret_cast<float4*>(o);
            t.x += a[0]; t.y += a[1 % VEC]; t.z += a[2 % VEC]; t.w += a[3 % VEC];
            *reinterpret_cast<float4*>(o) = t;
        } else {
            o[0] += a[0];
        }
    }
}
int launch_fpn_upsample_add_bwd(const float* ddst, float* dsrc, int n_seq, int Ld, int Ls, int C, cudaStream_t st) {
    const long long total = (long long)n_seq * Ls * C;
    if (total <= 0) return 0;
    const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(ddst) & 15) == 0 && (reinterpret_cast<uintptr_t>(dsrc) & 15) == 0;
    if (vec) launch_k(fpn_upsample_add_bwd_kernel<4>, GRID1D(total / 4, 256), 256, 0, st, ddst, dsrc, n_seq, Ld, Ls, C);
    else launch_k(fpn_upsample_add_bwd_kernel<1>, GRID1D(total, 256), 256, 0, st, ddst, dsrc, n_seq, Ld, Ls, C);
    RIFT_LAUNCH_OK();
    return 0;
}

// out[g, c] (+)= sum_{i<n} x[(g*n + i), c]        (broadcast-over-points / over-modes terms)
__global__ void groupsum_kernel(const float* __restrict__ x, long long ldx, int groups, int n, int C, float* __restrict__ out,
                                long long ldo, int accumulate) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)groups * C) {
        const int c = (int)(e % C);
        const long long g = e / C;
        float a = 0.f;
        for (int i = 0; i < n; ++i) a += x[(g * n + i) * ldx + c];
        float* o = out + g * ldo + c;
        *o = accumulate ? *o + a : a;
    }
}
int launch_groupsum(const float* x, long long ldx, int groups, int n, int C, float* out, long long ldo, int accumulate,
                    cudaStream_t st) {
    if (groups <= 0 || C <= 0) return 0;
    launch_k(groupsum_kernel, GRID1D((long long)groups * C, 256), 256, 0, st, x, ldx, groups, n, C, out, ldo, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}

// out[m, c] (+)= sum_{rows r with r % mod == m} x[r, c]   (per-mode parameters m_pos / m_emb, pos_embed)
__global__ void __launch_bounds__(256)
modsum_kernel(const float* __restrict__ x, long long ldx, long long rows, int C, int mod, float* __restrict__ out, int accumulate) {
    pdl_grid_sync();
    // block = (m, 32-column chunk); 8 row lanes x 4 independent accumulators, combined in a fixed order -> deterministic
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m = blockIdx.x, c = blockIdx.y * 32 + tx;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (c < C) {
        const long long step = 8LL * mod;
        long long r = m + (long long)ty * mod;
        for (; r + 3 * step < rows; r += 4 * step) {
            a0 += x[r * ldx + c];
            a1 += x[(r + step) * ldx + c];
            a2 += x[(r + 2 * step) * ldx + c];
            a3 += x[(r + 3 * step) * ldx + c];
        }
        for (; r < rows; r += step) a0 += x[r * ldx + c];
    }
    sm[ty][tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i][tx];
        float* o = out + (long long)m * C + c;
        *o = accumulate ? *o + s : s;
    }
}
int launch_modsum(const float* x, long long ldx, long long rows, int C, int mod, float* out, int accumulate, cudaStream_t st) {
    if (rows <= 0 || C <= 0) return 0;
    launch_k(modsum_kernel, dim3(mod, cdiv(C, 32)), 256, 0, st, x, ldx, rows, C, mod, out, accumulate);
    RIFT_LAUNCH_OK();
    return 0;
}

// embedding tables: demb[k, c] += sum_{rows with idx == k} dy[row, c]; idx may be int8 or a 0/1 byte mask.
// Two-stage like colsum: block (32 columns x 8 row lanes) per (column chunk, row slab, table entry k).
// rows are addressed as dy[(row / inner) * outer_stride_rows + row % inner + row_offset]
__global__ void __launch_bounds__(256)
embedding_bwd_partial_kernel(const float* __restrict__ dy, long long lddy, long long row_offset, int inner, int outer_stride_rows,
                             const int8_t* __restrict__ idx, long long rows, int C, int invert_mask, float* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx, k = blockIdx.z;
    float a = 0.f;
    if (c < C) {
        for (long long r = blockIdx.y * 8 + ty; r < rows; r += (long long)gridDim.y * 8) {
            int v = (int)idx[r];
            if (invert_mask) v = v ? -1 : 0;            // "unknown" embedding: selected where the mask is 0
            if (v == k) a += dy[((r / inner) * outer_stride_rows + (r % inner) + row_offset) * lddy + c];
        }
    }
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i][tx];
        partial[((long long)k * gridDim.y + blockIdx.y) * C + c] = s;
    }
}
__global__ void embedding_bwd_final_kernel(const float* __restrict__ partial, int slabs, int C, int n_emb, float* __restrict__ demb) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)n_emb * C) {
        const int c = (int)(e % C);
        const int k = (int)(e / C);
        float a = 0.f;
        for (int s = 0; s < slabs; ++s) a += partial[((long long)k * slabs + s) * C + c];
        demb[e] += a;
    }
}
int launch_embedding_bwd(const float* dy, long long lddy, long long row_offset, int inner, int outer_stride_rows,
                         const int8_t* idx, long long rows, int C, int n_emb, float* demb, int invert_mask, float* scratch,
                         cudaStream_t st) {
    if (rows <= 0) return 0;
    const int slabs = (int)max(1LL, min(148LL / n_emb, (rows + 63) / 64));
    launch_k(embedding_bwd_partial_kernel, dim3(cdiv(C, 32), slabs, n_emb), 256, 0, st, dy, lddy, row_offset, inner, outer_stride_rows,
                                                                                 idx, rows, C, invert_mask, scratch);
    RIFT_LAUNCH_OK();
    launch_k(embedding_bwd_final_kernel, GRID1D((long long)n_emb * C, 256), 256, 0, st, scratch, slabs, C, n_emb, demb);
    RIFT_LAUNCH_OK();
    return 0;
}

// out[r, c] = keep[r] ? src[(r / inner) * outer_stride_rows + r % inner + row_offset, c] : 0
__global__ void masked_gather_rows_kernel(const float* __restrict__ src, long long lds, long long row_offset, int inner,
                                          int outer_stride_rows, const uint8_t* __restrict__ keep, long long rows, int C,
                                          float* __restrict__ out, int zero_inner0) {
    pdl_grid_sync();
    FOR_GRID(e, rows * C) {
        const int c = (int)(e % C);
        const long long r = e / C;
        bool k = keep ? keep[r] != 0 : true;
        if (zero_inner0 && (r % inner) == 0) k = false;
        out[e] = k ? src[((r / inner) * outer_stride_rows + (r % inner) + row_offset) * lds + c] : 0.f;
    }
}
int launch_masked_gather_rows(const float* src, long long lds, long long row_offset, int inner, int outer_stride_rows,
                              const uint8_t* keep, long long rows, int C, float* out, int zero_inner0, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(masked_gather_rows_kernel, GRID1D(rows * C, 256), 256, 0, st, src, lds, row_offset, inner, outer_stride_rows, keep, rows,
                                                                    C, out, zero_inner0);
    RIFT_LAUNCH_OK();
    return 0;
}

// dy[r, c] *= colscale[c]
__global__ void scale_cols_kernel(float* __restrict__ dy, const float* __restrict__ s, long long rows, int C) {
    pdl_grid_sync();
    FOR_GRID(e, rows * C) dy[e] *= s[e % C];
}
int launch_scale_cols(float* dy, const float* colscale, long long rows, int C, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(scale_cols_kernel, GRID1D(rows * C, 256), 256, 0, st, dy, colscale, rows, C);
    RIFT_LAUNCH_OK();
    return 0;
}

// eval-mode BatchNorm affine parameters: v = xhat * gamma + beta (pre-ReLU value saved by the forward)
//   dgamma[c] += sum_r dz[r,c] * (v[r,c] - beta[c]) / gamma[c] ; dbeta[c] += sum_r dz[r,c]
__global__ void __launch_bounds__(256)
bn_affine_bwd_partial_kernel(const float* __restrict__ dz, const float* __restrict__ v, long long rows, int C,
                             const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ partial) {
    pdl_grid_sync();
    for (int c = threadIdx.x; c < C; c += 256) {
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        const float g = gamma[c], b = beta[c];
        const long long st = gridDim.x;
        long long r = blockIdx.x;
        for (; r + 3 * st < rows; r += 4 * st) {
            const float d0 = dz[r * C + c], d1 = dz[(r + st) * C + c], d2 = dz[(r + 2 * st) * C + c], d3 = dz[(r + 3 * st) * C + c];
            g0 = fmaf(d0, v[r * C + c] - b, g0); g1 = fmaf(d1, v[(r + st) * C + c] - b, g1);
            g2 = fmaf(d2, v[(r + 2 * st) * C + c] - b, g2); g3 = fmaf(d3, v[(r + 3 * st) * C + c] - b, g3);
            b0 += d0; b1 += d1; b2 += d2; b3 += d3;
        }
        for (; r < rows; r += st) { const float d0 = dz[r * C + c]; g0 = fmaf(d0, v[r * C + c] - b, g0); b0 += d0; }
        // (v - beta) / gamma is the normalised input; the division is applied once to the sum
        partial[((long long)blockIdx.x * 2 + 0) * C + c] = ((g0 + g1) + (g2 + g3)) / g;
        partial[((long long)blockIdx.x * 2 + 1) * C + c] = (b0 + b1) + (b2 + b3);
    }
}
// second stage: fixed-order sum over the nb partial rows ([nb][2][C]), 8 row lanes per column; blockIdx.y: 0 dgamma, 1 dbeta
__global__ void __launch_bounds__(256)
bn_affine_bwd_final_kernel(const float* __restrict__ partial, int nb, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_grid_sync();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const float* p = partial + (long long)blockIdx.y * C;
    float a = 0.f;
    if (c < C) for (int b = ty; b < nb; b += 8) a += p[(long long)b * 2 * C + c];
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += sm[i][tx];
        float* out = blockIdx.y ? dbeta : dgamma;
        out[c] += s;
    }
}
// vector form for C % 4 == 0, C <= 1024: a thread owns one column quad, 1024 / (C / 4) row lanes per block each stream rows
// with 16-byte loads (four times the bytes in flight of the scalar form); the row lanes are combined in shared memory in a
// fixed order.  Same partial layout as above.
__global__ void __launch_bounds__(1024)
bn_affine_bwd_partial4_kernel(const float* __restrict__ dz, const float* __restrict__ v, long long rows, int C,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ partial) {
    pdl_grid_sync();
    extern __shared__ float4 sm4[];                        // [lanes][2][C / 4]
    const int C4 = C >> 2, lanes = 1024 / C4;
    const int q = threadIdx.x % C4, lane = threadIdx.x / C4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), b = g;
    if (lane < lanes) {
        const float4 bt = *reinterpret_cast<const float4*>(beta + 4 * q);
        for (long long r = (long long)blockIdx.x * lanes + lane; r < rows; r += (long long)gridDim.x * lanes) {
            const float4 d = *reinterpret_cast<const float4*>(dz + r * C + 4 * q);
            const float4 x = *reinterpret_cast<const float4*>(v + r * C + 4 * q);
            g.x = fmaf(d.x, x.x - bt.x, g.x); g.y = fmaf(d.y, x.y - bt.y, g.y); g.z = fmaf(d.z, x.z - bt.z, g.z); g.w = fmaf(d.w, x.w - bt.w, g.w);
            b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
        }
        sm4[(lane * 2 + 0) * C4 + q] = g;
        sm4[(lane * 2 + 1) * C4 + q] = b;
    }
    __syncthreads();
    if (threadIdx.x < 2 * C4) {
        const int which = threadIdx.x / C4, qq = threadIdx.x % C4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { const float4 t = sm4[(l * 2 + which) * C4 + qq]; a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        if (which == 0) {
            const float4 gm = *reinterpret_cast<const float4*>(gamma + 4 * qq);
            a.x /= gm.x; a.y /= gm.y; a.z /= gm.z; a.w /= gm.w;
        }
        *reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + which) * C + 4 * qq) = a;
    }
}
int launch_bn_affine_bwd(const float* dz, const float* v, long long rows, int C, const float* gamma, const float* beta,
                         float* dgamma, float* dbeta, float* scratch, cudaStream_t st) {
    if (rows <= 0) return 0;
    const int nb = (int)min((long long)148, rows);          // scratch holds 148 x 2 x C partials
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((C & 3) == 0 && C >= 16 && C <= 1024 && al16(dz) && al16(v) && al16(gamma) && al16(beta) && al16(scratch)) {
        const int lanes = 1024 / (C / 4);
        launch_k(bn_affine_bwd_partial4_kernel, nb, 1024, (size_t)lanes * 2 * C * sizeof(float), st, dz, v, rows, C, gamma, beta, scratch);
        RIFT_LAUNCH_OK();
        launch_k(bn_affine_bwd_final_kernel, dim3(cdiv(C, 32), 2), 256, 0, st, scratch, nb, C, dgamma, dbeta);
        RIFT_LAUNCH_OK();
        return 0;
    }
    launch_k(bn_affine_bwd_partial_kernel, nb, 256, 0, st, dz, v, rows, C, gamma, beta, scratch);
    RIFT_LAUNCH_OK();
    launch_k(bn_affine_bwd_final_kernel, dim3(cdiv(C, 32), 2), 256, 0, st, scratch, nb, C, dgamma, dbeta);
    RIFT_LAUNCH_OK();
    return 0;
}

// FourierEmbedding frequencies: feat = [cos(a_j), sin(a_j), x], a_j = 2 pi f_j x
//   contrib[r, j] = (-sin(a_j) dfeat[r, j] + cos(a_j) dfeat[r, nfreq + j]) * 2 pi x[r]   (then column-summed)
__global__ void fourier_freq_bwd_kernel(const float* __restrict__ x, int rows, int d, int dsel, const float* __restrict__ freqs,
                                        int nfreq, const float* __restrict__ dfeat, int ldf, float* __restrict__ contrib) {
    pdl_grid_sync();
    FOR_GRID(e, (long long)rows * nfreq) {
        const int j = (int)(e % nfreq);
        const long long r = e / nfreq;
        const float xv = x[r * d + dsel];
        const float ang = ((xv * freqs[dsel * nfreq + j]) * 2.f) * 3.141592653589793f;
        contrib[e] = (-sinf(ang) * dfeat[r * ldf + j] + cosf(ang) * dfeat[r * ldf + nfreq + j]) * 6.283185307179586f * xv;
    }
}
int launch_fourier_freq_bwd(const float* x, int rows, int d, int dsel, const float* freqs, int nfreq, const float* dfeat,
                            int ldf, float* contrib, cudaStream_t st) {
    if (rows <= 0) return 0;
    launch_k(fourier_freq_bwd_kernel, GRID1D((long long)rows * nfreq, 256), 256, 0, st, x, rows, d, dsel, freqs, nfreq, dfeat, ldf,
                                                                                 contrib);
    RIFT_LAUNCH_OK();
    return 0;
}

// StateAttentionEncoder token linears (Linear(1, D) each): dw_i[c] += sum_b cur[b,i] dtok[b,i,c] ; db_i[c] += sum_b dtok[b,i,c]
__global__ void __launch_bounds__(256)
state_tokens_bwd_kernel(const float* __restrict__ cur, int cs_stride, int bs, int n_tok, int D, const float* __restrict__ dtok,
                        float* __restrict__ dw_all /*[n_tok][D]*/, float* __restrict__ db_all /*[n_tok][D]*/) {
    pdl_grid_sync();
    const int i = blockIdx.x;
    for (int c = blockIdx.y * 256 + threadIdx.x; c < D; c += gridDim.y * 256) {
        float aw = 0.f, ab = 0.f;
        for (int b = 0; b < bs; ++b) {
            const float g = dtok[((long long)b * n_tok + i) * D + c];
            aw += cur[(long long)b * cs_stride + i] * g;
            ab += g;
        }
        dw_all[(long long)i * D + c] = aw;
        db_all[(long long)i * D + c] = ab;
    }
}
int launch_state_tokens_bwd(const float* cur, int cs_stride, int bs, int n_tok, int D, const float* dtok, float* dw_all,
                            float* db_all, cudaStream_t st) {
    if (bs <= 0) return 0;
    launch_k(state_tokens_bwd_kernel, dim3(n_tok, cdiv(D, 256)), 256, 0, st, cur, cs_stride, bs, n_tok, D, dtok, dw_all, db_all);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// Weight gradient of a first-layer Linear with a narrow input (K <= 32: raw features, im2col of 9 channels):
// dW[N, K] (+)= dY[M, N]^T X[M, K].  The tensor-core path does not take K < 32; the generic SIMT GEMM runs this shape
// on a handful of CTAs.  Here: block = (row slab, 128 output features) x 2 row lanes, thread = one output feature with
// its K accumulators in registers (X rows are warp-uniform broadcast loads); per-slab partials [slabs][N * K] are
// then summed in a fixed order by launch_colsum_final.
// =====================================================================================
__global__ void __launch_bounds__(256)
wgrad_narrow_kernel(const float* __restrict__ dY, long long lddy, const float* __restrict__ X, long long ldx, int M, int N, int K,
                    int rows_per_slab, float* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ float sm[128][33];
    const int nl = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int n = blockIdx.y * 128 + nl;
    const int r0 = blockIdx.x * rows_per_slab, r1 = min(M, r0 + rows_per_slab);
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.f;
    if (n < N) {
        for (int r = r0 + half; r < r1; r += 2) {
            const float d = dY[(long long)r * lddy + n];
            const float* xr = X + (long long)r * ldx;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < K) acc[k] = fmaf(d, __ldg(xr + k), acc[k]);
        }
    }
    if (half == 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k) sm[nl][k] = acc[k];
    }
    __syncthreads();
    if (half == 0 && n < N) {
        float* o = partial + (long long)blockIdx.x * N * K + (long long)n * K;
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (k < K) o[k] = acc[k] + sm[nl][k];
    }
}

// partial: slabs * N * K floats with slabs = wgrad_narrow_slabs(M)
int wgrad_narrow_slabs(int M) { return max(1, min(296, cdiv(M, 64))); }
int launch_wgrad_narrow(const float* dY, long long lddy, const float* X, long long ldx, int M, int N, int K, float* dW,
                        float* partial, cudaStream_t st) {
    RIFT_REQUIRE(K >= 1 && K <= 32, "wgrad_narrow: K must be in [1, 32]");
    if (M <= 0 || N <= 0) return 0;
    const int slabs = wgrad_narrow_slabs(M);
    const int rps = cdiv(M, slabs);
    launch_k(wgrad_narrow_kernel, dim3(slabs, cdiv(N, 128)), 256, 0, st, dY, lddy, X, ldx, M, N, K, rps, partial);
    RIFT_LAUNCH_OK();
    return launch_colsum_final(partial, slabs, N * K, dW, 1, st);
}

}  // namespace rift
