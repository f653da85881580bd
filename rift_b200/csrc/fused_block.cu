// Fused transformer sub-block kernels for sm_100a: one launch per pre-LN sub-block instead of one per operator.
//
// A cluster of C CTAs owns one tile of 128 rows of the residual stream (rows of different samples never interact in
// an MLP) and splits the sub-block "tensor-parallel":
//
//     Y = X + b2 + sum_c  act( LN(X) W1[c]^T + b1[c] ) W2[:, c]^T          c = 0 .. C-1   (hidden slice of N1 = Hd / C)
//
//   phase 1  compute warps : LayerNorm of the tile from global fp32 rows -> split-bf16 operand planes in shared
//                            memory, written by hand in the UMMA canonical layout (K-major, SWIZZLE_128B: 64-column
//                            panels of 128 rows x 128 B, 16-byte chunk index XOR (row & 7))
//   phase 2  MMA thread    : GEMM1  acc1[128, N1] = A x W1[c]^T ; weights stream through a 3-stage TMA ring of
//                            128 x 64 boxes (hi + lo), three tcgen05.mma kind::f16 per k-step (split-bf16 x3)
//   phase 3  compute warps : tcgen05.ld acc1 -> + bias -> activation -> hi / lo planes of the hidden slice, written
//                            over the (dead) LayerNorm planes: the 4 D-wide hidden tensor never leaves the SM
//   phase 4  MMA thread    : GEMM2  acc2[128, D] = H x W2[:, c]^T   (second TMEM accumulator)
//   phase 5  compute warps : acc2 -> fp32 partial tile in shared memory; cluster barrier
//   phase 6  compute warps : CTA c sums the C partial tiles of its D / C output columns through distributed shared
//                            memory (fixed order -> deterministic), adds b2 and the residual, stores Y
//
// What the backward needs (LayerNorm statistics, LN(X) planes, fc1 pre-activation, hidden planes) is written on the
// way when asked for.  The 4 x (rows x Hd) fp32 / bf16 round trips through HBM of the unfused path shrink to those
// optional writes, the L2 -> SMEM operand traffic halves (the activation tile is loaded once, not once per n-tile),
// and five dependent launches (LayerNorm, fc1, fc2 + pack / epilogue kernels) become one.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "fused_block.h"
#include "split.cuh"
#include "tc_ptx.cuh"

namespace rift {

int tc_plane_map(const void* plane, int rows, int pitch, const CUtensorMap** out);      // gemm_tc.cu

constexpr int FB_THREADS = 320;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: compute
constexpr int FB_CWARPS = 8;
constexpr int FB_CTHREADS = 32 * FB_CWARPS;
constexpr int FB_BM = 128;
constexpr int FB_PANEL = FB_BM * 128;    // bytes of one 64-column bf16 panel of a 128-row operand
constexpr int FB_BOX = 64 * 64 * 2;      // one 64 x 64 bf16 TMA box
constexpr int FB_STAGE = 4 * FB_BOX;     // 128 weight rows x 64 k: [hi box0, box1 | lo box0, box1]
constexpr int FB_NSTAGE = 3;

__host__ __device__ constexpr uint32_t fb_idesc(int n) {      // D = f32, A = B = bf16, both K-major, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FB_BM >> 4) << 24);
}

struct FbMlpKernelArgs {
    const float* X; long long ldx; float* Y; long long ldy;
    int rows, D, N1, C, act;
    const float* ln_g; const float* ln_b; const float* b1; const float* b2;
    float* ln_mean; float* ln_rstd; Planes t2p; float* hpre; long long ldh; Planes hmp;
    unsigned long long* trace;       // profiling aid: CTA 0 writes %globaltimer stamps at the phase boundaries
};
__device__ __forceinline__ unsigned long long fb_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FB_TRACE(slot) do { if (g.trace && blockIdx.x == 0) g.trace[(slot)] = fb_gtimer(); } while (0)

// byte offset of the 16-byte chunk (row, columns [8 * col8, 8 * col8 + 8)) inside one operand plane
__device__ __forceinline__ uint32_t op_chunk_off(int row, int col8) {
    return (uint32_t)((col8 >> 3) * FB_PANEL + (row >> 3) * 1024 + (row & 7) * 128 + (((col8 & 7) ^ (row & 7)) << 4));
}
// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (packed pairs)
__device__ __forceinline__ void split8(const float* v, uint32_t (&h)[4], uint32_t (&l)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
        __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
        __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - __bfloat162float(h0), v[2 * i + 1] - __bfloat162float(h1));
        h[i] = *reinterpret_cast<uint32_t*>(&hh);
        l[i] = *reinterpret_cast<uint32_t*>(&ll);
    }
}

template <int DT>
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                 const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo, FbMlpKernelArgs g,
                 int region_bytes, int tmem_cols) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* region = smem;                                  // LN planes -> hidden planes -> fp32 partial tile
    uint8_t* ring = smem + region_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + FB_NSTAGE * FB_STAGE);
    uint64_t* full = bars;                                   // [3] TMA bytes landed
    uint64_t* empty = bars + FB_NSTAGE;                      // [3] tcgen05.commit
    uint64_t* a_ready = bars + 2 * FB_NSTAGE;                // LayerNorm planes written
    uint64_t* acc1_full = a_ready + 1;
    uint64_t* h_ready = a_ready + 2;
    uint64_t* acc2_full = a_ready + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int D = DT;                                    // model dim: compile time, so the LayerNorm reductions unroll
    const int C = g.C, N1 = g.N1;
    const int c = C > 1 ? (int)cluster_ctarank() : 0;
    const int tile = blockIdx.x / C;
    const int row0 = tile * FB_BM;
    const int KB1 = D >> 6, KB2 = N1 >> 6;                    // k-blocks of GEMM1 / GEMM2
    const int NCH1 = (N1 + 127) >> 7, NCH2 = (D + 127) >> 7;  // 128-wide n-chunks
    const int ACC2 = N1;                                      // TMEM column of the second accumulator

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < FB_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(a_ready, 1); mbar_init(acc1_full, 1); mbar_init(h_ready, 1); mbar_init(acc2_full, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2_lo) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) FB_TRACE(0);
    pdl_wait();
    if (threadIdx.x == 0) FB_TRACE(1);

    if (warp == 0) {
        // ===================== weight stream (TMA) =====================
        if (lane == 0) {
            int it = 0;
            for (int gemm = 0; gemm < 2; ++gemm) {
                const int nch = gemm ? NCH2 : NCH1, kbs = gemm ? KB2 : KB1, ntot = gemm ? D : N1;
                const CUtensorMap* mh = gemm ? &tmW2_hi : &tmW1_hi;
                const CUtensorMap* ml = gemm ? &tmW2_lo : &tmW1_lo;
                for (int j = 0; j < nch; ++j) {
                    const int nb = min(2, (ntot - j * 128) >> 6);
                    for (int kb = 0; kb < kbs; ++kb, ++it) {
                        const int s = it % FB_NSTAGE;
                        const uint32_t ph = (it / FB_NSTAGE) & 1;
                        mbar_wait(&empty[s], ph ^ 1);
                        uint8_t* st = ring + s * FB_STAGE;
                        mbar_arrive_expect_tx(&full[s], (uint32_t)(nb * 2 * FB_BOX));
                        for (int b = 0; b < nb; ++b) {
                            // GEMM1: rows = hidden units of this CTA's slice, k = model dim; GEMM2: rows = model dim, k = slice
                            const int n = gemm ? j * 128 + b * 64 : c * N1 + j * 128 + b * 64;
                            const int k = gemm ? c * N1 + kb * 64 : kb * 64;
                            tma_load_2d(st + b * FB_BOX, mh, &full[s], k, n);
                            tma_load_2d(st + 2 * FB_BOX + b * FB_BOX, ml, &full[s], k, n);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int it = 0;
            for (int gemm = 0; gemm < 2; ++gemm) {
                const int nch = gemm ? NCH2 : NCH1, kbs = gemm ? KB2 : KB1, ntot = gemm ? D : N1;
                mbar_wait(gemm ? h_ready : a_ready, 0);
                tc_fence_after();
                FB_TRACE(gemm ? 6 : 3);                      // operand planes ready
                const uint32_t a_base = smem_u32(region);
                const uint32_t a_lo_off = (uint32_t)(kbs * FB_PANEL);
                for (int j = 0; j < nch; ++j) {
                    const int nb = min(2, (ntot - j * 128) >> 6);
                    const uint32_t idesc = fb_idesc(nb * 64);
                    const uint32_t tacc = tmem_base + (uint32_t)((gemm ? ACC2 : 0) + j * 128);
                    for (int kb = 0; kb < kbs; ++kb, ++it) {
                        const int s = it % FB_NSTAGE;
                        const uint32_t ph = (it / FB_NSTAGE) & 1;
                        mbar_wait(&full[s], ph);
                        tc_fence_after();
                        const uint32_t a_hi = a_base + (uint32_t)(kb * FB_PANEL), a_lo = a_hi + a_lo_off;
                        const uint32_t b_hi = smem_u32(ring + s * FB_STAGE), b_lo = b_hi + 2 * FB_BOX;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t dah = make_sdesc(a_hi + k * 32u, 0), dal = make_sdesc(a_lo + k * 32u, 0);
                            const uint64_t dbh = make_sdesc(b_hi + k * 32u, 0), dbl = make_sdesc(b_lo + k * 32u, 0);
                            umma_bf16(tacc, dal, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);     // small terms first
                            umma_bf16(tacc, dah, dbl, idesc, 1);
                            umma_bf16(tacc, dah, dbh, idesc, 1);
                        }
                        umma_commit(&empty[s]);
                    }
                }
                umma_commit(gemm ? acc2_full : acc1_full);
                FB_TRACE(gemm ? 7 : 4);                      // MMAs of this GEMM issued
            }
        }
    } else {
        // ===================== compute warps =====================
        const int w = warp - 2;
        const int tc = threadIdx.x - 64;                     // 0 .. 255
        const uint32_t reg_u = smem_u32(region);
        // ---- phase 1: LayerNorm -> operand planes
        {
            constexpr int LPR = D >> 3, RPW = 32 / LPR;       // lanes per row (8 columns each), rows per pass
            const int sl = lane % LPR, sub = lane / LPR;
            const uint32_t lo_off = (uint32_t)(KB1 * FB_PANEL);
            float gm[8], bt[8];
            {
                const float4 g0 = *reinterpret_cast<const float4*>(g.ln_g + 8 * sl), g1 = *reinterpret_cast<const float4*>(g.ln_g + 8 * sl + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(g.ln_b + 8 * sl), b1 = *reinterpret_cast<const float4*>(g.ln_b + 8 * sl + 4);
                gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
                bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
            }
            const float invD = 1.f / (float)D;
            // eight row passes per batch: their global loads are issued back to back, unconditionally (row index
            // clamped; rows beyond the valid range are zeroed afterwards), so one L2 / HBM latency covers the batch
            constexpr int NB = 8;
            for (int rb = w * 16; rb < w * 16 + 16; rb += NB * RPW) {
                float4 xa[NB], xb[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const int r = min(rb + u * RPW + sub, FB_BM - 1);
                    const long long grow = min((long long)row0 + r, (long long)g.rows - 1);
                    const float* xp = g.X + grow * g.ldx + 8 * sl;
                    xa[u] = __ldg(reinterpret_cast<const float4*>(xp));
                    xb[u] = __ldg(reinterpret_cast<const float4*>(xp + 4));
                }
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    if (rb + u * RPW >= w * 16 + 16) break;              // warp-uniform (narrow rows: one batch covers the 16 rows)
                    const int r = rb + u * RPW + sub;
                    const long long grow = (long long)row0 + r;
                    const bool ok = grow < g.rows;
                    const float v[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
                    float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
#pragma unroll
                    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    const float mean = s * invD;
                    float q = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
                    for (int o = LPR >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                    const float rstd = rsqrtf(q * invD + 1e-5f);
                    float y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = ok ? (v[i] - mean) * rstd * gm[i] + bt[i] : 0.f;
                    uint32_t h[4], l[4];
                    split8(y, h, l);
                    const uint32_t off = op_chunk_off(r, sl);
                    sts4u(reg_u + off, h[0], h[1], h[2], h[3]);
                    sts4u(reg_u + lo_off + off, l[0], l[1], l[2], l[3]);
                    if (ok && c == 0) {
                        if (g.t2p.on()) {
                            *reinterpret_cast<uint4*>(g.t2p.hi + grow * g.t2p.Kp + 8 * sl) = make_uint4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<uint4*>(g.t2p.lo + grow * g.t2p.Kp + 8 * sl) = make_uint4(l[0], l[1], l[2], l[3]);
                        }
                        if (sl == 0 && g.ln_mean) { g.ln_mean[grow] = mean; g.ln_rstd[grow] = rstd; }
                    }
                }
            }
            fence_async_smem();
            named_bar_sync(1, FB_CTHREADS);
            if (tc == 0) { mbar_arrive(a_ready); FB_TRACE(2); }
        }
        const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = w >> 2;
        const int r = quarter * 32 + lane;                   // tile row of this thread in the TMEM (row-per-lane) view
        const long long grow = (long long)row0 + r;
        const bool ok = grow < g.rows;
        const uint32_t lane_bits = (uint32_t)(quarter * 32) << 16;
        // ---- phase 3: acc1 -> + bias -> activation -> hidden planes (over the LayerNorm planes)
        {
            mbar_wait(acc1_full, 0);
            tc_fence_after();
            if (tc == 0) FB_TRACE(5);
            const int nhalf = N1 >> 1;
            const uint32_t lo_off = (uint32_t)(KB2 * FB_PANEL);
#pragma unroll 1
            for (int cc = 0; cc < nhalf; cc += 32) {
                const int col0 = half * nhalf + cc;           // column inside this CTA's hidden slice
                const int gcol = c * N1 + col0;               // hidden unit
                uint32_t rr[32];
                tmem_ld_32(tmem_base + lane_bits + (uint32_t)col0, rr);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(g.b1 + gcol + j);
                    v[j] = __uint_as_float(rr[j]) + b.x; v[j + 1] = __uint_as_float(rr[j + 1]) + b.y;
                    v[j + 2] = __uint_as_float(rr[j + 2]) + b.z; v[j + 3] = __uint_as_float(rr[j + 3]) + b.w;
                }
                if (g.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (g.act == ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                }
                if (!ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t h[4], l[4];
                    split8(v + 8 * q, h, l);
                    const uint32_t off = op_chunk_off(r, (col0 >> 3) + q);
                    sts4u(reg_u + off, h[0], h[1], h[2], h[3]);
                    sts4u(reg_u + lo_off + off, l[0], l[1], l[2], l[3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            named_bar_sync(1, FB_CTHREADS);
            if (tc == 0) mbar_arrive(h_ready);
            // ---- what the backward reads, written while GEMM2 runs (it only READS the hidden planes; acc1 stays intact)
            if (g.hmp.on()) {
                // hidden planes: shared memory -> global, 4 rows x 128 B per warp instruction (lane = row l / 8, chunk l % 8)
                const int rl = lane >> 3, chk = lane & 7;
                for (int pi = w; pi < 2 * KB2 * (FB_BM / 4); pi += FB_CWARPS) {
                    const int plane = pi / (KB2 * (FB_BM / 4)), rem = pi % (KB2 * (FB_BM / 4));
                    const int panel = rem / (FB_BM / 4), rr4 = (rem % (FB_BM / 4)) * 4 + rl;
                    const long long gr = (long long)row0 + rr4;
                    if (gr < g.rows) {
                        const uint32_t sa = reg_u + (plane ? lo_off : 0u) + (uint32_t)(panel * FB_PANEL + (rr4 >> 3) * 1024 + (rr4 & 7) * 128 +
                                                                                      ((chk ^ (rr4 & 7)) << 4));
                        uint4 val;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(sa));
                        uint16_t* dst = (plane ? g.hmp.lo : g.hmp.hi) + gr * g.hmp.Kp + c * N1 + panel * 64 + chk * 8;
                        *reinterpret_cast<uint4*>(dst) = val;
                    }
                }
            }
            if (g.hpre) {                                    // (warp-uniform: tcgen05.ld is a whole-warp instruction)
#pragma unroll 1
                for (int cc = 0; cc < nhalf; cc += 32) {
                    const int col0 = half * nhalf + cc, gcol = c * N1 + col0;
                    uint32_t rr[32];
                    tmem_ld_32(tmem_base + lane_bits + (uint32_t)col0, rr);
                    tmem_ld_wait();
                    if (!ok) continue;
                    float* hp = g.hpre + grow * g.ldh + gcol;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(g.b1 + gcol + j);
                        *reinterpret_cast<float4*>(hp + j) = make_float4(__uint_as_float(rr[j]) + b.x, __uint_as_float(rr[j + 1]) + b.y,
                                                                       __uint_as_float(rr[j + 2]) + b.z, __uint_as_float(rr[j + 3]) + b.w);
                    }
                }
            }
            named_bar_sync(1, FB_CTHREADS);          // every reader of the hidden planes is done before phase 5 overwrites them
        }
        // ---- phase 5: acc2 -> fp32 partial tile [128, D] in shared memory (16-byte chunks XOR-swizzled with row & 7)
        {
            mbar_wait(acc2_full, 0);
            tc_fence_after();
            if (tc == 0) FB_TRACE(8);
            const int dhalf = D >> 1;
            const uint32_t row_u = reg_u + (uint32_t)r * (uint32_t)(D * 4);
#pragma unroll 1
            for (int cc = 0; cc < dhalf; cc += 32) {
                const int col0 = half * dhalf + cc;
                uint32_t rr[32];
                tmem_ld_32(tmem_base + lane_bits + (uint32_t)(ACC2 + col0), rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const uint32_t ch = (uint32_t)((col0 + j) >> 2);
                    sts4u(row_u + ((ch ^ (uint32_t)(r & 7)) << 4), rr[j], rr[j + 1], rr[j + 2], rr[j + 3]);
                }
            }
            tc_fence_before();
        }
    }
    // every thread of the cluster: the partial tiles are complete
    __syncwarp();
    if (threadIdx.x == 64) FB_TRACE(9);
    if (C > 1) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x == 64) FB_TRACE(10);
    if (warp >= 2) {
        // ---- phase 6: sum the C partial tiles of this CTA's D / C columns, + bias + residual -> Y
        const int tc = threadIdx.x - 64;
        const uint32_t reg_u = smem_u32(region);
        const int cols_per = D / C, cpr = cols_per >> 2;      // 16-byte chunks per row in this CTA's column range
        const int ch0 = (c * cols_per) >> 2;
        constexpr int U = 4;
        for (int idx0 = tc; idx0 < FB_BM * cpr; idx0 += U * FB_CTHREADS) {
            float4 part[U][4], bx[U], xx[U];
            bool okk[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = idx0 + u * FB_CTHREADS;
                const int r = min(idx / cpr, FB_BM - 1), ch = ch0 + idx % cpr;
                const long long grow = (long long)row0 + r;
                okk[u] = idx < FB_BM * cpr && grow < g.rows;
                const uint32_t off = (uint32_t)r * (uint32_t)(D * 4) + (((uint32_t)ch ^ (uint32_t)(r & 7)) << 4);
                if (C > 1) {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        part[u][p] = p < C ? ld_dsmem4(mapa_shared(reg_u + off, (uint32_t)p)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    part[u][0] = lds4(reg_u + off);
                    part[u][1] = part[u][2] = part[u][3] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const long long gc = min(grow, (long long)g.rows - 1);
                bx[u] = __ldg(reinterpret_cast<const float4*>(g.b2 + (ch << 2)));
                xx[u] = __ldg(reinterpret_cast<const float4*>(g.X + gc * g.ldx + (ch << 2)));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!okk[u]) continue;
                const int idx = idx0 + u * FB_CTHREADS;
                const int r = idx / cpr, ch = ch0 + idx % cpr;
                const long long grow = (long long)row0 + r;
                float4 acc = part[u][0];                   // fixed summation order over the cluster ranks
#pragma unroll
                for (int p = 1; p < 4; ++p) { acc.x += part[u][p].x; acc.y += part[u][p].y; acc.z += part[u][p].z; acc.w += part[u][p].w; }
                acc.x += bx[u].x + xx[u].x; acc.y += bx[u].y + xx[u].y; acc.z += bx[u].z + xx[u].z; acc.w += bx[u].w + xx[u].w;
                *reinterpret_cast<float4*>(g.Y + grow * g.ldy + (ch << 2)) = acc;
            }
        }
    }
    __syncwarp();
    if (threadIdx.x == 64) FB_TRACE(11);
    if (C > 1) cluster_sync_all();       // nobody leaves (and frees its shared memory) while a peer still reads it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
        if (lane == 0) FB_TRACE(12);
    }
}

static unsigned long long* g_fb_trace = nullptr;
void set_fused_trace(void* dev_buf) { g_fb_trace = static_cast<unsigned long long*>(dev_buf); }

// ---------------------------------------------------------------------------------- host side
static bool fb_plan(int D, int Hd, int* C_out, int* N1_out) {
    if (!(D == 64 || D == 128 || D == 256)) return false;
    for (int C : {1, 2, 4}) {
        if (Hd % C) continue;
        const int N1 = Hd / C;
        if (N1 > 256 || (N1 % 64) != 0) continue;
        if (D % C || ((D / C) % 4) != 0) continue;
        if (C > 1 && D / C < 16) continue;
        *C_out = C; *N1_out = N1;
        return true;
    }
    return false;
}

bool fused_mlp_shape_ok(int rows, int D, int Hd) {
    int C, N1;
    return rows >= 64 && fb_plan(D, Hd, &C, &N1);
}
// The forward schedule uses the fused sub-block kernels only when asked to (RIFT_B200_FUSED=1): measured on the cfg2 step
// they are level with the unfused LayerNorm + two tcgen05 GEMMs at these row counts (9.99 vs 9.83 ms, profiles/r2b_*),
// because one CTA walks the phases of its tile serially while the unfused kernels spread each phase over all SMs.
bool fused_blocks_enabled() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_FUSED"); return e && atoi(e) != 0; }();
    return on;
}

int launch_fused_mlp(const FusedMlpArgs& a, const TcWeight& w1, const TcWeight& w2, cudaStream_t st) {
    int C = 0, N1 = 0;
    RIFT_REQUIRE(fb_plan(a.D, a.Hd, &C, &N1), "fused_mlp: unsupported (D, hidden) shape");
    RIFT_REQUIRE(w1.N == a.Hd && w1.K == a.D && w2.N == a.D && w2.K == a.Hd && !w1.transpose && !w2.transpose,
                 "fused_mlp: weight planes do not match the block shape");
    RIFT_REQUIRE(a.X && a.Y && a.ln_g && a.ln_b && a.b1 && a.b2 && (a.ldx % 4) == 0 && (a.ldy % 4) == 0, "fused_mlp: bad arguments");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    RIFT_REQUIRE(al16(a.X) && al16(a.Y) && al16(a.ln_g) && al16(a.ln_b) && al16(a.b1) && al16(a.b2) && al16(a.hpre),
                 "fused_mlp: pointers must be 16-byte aligned");
    if (a.rows <= 0) return 0;
    const CUtensorMap *m1h, *m1l, *m2h, *m2l;
    int r;
    if ((r = tc_plane_map(w1.hi, w1.N, w1.Kp, &m1h))) return r;
    if ((r = tc_plane_map(w1.lo, w1.N, w1.Kp, &m1l))) return r;
    if ((r = tc_plane_map(w2.hi, w2.N, w2.Kp, &m2h))) return r;
    if ((r = tc_plane_map(w2.lo, w2.N, w2.Kp, &m2l))) return r;
    const int region = 512 * (a.D > N1 ? a.D : N1);
    const size_t smem = 1024 + (size_t)region + FB_NSTAGE * FB_STAGE + 256;
    const int need = N1 + a.D;
    const int tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    static bool attr = false;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RIFT_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RIFT_CUDA_OK(cudaFuncSetAttribute(fused_mlp_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    FbMlpKernelArgs g;
    g.X = a.X; g.ldx = a.ldx; g.Y = a.Y; g.ldy = a.ldy; g.rows = a.rows; g.D = a.D; g.N1 = N1; g.C = C; g.act = a.act;
    g.ln_g = a.ln_g; g.ln_b = a.ln_b; g.b1 = a.b1; g.b2 = a.b2;
    g.ln_mean = a.ln_mean; g.ln_rstd = a.ln_rstd; g.t2p = a.t2p; g.hpre = a.hpre; g.ldh = a.Hd; g.hmp = a.hmp;
    g.trace = g_fb_trace;
    const int tiles = cdiv(a.rows, FB_BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles * C); cfg.blockDim = dim3(FB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    if (pdl_enabled()) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (C > 1) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = C; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attrs; cfg.numAttrs = na;
    if (a.D == 64) RIFT_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_mlp_kernel<64>, *m1h, *m1l, *m2h, *m2l, g, region, tmem_cols));
    else if (a.D == 128) RIFT_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_mlp_kernel<128>, *m1h, *m1l, *m2h, *m2l, g, region, tmem_cols));
    else RIFT_CUDA_OK(cudaLaunchKernelEx(&cfg, fused_mlp_kernel<256>, *m1h, *m1l, *m2h, *m2l, g, region, tmem_cols));
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift
