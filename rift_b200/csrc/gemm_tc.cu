// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] x W[N,K]^T ), fp32 in / fp32 out.
//
// Numerics: "split-bf16 x3".  Every fp32 operand x is carried as two bf16 planes hi = bf16(x),
// lo = bf16(x - hi) (16 mantissa bits together) and the product is accumulated in fp32 in TMEM as
//     A_lo*W_hi + A_hi*W_lo + A_hi*W_hi
// i.e. three kind::f16 (bf16) UMMAs per k-step.  Measured end to end (tools_precision_study.py):
// plain bf16 operands give 1e-2 relative logit error and single-pass tf32 1-2e-3, both outside the
// 1e-3 parity tolerance of BASELINE.json; the split form stays at the 1e-5 level for 3x the MMA work
// of bf16 (= 1.5x tf32), which these memory-bound shapes (K <= 1024, N <= 1024) hide.
//
// Structure: persistent, warp-specialised, one CTA per SM, 320 threads
//   warp 0      TMA producer: A_hi / A_lo and B_hi / B_lo tiles as 64 x 64 boxes, SWIZZLE_128B,
//               3-stage ring of 64 KB stages, mbarrier complete_tx
//   warp 1      TMEM allocation (2 x BN columns: double-buffered accumulator) + single-thread tcgen05.mma
//               issue; tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2-9   epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), fused
//               alpha / pre-add / BatchNorm scale / bias / ReLU|GELU / residual / beta*C, float4 stores;
//               the epilogue of tile i overlaps the main loop of tile i+1 through the second accumulator
// The A planes are produced by pack_split (below) or directly by the producing kernel (LayerNorm etc.);
// the W planes (and their transposes, for dX = dY W) are refreshed after every optimizer step.
// Operand majorness is a template switch: the forward and the data-gradient products read K-major
// tiles, the weight-gradient product dW = dY^T X reads the SAME [rows, features] planes as MN-major
// tiles (reduction over rows), with split-K over the row range and a deterministic partial reduce.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include <stdlib.h>
#include <string.h>

#include <unordered_map>

#include "common.cuh"
#include "gemm_tc.h"
#include "split.cuh"
#include "tc_ptx.cuh"

namespace rift {

// RIFT_TC_LITE (experiment): two operand stages, four epilogue warps and 64-wide tiles only -> < 114 KB of shared memory per
// CTA, so TWO CTAs (of the same or of different kernels / streams) share an SM and hide each other's TMA / MMA / epilogue latencies
#ifdef RIFT_TC_LITE
constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 2;
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_CTAS_PER_SM = 2;
#else
constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_CTAS_PER_SM = 1;
#endif
constexpr int TC_EPI_HALVES = TC_EPI_WARPS / 4;
constexpr int TC_EPI_PITCH = 32;         // floats per staged row (128 B); 16-byte chunks are XOR-swizzled with (row & 7)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, M = 128, N = BN,
// bits 15 / 16 = A / B major (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_idesc(int bn, bool mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(bn >> 3) << 17) |
           ((uint32_t)(TC_BM >> 4) << 24);
}

struct TcEpilogue {
    const float* bias; const float* colscale;
    const float* pre; long long ldpre; int pre_div;
    const float* res; long long ldres; int res_div; int res_mod;
    int act; float beta; float alpha;
    float* preact;
    const float* dact_ref; long long lddact; int dact;     // multiply by act'(dact_ref[m, n]) (fused activation backward)
};

struct TcKernelArgs {
    float* C; long long ldc;
    int M, N, K;
    int a_mn0, a_k0, b_mn0, b_k0;   // origins of the operand tiles inside their planes
    int splits, kb_per_split; long long split_stride;     // split-K: slab `s` writes raw sums to C + s * split_stride
    TcEpilogue ep;
    Planes out;              // optional split-bf16 copy of the result for the next GEMM (C may then be null)
    int atomic;              // 1: every (tile, split) ADDS its raw sums into C with red.global.add (weight gradients: no partial
                             // slabs, no reduce kernel; C was zeroed by the caller)
    int n_store;             // atomic mode, > 0: only columns [0, n_store) exist in C (arbitrary ldc): scalar reds
    float* colsum;           // MN-major (dW = dY^T X) only: colsum[m] += sum_k A[k, m] (bias gradient), from the tensor core
    int terms;               // 3: split-bf16 x3 (A_lo B_hi + A_hi B_lo + A_hi B_hi); 1: plain bf16 (hi planes only: a third of the
                             // MMA work, half the operand bytes) - backward products whose tolerance allows it
    int late;                // issue griddepcontrol.launch_dependents after the last tile's MMAs (RIFT_B200_PDL_LATE)
    int dbg;                 // bottleneck experiments only (RIFT_B200_TC_DBG)
    unsigned long long* trace;   // profiling aid (rift_b200_debug_gemm_trace): CTA 0 writes %globaltimer stamps
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TC_TRACE(slot) do { if (g.trace && blockIdx.x == 0) g.trace[(slot)] = gtimer(); } while (0)

// Late programmatic-launch trigger: once a CTA has ISSUED the MMAs of its last tile only the epilogue of that tile is left, so
// the stream's next kernel may be scheduled now - its launch latency and set-up then overlap our epilogue and tear-down; it
// still waits in griddepcontrol.wait for this grid to complete, so memory ordering is unchanged.  (Triggering at kernel START
// was measured slower: dependents park on SMs the later waves of this grid and the other streams need.)
__device__ __forceinline__ void pdl_trigger_late(int on) {
    if (on) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
static int tc_late_trigger() {
    static const int v = [] { const char* e = getenv("RIFT_B200_PDL_LATE"); return (e && atoi(e) == 0) ? 0 : 1; }();
    return v;
}

template <int BN>
struct TcSmem {
    static constexpr int A_TILE = TC_BM * TC_BK * 2;     // bytes per bf16 plane
    static constexpr int B_TILE = BN * TC_BK * 2;
    static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    static constexpr int EPI = TC_EPI_WARPS * 32 * TC_EPI_PITCH * 4;       // per-warp 32 x 32 fp32 staging tiles
    static constexpr int ONES = 512;                     // all-ones bf16 B operand of the bias-gradient MMA (MN kernels)
    static constexpr int TOTAL = TC_STAGES * STAGE + (TC_CTAS_PER_SM > 1 ? 0 : 1024) /*align*/ + 256 /*barriers*/ + EPI + ONES;
};
constexpr int TC_MAX_BN = (TC_CTAS_PER_SM > 1) ? 64 : 128;
// per SM: 228 KB, of which 1 KB is reserved per resident CTA
static_assert(TC_CTAS_PER_SM * (TcSmem<TC_MAX_BN>::TOTAL + 1024) <= 228 * 1024, "tcgen05 GEMM shared memory");
constexpr int TC_BOX = 64 * 64 * 2;      // bytes of one 64 x 64 bf16 TMA box

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
// Bias gradient on the tensor core (MN-major weight-gradient kernels): next to dW[m, n] += sum_k dY[k, m] X[k, n] a second
// accumulator of 16 columns takes dY^T x ONES, so every column of it is colsum_k dY[k, m].  The all-ones B operand is a
// 512-byte no-swizzle K-major tile (2 x 2 core matrices of 8 rows x 16 B: LBO = 128, SBO = 256); being constant its
// layout is irrelevant.
constexpr int TC_CS_COLS = 16;
__device__ __forceinline__ uint64_t make_sdesc_ones(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc_colsum() {      // A MN-major (bit 15), B K-major, N = 16, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(TC_CS_COLS >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
constexpr uint32_t tmem_cols_pow2(uint32_t n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// DA: the epilogue multiplies by act'(dact_ref) (fused activation backward) - a separate instantiation so that the
// common kernels do not carry its code
template <int BN, bool MN, bool DA>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcKernelArgs g) {
    pdl_trigger();        // the dependent kernel may be scheduled; our own dependency is awaited after the set-up below
#ifdef RIFT_TC_LITE
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // room for a second CTA per SM: let the next kernel's set-up overlap
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;                            // no static shared memory in this kernel: the dynamic window starts 1024-aligned
#else
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
#endif
    using SM = TcSmem<BN>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * SM::STAGE);
    uint64_t* full = bars;                          // [stages] 1 arrival + TMA bytes
    uint64_t* empty = bars + TC_STAGES;             // [stages] tcgen05.commit
    uint64_t* acc_full = bars + 2 * TC_STAGES;      // [2] accumulator complete (tcgen05.commit)
    uint64_t* acc_empty = bars + 2 * TC_STAGES + 2; // [2] accumulator drained (one arrival per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    float* stage_t = reinterpret_cast<float*>(smem + TC_STAGES * SM::STAGE + 256);
    uint8_t* ones = smem + TC_STAGES * SM::STAGE + 256 + SM::EPI;
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * BN + (MN ? 2 * TC_CS_COLS : 0));
    const bool do_colsum = MN && g.colsum != nullptr;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (MN && threadIdx.x >= 64 && threadIdx.x < 64 + SM::ONES / 4)
        reinterpret_cast<uint32_t*>(ones)[threadIdx.x - 64] = 0x3F803F80u;          // bf16 1.0 pairs
    const int tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + TC_BM - 1) / TC_BM;
    const int tiles_mn = tiles_m * tiles_n;
    const int n_tiles = tiles_mn * g.splits;
    const int num_kb = (g.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], TC_EPI_WARPS); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    if (MN) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the ones tile is read by the async (UMMA) proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();           // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail
    if (threadIdx.x == 0) { if (g.trace && blockIdx.x == 0) g.trace[0] = gtimer(); TC_TRACE(1); }

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int sp = tile / tiles_mn, tmn = tile - sp * tiles_mn;
                const int m0 = (tmn / tiles_n) * TC_BM, n0 = (tmn % tiles_n) * BN;
                const int kb0 = sp * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (it / TC_STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_hi = smem + s * SM::STAGE;
                    uint8_t* a_lo = a_hi + SM::A_TILE;
                    uint8_t* b_hi = a_lo + SM::A_TILE;
                    uint8_t* b_lo = b_hi + SM::B_TILE;
                    if (g.dbg & 2) { mbar_arrive(&full[s]); continue; }
                    const bool lo = g.terms != 1;
                    mbar_arrive_expect_tx(&full[s], lo ? SM::STAGE : SM::STAGE / 2);
                    const int ka = g.a_k0 + kb * TC_BK, kbb = g.b_k0 + kb * TC_BK;
#pragma unroll
                    for (int rb = 0; rb < TC_BM / 64; ++rb) {
                        const int mn = g.a_mn0 + m0 + rb * 64;
                        tma_load_2d(a_hi + rb * TC_BOX, &tmA_hi, &full[s], MN ? mn : ka, MN ? ka : mn);
                        if (lo) tma_load_2d(a_lo + rb * TC_BOX, &tmA_lo, &full[s], MN ? mn : ka, MN ? ka : mn);
                    }
#pragma unroll
                    for (int rb = 0; rb < BN / 64; ++rb) {
                        const int mn = g.b_mn0 + n0 + rb * 64;
                        tma_load_2d(b_hi + rb * TC_BOX, &tmB_hi, &full[s], MN ? mn : kbb, MN ? kbb : mn);
                        if (lo) tma_load_2d(b_lo + rb * TC_BOX, &tmB_lo, &full[s], MN ? mn : kbb, MN ? kbb : mn);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, MN);
            constexpr uint32_t kstep = MN ? 2048u : 32u;          // bytes per UMMA_K = 16 step
            constexpr uint32_t lbo = MN ? (uint32_t)TC_BOX : 0u;
            int it = 0, ti = 0;
            const uint64_t d_ones = make_sdesc_ones(smem_u32(ones));
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
                const int sp = tile / tiles_mn, tmn = tile - sp * tiles_mn;
                const int kb0 = sp * g.kb_per_split, kb1 = min(num_kb, kb0 + g.kb_per_split);
                const int buf = ti & 1;
                mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                const bool cs_tile = do_colsum && (tmn % tiles_n) == 0;       // the first n-tile of each m-tile carries the bias sum
                const uint32_t tcs = tmem_base + (uint32_t)(2 * BN + buf * TC_CS_COLS);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (it / TC_STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    if (it == 0) TC_TRACE(2);
                    const uint32_t a_hi = smem_u32(smem + s * SM::STAGE);
                    const uint32_t a_lo = a_hi + SM::A_TILE;
                    const uint32_t b_hi = a_lo + SM::A_TILE;
                    const uint32_t b_lo = b_hi + SM::B_TILE;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        if (g.dbg & 4) break;
                        const uint64_t dah = make_sdesc(a_hi + k * kstep, lbo), dal = make_sdesc(a_lo + k * kstep, lbo);
                        const uint64_t dbh = make_sdesc(b_hi + k * kstep, lbo), dbl = make_sdesc(b_lo + k * kstep, lbo);
                        const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
                        if (g.terms != 1) {
                            umma_bf16(tacc, dal, dbh, idesc, first);     // small terms first
                            umma_bf16(tacc, dah, dbl, idesc, 1);
                            umma_bf16(tacc, dah, dbh, idesc, 1);
                        } else {
                            umma_bf16(tacc, dah, dbh, idesc, first);
                        }
                        if (MN && cs_tile) {
                            if (g.terms != 1) {
                                umma_bf16(tcs, dal, d_ones, make_idesc_colsum(), first);
                                umma_bf16(tcs, dah, d_ones, make_idesc_colsum(), 1);
                            } else {
                                umma_bf16(tcs, dah, d_ones, make_idesc_colsum(), first);
                            }
                        }
                    }
                    umma_commit(&empty[s]);                      // frees the stage once these MMAs retire
                }
                umma_commit(&acc_full[buf]);
                if (ti == 0) TC_TRACE(3);
            }
            pdl_trigger_late(g.late);
        }
    } else {
        // ===================== epilogue (8 warps) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int half = ew >> 2;                    // column slice of the tile (TC_EPI_HALVES slices)
        const TcEpilogue& e = g.ep;
        int ti = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            const int sp = tile / tiles_mn, tmn = tile - sp * tiles_mn;
            const int m0 = (tmn / tiles_n) * TC_BM, n0 = (tmn % tiles_n) * BN;
            const int buf = ti & 1;
            mbar_wait(&acc_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0 && ti < 24) TC_TRACE(4 + 2 * ti);
            // Row phase: TMEM hands every thread one accumulator ROW.  The fused epilogue math runs there as
            // straight-line, branch-free code (operand addresses are clamped, out-of-range results are zeroed by a
            // select) so the eight 4-column groups of a chunk are independent instruction streams the scheduler can
            // interleave; the row is then parked in a 32 x 32 shared-memory tile whose 16-byte chunks are
            // XOR-swizzled with (row & 7) - conflict free for both phases.  Store phase: lane = (row l/8 + 4*it,
            // column quad l%8), so one warp instruction moves 4 rows x 128 B of fp32 (or 4 x 64 B of each bf16
            // plane) with 16-byte accesses.
            const int mrow0 = m0 + quarter * 32;
            const int m = mrow0 + lane;
            const bool row_ok = m < g.M;
            const int mc = min(m, g.M - 1);                  // clamped row for operand addresses
            const float* pre_row = e.pre ? e.pre + (long long)(mc / e.pre_div) * e.ldpre : nullptr;
            const float* res_row = nullptr;
            if (e.res) res_row = e.res + (long long)(e.res_mod > 0 ? (mc % e.res_mod) : (mc / e.res_div)) * e.ldres;
            const uint32_t tb = smem_u32(stage_t + ew * (32 * TC_EPI_PITCH));
            const uint32_t tb_row = tb + (uint32_t)lane * (TC_EPI_PITCH * 4);              // row phase: my row
            const uint32_t sw_row = (uint32_t)(lane & 7);
            const int srow = lane >> 3, sq = (lane & 7) * 4;                               // store phase: my row offset / quad
            // store-phase row = srow + 4 * it, so (row & 7) = (srow + 4 * (it & 1)) & 7
            const uint32_t tb_st0 = tb + (uint32_t)srow * (TC_EPI_PITCH * 4) + ((((uint32_t)lane & 7) ^ ((uint32_t)srow & 7)) << 4);
            const uint32_t tb_st1 = tb + (uint32_t)srow * (TC_EPI_PITCH * 4) + ((((uint32_t)lane & 7) ^ ((uint32_t)(srow + 4) & 7)) << 4);
            const int rows_here = min(32, g.M - mrow0);
            const long long row_step = 4LL * g.ldc;
            float* c_row = g.C ? g.C + (long long)sp * g.split_stride + (long long)(mrow0 + srow) * g.ldc : nullptr;
            float* pa_row = e.preact ? e.preact + (long long)(mrow0 + srow) * g.ldc : nullptr;
#pragma unroll 1
            for (int cc = 0; cc < BN / TC_EPI_HALVES; cc += 32) {
                const int c0 = half * (BN / TC_EPI_HALVES) + cc;
                uint32_t r[32];
                tmem_ld_32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c0), r);
                tmem_ld_wait();
                if (g.dbg & 1) continue;
                const int ncol = n0 + c0 + sq;                   // this lane's first column in the store phase
                const bool col_ok = ncol < g.N;
                // ---- linear part: r <- (alpha * acc + pre) * colscale + bias
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = min(n0 + c0 + j, g.N - 4);     // clamped (N % 4 == 0, N >= 16)
                    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (e.alpha != 1.f) { v.x *= e.alpha; v.y *= e.alpha; v.z *= e.alpha; v.w *= e.alpha; }
                    if (pre_row) { const float4 t = ld4(pre_row + n); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    if (e.colscale) { const float4 t = ld4(e.colscale + n); v.x *= t.x; v.y *= t.y; v.z *= t.z; v.w *= t.w; }
                    if (e.bias) { const float4 t = ld4(e.bias + n); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    r[j] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y);
                    r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
                }
                if (pa_row) {                                    // value before the activation (saved for backward)
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        sts4(tb_row + ((((uint32_t)j >> 2) ^ sw_row) << 4),
                             make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                         __uint_as_float(r[j + 3])));
                    __syncwarp();
                    float4 w[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) w[it] = lds4(((it & 1) ? tb_st1 : tb_st0) + it * (4 * TC_EPI_PITCH * 4));
                    if (col_ok) {
                        float* pp = pa_row + ncol;
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (srow + 4 * it < rows_here) *reinterpret_cast<float4*>(pp + it * row_step) = w[it];
                    }
                    __syncwarp();
                }
                // ---- activation, residual, zero outside the valid range (uniform switches hoisted out of the loops)
                if (e.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
                } else if (e.act == ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(gelu_erf(__uint_as_float(r[j])));
                }
                if constexpr (DA) {                              // fused activation backward: v *= act'(ref[m, n])
                    const float* dr = e.dact_ref + (long long)mc * e.lddact;
                    float4 t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = ld4(dr + min(n0 + c0 + 4 * j, g.N - 4));
                    if (e.dact == ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (!(t[j].x > 0.f)) r[4 * j] = 0u;
                            if (!(t[j].y > 0.f)) r[4 * j + 1] = 0u;
                            if (!(t[j].z > 0.f)) r[4 * j + 2] = 0u;
                            if (!(t[j].w > 0.f)) r[4 * j + 3] = 0u;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * gelu_erf_grad(t[j].x));
                            r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * gelu_erf_grad(t[j].y));
                            r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * gelu_erf_grad(t[j].z));
                            r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * gelu_erf_grad(t[j].w));
                        }
                    }
                }
                if (res_row) {
                    float4 t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = ld4(res_row + min(n0 + c0 + 4 * j, g.N - 4));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + t[j].x);
                        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + t[j].y);
                        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + t[j].z);
                        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + t[j].w);
                    }
                }
                if (mrow0 + 32 <= g.M && n0 + c0 + 32 <= g.N) {      // interior chunk (warp-uniform): nothing to zero
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        sts4(tb_row + ((((uint32_t)j >> 2) ^ sw_row) << 4),
                             make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                         __uint_as_float(r[j + 3])));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const bool ok = row_ok && (n0 + c0 + j < g.N);
                        sts4(tb_row + ((((uint32_t)j >> 2) ^ sw_row) << 4),
                             make_float4(ok ? __uint_as_float(r[j]) : 0.f, ok ? __uint_as_float(r[j + 1]) : 0.f,
                                         ok ? __uint_as_float(r[j + 2]) : 0.f, ok ? __uint_as_float(r[j + 3]) : 0.f));
                    }
                }
                __syncwarp();
                if (!(g.dbg & 8)) {
                    float4 w[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) w[it] = lds4(((it & 1) ? tb_st1 : tb_st0) + it * (4 * TC_EPI_PITCH * 4));
                    if (c_row && g.atomic) {
                        if (g.n_store > 0) {             // padded column count / unaligned C rows: scalar reds on the real columns
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                if (srow + 4 * it < rows_here) {
                                    float* cp = c_row + it * row_step + ncol;
                                    if (ncol < g.n_store) red_add_f32(cp, w[it].x);
                                    if (ncol + 1 < g.n_store) red_add_f32(cp + 1, w[it].y);
                                    if (ncol + 2 < g.n_store) red_add_f32(cp + 2, w[it].z);
                                    if (ncol + 3 < g.n_store) red_add_f32(cp + 3, w[it].w);
                                }
                        } else if (col_ok) {
                            float* cp = c_row + ncol;
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                if (srow + 4 * it < rows_here) red_add_v4(cp + it * row_step, w[it]);
                        }
                    } else if (c_row && col_ok) {
                        float* cp = c_row + ncol;
                        if (e.beta != 0.f) {
                            float4 o[8];
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                o[it] = srow + 4 * it < rows_here ? *reinterpret_cast<const float4*>(cp + it * row_step)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                if (srow + 4 * it < rows_here)
                                    *reinterpret_cast<float4*>(cp + it * row_step) =
                                        make_float4(w[it].x + e.beta * o[it].x, w[it].y + e.beta * o[it].y,
                                                    w[it].z + e.beta * o[it].z, w[it].w + e.beta * o[it].w);
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it)
                                if (srow + 4 * it < rows_here) *reinterpret_cast<float4*>(cp + it * row_step) = w[it];
                        }
                    }
                    if (g.out.on()) {
                        // columns >= N hold zeros in the tile, which is exactly the planes' zero padding up to Kp
                        const int pc = n0 + c0 + sq;
                        if (pc < g.out.Kp) {
                            uint16_t* dh = g.out.hi + (long long)(mrow0 + srow) * g.out.Kp + pc;
                            uint16_t* dl = g.out.lo + (long long)(mrow0 + srow) * g.out.Kp + pc;
                            const long long pstep = 4LL * g.out.Kp;
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                if (srow + 4 * it < rows_here) {
                                    const float4 x = w[it];
                                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y),
                                                        h2 = __float2bfloat16_rn(x.z), h3 = __float2bfloat16_rn(x.w);
                                    __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
                                    __nv_bfloat162 c2 = __floats2bfloat162_rn(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
                                    __nv_bfloat162 d2 = __floats2bfloat162_rn(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
                                    uint2 uh, ul;
                                    uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b);
                                    ul.x = *reinterpret_cast<uint32_t*>(&c2); ul.y = *reinterpret_cast<uint32_t*>(&d2);
                                    *reinterpret_cast<uint2*>(dh + it * pstep) = uh;
                                    *reinterpret_cast<uint2*>(dl + it * pstep) = ul;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (MN && do_colsum && half == 0 && n0 == 0) {
                // bias gradient: column 0 of the 16-column side accumulator, one out-feature per lane
                const uint32_t v = tmem_ld_1(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * BN + buf * TC_CS_COLS));
                tmem_ld_wait();
                if (row_ok) red_add_f32(g.colsum + m, __uint_as_float(v));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (ew == 0 && lane == 0 && ti < 24) TC_TRACE(5 + 2 * ti);
        }
    }
    if (threadIdx.x == 0) TC_TRACE(60);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
        if (lane == 0) TC_TRACE(61);
    }
}

// ---------------------------------------------------------------------------------- v2: tall GEMMs
// Variant for GEMMs with many M tiles per CTA and K <= KB_MAX * 64 (the history / PointNet encoders: 10^4..10^5 rows,
// K <= 256).  Two changes against gemm_tc_kernel, which at these shapes sits at ~3 us per tile between the per-SM
// L2 -> SMEM rate (256 KB of operand tiles per 128 x 128 tile) and the epilogue's instruction latency:
//   * the CTA keeps ITS weight tile (all k-blocks of B hi / lo for one n-tile) resident in shared memory and walks the
//     M tiles of that n-tile only, so the main loop streams just the A planes (half the bytes);
//   * the epilogue leaves through TMA: the row phase parks the chunk in a swizzled shared-memory tile and one lane
//     issues cp.async.bulk.tensor stores (cp.reduce ... add for beta = 1); M / N tails are clipped by the tensor map,
//     so there are no per-lane loads, stores, predicates or address arithmetic in the store phase.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(smem_src), "r"(c0),
                 "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int TC2_MAX_STAGES = 4;
constexpr int TC2_EPI_BYTES = TC_EPI_WARPS * 4096;          // per-warp 32 x 32 fp32 (or 2 x 32 x 32 bf16) staging tile
constexpr int TC2_A_STAGE = 2 * TC_BM * TC_BK * 2;          // A hi + lo of one k-block

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmOh, const __grid_constant__ CUtensorMap tmOl, TcKernelArgs g, int stages) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int B_PLANE = BN * TC_BK * 2;                 // one bf16 plane of one k-block of the weight tile
    constexpr int B_KB = 2 * B_PLANE;
    const int num_kb = (g.K + TC_BK - 1) / TC_BK;
    uint8_t* epi = smem;                                     // [8 warps][4 KB], 1024-aligned tiles
    uint8_t* bres = smem + TC2_EPI_BYTES;                    // [num_kb][hi | lo]
    uint8_t* astg = bres + num_kb * B_KB;                    // [stages][A hi | A lo]
    uint64_t* bars = reinterpret_cast<uint64_t*>(astg + stages * TC2_A_STAGE);
    uint64_t* full = bars;                                   // [TC2_MAX_STAGES]
    uint64_t* empty = bars + TC2_MAX_STAGES;                 // [TC2_MAX_STAGES]
    uint64_t* b_full = bars + 2 * TC2_MAX_STAGES;            // weight tile landed
    uint64_t* acc_full = b_full + 1;                         // [2]
    uint64_t* acc_empty = acc_full + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (g.N + BN - 1) / BN;
    const int tiles_m = (g.M + TC_BM - 1) / TC_BM;
    const int nt = blockIdx.x % tiles_n;                     // this CTA's n-tile (gridDim.x is a multiple of tiles_n)
    const int mt0 = blockIdx.x / tiles_n, mt_step = gridDim.x / tiles_n;
    const int n0 = nt * BN;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(b_full, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], TC_EPI_WARPS); }
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(b_full, (uint32_t)(num_kb * B_KB));
            for (int kb = 0; kb < num_kb; ++kb) {
                const int kk = g.b_k0 + kb * TC_BK;
#pragma unroll
                for (int rb = 0; rb < BN / 64; ++rb) {
                    const int mn = g.b_mn0 + n0 + rb * 64;
                    tma_load_2d(bres + kb * B_KB + rb * TC_BOX, &tmB_hi, b_full, kk, mn);
                    tma_load_2d(bres + kb * B_KB + B_PLANE + rb * TC_BOX, &tmB_lo, b_full, kk, mn);
                }
            }
            int it = 0;
            for (int mt = mt0; mt < tiles_m; mt += mt_step) {
                const int m0 = mt * TC_BM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_hi = astg + s * TC2_A_STAGE;
                    uint8_t* a_lo = a_hi + TC_BM * TC_BK * 2;
                    mbar_arrive_expect_tx(&full[s], TC2_A_STAGE);
                    const int ka = g.a_k0 + kb * TC_BK;
#pragma unroll
                    for (int rb = 0; rb < TC_BM / 64; ++rb) {
                        const int mn = g.a_mn0 + m0 + rb * 64;
                        tma_load_2d(a_hi + rb * TC_BOX, &tmA_hi, &full[s], ka, mn);
                        tma_load_2d(a_lo + rb * TC_BOX, &tmA_lo, &full[s], ka, mn);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, false);
            mbar_wait(b_full, 0);
            tc_fence_after();
            int it = 0, ti = 0;
            for (int mt = mt0; mt < tiles_m; mt += mt_step, ++ti) {
                const int buf = ti & 1;
                mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(astg + s * TC2_A_STAGE);
                    const uint32_t a_lo = a_hi + TC_BM * TC_BK * 2;
                    const uint32_t b_hi = smem_u32(bres + kb * B_KB);
                    const uint32_t b_lo = b_hi + B_PLANE;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t dah = make_sdesc(a_hi + k * 32u, 0), dal = make_sdesc(a_lo + k * 32u, 0);
                        const uint64_t dbh = make_sdesc(b_hi + k * 32u, 0), dbl = make_sdesc(b_lo + k * 32u, 0);
                        umma_bf16(tacc, dal, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);     // small terms first
                        umma_bf16(tacc, dah, dbl, idesc, 1);
                        umma_bf16(tacc, dah, dbh, idesc, 1);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
            pdl_trigger_late(g.late);
        }
    } else {
        // ===================== epilogue (8 warps) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        const TcEpilogue& e = g.ep;
        const uint32_t tb = smem_u32(epi + ew * 4096);
        const uint32_t tb_row = tb + (uint32_t)lane * 128u;                      // fp32 tile: my row (128 B, SWIZZLE_128B)
        const uint32_t sw128 = (uint32_t)(lane & 7);
        const uint32_t tb_hrow = tb + (uint32_t)lane * 64u;                      // bf16 tiles: my row (64 B, SWIZZLE_64B)
        const uint32_t sw64 = (uint32_t)(lane >> 1) & 3u;
        const bool has_c = g.C != nullptr, has_p = e.preact != nullptr, has_o = g.out.on();
        int ti = 0;
        for (int mt = mt0; mt < tiles_m; mt += mt_step, ++ti) {
            const int m0 = mt * TC_BM;
            const int buf = ti & 1;
            mbar_wait(&acc_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            const int mrow0 = m0 + quarter * 32;
            const int mc = min(mrow0 + lane, g.M - 1);
            const float* pre_row = e.pre ? e.pre + (long long)(mc / e.pre_div) * e.ldpre : nullptr;
            const float* res_row = nullptr;
            if (e.res) res_row = e.res + (long long)(e.res_mod > 0 ? (mc % e.res_mod) : (mc / e.res_div)) * e.ldres;
#pragma unroll 1
            for (int cc = 0; cc < BN / 2; cc += 32) {
                const int c0 = half * (BN / 2) + cc;
                uint32_t r[32];
                tmem_ld_32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = min(n0 + c0 + j, g.N - 4);
                    float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                           __uint_as_float(r[j + 3]));
                    if (pre_row) { const float4 t = ld4(pre_row + n); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    if (e.colscale) { const float4 t = ld4(e.colscale + n); v.x *= t.x; v.y *= t.y; v.z *= t.z; v.w *= t.w; }
                    if (e.bias) { const float4 t = ld4(e.bias + n); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    r[j] = __float_as_uint(v.x); r[j + 1] = __float_as_uint(v.y);
                    r[j + 2] = __float_as_uint(v.z); r[j + 3] = __float_as_uint(v.w);
                }
                if (has_p) {                                     // value before the activation (saved for backward)
                    if (lane == 0) tma_wait_read();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) sts4u(tb_row + ((((uint32_t)j >> 2) ^ sw128) << 4), r[j], r[j + 1], r[j + 2], r[j + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmP, tb, n0 + c0, mrow0); tma_commit(); }
                }
                if (e.act == ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
                } else if (e.act == ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(gelu_erf(__uint_as_float(r[j])));
                }
                if (res_row) {
                    float4 t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t[j] = ld4(res_row + min(n0 + c0 + 4 * j, g.N - 4));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + t[j].x);
                        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + t[j].y);
                        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + t[j].z);
                        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + t[j].w);
                    }
                }
                if (has_c) {
                    if (lane == 0) tma_wait_read();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) sts4u(tb_row + ((((uint32_t)j >> 2) ^ sw128) << 4), r[j], r[j + 1], r[j + 2], r[j + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (e.beta != 0.f) tma_reduce_add_2d(&tmC, tb, n0 + c0, mrow0);
                        else tma_store_2d(&tmC, tb, n0 + c0, mrow0);
                        tma_commit();
                    }
                }
                if (has_o) {
                    // split-bf16 planes of the result: hi tile at tb, lo tile at tb + 2 KB (32 rows x 64 B each);
                    // columns >= N must be zero (they are the planes' padding up to Kp)
                    if (n0 + c0 + 32 > g.N) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n0 + c0 + j >= g.N) r[j] = 0u;
                    }
                    if (lane == 0) tma_wait_read();
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a = __uint_as_float(r[8 * q + 2 * u]), b = __uint_as_float(r[8 * q + 2 * u + 1]);
                            const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
                            __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
                            __nv_bfloat162 ll = __floats2bfloat162_rn(a - __bfloat162float(h0), b - __bfloat162float(h1));
                            hw[u] = *reinterpret_cast<uint32_t*>(&hh);
                            lw[u] = *reinterpret_cast<uint32_t*>(&ll);
                        }
                        const uint32_t off = (((uint32_t)q ^ sw64) << 4);
                        sts4u(tb_hrow + off, hw[0], hw[1], hw[2], hw[3]);
                        sts4u(tb_hrow + 2048u + off, lw[0], lw[1], lw[2], lw[3]);
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOh, tb, n0 + c0, mrow0);
                        tma_store_2d(&tmOl, tb + 2048u, n0 + c0, mrow0);
                        tma_commit();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (lane == 0) tma_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// ---------------------------------------------------------------------------------- operand planes
// fp32 [M, K] (row pitch ld) -> bf16 planes hi / lo [M, Kp], zero padded to Kp
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ src, long long ld, int M, int K, int Kp, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, int vec) {
    pdl_grid_sync_sel();
    const int kq = Kp >> 2;
    const long long total = (long long)M * kq;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        const int m = (int)(e / kq), k = (int)(e - (long long)m * kq) << 2;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = src + (long long)m * ld + k;
        if (vec && k + 3 < K) x = __ldg(reinterpret_cast<const float4*>(p));
        else {
            if (k < K) x.x = __ldg(p);
            if (k + 1 < K) x.y = __ldg(p + 1);
            if (k + 2 < K) x.z = __ldg(p + 2);
            if (k + 3 < K) x.w = __ldg(p + 3);
        }
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2 = __float2bfloat16_rn(x.z),
                            h3 = __float2bfloat16_rn(x.w);
        __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
        __nv_bfloat162 c = __floats2bfloat162_rn(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
        __nv_bfloat162 d = __floats2bfloat162_rn(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
        uint2 uh, ul;
        uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b);
        ul.x = *reinterpret_cast<uint32_t*>(&c); ul.y = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint2*>(hi + (long long)m * Kp + k) = uh;
        *reinterpret_cast<uint2*>(lo + (long long)m * Kp + k) = ul;
    }
}

// pack_split fused with the first stage of a column sum (bias gradient): one pass over dY writes its planes and
// per-slab column partials [slabs][K] (second stage: launch_colsum_final).  block = 8 row lanes x 32 column quads.
__global__ void __launch_bounds__(256)
pack_split_colsum_kernel(const float* __restrict__ src, long long ld, int M, int K, int Kp, __nv_bfloat16* __restrict__ hi,
                         __nv_bfloat16* __restrict__ lo, float* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ float4 sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = (blockIdx.x * 32 + tx) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < Kp) {
        for (int m = blockIdx.y * 8 + ty; m < M; m += gridDim.y * 8) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* p = src + (long long)m * ld + k;
            if (k + 3 < K) x = __ldg(reinterpret_cast<const float4*>(p));
            else {
                if (k < K) x.x = __ldg(p);
                if (k + 1 < K) x.y = __ldg(p + 1);
                if (k + 2 < K) x.z = __ldg(p + 2);
            }
            const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2 = __float2bfloat16_rn(x.z),
                                h3 = __float2bfloat16_rn(x.w);
            __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
            __nv_bfloat162 c = __floats2bfloat162_rn(x.x - __bfloat162float(h0), x.y - __bfloat162float(h1));
            __nv_bfloat162 d = __floats2bfloat162_rn(x.z - __bfloat162float(h2), x.w - __bfloat162float(h3));
            uint2 uh, ul;
            uh.x = *reinterpret_cast<uint32_t*>(&a); uh.y = *reinterpret_cast<uint32_t*>(&b);
            ul.x = *reinterpret_cast<uint32_t*>(&c); ul.y = *reinterpret_cast<uint32_t*>(&d);
            *reinterpret_cast<uint2*>(hi + (long long)m * Kp + k) = uh;
            *reinterpret_cast<uint2*>(lo + (long long)m * Kp + k) = ul;
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
    }
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && k < K) {
        float4 s = sm[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) { const float4 t = sm[i][tx]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        float* o = partial + (long long)blockIdx.y * K + k;
        o[0] = s.x;
        if (k + 1 < K) o[1] = s.y;
        if (k + 2 < K) o[2] = s.z;
        if (k + 3 < K) o[3] = s.w;
    }
}

int pack_colsum_slabs(int M, int Kp) {
    const int chunks = cdiv(Kp, 128);
    return max(1, min(cdiv(M, 16), cdiv(148 * 4, chunks)));
}
// returns the number of partial slabs through *slabs_out (0: the fused form does not apply, use the separate kernels)
int launch_pack_split_colsum(const float* src, long long ld, int M, int K, int Kp, void* hi, void* lo, float* partial,
                             int* slabs_out, cudaStream_t st) {
    *slabs_out = 0;
    if (M <= 0) return 0;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (!vec) return 0;
    const int chunks = cdiv(Kp, 128);
    const int slabs = pack_colsum_slabs(M, Kp);      // four CTAs per SM in total: narrow inputs (one 128-column chunk) get up to 592 row slabs
    launch_k(pack_split_colsum_kernel, dim3(chunks, slabs), 256, 0, st, src, ld, M, K, Kp, static_cast<__nv_bfloat16*>(hi),
                                                                 static_cast<__nv_bfloat16*>(lo), partial);
    RIFT_LAUNCH_OK();
    *slabs_out = slabs;
    return 0;
}

int launch_pack_split(const float* src, long long ld, int M, int K, int Kp, void* hi, void* lo, cudaStream_t st) {
    if (M <= 0) return 0;
    const int vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const long long total = (long long)M * (Kp >> 2);
    launch_k(pack_split_kernel, (int)min((long long)148 * 16, (total + 255) / 256), 256, 0, st, 
        src, ld, M, K, Kp, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), vec);
    RIFT_LAUNCH_OK();
    return 0;
}

// one weight matrix -> planes.  transpose = 0: planes [N, Kp] of W ; transpose = 1: planes [K, Kp] of W^T with
// Kp = pitch over N (N, K, Kp below are then the plane's rows / valid columns / pitch, src is read transposed).
// Work unit = one 32 x 64 tile of a plane; `first` = index of the job's first tile in the launch.
struct SplitJob { const float* src; long long ld; int N, K, Kp; __nv_bfloat16* hi; __nv_bfloat16* lo; long long first; int transpose; };

long long split_job_units(int N, int Kp) { return (long long)((N + 31) / 32) * (Kp / 64); }

__global__ void __launch_bounds__(256)
split_weights_kernel(const SplitJob* __restrict__ jobs, int n_jobs, long long total) {
    pdl_grid_sync();
    __shared__ float tile[64][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        int lo_j = 0, hi_j = n_jobs - 1;                 // last job whose first <= u (uniform across the block)
        while (lo_j < hi_j) {
            const int mid = (lo_j + hi_j + 1) >> 1;
            if (jobs[mid].first <= u) lo_j = mid; else hi_j = mid - 1;
        }
        const SplitJob j = jobs[lo_j];
        const int tiles_c = j.Kp >> 6;
        const int t = (int)(u - j.first);
        const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 64;
        const int c = c0 + 2 * tx;
        if (!j.transpose) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + ty + 8 * i;
                if (r < j.N) {
                    const float* p = j.src + (long long)r * j.ld + c;
                    const float x0 = c < j.K ? __ldg(p) : 0.f, x1 = c + 1 < j.K ? __ldg(p + 1) : 0.f;
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                    *reinterpret_cast<__nv_bfloat162*>(j.hi + (long long)r * j.Kp + c) = __halves2bfloat162(h0, h1);
                    *reinterpret_cast<__nv_bfloat162*>(j.lo + (long long)r * j.Kp + c) =
                        __floats2bfloat162_rn(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
                }
            }
        } else {
            // plane[r][c] = src[c * ld + r]: read 64 source rows x 32 contiguous floats, transpose through shared memory
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cc = c0 + ty + 8 * i, rr = r0 + tx;
                tile[ty + 8 * i][tx] = (cc < j.K && rr < j.N) ? __ldg(j.src + (long long)cc * j.ld + rr) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rl = ty + 8 * i, r = r0 + rl;
                if (r < j.N) {
                    const float x0 = tile[2 * tx][rl], x1 = tile[2 * tx + 1][rl];
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                    *reinterpret_cast<__nv_bfloat162*>(j.hi + (long long)r * j.Kp + c) = __halves2bfloat162(h0, h1);
                    *reinterpret_cast<__nv_bfloat162*>(j.lo + (long long)r * j.Kp + c) =
                        __floats2bfloat162_rn(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
                }
            }
        }
    }
}

int launch_split_weights(const void* jobs_dev, int n_jobs, long long total, cudaStream_t st) {
    if (n_jobs <= 0 || total <= 0) return 0;
    launch_k(split_weights_kernel, (int)min((long long)148 * 16, total), 256, 0, st, static_cast<const SplitJob*>(jobs_dev), n_jobs, total);
    RIFT_LAUNCH_OK();
    return 0;
}

size_t split_job_bytes() { return sizeof(SplitJob); }
void fill_split_job(void* dst, const float* src, long long ld, int N, int K, int Kp, void* hi, void* lo, long long first,
                    int transpose) {
    SplitJob j{src, ld, N, K, Kp, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), first, transpose};
    *static_cast<SplitJob*>(dst) = j;
}

// ---------------------------------------------------------------------------------- host side
static unsigned long long* g_tc_trace = nullptr;      // device buffer of >= 64 u64, or null (rift_b200_debug_gemm_trace)
void set_gemm_tc_trace(void* dev_buf) { g_tc_trace = static_cast<unsigned long long*>(dev_buf); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// bf16 plane [rows, pitch] -> 2-D tensor map with a 64 x 64 box and 128 B swizzle
static int encode_plane_map(void* map_out, const void* plane, int rows, int pitch) {
    auto enc = get_encode();
    RIFT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(plane), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RIFT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

// planes live at fixed addresses (weight cache, workspace), so descriptors are encoded once and reused
struct MapKey {
    const void* p; int rows, pitch;
    bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && pitch == o.pitch; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        return std::hash<const void*>()(k.p) ^ (std::hash<long long>()(((long long)k.rows << 20) ^ k.pitch) * 1000003u);
    }
};
struct MapVal { alignas(64) unsigned char m[128]; };

static int plane_map(const void* plane, int rows, int pitch, const CUtensorMap** out) {
    static std::unordered_map<MapKey, MapVal, MapKeyHash> cache;
    MapKey key{plane, rows, pitch};
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 65536) cache.clear();
        MapVal v;
        int r = encode_plane_map(v.m, plane, rows, pitch);
        if (r) return r;
        it = cache.emplace(key, v).first;
    }
    *out = reinterpret_cast<const CUtensorMap*>(it->second.m);
    return 0;
}

// the same cache for the other tcgen05 kernels (fused_block.cu)
int tc_plane_map(const void* plane, int rows, int pitch, const CUtensorMap** out) { return plane_map(plane, rows, pitch, out); }

// output-side tensor maps of the v2 kernel: fp32 [rows, cols] (row pitch ld floats) with a 32 x 32 box / 128 B swizzle,
// bf16 plane [rows, Kp] with a 32 x 32 box / 64 B swizzle
struct OutMapKey {
    const void* p; long long rows, cols, pitch; int kind;
    bool operator==(const OutMapKey& o) const { return p == o.p && rows == o.rows && cols == o.cols && pitch == o.pitch && kind == o.kind; }
};
struct OutMapKeyHash {
    size_t operator()(const OutMapKey& k) const {
        return std::hash<const void*>()(k.p) ^ (std::hash<long long>()((k.rows << 24) ^ (k.cols << 8) ^ k.pitch ^ k.kind) * 1000003u);
    }
};
static int out_map(const void* ptr, long long rows, long long cols, long long pitch_elems, int kind /*0 fp32, 1 bf16*/,
                   const CUtensorMap** out) {
    static std::unordered_map<OutMapKey, MapVal, OutMapKeyHash> cache;
    OutMapKey key{ptr, rows, cols, pitch_elems, kind};
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() > 65536) cache.clear();
        auto enc = get_encode();
        RIFT_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
        MapVal v;
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * (kind ? 2 : 4)};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(reinterpret_cast<CUtensorMap*>(v.m), kind ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         kind ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RIFT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (output) failed (" + std::to_string((int)r) + ")");
        it = cache.emplace(key, v).first;
    }
    *out = reinterpret_cast<const CUtensorMap*>(it->second.m);
    return 0;
}

// does the weight-resident / TMA-store variant take this GEMM?  (tall: at least two M tiles per CTA; K small enough
// for the weight tile to stay in shared memory; only the epilogue forms it implements)
static bool tc2_takes(const GemmArgs& a, int BN, int sms) {
#ifdef RIFT_TC_LITE
    return false;
#endif
    static const bool off = [] { const char* e = getenv("RIFT_B200_GEMM_V2"); return e && atoi(e) == 0; }();
    if (off) return false;
    const int num_kb = cdiv(a.K, TC_BK);
    const long long tiles = (long long)cdiv(a.M, TC_BM) * cdiv(a.N, BN);
    const int b_bytes = num_kb * 2 * BN * TC_BK * 2;
    const int stages_fit = (227 * 1024 - 1024 - 256 - TC2_EPI_BYTES - b_bytes) / TC2_A_STAGE;
    static const int min_stages = [] { const char* e = getenv("RIFT_B200_GEMM_V2_MIN_STAGES"); return e ? atoi(e) : 3; }();
    static const int min_waves = [] { const char* e = getenv("RIFT_B200_GEMM_V2_MIN_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 2; }();
    return a.terms != 1 && tiles >= (long long)min_waves * sms && stages_fit >= min_stages && a.alpha == 1.f && (a.beta == 0.f || a.beta == 1.f) && !a.dact_ref && a.n_store <= 0 &&
           cdiv(a.N, BN) <= sms;
}

template <int BN>
static int launch_tc2(const GemmArgs& a, const PlaneOp& A, const PlaneOp& B, cudaStream_t st) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int num_kb = cdiv(a.K, TC_BK);
    const int b_bytes = num_kb * 2 * BN * TC_BK * 2;
    int stages = (227 * 1024 - 1024 - 256 - TC2_EPI_BYTES - b_bytes) / TC2_A_STAGE;
    if (stages > TC2_MAX_STAGES) stages = TC2_MAX_STAGES;
    const size_t smem = 1024 + TC2_EPI_BYTES + (size_t)b_bytes + (size_t)stages * TC2_A_STAGE + 256;
    static bool attr = false;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    const CUtensorMap *ma_hi, *ma_lo, *mb_hi, *mb_lo, *mc = nullptr, *mp = nullptr, *moh = nullptr, *mol = nullptr;
    int r;
    if ((r = plane_map(A.hi, A.rows, A.pitch, &ma_hi))) return r;
    if ((r = plane_map(A.lo, A.rows, A.pitch, &ma_lo))) return r;
    if ((r = plane_map(B.hi, B.rows, B.pitch, &mb_hi))) return r;
    if ((r = plane_map(B.lo, B.rows, B.pitch, &mb_lo))) return r;
    if (a.C && (r = out_map(a.C, a.M, a.N, a.ldc, 0, &mc))) return r;
    if (a.preact && (r = out_map(a.preact, a.M, a.N, a.ldc, 0, &mp))) return r;
    if (a.out_planes.on()) {
        if ((r = out_map(a.out_planes.hi, a.M, a.out_planes.Kp, a.out_planes.Kp, 1, &moh))) return r;
        if ((r = out_map(a.out_planes.lo, a.M, a.out_planes.Kp, a.out_planes.Kp, 1, &mol))) return r;
    }
    // unused maps still need a valid object to copy into parameter space
    if (!mc) mc = ma_hi;
    if (!mp) mp = ma_hi;
    if (!moh) moh = ma_hi;
    if (!mol) mol = ma_hi;
    TcKernelArgs g;
    g.M = a.M; g.N = a.N; g.K = a.K;
    g.a_mn0 = A.mn0; g.a_k0 = A.k0; g.b_mn0 = B.mn0; g.b_k0 = B.k0;
    g.out = a.out_planes;
    g.trace = nullptr; g.dbg = 0; g.atomic = 0; g.n_store = 0; g.colsum = nullptr; g.terms = 3; g.late = tc_late_trigger();
    g.splits = 1; g.kb_per_split = num_kb; g.split_stride = 0;
    g.C = a.C; g.ldc = a.ldc;
    g.ep = TcEpilogue{a.bias, a.colscale, a.pre, a.ldpre, a.pre_div, a.res, a.ldres, a.res_div, a.res_mod, a.act, a.beta,
                      a.alpha, a.preact, nullptr, 0, 0};
    const int tiles_n = cdiv(a.N, BN);
    const int grid = (sms / tiles_n) * tiles_n;              // every n-tile gets the same number of CTAs
    launch_k(gemm_tc2_kernel<BN>, grid, TC_THREADS, smem, st, *ma_hi, *ma_lo, *mb_hi, *mb_lo, *mc, *mp, *moh, *mol, g, stages);
    RIFT_LAUNCH_OK();
    return 0;
}

bool gemm_tc_shape_ok(int M, int N, int K) { return K >= 32 && N >= 16 && (N % 4) == 0 && M >= 64; }

bool gemm_tc_eligible(const GemmArgs& a) {
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (a.C == nullptr && !a.out_planes.on()) return false;
    if (a.beta != 0.f && a.C == nullptr) return false;
    if (a.n_store > 0) return gemm_tc_shape_ok(a.M, a.N, a.K) && a.C != nullptr && a.n_store <= a.N;     // scalar reduce-side epilogue
    return gemm_tc_shape_ok(a.M, a.N, a.K) && (a.ldc % 4) == 0 && al16(a.C) && al16(a.bias) && al16(a.colscale) &&
           al16(a.pre) && (a.ldpre % 4) == 0 && al16(a.res) && (a.ldres % 4) == 0 && al16(a.preact) && al16(a.dact_ref) &&
           (a.lddact % 4) == 0;
}

template <int BN, bool MN, bool DA = false>
static int launch_tc(const GemmArgs& a, const PlaneOp& A, const PlaneOp& B, int splits, float* partials, cudaStream_t st) {
    static bool attr = false;
    static int sms = 148;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN, DA>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<BN>::TOTAL));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        attr = true;
    }
    const CUtensorMap *ma_hi, *ma_lo, *mb_hi, *mb_lo;
    int r;
    if ((r = plane_map(A.hi, A.rows, A.pitch, &ma_hi))) return r;
    if ((r = plane_map(A.lo, A.rows, A.pitch, &ma_lo))) return r;
    if ((r = plane_map(B.hi, B.rows, B.pitch, &mb_hi))) return r;
    if ((r = plane_map(B.lo, B.rows, B.pitch, &mb_lo))) return r;
    const int num_kb = cdiv(a.K, TC_BK);
    TcKernelArgs g;
    g.M = a.M; g.N = a.N; g.K = a.K;
    g.a_mn0 = A.mn0; g.a_k0 = A.k0; g.b_mn0 = B.mn0; g.b_k0 = B.k0;
    g.out = a.out_planes;
    g.trace = g_tc_trace; g.late = tc_late_trigger();
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("RIFT_B200_TC_DBG"); dbg = e ? atoi(e) : 0; } g.dbg = dbg; }
    const bool via_ws = !a.atomic_out && (splits > 1 || a.n_store > 0);       // raw sums into the partial buffer, epilogue in the reduce
    g.atomic = 0; g.n_store = 0; g.colsum = nullptr; g.terms = a.terms == 1 ? 1 : 3;
    if (a.atomic_out) {
        // weight gradients: every (tile, split) adds its raw sums straight into C (zeroed by the caller)
        if (splits < 1) splits = 1;
        g.splits = splits; g.kb_per_split = cdiv(num_kb, splits);
        g.splits = cdiv(num_kb, g.kb_per_split);
        g.C = a.C; g.ldc = a.ldc; g.split_stride = 0;
        g.ep = TcEpilogue{nullptr, nullptr, nullptr, 0, 1, nullptr, 0, 1, 0, ACT_NONE, 0.f, 1.f, nullptr, nullptr, 0, 0};
        g.out = Planes();
        g.atomic = 1; g.n_store = a.n_store; g.colsum = MN ? a.colsum_out : nullptr;
    } else if (via_ws) {
        if (splits < 1) splits = 1;
        g.splits = splits; g.kb_per_split = cdiv(num_kb, splits);
        g.splits = cdiv(num_kb, g.kb_per_split);
        g.C = partials; g.ldc = a.N; g.split_stride = (long long)a.M * a.N;
        g.ep = TcEpilogue{nullptr, nullptr, nullptr, 0, 1, nullptr, 0, 1, 0, ACT_NONE, 0.f, 1.f, nullptr, nullptr, 0, 0};
        g.out = Planes();
    } else {
        g.splits = 1; g.kb_per_split = num_kb; g.split_stride = 0;
        g.C = a.C; g.ldc = a.ldc;
        g.ep = TcEpilogue{a.bias, a.colscale, a.pre, a.ldpre, a.pre_div, a.res, a.ldres, a.res_div, a.res_mod, a.act, a.beta,
                          a.alpha, a.preact, a.dact_ref, a.lddact, a.dact};
    }
    const int n_tiles = cdiv(a.N, BN) * cdiv(a.M, TC_BM) * g.splits;
    launch_k(gemm_tc_kernel<BN, MN, DA>, min(n_tiles, TC_CTAS_PER_SM * sms), TC_THREADS, TcSmem<BN>::TOTAL, st, *ma_hi, *ma_lo, *mb_hi, *mb_lo, g);
    RIFT_LAUNCH_OK();
    if (via_ws) {
        if (a.n_store > 0) { GemmArgs b = a; b.N = a.n_store; return launch_splitk_reduce(partials, g.splits, b, st, a.N); }
        return launch_splitk_reduce(partials, g.splits, a, st);
    }
    return 0;
}

int launch_gemm_tc_ex(const GemmArgs& a, const PlaneOp& A, const PlaneOp& B, bool mn_major, int splits, float* partials,
                      cudaStream_t st) {
    RIFT_REQUIRE(gemm_tc_eligible(a), "gemm_tc: shape / layout not eligible");
    RIFT_REQUIRE(A.pitch % 64 == 0 && B.pitch % 64 == 0, "gemm_tc: plane pitches must be multiples of 64");
    RIFT_REQUIRE(a.atomic_out || (splits <= 1 && a.n_store <= 0) || partials != nullptr, "gemm_tc: split-K / padded-N needs a partial buffer");
    RIFT_REQUIRE(!a.atomic_out || (a.C != nullptr && a.alpha == 1.f && !a.bias && !a.res && !a.pre && !a.dact_ref && !a.out_planes.on()),
                 "gemm_tc: the atomic-accumulate form takes no epilogue");
    RIFT_REQUIRE(a.colsum_out == nullptr || (mn_major && a.atomic_out), "gemm_tc: colsum_out needs the MN-major atomic form");
    if (a.M <= 0 || a.N <= 0) return 0;
    // tile width: 64-wide tiles when they shorten the longest per-CTA queue (one persistent CTA per SM): small grids
    // that leave SMs idle with 128-wide tiles, and N that is not a multiple of 128 (e.g. 192 = 3 x 64)
    bool narrow = a.N <= 64;
#ifdef RIFT_TC_LITE
    narrow = true;
#endif
    if (!narrow && splits <= 1) {
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
        const long long t128 = (long long)cdiv(a.M, TC_BM) * cdiv(a.N, 128), t64 = (long long)cdiv(a.M, TC_BM) * cdiv(a.N, 64);
        const double w128 = (double)((t128 + sms - 1) / sms), w64 = 0.5 * (double)((t64 + sms - 1) / sms);
        narrow = w64 + 0.2 < w128;
    }
    if (mn_major) {
        if (narrow) return launch_tc<64, true>(a, A, B, splits, partials, st);
        return launch_tc<128, true>(a, A, B, splits, partials, st);
    }
    if (!mn_major && splits <= 1) {
        static int sms2 = 0;
        if (!sms2) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms2, cudaDevAttrMultiProcessorCount, dev); if (sms2 <= 0) sms2 = 148; }
        // the weight tile must leave room for >= 3 operand stages: fall back to 64-wide tiles when 128 do not fit
        if (!narrow && tc2_takes(a, 128, sms2)) return launch_tc2<128>(a, A, B, st);
        // (measured: 46080 x 256 x 256 runs 30.4 us on gemm_tc_kernel, 34.6 us here with 128-wide tiles / 2 stages and
        //  38.5 us with 64-wide tiles / 4 stages, so that fallback is off unless asked for)
        static const bool try64 = [] { const char* e = getenv("RIFT_B200_GEMM_V2_TRY64"); return e && atoi(e) != 0; }();
        if ((narrow || try64) && tc2_takes(a, 64, sms2)) return launch_tc2<64>(a, A, B, st);
    }
    if (a.dact_ref) {
        RIFT_REQUIRE(splits <= 1 && a.n_store <= 0, "gemm_tc: fused activation backward needs the direct epilogue");
        if (narrow) return launch_tc<64, false, true>(a, A, B, splits, partials, st);
        return launch_tc<128, false, true>(a, A, B, splits, partials, st);
    }
    if (narrow) return launch_tc<64, false>(a, A, B, splits, partials, st);
    return launch_tc<128, false>(a, A, B, splits, partials, st);
}

// ---------------------------------------------------------------------------------- grouped weight gradients
// Up to WG_MAXP independent products dW_p[M_p, N_p] += dY_p^T X_p (+ bias gradient colsum(dY_p)) in ONE persistent launch.
// The backward of a transformer block produces ~10 such products, each a grid of 24 .. 96 CTAs that holds 200 KB of shared
// memory per SM for 3 k-blocks; queued behind each other on the side stream they cost ~20 us apiece (launch, pipeline fill,
// drain) and compete SM by SM with the data-gradient chain.  Here their (tile, split) work items form one queue: 148 CTAs
// walk it with the operand ring and the two TMEM accumulators staying busy across items of different products.
// Same roles / barriers / MMA issue as gemm_tc_kernel<128, true>; the epilogue is the atomic one only.
// Every product: MN-major operands (planes [rows, features]), N_p % 4 == 0, red.global.add.v4.f32 into dW (zeroed by the caller).
constexpr int WG_MAXP = 12;                      // 12 x 640 B of descriptors + header = 7.7 KB of kernel parameters (CUDA >= 12.1: 32 KB limit)
struct alignas(64) WgProblem {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    float* C; float* colsum; long long ldc;
    int M, N, K;
    int tiles_n, tiles_mn, kb_per_split, num_kb, item0;
    int pad_[3];
};
struct alignas(64) WgGroup {
    WgProblem p[WG_MAXP];
    int n_problems, n_items, terms, late, pad_[12];
};
static_assert(sizeof(WgGroup) <= 8192, "grouped weight-gradient parameters");

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_group_kernel(const __grid_constant__ WgGroup grp) {
    pdl_trigger();
    constexpr int BN = 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using SM = TcSmem<BN>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * SM::STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC_STAGES;
    uint64_t* acc_full = bars + 2 * TC_STAGES;
    uint64_t* acc_empty = bars + 2 * TC_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    float* stage_t = reinterpret_cast<float*>(smem + TC_STAGES * SM::STAGE + 256);
    uint8_t* ones = smem + TC_STAGES * SM::STAGE + 256 + SM::EPI;
    constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * BN + 2 * TC_CS_COLS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x >= 64 && threadIdx.x < 64 + SM::ONES / 4) reinterpret_cast<uint32_t*>(ones)[threadIdx.x - 64] = 0x3F803F80u;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], TC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    const int n_items = grp.n_items, np = grp.n_problems;
    // item -> problem (items of a problem are contiguous; <= WG_MAXP problems: linear scan)
    auto find = [&](int item) { int pi = 0; while (pi + 1 < np && item >= grp.p[pi + 1].item0) ++pi; return pi; };

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const WgProblem& P = grp.p[find(item)];
                const int loc = item - P.item0;
                const int sp = loc / P.tiles_mn, tmn = loc - sp * P.tiles_mn;
                const int m0 = (tmn / P.tiles_n) * TC_BM, n0 = (tmn % P.tiles_n) * BN;
                const int kb0 = sp * P.kb_per_split, kb1 = min(P.num_kb, kb0 + P.kb_per_split);
                const bool lo = grp.terms != 1;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (it / TC_STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* a_hi = smem + s * SM::STAGE;
                    uint8_t* a_lo = a_hi + SM::A_TILE;
                    uint8_t* b_hi = a_lo + SM::A_TILE;
                    uint8_t* b_lo = b_hi + SM::B_TILE;
                    mbar_arrive_expect_tx(&full[s], lo ? SM::STAGE : SM::STAGE / 2);
                    const int kr = kb * TC_BK;
#pragma unroll
                    for (int rb = 0; rb < TC_BM / 64; ++rb) {
                        tma_load_2d(a_hi + rb * TC_BOX, &P.a_hi, &full[s], m0 + rb * 64, kr);
                        if (lo) tma_load_2d(a_lo + rb * TC_BOX, &P.a_lo, &full[s], m0 + rb * 64, kr);
                    }
#pragma unroll
                    for (int rb = 0; rb < BN / 64; ++rb) {
                        tma_load_2d(b_hi + rb * TC_BOX, &P.b_hi, &full[s], n0 + rb * 64, kr);
                        if (lo) tma_load_2d(b_lo + rb * TC_BOX, &P.b_lo, &full[s], n0 + rb * 64, kr);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN, true);
            constexpr uint32_t kstep = 2048u, lbo = (uint32_t)TC_BOX;
            int it = 0, ti = 0;
            const uint64_t d_ones = make_sdesc_ones(smem_u32(ones));
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
                const WgProblem& P = grp.p[find(item)];
                const int loc = item - P.item0;
                const int sp = loc / P.tiles_mn, tmn = loc - sp * P.tiles_mn;
                const int kb0 = sp * P.kb_per_split, kb1 = min(P.num_kb, kb0 + P.kb_per_split);
                const int buf = ti & 1;
                mbar_wait(&acc_empty[buf], ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
                const bool cs_tile = P.colsum != nullptr && (tmn % P.tiles_n) == 0;
                const uint32_t tcs = tmem_base + (uint32_t)(2 * BN + buf * TC_CS_COLS);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (it / TC_STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * SM::STAGE);
                    const uint32_t a_lo = a_hi + SM::A_TILE;
                    const uint32_t b_hi = a_lo + SM::A_TILE;
                    const uint32_t b_lo = b_hi + SM::B_TILE;
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t dah = make_sdesc(a_hi + k * kstep, lbo), dal = make_sdesc(a_lo + k * kstep, lbo);
                        const uint64_t dbh = make_sdesc(b_hi + k * kstep, lbo), dbl = make_sdesc(b_lo + k * kstep, lbo);
                        const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
                        if (grp.terms != 1) {
                            umma_bf16(tacc, dal, dbh, idesc, first);
                            umma_bf16(tacc, dah, dbl, idesc, 1);
                            umma_bf16(tacc, dah, dbh, idesc, 1);
                        } else {
                            umma_bf16(tacc, dah, dbh, idesc, first);
                        }
                        if (cs_tile) {
                            if (grp.terms != 1) {
                                umma_bf16(tcs, dal, d_ones, make_idesc_colsum(), first);
                                umma_bf16(tcs, dah, d_ones, make_idesc_colsum(), 1);
                            } else {
                                umma_bf16(tcs, dah, d_ones, make_idesc_colsum(), first);
                            }
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
            pdl_trigger_late(grp.late);
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        int ti = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ti) {
            const WgProblem& P = grp.p[find(item)];
            const int loc = item - P.item0;
            const int tmn = loc % P.tiles_mn;
            const int m0 = (tmn / P.tiles_n) * TC_BM, n0 = (tmn % P.tiles_n) * BN;
            const int buf = ti & 1;
            mbar_wait(&acc_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            const int mrow0 = m0 + quarter * 32;
            const int m = mrow0 + lane;
            const bool row_ok = m < P.M;
            const uint32_t tb = smem_u32(stage_t + ew * (32 * TC_EPI_PITCH));
            const uint32_t tb_row = tb + (uint32_t)lane * (TC_EPI_PITCH * 4);
            const uint32_t sw_row = (uint32_t)(lane & 7);
            const int srow = lane >> 3, sq = (lane & 7) * 4;
            const uint32_t tb_st0 = tb + (uint32_t)srow * (TC_EPI_PITCH * 4) + ((((uint32_t)lane & 7) ^ ((uint32_t)srow & 7)) << 4);
            const uint32_t tb_st1 = tb + (uint32_t)srow * (TC_EPI_PITCH * 4) + ((((uint32_t)lane & 7) ^ ((uint32_t)(srow + 4) & 7)) << 4);
            const int rows_here = min(32, P.M - mrow0);
            const long long row_step = 4LL * P.ldc;
            float* c_row = P.C + (long long)(mrow0 + srow) * P.ldc;
#pragma unroll 1
            for (int cc = 0; cc < BN / TC_EPI_HALVES; cc += 32) {
                const int c0 = half * (BN / TC_EPI_HALVES) + cc;
                if (n0 + c0 >= P.N) break;                       // warp-uniform: nothing of this chunk exists
                uint32_t r[32];
                tmem_ld_32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + c0), r);
                tmem_ld_wait();
                const int ncol = n0 + c0 + sq;
                const bool col_ok = ncol < P.N;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts4(tb_row + ((((uint32_t)j >> 2) ^ sw_row) << 4),
                         make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                __syncwarp();
                float4 w[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) w[it] = lds4(((it & 1) ? tb_st1 : tb_st0) + it * (4 * TC_EPI_PITCH * 4));
                if (col_ok) {
                    float* cp = c_row + ncol;
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        if (srow + 4 * it < rows_here) red_add_v4(cp + it * row_step, w[it]);
                }
                __syncwarp();
            }
            if (P.colsum != nullptr && half == 0 && n0 == 0) {
                const uint32_t v = tmem_ld_1(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * BN + buf * TC_CS_COLS));
                tmem_ld_wait();
                if (row_ok) red_add_f32(P.colsum + m, __uint_as_float(v));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

bool wgrad_group_takes(int M, int N, int K, long long ldc, const void* C) {
    return M >= 64 && N > 64 && (N % 4) == 0 && K >= 64 && (ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0;
}

int launch_wgrad_group(const WgradItem* items, int n, int terms, cudaStream_t st) {
    if (n <= 0) return 0;
    RIFT_REQUIRE(n <= WG_MAXP, "wgrad_group: too many products for one launch");
    static bool attr = false;
    static int sms = 148;
    if (!attr) {
        RIFT_CUDA_OK(cudaFuncSetAttribute(wgrad_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::TOTAL));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        attr = true;
    }
    WgGroup g;
    memset(&g, 0, sizeof(g));
    long long kb_tiles = 0;
    for (int i = 0; i < n; ++i) {
        const WgradItem& w = items[i];
        RIFT_REQUIRE(wgrad_group_takes(w.M, w.N, w.K, w.ldc, w.C), "wgrad_group: product not eligible");
        RIFT_REQUIRE(w.A.pitch % 64 == 0 && w.B.pitch % 64 == 0 && w.A.mn0 == 0 && w.A.k0 == 0 && w.B.mn0 == 0 && w.B.k0 == 0,
                     "wgrad_group: operand planes must start at the origin with a pitch that is a multiple of 64");
        kb_tiles += (long long)cdiv(w.M, TC_BM) * cdiv(w.N, 128) * cdiv(w.K, TC_BK);
    }
    // k-blocks per work item: about two items per CTA over the whole group (measured: 2 -> 8.44, 4 -> 8.46, 8 -> 8.51 ms per step),
    // at least 2 and at most 16 k-blocks each
    static const int per_cta = [] { const char* e = getenv("RIFT_B200_WGRAD_GROUP_ITEMS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 2; }();
    int kbs = (int)((kb_tiles + (long long)sms * per_cta - 1) / ((long long)sms * per_cta));
    kbs = kbs < 2 ? 2 : (kbs > 16 ? 16 : kbs);
    int item0 = 0;
    for (int i = 0; i < n; ++i) {
        const WgradItem& w = items[i];
        WgProblem& P = g.p[i];
        const CUtensorMap* mp;
        int r;
        if ((r = plane_map(w.A.hi, w.A.rows, w.A.pitch, &mp))) return r; P.a_hi = *mp;
        if ((r = plane_map(w.A.lo, w.A.rows, w.A.pitch, &mp))) return r; P.a_lo = *mp;
        if ((r = plane_map(w.B.hi, w.B.rows, w.B.pitch, &mp))) return r; P.b_hi = *mp;
        if ((r = plane_map(w.B.lo, w.B.rows, w.B.pitch, &mp))) return r; P.b_lo = *mp;
        P.C = w.C; P.colsum = w.colsum; P.ldc = w.ldc;
        P.M = w.M; P.N = w.N; P.K = w.K;
        P.tiles_n = cdiv(w.N, 128);
        P.tiles_mn = P.tiles_n * cdiv(w.M, TC_BM);
        P.num_kb = cdiv(w.K, TC_BK);
        P.kb_per_split = min(kbs, P.num_kb);
        P.item0 = item0;
        item0 += P.tiles_mn * cdiv(P.num_kb, P.kb_per_split);
    }
    g.n_problems = n; g.n_items = item0; g.terms = terms == 1 ? 1 : 3; g.late = tc_late_trigger();
    // grid: one persistent CTA per SM walking the queue, or (RIFT_B200_WGRAD_GROUP_ONESHOT=1) one CTA per work item - short-lived
    // CTAs hand their SM back after every item, so a higher-priority stream gets in between
    static const bool oneshot = [] { const char* e = getenv("RIFT_B200_WGRAD_GROUP_ONESHOT"); return e && atoi(e) != 0; }();
    static const int grid_cap = [] { const char* e = getenv("RIFT_B200_WGRAD_GROUP_CTAS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 0; }();
    int grid = oneshot ? item0 : min(item0, sms);
    if (grid_cap > 0 && !oneshot) grid = min(grid, grid_cap);
    launch_k(wgrad_group_kernel, grid, TC_THREADS, TcSmem<128>::TOTAL, st, g);
    RIFT_LAUNCH_OK();
    return 0;
}

int launch_gemm_tc(const GemmArgs& a, const void* a_hi, const void* a_lo, int Kp, const TcWeight& w, int n0, int k0,
                   cudaStream_t st) {
    RIFT_REQUIRE(n0 + a.N <= w.N && k0 + a.K <= w.Kp, "gemm_tc: weight slice out of range");
    RIFT_REQUIRE(Kp >= a.K, "gemm_tc: bad activation plane pitch");
    PlaneOp A{a_hi, a_lo, a.M, Kp, 0, 0};
    PlaneOp B{w.hi, w.lo, w.N, w.Kp, n0, k0};
    return launch_gemm_tc_ex(a, A, B, false, 1, nullptr, st);
}

}  // namespace rift
