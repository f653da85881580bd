// GRPO / RIFT / PPO / REINFORCE objective kernels and the group-relative advantage kernel.
//
// Reference semantics (file:line relative to /root/reference):
//   group advantage        rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:466-470
//   RIFT objective         rift/cbv/planning/fine_tuner/rlft/rift_pluto/rift_trainer.py:140-182
//   GRPO objective         rift/cbv/planning/fine_tuner/rlft/grpo_pluto/grpo_trainer.py:140-194
//   PPO objective          rift/cbv/planning/fine_tuner/rlft/ppo_pluto/ppo_trainer.py:161-183
//   REINFORCE objective    rift/cbv/planning/fine_tuner/rlft/reinforce_pluto/reinforce_trainer.py:120-170
//   GAE / returns          .../ppo_pluto/ppo_datamodule.py:22-37,163 ; .../reinforce_datamodule.py:19-38
//
// All of these are HBM-bound byte movers: coalesced (vectorised where alignment allows) loads,
// shared-memory staging of the group block, warp-shuffle reductions, fp64 with explicit
// round-to-nearest intrinsics (no FMA contraction) where the reference computes in numpy float64.
#include <stdlib.h>

#include "common.cuh"
#include "ops.h"

namespace rift {

// =====================================================================================
// group-relative advantage, bit-exact with numpy float64 mean/std (pairwise summation)
// =====================================================================================
// numpy's pairwise_sum (numpy/_core/src/umath/loops_utils.h.src) for a contiguous array:
//   n < 8      : sequential from 0
//   n <= 128   : eight accumulators over strides of 8, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail sequential
//   n > 128    : split at n/2 rounded down to a multiple of 8, recurse, add the halves
// Eight lanes of a warp play the eight accumulators; the butterfly (xor 1,2,4) reproduces the
// combine tree exactly because IEEE addition is commutative.

template <bool SQ>
__device__ __forceinline__ double pw_term(const double* a, int i) {
    const double v = a[i];
    return SQ ? __dmul_rn(v, v) : v;
}

template <bool SQ>
__device__ __forceinline__ double pw_leaf8(const double* a, int n, int lane8, unsigned gmask) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, pw_term<SQ>(a, i));
        return res;
    }
    double r = pw_term<SQ>(a, lane8);
    const int nfull = n - (n & 7);
    for (int i = 8; i < nfull; i += 8) r = __dadd_rn(r, pw_term<SQ>(a, i + lane8));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    for (int i = nfull; i < n; ++i) r = __dadd_rn(r, pw_term<SQ>(a, i));
    return r;
}

template <bool SQ>
__device__ double pw_sum8(const double* a, int n, int lane8, unsigned gmask) {
    if (n <= 128) return pw_leaf8<SQ>(a, n, lane8, gmask);
    // explicit post-order walk of the split tree (depth <= 24 covers n < 2^31)
    int s_off[26], s_n[26], s_state[26];
    double vals[26];
    int top = 0, vtop = 0;
    s_off[0] = 0; s_n[0] = n; s_state[0] = 0; top = 1;
    while (top > 0) {
        const int f = top - 1;
        if (s_n[f] <= 128) {
            vals[vtop++] = pw_leaf8<SQ>(a + s_off[f], s_n[f], lane8, gmask);
            --top;
            continue;
        }
        int n2 = s_n[f] / 2;
        n2 -= n2 % 8;
        if (s_state[f] == 0) {
            s_state[f] = 1;
            s_off[top] = s_off[f]; s_n[top] = n2; s_state[top] = 0; ++top;
        } else if (s_state[f] == 1) {
            s_state[f] = 2;
            s_off[top] = s_off[f] + n2; s_n[top] = s_n[f] - n2; s_state[top] = 0; ++top;
        } else {
            const double r = __dadd_rn(vals[vtop - 2], vals[vtop - 1]);
            vtop -= 2;
            vals[vtop++] = r;
            --top;
        }
    }
    return vals[0];
}

// x / d, correctly rounded (== __ddiv_rn), for a divisor shared by many quotients: y = RN(1 / d) once, then
// q0 = RN(x y) and two FMA residual corrections.  q1 is a faithful quotient, so by Markstein's theorem q2 is the correctly
// rounded one, provided y is the correctly rounded reciprocal, d's significand is not all ones and nothing leaves the
// normal range (`ok` = the divisor passed that test; an operand outside +-2^400 takes the division itself).
__device__ __forceinline__ bool div_shared_ok(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    const int e = (int)((b >> 52) & 0x7ff);
    return e > 1023 - 400 && e < 1023 + 400 && (b & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull;
}
__device__ __forceinline__ double div_shared(double x, double d, double y, bool ok) {
    const int xe = (int)(((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ff);
    if (ok && xe > 1023 - 400 && xe < 1023 + 400) {
        const double q0 = __dmul_rn(x, y);
        const double q1 = __fma_rn(__fma_rn(-d, q0, x), y, q0);
        return __fma_rn(__fma_rn(-d, q1, x), y, q1);
    }
    return __ddiv_rn(x, d);                              // zero, subnormal, huge, inf, nan
}

// one group held in (shared or global) memory at x[0..n): writes adv to out[0..n)
__device__ __forceinline__ void advantage_group8(double* x, double* out, int n, int lane8, unsigned gmask) {
    const double dn = (double)n;
    const bool nok = div_shared_ok(dn);
    const double yn = __drcp_rn(dn);
    const double mean = div_shared(pw_sum8<false>(x, n, lane8, gmask), dn, yn, nok);
    for (int i = lane8; i < n; i += 8) out[i] = __dadd_rn(x[i], -mean);   // x - mean
    __syncwarp(gmask);
    const double var = div_shared(pw_sum8<true>(out, n, lane8, gmask), dn, yn, nok);
    const double sd = __dadd_rn(__dsqrt_rn(var), 1e-5);
    __syncwarp(gmask);
    // __ddiv_rn costs ~25 dependent fp64 instructions per quotient; the shared-divisor form 5
    const bool sok = div_shared_ok(sd);
    const double ys = sok ? __drcp_rn(sd) : 0.0;
    for (int i = lane8; i < n; i += 8) out[i] = div_shared(out[i], sd, ys, sok);
}

// Fixed group size: a block stages GPB groups through shared memory with 16-byte loads.
constexpr int ADV_THREADS = 256;
constexpr int ADV_GPB = ADV_THREADS / 8;   // 32 groups per block pass

__global__ void __launch_bounds__(ADV_THREADS)
group_advantage_fixed_kernel(const double* __restrict__ ret, double* __restrict__ adv,
                             long long n_groups, int G, int Gp) {
    pdl_grid_sync();
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    const int lane8 = tid & 7;
    const int gl = tid >> 3;
    const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
    const float invG = 1.f / (float)G;
    const long long n_chunks = (n_groups + ADV_GPB - 1) / ADV_GPB;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long g0 = chunk * ADV_GPB;
        const int ng = (int)min((long long)ADV_GPB, n_groups - g0);
        const long long base = g0 * G;
        const int total = ng * G;
        // ---- coalesced global -> shared (double2 when the chunk base is 16B aligned)
        if (((base & 1) == 0) && ((G & 1) == 0)) {
            const double2* src = reinterpret_cast<const double2*>(ret + base);
            for (int e = tid; e < total / 2; e += ADV_THREADS) {
                const double2 v = __ldg(src + e);
                const int i = 2 * e;
                const int g = (int)(((float)i + 0.5f) * invG), k = i - g * G;   // exact for i < 2^14 (no integer division)
                sm[g * Gp + k] = v.x;
                sm[g * Gp + k + 1] = v.y;
            }
        } else {
            for (int i = tid; i < total; i += ADV_THREADS) {
                const int g = (int)(((float)i + 0.5f) * invG), k = i - g * G;   // exact for i < 2^14 (no integer division)
                sm[g * Gp + k] = __ldg(ret + base + i);
            }
        }
        __syncthreads();
        if (gl < ng) {
            double* x = sm + gl * Gp;
            advantage_group8(x, x, G, lane8, gmask);
        }
        __syncthreads();
        // ---- shared -> global, coalesced
        if (((base & 1) == 0) && ((G & 1) == 0)) {
            double2* dst = reinterpret_cast<double2*>(adv + base);
            for (int e = tid; e < total / 2; e += ADV_THREADS) {
                const int i = 2 * e;
                const int g = (int)(((float)i + 0.5f) * invG), k = i - g * G;   // exact for i < 2^14 (no integer division)
                dst[e] = make_double2(sm[g * Gp + k], sm[g * Gp + k + 1]);
            }
        } else {
            for (int i = tid; i < total; i += ADV_THREADS) {
                const int g = (int)(((float)i + 0.5f) * invG), k = i - g * G;   // exact for i < 2^14 (no integer division)
                adv[base + i] = sm[g * Gp + k];
            }
        }
        __syncthreads();
    }
}

// Fixed group size <= 128, warp-autonomous form: a warp owns four consecutive groups per pass (eight lanes each), stages
// them through its PRIVATE slice of shared memory with fully coalesced 8-byte loads / stores and synchronises with
// __syncwarp only - no block-wide barrier separates the load / compute / store phases of different warps - while the
// next pass's elements are already in flight in registers (software prefetch).
constexpr int ADVW_MAXJ = 16;                 // elements per lane per pass: ceil(4 * G / 32) <= 16  <=>  G <= 128

// NJ = ceil(4 G / 32) is a template parameter: the staging loops unroll to exactly the passes this group size needs (ncu of
// the run-time form at G = 72: 89 issued instructions per element, 116 registers -> 16 warps / SM; profiles/r2_advantage_ncu.md)
template <int NJ>
__global__ void __launch_bounds__(ADV_THREADS, 3)
group_advantage_warp_kernel(const double* __restrict__ ret, double* __restrict__ adv, long long n_groups, int G, int Gp) {
    pdl_grid_sync();
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int lane8 = lane & 7, gl = lane >> 3;
    const unsigned gmask = 0xFFu << (lane & ~7);
    double* ws = sm + (size_t)wib * 4 * Gp;                        // this warp's four groups
    const float invG = 1.f / (float)G;
    const int per = 4 * G;                                         // elements per pass
    const long long n_quads = (n_groups + 3) >> 2;
    const long long wstride = (long long)gridDim.x * (ADV_THREADS / 32);
    long long quad = (long long)blockIdx.x * (ADV_THREADS / 32) + wib;
    const long long total = n_groups * G;
    double v[NJ];
    // element e of a pass -> (group e / G, position e % G) in the staged layout (no integer division: e < 2^10)
    int slot[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int e = lane + 32 * j;
        const int g = (int)(((float)e + 0.5f) * invG);
        slot[j] = e < per ? g * Gp + (e - g * G) : -1;
    }
    if (quad < n_quads) {
        const long long base = quad * per;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const long long i = base + lane + 32 * j;
            v[j] = (slot[j] >= 0 && i < total) ? __ldg(ret + i) : 0.0;
        }
    }
    for (; quad < n_quads; quad += wstride) {
        const long long base = quad * per;
        const int ng = (int)min(4LL, n_groups - quad * 4);
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (slot[j] >= 0) ws[slot[j]] = v[j];
        __syncwarp();
        const long long nq = quad + wstride;                       // prefetch the next pass while this one computes
        if (nq < n_quads) {
            const long long nb = nq * per;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const long long i = nb + lane + 32 * j;
                v[j] = (slot[j] >= 0 && i < total) ? __ldg(ret + i) : 0.0;
            }
        }
        if (gl < ng) {
            double* x = ws + gl * Gp;
            advantage_group8(x, x, G, lane8, gmask);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const long long i = base + lane + 32 * j;
            if (slot[j] >= 0 && i < total) adv[i] = ws[slot[j]];
        }
        __syncwarp();
    }
}

// Fixed group size 8 <= G <= 128, REGISTER form: the eight lanes of a group read the group straight from global memory in
// numpy's own order - lane k owns elements k, k + 8, k + 16, ... = exactly the terms of numpy's k-th accumulator - so the
// whole group lives in NE = ceil(G / 8) registers per lane from the load to the store: no shared-memory staging, no
// index arithmetic, no __syncwarp, NE independent 8-byte loads in flight per lane (every load instruction of a warp
// covers four 64-byte segments).  The pairwise order is the one of pw_leaf8: per-lane accumulation over the full blocks
// of eight, xor-butterfly ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n % 8 tail elements added one by one.
// (ncu of the staged form at G = 72: 89 issued instructions per element, issue-bound at 0.50 of the copy bandwidth.)
template <int NE, bool SQ>
__device__ __forceinline__ double pw_regs(const double (&x)[NE], int n, int lane8, unsigned gmask) {
    const int nb = n >> 3;                                   // full blocks of eight
    double r = SQ ? __dmul_rn(x[0], x[0]) : x[0];
#pragma unroll
    for (int j = 1; j < NE; ++j)
        if (j < nb) r = __dadd_rn(r, SQ ? __dmul_rn(x[j], x[j]) : x[j]);
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    // tail: elements 8 * nb + t live in lane t's register nb (NE - 1 when n is not a multiple of 8)
    const int tail = n & 7;
    if (tail) {
        const double mine = SQ ? __dmul_rn(x[NE - 1], x[NE - 1]) : x[NE - 1];
        const int base = (threadIdx.x & 31) & ~7;
        for (int t = 0; t < tail; ++t) r = __dadd_rn(r, __shfl_sync(gmask, mine, base + t));
    }
    return r;
}

template <int NE>
__global__ void __launch_bounds__(256)
group_advantage_reg_kernel(const double* __restrict__ ret, double* __restrict__ adv, long long n_groups, int G) {
    pdl_grid_sync();
    const int lane8 = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const double dn = (double)G;
    const bool nok = div_shared_ok(dn);
    const double yn = __drcp_rn(dn);
    const long long stride = (long long)gridDim.x * 32;      // groups per grid pass (256 threads = 32 groups per block)
    // warp-uniform loop over the first group of the warp's four: the shuffles below need all lanes of a live warp
    for (long long gw = (long long)blockIdx.x * 32 + (threadIdx.x >> 5) * 4; gw < n_groups; gw += stride) {
        const long long g0 = gw + ((threadIdx.x & 31) >> 3);
        const bool live = g0 < n_groups;
        const double* src = ret + (live ? g0 : 0) * G;
        double x[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) x[j] = (live && lane8 + 8 * j < G) ? __ldg(src + lane8 + 8 * j) : 0.0;
        const double mean = div_shared(pw_regs<NE, false>(x, G, lane8, gmask), dn, yn, nok);
#pragma unroll
        for (int j = 0; j < NE; ++j) x[j] = __dadd_rn(x[j], -mean);
        const double var = div_shared(pw_regs<NE, true>(x, G, lane8, gmask), dn, yn, nok);
        const double sd = __dadd_rn(__dsqrt_rn(var), 1e-5);
        const bool sok = div_shared_ok(sd);
        const double ys = sok ? __drcp_rn(sd) : 0.0;
        if (live) {
            double* dst = adv + g0 * G;
#pragma unroll
            for (int j = 0; j < NE; ++j)
                if (lane8 + 8 * j < G) dst[lane8 + 8 * j] = div_shared(x[j], sd, ys, sok);
        }
    }
}

// Fixed group size <= 128, ONE LANE PER GROUP: the per-group scalar work (two divisions by n, the square root, the
// reciprocal - about half of all fp64 instructions when eight lanes share a group, because a warp instruction then
// serves only four groups) is amortised over 32 groups per warp instruction, and numpy's eight accumulators are simply
// eight registers of the lane (no shuffles).  A warp stages its 32 groups through a private shared-memory slice with
// coalesced 8-byte loads / stores (row pitch Gp odd: lane-per-row accesses are conflict free); only __syncwarp is used.
template <bool SQ>
__device__ __forceinline__ double pw_leaf_lane(const double* x, int n) {        // numpy pairwise leaf, n <= 128
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, pw_term<SQ>(x, i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = pw_term<SQ>(x, k);
    const int nfull = n - (n & 7);
    for (int i = 8; i < nfull; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], pw_term<SQ>(x, i + k));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (int i = nfull; i < n; ++i) res = __dadd_rn(res, pw_term<SQ>(x, i));
    return res;
}

constexpr int ADVL_THREADS = 128;

// GT = group size at compile time: the staging loops unroll completely, so the GT coalesced 8-byte loads of a pass are all in
// flight before the first one is used (the run-time form issued them one by one, each waiting for its own HBM round trip:
// 0.34 of the copy bandwidth at G = 12), and the element -> (group, position) map costs a multiply instead of a loop.
template <int GT>
__global__ void __launch_bounds__(ADVL_THREADS)
group_advantage_lane_kernel(const double* __restrict__ ret, double* __restrict__ adv, long long n_groups) {
    pdl_grid_sync();
    constexpr int G = GT, Gp = GT | 1;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* ws = sm + (size_t)wib * 32 * Gp;                       // this warp's 32 groups
    const long long n_pass = (n_groups + 31) >> 5;
    const long long wstride = (long long)gridDim.x * (ADVL_THREADS / 32);
    const double dn = (double)G;
    const bool nok = div_shared_ok(dn);
    const double yn = __drcp_rn(dn);
    for (long long pass = (long long)blockIdx.x * (ADVL_THREADS / 32) + wib; pass < n_pass; pass += wstride) {
        const long long g0 = pass << 5;
        const int ng = (int)min(32LL, n_groups - g0);
        const long long base = g0 * G;
        const int cnt = ng * G;
        double v[GT];
#pragma unroll
        for (int j = 0; j < GT; ++j) {
            const int e = lane + 32 * j;
            v[j] = e < cnt ? __ldg(ret + base + e) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < GT; ++j) {
            const int e = lane + 32 * j;
            ws[(e / G) * Gp + (e % G)] = v[j];
        }
        __syncwarp();
        {
            double* x = ws + lane * Gp;
            const double mean = div_shared(pw_leaf_lane<false>(x, G), dn, yn, nok);
#pragma unroll
            for (int i = 0; i < G; ++i) x[i] = __dadd_rn(x[i], -mean);
            const double var = div_shared(pw_leaf_lane<true>(x, G), dn, yn, nok);
            const double sd = __dadd_rn(__dsqrt_rn(var), 1e-5);
            const bool sok = div_shared_ok(sd);
            const double ys = sok ? __drcp_rn(sd) : 0.0;
#pragma unroll
            for (int i = 0; i < G; ++i) x[i] = div_shared(x[i], sd, ys, sok);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < GT; ++j) {
            const int e = lane + 32 * j;
            if (e < cnt) adv[base + e] = ws[(e / G) * Gp + (e % G)];
        }
        __syncwarp();
    }
}

// Ragged groups (offsets[n_groups+1]); eight lanes per group straight from global memory.
__global__ void __launch_bounds__(ADV_THREADS)
group_advantage_ragged_kernel(const double* __restrict__ ret, const long long* __restrict__ offsets,
                              double* __restrict__ adv, long long n_groups, int G) {
    pdl_grid_sync();
    const int tid = threadIdx.x;
    const int lane8 = tid & 7;
    const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
    const long long g = (long long)blockIdx.x * ADV_GPB + (tid >> 3);
    if (g >= n_groups) return;
    const long long o0 = offsets ? offsets[g] : g * G, o1 = offsets ? offsets[g + 1] : (g + 1) * G;
    const int n = (int)(o1 - o0);
    if (n <= 0) return;
    // x - mean is written to adv first, so the variance pass re-reads our own output (L1/L2 hit)
    advantage_group8(const_cast<double*>(ret) + o0, adv + o0, n, lane8, gmask);
}

// advantage_group8 writes `out` before it finished reading `x` only when out == x, which is the
// staged case where the in-place update is element-wise safe; the ragged kernel has out != x.

int launch_group_advantage(const double* ret, const long long* offsets, long long n_groups, int G,
                           double* adv, cudaStream_t st) {
    if (n_groups <= 0) return 0;
    if (offsets == nullptr) {
        RIFT_REQUIRE(G > 0, "group size must be positive");
        const int Gp = G | 1;                                   // odd stride: conflict-free 64-bit lanes
        const size_t smem = (size_t)ADV_GPB * Gp * sizeof(double);
        static const bool block_form = getenv("RIFT_B200_ADV_BLOCK") != nullptr;
        static const bool warp_form = getenv("RIFT_B200_ADV_WARP") != nullptr;
        if (G <= 32 && n_groups >= 32768 && !block_form && !warp_form) {
            // one lane per group: 4 warps x 32 groups staged per CTA.  Measured at 2^18 groups: G = 12: 31 vs 74 us,
            // G = 24: 51 vs 76 us for the eight-lanes-per-group form below; from G = 36 on (94 vs 89 us) the staged
            // rows cost too much occupancy and the eight-lane form wins (G = 72: 351 vs 105 us)
            const size_t smem_l = (size_t)(ADVL_THREADS / 32) * 32 * Gp * sizeof(double);
            static bool attr_l = false;
            (void)attr_l;       // 4 warps x 32 groups x 33 doubles = 33 KB at most: inside the default dynamic shared-memory limit
            const long long n_pass = (n_groups + 31) / 32;
            const int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(220 * 1024) / (smem_l + 1024)));
            const int grid = (int)min((n_pass + 3) / 4, (long long)148 * per_sm);
#define RIFT_ADVL(N) case N: launch_k(group_advantage_lane_kernel<N>, grid, ADVL_THREADS, smem_l, st, ret, adv, n_groups); break;
            switch (G) {
                RIFT_ADVL(1) RIFT_ADVL(2) RIFT_ADVL(3) RIFT_ADVL(4) RIFT_ADVL(5) RIFT_ADVL(6) RIFT_ADVL(7) RIFT_ADVL(8)
                RIFT_ADVL(9) RIFT_ADVL(10) RIFT_ADVL(11) RIFT_ADVL(12) RIFT_ADVL(13) RIFT_ADVL(14) RIFT_ADVL(15) RIFT_ADVL(16)
                RIFT_ADVL(17) RIFT_ADVL(18) RIFT_ADVL(19) RIFT_ADVL(20) RIFT_ADVL(21) RIFT_ADVL(22) RIFT_ADVL(23) RIFT_ADVL(24)
                RIFT_ADVL(25) RIFT_ADVL(26) RIFT_ADVL(27) RIFT_ADVL(28) RIFT_ADVL(29) RIFT_ADVL(30) RIFT_ADVL(31) RIFT_ADVL(32)
                default: set_last_error("internal: advantage lane form"); return -1;
            }
#undef RIFT_ADVL
            RIFT_LAUNCH_OK();
            return 0;
        }
        static const bool staged_form = getenv("RIFT_B200_ADV_STAGED") != nullptr;
        if (G >= 8 && G <= 128 && !block_form && !warp_form && !staged_form) {
            const int ne = (G + 7) / 8;
            const int grid = (int)min((n_groups + 31) / 32, (long long)148 * 8);
#define RIFT_ADVR(N) case N: launch_k(group_advantage_reg_kernel<N>, grid, 256, 0, st, ret, adv, n_groups, G); break;
            switch (ne) {
                RIFT_ADVR(1) RIFT_ADVR(2) RIFT_ADVR(3) RIFT_ADVR(4) RIFT_ADVR(5) RIFT_ADVR(6) RIFT_ADVR(7) RIFT_ADVR(8)
                RIFT_ADVR(9) RIFT_ADVR(10) RIFT_ADVR(11) RIFT_ADVR(12) RIFT_ADVR(13) RIFT_ADVR(14) RIFT_ADVR(15) RIFT_ADVR(16)
                default: set_last_error("internal: advantage register count"); return -1;
            }
#undef RIFT_ADVR
            RIFT_LAUNCH_OK();
            return 0;
        }
        if (G <= 128 && !block_form) {
            const long long n_quads = (n_groups + 3) / 4;
            const int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(200 * 1024) / (smem + 1024)));
            const int grid = (int)min((n_quads + 7) / 8, (long long)148 * per_sm);
            const int nj = (4 * G + 31) / 32;
#define RIFT_ADVW(N) case N: launch_k(group_advantage_warp_kernel<N>, grid, ADV_THREADS, smem, st, ret, adv, n_groups, G, Gp); break;
            switch (nj) {
                RIFT_ADVW(1) RIFT_ADVW(2) RIFT_ADVW(3) RIFT_ADVW(4) RIFT_ADVW(5) RIFT_ADVW(6) RIFT_ADVW(7) RIFT_ADVW(8)
                RIFT_ADVW(9) RIFT_ADVW(10) RIFT_ADVW(11) RIFT_ADVW(12) RIFT_ADVW(13) RIFT_ADVW(14) RIFT_ADVW(15) RIFT_ADVW(16)
                default: set_last_error("internal: advantage pass count"); return -1;
            }
#undef RIFT_ADVW
            RIFT_LAUNCH_OK();
            return 0;
        }
        if (smem <= 200 * 1024) {
            static bool attr_done = false;
            if (!attr_done) {
                RIFT_CUDA_OK(cudaFuncSetAttribute(group_advantage_fixed_kernel,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_done = true;
            }
            const long long n_chunks = (n_groups + ADV_GPB - 1) / ADV_GPB;
            const int per_sm = (int)max((size_t)1, min((size_t)8, (size_t)(220 * 1024) / (smem + 1024)));
            const int grid = (int)min(n_chunks, (long long)148 * per_sm);
            launch_k(group_advantage_fixed_kernel, grid, ADV_THREADS, smem, st, ret, adv, n_groups, G, Gp);
            RIFT_LAUNCH_OK();
            return 0;
        }
        // very large groups: straight from global memory (second pass re-reads through L1/L2)
    }
    const int grid = (int)((n_groups + ADV_GPB - 1) / ADV_GPB);
    launch_k(group_advantage_ragged_kernel, grid, ADV_THREADS, 0, st, ret, offsets, adv, n_groups, G);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// RIFT / GRPO group objective: forward value + d(loss)/d(logits) in one pass over the batch
// =====================================================================================
// One warp per sample (G = R*Mo candidates, typically 12..144).  Outputs per sample:
//   part_sum[b]  (double)  sum of the objective over valid candidates
//   part_cnt[b]  (int)     number of valid candidates
//   dlogits[b,:] (float)   d(-sum_valid obj)/d logits  -- NOT yet divided by the global valid count
// The division by the (all-rank) valid count happens in finalize / in the optimizer kernel, which
// is what makes the data-parallel loss an exact global masked mean (rift_trainer.py:173-178).

constexpr int LOSS_MAX_PER_LANE = 16;   // G <= 512

template <bool GRPO>
__global__ void __launch_bounds__(128)
group_objective_kernel(const float* __restrict__ logits, const float* __restrict__ old_logits,
                       const float* __restrict__ ref_logits, const double* __restrict__ adv,
                       const uint8_t* __restrict__ valid, const uint8_t* __restrict__ r_pad,
                       int bs, int R, int Mo, float clip_lo, float clip_hi, float dual_clip, float kl_w,
                       double* __restrict__ part_sum, int* __restrict__ part_cnt,
                       float* __restrict__ dlogits) {
    pdl_grid_sync();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= bs) return;
    const int G = R * Mo;
    const long long base = (long long)warp * G;
    float z[LOSS_MAX_PER_LANE], o[LOSS_MAX_PER_LANE], f[LOSS_MAX_PER_LANE];
    bool pad[LOSS_MAX_PER_LANE];
    float zmax = -INFINITY, omax = -INFINITY, fmax_ = -INFINITY;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int i = lane + 32 * t;
        z[t] = o[t] = f[t] = -INFINITY;
        pad[t] = true;
        if (i < G) {
            const int r = i / Mo;
            bool p;
            if (r_pad) {
                p = r_pad[(long long)warp * R + r] != 0;
            } else {                                   // r_padding = ~valid.any(-1)
                p = true;
                for (int m = 0; m < Mo; ++m) p = p && (valid[base + r * Mo + m] == 0);
            }
            pad[t] = p;
            z[t] = p ? -1e8f : logits[base + i];
            o[t] = p ? -1e8f : old_logits[base + i];
            if (GRPO) f[t] = p ? -1e8f : ref_logits[base + i];
            zmax = fmaxf(zmax, z[t]);
            omax = fmaxf(omax, o[t]);
            if (GRPO) fmax_ = fmaxf(fmax_, f[t]);
        }
    }
    zmax = warp_max(zmax); omax = warp_max(omax);
    if (GRPO) fmax_ = warp_max(fmax_);
    float zs = 0.f, os = 0.f, fs = 0.f;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        if (lane + 32 * t < G) {
            zs += expf(z[t] - zmax);
            os += expf(o[t] - omax);
            if (GRPO) fs += expf(f[t] - fmax_);
        }
    }
    zs = warp_sum(zs); os = warp_sum(os);
    if (GRPO) fs = warp_sum(fs);
    const float zl = logf(zs), ol = logf(os);

    double lsum = 0.0;
    int lcnt = 0;
    float gsum = 0.f;                 // sum_i g_i (g = d obj / d lp)
    float g[LOSS_MAX_PER_LANE], p[LOSS_MAX_PER_LANE];
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int i = lane + 32 * t;
        g[t] = 0.f; p[t] = 0.f;
        if (i < G) {
            const float lp = (z[t] - zmax) - zl;
            const float lo = (o[t] - omax) - ol;
            p[t] = expf(lp);
            const bool v = valid[base + i] != 0;
            if (v) {
                const float ratio = expf(lp - lo);
                const double a = adv[base + i];
                const float rc = fminf(fmaxf(ratio, clip_lo), clip_hi);
                const double t1 = a * (double)ratio;
                const double t2 = a * (double)rc;
                const double mn = fmin(t1, t2);
                // torch sub-gradients: min/max split 0.5/0.5 on ties, clamp passes lo<=x<=hi
                const double w1 = t1 < t2 ? 1.0 : (t1 == t2 ? 0.5 : 0.0);
                const double w2 = (1.0 - w1) * ((ratio >= clip_lo && ratio <= clip_hi) ? 1.0 : 0.0);
                double obj = mn, wmn = 1.0;
                if (!GRPO) {
                    if (a < 0.0) {
                        const double dc = a * (double)dual_clip;
                        obj = fmax(mn, dc);
                        wmn = mn > dc ? 1.0 : (mn == dc ? 0.5 : 0.0);
                    }
                }
                double dobj_dlp = wmn * (w1 + w2) * a * (double)ratio;
                if (GRPO) {
                    const float pf = expf((f[t] - fmax_) - logf(fs));
                    // F.kl_div(input=lp, target=pf): xlogy(pf, pf) - pf * lp   (fp32)
                    const float kl = (pf > 0.f ? pf * logf(pf) : 0.f) - pf * lp;
                    obj = mn - (double)(kl_w * kl);
                    dobj_dlp += (double)(kl_w * pf);
                }
                lsum += obj;
                lcnt += 1;
                g[t] = (float)dobj_dlp;
                gsum += g[t];
            }
        }
    }
    lsum = warp_sum_d(lsum);
    gsum = warp_sum(gsum);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) lcnt += __shfl_xor_sync(0xffffffffu, lcnt, o2);
    if (lane == 0) { part_sum[warp] = lsum; part_cnt[warp] = lcnt; }
    if (dlogits) {
#pragma unroll
        for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
            const int i = lane + 32 * t;
            // loss = -sum obj / N  ->  d/dz_j = -(g_j - softmax_j * sum_i g_i) / N ; padded rows get 0
            if (i < G) dlogits[base + i] = pad[t] ? 0.f : -(g[t] - p[t] * gsum);
        }
    }
}

// Deterministic finalize: fixed-order tree over the per-sample partials, one block.
// out[0] = loss (= -sum/count, 0 when count == 0), out[1] = sum of objectives, out[2] = count.
// If `scale_dlogits` the gradient is divided by the local count (single-GPU drop-in use).
__global__ void __launch_bounds__(256)
objective_finalize_kernel(const double* __restrict__ part_sum, const int* __restrict__ part_cnt, int bs,
                          double* __restrict__ out, float* __restrict__ dlogits, long long n_dlogits,
                          int scale_dlogits) {
    pdl_grid_sync();
    __shared__ double ssum[256];
    __shared__ long long scnt[256];
    double s = 0.0;
    long long c = 0;
    for (int i = threadIdx.x; i < bs; i += 256) { s += part_sum[i]; c += part_cnt[i]; }
    ssum[threadIdx.x] = s; scnt[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
        __syncthreads();
    }
    const double tot = ssum[0];
    const long long cnt = scnt[0];
    if (threadIdx.x == 0) {
        out[0] = cnt > 0 ? -(tot / (double)cnt) : 0.0;
        out[1] = tot;
        out[2] = (double)cnt;
    }
    if (scale_dlogits && dlogits) {
        const float inv = cnt > 0 ? (float)(1.0 / (double)cnt) : 0.f;
        for (long long i = threadIdx.x; i < n_dlogits; i += 256) dlogits[i] *= inv;
    }
}

__global__ void scale_f32_by_count_kernel(float* __restrict__ x, long long n, const double* __restrict__ out3) {
    pdl_grid_sync();
    const double cnt = out3[2];
    const float inv = cnt > 0 ? (float)(1.0 / cnt) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] *= inv;
}

int launch_group_objective(int algo, const float* logits, const float* old_logits, const float* ref_logits,
                           const double* adv, const uint8_t* valid, const uint8_t* r_pad, int bs, int R, int Mo,
                           float clip_lo, float clip_hi, float dual_clip, float kl_w,
                           double* part_sum, int* part_cnt, double* out3, float* dlogits, int scale_dlogits,
                           cudaStream_t st) {
    RIFT_REQUIRE(R * Mo <= 32 * LOSS_MAX_PER_LANE, "R*num_modes exceeds 512 candidates per sample");
    RIFT_REQUIRE(algo == 0 || ref_logits != nullptr, "GRPO needs ref_logits");
    if (bs <= 0) return 0;
    const int grid = cdiv((long long)bs * 32, 128);
    if (algo == 0)
        launch_k(group_objective_kernel<false>, grid, 128, 0, st, logits, old_logits, ref_logits, adv, valid, r_pad, bs, R, Mo,
                                                            clip_lo, clip_hi, dual_clip, kl_w, part_sum, part_cnt, dlogits);
    else
        launch_k(group_objective_kernel<true>, grid, 128, 0, st, logits, old_logits, ref_logits, adv, valid, r_pad, bs, R, Mo,
                                                           clip_lo, clip_hi, dual_clip, kl_w, part_sum, part_cnt, dlogits);
    RIFT_LAUNCH_OK();
    const long long n = (long long)bs * R * Mo;
    // small batches: the single finalize block also rescales; large ones use a wide kernel
    const bool wide = scale_dlogits && dlogits && n > 65536;
    launch_k(objective_finalize_kernel, 1, 256, 0, st, part_sum, part_cnt, bs, out3, dlogits, n, scale_dlogits && !wide);
    RIFT_LAUNCH_OK();
    if (wide) {
        launch_k(scale_f32_by_count_kernel, min(cdiv(n, 256), 148 * 8), 256, 0, st, dlogits, n, out3);
        RIFT_LAUNCH_OK();
    }
    return 0;
}

// =====================================================================================
// PPO (chosen action + entropy) and REINFORCE (argmax action) objectives
// =====================================================================================
// mode 0: PPO   per-sample term  -(min(A rho, A clamp(rho)) + lambda * H)          (ppo_trainer.py:161-183)
// mode 1: REINFORCE per-sample   -(lp[argmax] * return)                            (reinforce_trainer.py:154-170)
// part[b] = per-sample term (float, the reference computes these in fp32); the mean over the
// (global) batch and the matching gradient scale `inv_n` are applied here directly because the
// batch size is known on the host.
__global__ void __launch_bounds__(128)
action_objective_kernel(int mode, const float* __restrict__ logits, const uint8_t* __restrict__ r_pad,
                        const long long* __restrict__ action_mode, const float* __restrict__ weight,   // advantage | return
                        const float* __restrict__ old_log_prob, int bs, int R, int Mo, float eps_clip,
                        float lambda_entropy, float inv_n, float* __restrict__ part, float* __restrict__ dlogits,
                        int* __restrict__ chosen) {
    pdl_grid_sync();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= bs) return;
    const int G = R * Mo;
    const long long base = (long long)warp * G;
    float z[LOSS_MAX_PER_LANE];
    bool pad[LOSS_MAX_PER_LANE];
    float zmax = -INFINITY;
    int amax = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int i = lane + 32 * t;
        z[t] = -INFINITY; pad[t] = true;
        if (i < G) {
            pad[t] = r_pad[(long long)warp * R + i / Mo] != 0;
            z[t] = pad[t] ? -1e8f : logits[base + i];
            if (z[t] > zmax) { zmax = z[t]; amax = i; }     // first maximum within the lane
        }
    }
    // warp arg-max with first-index tie-break (torch.argmax returns the first maximal index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, zmax, o);
        const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
        if (om > zmax || (om == zmax && oi < amax)) { zmax = om; amax = oi; }
    }
    float zs = 0.f;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t)
        if (lane + 32 * t < G) zs += expf(z[t] - zmax);
    zs = warp_sum(zs);
    const float zl = logf(zs);
    int idx;
    if (mode == 0) idx = (int)(action_mode[2 * warp] * Mo + action_mode[2 * warp + 1]);
    else idx = amax;
    float ent = 0.f, lp_idx = 0.f;
    float lp[LOSS_MAX_PER_LANE];
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int i = lane + 32 * t;
        lp[t] = 0.f;
        if (i < G) {
            lp[t] = (z[t] - zmax) - zl;
            ent -= expf(lp[t]) * lp[t];
            if (i == idx) lp_idx = lp[t];
        }
    }
    ent = warp_sum(ent);
    lp_idx = warp_sum(lp_idx);
    const float w = weight[warp];
    float term, coef;             // term = per-sample loss contribution ; coef = d term / d lp[idx]
    if (mode == 0) {
        const float ratio = expf(lp_idx - old_log_prob[warp]);
        const float lo = 1.f - eps_clip, hi = 1.f + eps_clip;
        const float rc = fminf(fmaxf(ratio, lo), hi);
        const float t1 = w * ratio, t2 = w * rc;
        const float w1 = t1 < t2 ? 1.f : (t1 == t2 ? 0.5f : 0.f);
        const float w2 = (1.f - w1) * ((ratio >= lo && ratio <= hi) ? 1.f : 0.f);
        term = -(fminf(t1, t2) + lambda_entropy * ent);
        coef = -(w1 + w2) * w * ratio;
    } else {
        term = -(lp_idx * w);
        coef = -w;
    }
    if (lane == 0) { part[warp] = term; if (chosen) chosen[warp] = idx; }
    if (dlogits) {
        // g_i = d term / d lp_i ; dz_j = g_j - softmax_j * sum_i g_i
        float g[LOSS_MAX_PER_LANE], gsum = 0.f;
#pragma unroll
        for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
            const int i = lane + 32 * t;
            g[t] = 0.f;
            if (i < G) {
                if (i == idx) g[t] += coef;
                if (mode == 0) g[t] += lambda_entropy * expf(lp[t]) * (lp[t] + 1.f);   // d(-lambda H)/d lp
                gsum += g[t];
            }
        }
        gsum = warp_sum(gsum);
#pragma unroll
        for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
            const int i = lane + 32 * t;
            if (i < G) dlogits[base + i] = pad[t] ? 0.f : (g[t] - expf(lp[t]) * gsum) * inv_n;
        }
    }
}

// out[0] = inv_n * sum_b part[b] (+ extra[0] if given), sequential fixed order -> deterministic
__global__ void __launch_bounds__(256)
mean_finalize_kernel(const float* __restrict__ part, int bs, float inv_n, const float* __restrict__ extra,
                     float* __restrict__ out) {
    pdl_grid_sync();
    __shared__ float s[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < bs; i += 256) a += part[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0] * inv_n + (extra ? extra[0] : 0.f);
}

int launch_action_objective(int mode, const float* logits, const uint8_t* r_pad, const long long* action_mode,
                            const float* weight, const float* old_log_prob, int bs, int R, int Mo, float eps_clip,
                            float lambda_entropy, float inv_n, const float* extra_loss, float* part, float* loss_out,
                            float* dlogits, int* chosen, cudaStream_t st) {
    RIFT_REQUIRE(R * Mo <= 32 * LOSS_MAX_PER_LANE, "R*num_modes exceeds 512 candidates per sample");
    RIFT_REQUIRE(r_pad != nullptr, "r_pad is required");
    if (bs <= 0) return 0;
    launch_k(action_objective_kernel, cdiv((long long)bs * 32, 128), 128, 0, st, 
        mode, logits, r_pad, action_mode, weight, old_log_prob, bs, R, Mo, eps_clip, lambda_entropy, inv_n, part,
        dlogits, chosen);
    RIFT_LAUNCH_OK();
    launch_k(mean_finalize_kernel, 1, 256, 0, st, part, bs, inv_n, extra_loss, loss_out);
    RIFT_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// SFT / RTR teacher objective (fine_tuner/sft/sft_trainer.py:123-215): cross-entropy of the candidate logits against a
// one-hot label at (best reference line of the model, mode whose PID target speed is closest to the teacher's).
//   target speed of a candidate (pid_controller.py:121-136 on sft/utils.py:10-33): the trajectory is resampled every
//   `step` points, moved into the teacher's frame (origin, heading) and the mean segment length is taken (one point only:
//   its distance from the origin);  label mode = argmin over ALL (R, Mo) candidates of |target speed - teacher speed| mod Mo;
//   label line = argmax of the masked logits / Mo.  Only the logits carry gradient (the label is detached):
//   d loss / d z_j = weight * inv_n * (softmax(z')_j - [j == label]),  loss term = -log_softmax(z')[label].
// One warp per sample; dlogits is assigned (accumulate == 0) or added to (RTR: 5 * PPO term + teacher term).
__global__ void __launch_bounds__(128)
teacher_objective_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ r_pad, const float* __restrict__ traj,
                         const float* __restrict__ teacher, int bs, int R, int Mo, int T, int step, float scale, float* __restrict__ part,
                         float* __restrict__ dlogits, int accumulate, int* __restrict__ label_out) {
    pdl_grid_sync();
    const int b = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= bs) return;
    const int G = R * Mo;
    const float* zb = logits + (long long)b * G;
    const float t_speed = teacher[b * 5 + 0], ox = teacher[b * 5 + 1], oy = teacher[b * 5 + 2], hd = teacher[b * 5 + 3];
    const float ch = cosf(hd), sh = sinf(hd);
    const int t0 = T < step ? T - 1 : step - 1, tstep = T < step ? T : step;
    const int n_pts = T < step ? 1 : (T - (step - 1) + step - 1) / step;
    float z[LOSS_MAX_PER_LANE];
    float zmax = -INFINITY, best_d = INFINITY;
    int amax = 0x7fffffff, amin = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int g = lane + 32 * t;
        z[t] = -INFINITY;
        if (g < G) {
            const bool pad = r_pad[(long long)b * R + g / Mo] != 0;
            z[t] = pad ? -1e8f : zb[g];
            if (z[t] > zmax) { zmax = z[t]; amax = g; }
            // PID target speed of candidate g in the teacher's frame
            const float* tr = traj + ((long long)b * G + g) * T * 6;
            float px = 0.f, py = 0.f, acc = 0.f;
            for (int i = 0; i < n_pts; ++i) {
                const float vx = tr[(long long)(t0 + i * tstep) * 6] - ox, vy = tr[(long long)(t0 + i * tstep) * 6 + 1] - oy;
                const float lx = vx * ch + vy * sh, ly = vx * (-sh) + vy * ch;
                if (i > 0) { const float dx = lx - px, dy = ly - py; acc += sqrtf(dx * dx + dy * dy); }
                px = lx; py = ly;
            }
            const float speed = n_pts > 1 ? acc / (float)(n_pts - 1) : sqrtf(px * px + py * py);
            const float d = fabsf(speed - t_speed);
            if (d < best_d) { best_d = d; amin = g; }
        }
    }
    // warp arg-max (first maximal index) and arg-min (first minimal index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, zmax, o); const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
        if (ov > zmax || (ov == zmax && oi < amax)) { zmax = ov; amax = oi; }
        const float od = __shfl_xor_sync(0xffffffffu, best_d, o); const int odi = __shfl_xor_sync(0xffffffffu, amin, o);
        if (od < best_d || (od == best_d && odi < amin)) { best_d = od; amin = odi; }
    }
    const int label = (amax / Mo) * Mo + (amin % Mo);
    float se = 0.f;
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) if (lane + 32 * t < G) se += __expf(z[t] - zmax);
    se = warp_sum(se);
    const float lse = zmax + logf(se);
#pragma unroll
    for (int t = 0; t < LOSS_MAX_PER_LANE; ++t) {
        const int g = lane + 32 * t;
        if (g < G) {
            if (g == label) part[b] = -(z[t] - lse);
            if (dlogits) {
                const float gr = scale * (__expf(z[t] - lse) - (g == label ? 1.f : 0.f));
                float* o = dlogits + (long long)b * G + g;
                *o = accumulate ? *o + gr : gr;
            }
        }
    }
    if (lane == 0 && label_out) label_out[b] = label;
}

// loss_out[0] (=|+=) weight * inv_n * sum_b part[b]
__global__ void __launch_bounds__(256)
weighted_mean_finalize_kernel(const float* __restrict__ part, int bs, float scale, int accumulate, float* __restrict__ out) {
    pdl_grid_sync();
    __shared__ float s[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < bs; i += 256) a += part[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + s[0] * scale;
}

int launch_teacher_objective(const float* logits, const uint8_t* r_pad, const float* traj, const float* teacher, int bs, int R,
                             int Mo, int T, int step, float inv_n, float weight, float* part, float* loss_out, float* dlogits,
                             int accumulate, int* label_out, cudaStream_t st) {
    RIFT_REQUIRE(R * Mo <= 32 * LOSS_MAX_PER_LANE, "R*num_modes exceeds 512 candidates per sample");
    RIFT_REQUIRE(step >= 1 && T >= 1, "teacher_objective: bad frame rate / horizon");
    if (bs <= 0) return 0;
    launch_k(teacher_objective_kernel, cdiv((long long)bs * 32, 128), 128, 0, st, logits, r_pad, traj, teacher, bs, R, Mo, T, step,
             weight * inv_n, part, dlogits, accumulate, label_out);
    RIFT_LAUNCH_OK();
    launch_k(weighted_mean_finalize_kernel, 1, 256, 0, st, part, bs, weight * inv_n, accumulate, loss_out);
    RIFT_LAUNCH_OK();
    return 0;
}

// SmoothL1(value, target) mean (beta = 1) and its gradient wrt value, one block (bs is a batch size).
__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float* __restrict__ value, const float* __restrict__ target, int n, float inv_n,
                 float* __restrict__ loss_out, float* __restrict__ dvalue) {
    pdl_grid_sync();
    __shared__ float s[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float d = value[i] - target[i];
        const float ad = fabsf(d);
        a += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
        if (dvalue) dvalue[i] = (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * inv_n;
    }
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[0] = s[0] * inv_n;
}

int launch_smooth_l1(const float* value, const float* target, int n, float inv_n, float* loss_out, float* dvalue,
                     cudaStream_t st) {
    launch_k(smooth_l1_kernel, 1, 256, 0, st, value, target, n, inv_n, loss_out, dvalue);
    RIFT_LAUNCH_OK();
    return 0;
}

// =====================================================================================
// PPO buffer pass: GAE reverse scan + normalisation ; REINFORCE discounted return
// =====================================================================================
// The scans are first-order linear recurrences over the (<= 4096-deep) buffer; they are evaluated
// in the reference's sequential fp32 order by one thread (4096 dependent FMAs-without-fusion is
// ~20 us) so results are bit-identical to ppo_datamodule.py:22-37 / reinforce_datamodule.py:19-38.
__global__ void gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ undones,
                                const float* __restrict__ values, const float* __restrict__ next_values,
                                const float* __restrict__ unterminated, int n, float gamma, float lam,
                                float* __restrict__ adv, float* __restrict__ reward_sum) {
    pdl_grid_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float a = 0.f;
    for (int t = n - 1; t >= 0; --t) {
        // delta = rewards[t] + unterminated[t] * gamma * next_values[t] - values[t]
        const float delta = __fsub_rn(__fadd_rn(rewards[t], __fmul_rn(__fmul_rn(unterminated[t], gamma), next_values[t])), values[t]);
        // advantage = delta + undones[t] * gamma * lambda * advantage
        a = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(undones[t], gamma), lam), a));
        adv[t] = a;
        if (reward_sum) reward_sum[t] = __fadd_rn(a, values[t]);
    }
}

// (x - mean) / (std_unbiased + 1e-5), fp32 like torch (ppo_datamodule.py:163); one block.
__global__ void __launch_bounds__(1024)
normalise_unbiased_kernel(const float* __restrict__ x, int n, float* __restrict__ y) {
    pdl_grid_sync();
    __shared__ double s[1024];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) a += x[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    const float mean = (float)(s[0] / n);
    __syncthreads();
    a = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) { const double d = (double)x[i] - (double)mean; a += d * d; }
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
    const float sd = (float)sqrt(s[0] / (n > 1 ? n - 1 : 1));
    for (int i = threadIdx.x; i < n; i += 1024) y[i] = (x[i] - mean) / (sd + 1e-5f);
}

__global__ void discounted_return_kernel(const float* __restrict__ rewards, const float* __restrict__ dones, int n,
                                         float gamma, float* __restrict__ out) {
    pdl_grid_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float g = 0.f;
    for (int t = n - 1; t >= 0; --t) {
        g = dones[t] == 1.f ? rewards[t] : __fadd_rn(rewards[t], __fmul_rn(gamma, g));
        out[t] = g;
    }
}

int launch_gae(const float* rewards, const float* undones, const float* values, const float* next_values,
               const float* unterminated, int n, float gamma, float lam, float* adv, float* reward_sum,
               float* adv_normalised, cudaStream_t st) {
    launch_k(gae_scan_kernel, 1, 32, 0, st, rewards, undones, values, next_values, unterminated, n, gamma, lam, adv, reward_sum);
    RIFT_LAUNCH_OK();
    if (adv_normalised) {
        launch_k(normalise_unbiased_kernel, 1, 1024, 0, st, adv, n, adv_normalised);
        RIFT_LAUNCH_OK();
    }
    return 0;
}

int launch_discounted_return(const float* rewards, const float* dones, int n, float gamma, float* out, cudaStream_t st) {
    launch_k(discounted_return_kernel, 1, 32, 0, st, rewards, dones, n, gamma, out);
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift
