// Per-step optimizer apply: global-norm clip (Lightning gradient_clip_val 0.5, L2) folded into a
// fused multi-tensor AdamW over the flat trainable arena.
//
// Reference: rift/cbv/planning/fine_tuner/rlft/config/lightning/custom_lightning.yaml:40-41 (clip),
// rift/cbv/planning/fine_tuner/rlft/rift_pluto/rift_trainer.py:333-351 (two AdamW groups: decay /
// no-decay, betas (0.9,0.999), eps 1e-8), torch.optim.AdamW update rule.
//
// Arena layout (decided on the host, see rift_b200/arena.py): [decay params | no-decay params | frozen ...]
// so one launch covers every trainable tensor and the weight-decay switch is a single index compare.
// HBM traffic: 16 B read (w, g, m, v) + 12 B written (w, m, v) per parameter = 28 B/param.
#include "common.cuh"
#include "ops.h"

namespace rift {

constexpr int NORM_THREADS = 256;

// stage 1: per-block sum of squares (double accumulate), fixed grid -> deterministic partials
__global__ void __launch_bounds__(NORM_THREADS)
sumsq_partial_kernel(const float* __restrict__ g, long long n, double* __restrict__ partial) {
    pdl_grid_sync();
    __shared__ double s[NORM_THREADS / 32];
    double a = 0.0;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = (long long)blockIdx.x * NORM_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * NORM_THREADS) {
        const float4 v = __ldg(g4 + i);
        a += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * NORM_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * NORM_THREADS)
        a += (double)g[i] * g[i];
    a = warp_sum_d(a);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NORM_THREADS / 32; ++w) t += s[w];
        partial[blockIdx.x] = t;
    }
}

// stage 2: one warp folds the partials in a fixed order.
// scal[0] = total L2 norm of the TRUE gradient (= grad_scale * raw), scal[1] = factor applied to raw
// gradients in the update = grad_scale * min(1, max_norm / (norm + 1e-6))   (torch clip_grad_norm_)
// grad_scale = 1 / count when `count` (device, double) is given (global masked-mean normalisation of
// the objective, see rl_kernels.cu), else 1.
__global__ void sumsq_finalize_kernel(const double* __restrict__ partial, int n_partial, const double* __restrict__ count,
                                      float max_norm, float* __restrict__ scal) {
    pdl_grid_sync();
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) t += partial[i];
    t = warp_sum_d(t);
    if (threadIdx.x == 0) {
        float gs = 1.f;
        if (count) gs = count[0] > 0.0 ? (float)(1.0 / count[0]) : 0.f;
        const float norm = (float)sqrt(t) * gs;
        float coef = 1.f;
        if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);
        scal[0] = norm;
        scal[1] = gs * coef;
    }
}

__global__ void __launch_bounds__(256)
adamw_apply_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   long long n, long long n_decay, const float* __restrict__ scal, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float bc1, float sqrt_bc2) {
    pdl_grid_sync();
    const float gscale = scal ? scal[1] : 1.f;
    const float step_size = lr / bc1;
    const float decay_keep = 1.f - lr * weight_decay;
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    auto upd = [&](float& pw, float gr, float& mm, float& vv, bool decay) {
        gr *= gscale;
        if (decay) pw *= decay_keep;
        mm = mm + (gr - mm) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
        vv = vv * beta2 + (1.f - beta2) * gr * gr;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
        const float denom = sqrtf(vv) / sqrt_bc2 + eps;
        pw = pw - step_size * (mm / denom);
    };
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 pw = p4[i], mm = m4[i], vv = v4[i];
        const float4 gr = __ldg(g4 + i);
        const long long e = i << 2;
        upd(pw.x, gr.x, mm.x, vv.x, e + 0 < n_decay);
        upd(pw.y, gr.y, mm.y, vv.y, e + 1 < n_decay);
        upd(pw.z, gr.z, mm.z, vv.z, e + 2 < n_decay);
        upd(pw.w, gr.w, mm.w, vv.w, e + 3 < n_decay);
        p4[i] = pw; m4[i] = mm; v4[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        float pw = p[i], mm = m[i], vv = v[i];
        upd(pw, g[i], mm, vv, i < n_decay);
        p[i] = pw; m[i] = mm; v[i] = vv;
    }
}

int optim_scratch_doubles() { return 148 * 4; }

int launch_clip_adamw(float* p, const float* g, float* m, float* v, long long n, long long n_decay,
                      const double* count, float max_norm, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, double* scratch, float* scal, cudaStream_t st) {
    RIFT_REQUIRE(step >= 1, "AdamW step counter starts at 1");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                 "optimizer arenas must be 16-byte aligned");
    if (n <= 0) return 0;
    const int nb = (int)min((long long)optim_scratch_doubles(), (n / 4 + NORM_THREADS - 1) / NORM_THREADS + 1);
    launch_k(sumsq_partial_kernel, nb, NORM_THREADS, 0, st, g, n, scratch);
    RIFT_LAUNCH_OK();
    launch_k(sumsq_finalize_kernel, 1, 32, 0, st, scratch, nb, count, max_norm, scal);
    RIFT_LAUNCH_OK();
    // torch computes the bias corrections in Python doubles
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const int grid = (int)min((long long)148 * 8, (n / 4 + 255) / 256 + 1);
    launch_k(adamw_apply_kernel, grid, 256, 0, st, p, g, m, v, n, n_decay, scal, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                             (float)sqrt(bc2));
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift
