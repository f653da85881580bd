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

// stage 2: one warp folds the partials in a fixed order and prepares every scalar of the update.
// scal[0] = total L2 norm of the TRUE gradient (= grad_scale * raw), scal[1] = factor applied to raw
// gradients in the update = grad_scale * min(1, max_norm / (norm + 1e-6))   (torch clip_grad_norm_)
// grad_scale = 1 / count when `count` (device, double) is given (global masked-mean normalisation of
// the objective, see rl_kernels.cu), else 1.
// scal[2] = 1 when the update is applied, 0 when it is skipped: a batch without a single valid objective term
// (count <= 0) has a constant loss in the reference, so Lightning takes no optimizer step - no weight decay, no
// moment decay, no step-counter increment.  scal[3] = lr / (1 - beta1^step), scal[4] = sqrt(1 - beta2^step),
// scal[5] = 1 - lr * weight_decay.
// `hyper` (device, optional): [0] learning rate, [1] number of updates applied so far (incremented here when the
// update is applied) - the graph-capturable form; without it lr / step come from the host arguments.
__global__ void sumsq_finalize_kernel(const double* __restrict__ partial, int n_partial, const double* __restrict__ count,
                                      float max_norm, float* __restrict__ scal, float* __restrict__ hyper, float lr_host,
                                      int step_host, float beta1, float beta2, float weight_decay) {
    pdl_grid_sync();
    double t = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += 32) t += partial[i];
    t = warp_sum_d(t);
    if (threadIdx.x == 0) {
        float gs = 1.f;
        bool apply = true;
        if (count) { apply = count[0] > 0.0; gs = apply ? (float)(1.0 / count[0]) : 0.f; }
        const float norm = (float)sqrt(t) * gs;
        float coef = 1.f;
        if (max_norm > 0.f) coef = fminf(max_norm / (norm + 1e-6f), 1.f);
        float lr = lr_host;
        int step = step_host;
        if (hyper) {
            lr = hyper[0];
            step = (int)hyper[1] + (apply ? 1 : 0);
            hyper[1] = (float)step;
        }
        if (step < 1) step = 1;
        // torch computes the bias corrections in Python doubles
        const double bc1 = 1.0 - pow((double)beta1, (double)step);
        const double bc2 = 1.0 - pow((double)beta2, (double)step);
        scal[0] = norm;
        scal[1] = gs * coef;
        scal[2] = apply ? 1.f : 0.f;
        scal[3] = lr / (float)bc1;
        scal[4] = (float)sqrt(bc2);
        scal[5] = 1.f - lr * weight_decay;
    }
}

__global__ void __launch_bounds__(256)
adamw_apply_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   long long n, long long n_decay, const float* __restrict__ scal, float beta1, float beta2, float eps) {
    pdl_grid_sync();
    if (scal[2] == 0.f) return;                               // skipped update (see sumsq_finalize_kernel)
    const float gscale = scal[1];
    const float step_size = scal[3];
    const float sqrt_bc2 = scal[4];
    const float decay_keep = scal[5];
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
                      float weight_decay, int step, double* scratch, float* scal, cudaStream_t st, float* hyper) {
    RIFT_REQUIRE(hyper != nullptr || step >= 1, "AdamW step counter starts at 1");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                 "optimizer arenas must be 16-byte aligned");
    if (n <= 0) return 0;
    const int nb = (int)min((long long)optim_scratch_doubles(), (n / 4 + NORM_THREADS - 1) / NORM_THREADS + 1);
    launch_k(sumsq_partial_kernel, nb, NORM_THREADS, 0, st, g, n, scratch);
    RIFT_LAUNCH_OK();
    launch_k(sumsq_finalize_kernel, 1, 32, 0, st, scratch, nb, count, max_norm, scal, hyper, lr, step, beta1, beta2, weight_decay);
    RIFT_LAUNCH_OK();
    const int grid = (int)min((long long)148 * 8, (n / 4 + 255) / 256 + 1);
    launch_k(adamw_apply_kernel, grid, 256, 0, st, p, g, m, v, n, n_decay, scal, beta1, beta2, eps);
    RIFT_LAUNCH_OK();
    return 0;
}

}  // namespace rift
