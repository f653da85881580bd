// extern "C" surface declared in include/rift_b200.h.
#include <stdlib.h>

#include <mutex>
#include <new>

#include "engine.h"
#include "fused_block.h"

namespace rift {
static thread_local std::string g_last_error;
long long g_kernel_launches = 0;
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }
}  // namespace rift

using namespace rift;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* rift_b200_last_error(void) { return get_last_error(); }
int rift_b200_version(void) { return 100; }
long long rift_b200_launch_count(void) { return g_kernel_launches; }
void rift_b200_debug_gemm_trace(void* dev_buf) { set_gemm_tc_trace(dev_buf); }
void rift_b200_debug_fused_trace(void* dev_buf) { set_fused_trace(dev_buf); }

int rift_b200_create(const rift_b200_model_config* cfg, const rift_b200_param_entry* entries, int n_entries,
                     rift_b200_engine** out) {
    RIFT_REQUIRE(cfg && entries && out && n_entries > 0, "create: null argument");
    RIFT_REQUIRE(cfg->dim % 32 == 0 && cfg->dim / 4 >= 32 / 2, "create: dim must be a multiple of 32");
    RIFT_REQUIRE(cfg->dim / cfg->num_heads == 32, "create: dim / num_heads must be 32");
    RIFT_REQUIRE(cfg->state_channel <= 8, "create: state_channel <= 8");
    rift_b200_engine* e = new (std::nothrow) rift_b200_engine();
    RIFT_REQUIRE(e != nullptr, "create: out of host memory");
    e->cfg = *cfg;
    long long lo = -1, hi = -1;
    for (int i = 0; i < n_entries; ++i) {
        RIFT_REQUIRE(entries[i].name != nullptr && entries[i].offset >= 0 && entries[i].numel >= 0, "create: bad entry");
        e->table[entries[i].name] = ParamRef{entries[i].offset, entries[i].numel, entries[i].trainable != 0};
        if (entries[i].trainable) {
            if (lo < 0 || entries[i].offset < lo) lo = entries[i].offset;
            if (entries[i].offset + entries[i].numel > hi) hi = entries[i].offset + entries[i].numel;
        }
    }
    e->train_lo = lo < 0 ? 0 : lo;
    e->train_hi = hi < 0 ? 0 : hi;
    *out = e;
    return 0;
}

void rift_b200_destroy(rift_b200_engine* e) { delete e; }

int rift_b200_bind_arena(rift_b200_engine* e, float* params, float* grads, long long numel) {
    RIFT_REQUIRE(e && params, "bind_arena: null argument");
    for (auto& kv : e->table)
        RIFT_REQUIRE(kv.second.offset + kv.second.numel <= numel, "bind_arena: entry beyond the arena: " + kv.first);
    e->params = params; e->grads = grads; e->numel = numel;
    int r = e->build_model();
    e->bound = (r == 0);
    return r;
}

size_t rift_b200_workspace_bytes(const rift_b200_engine* e_, const rift_b200_batch* shape) {
    if (!e_ || !shape || !e_->bound) return 0;
    rift_b200_engine* e = const_cast<rift_b200_engine*>(e_);
    rift_b200_outputs out;
    float* dummy = reinterpret_cast<float*>(0x100);
    out.probability = dummy; out.trajectory = dummy; out.prediction = dummy; out.hidden = dummy;
    out.ref_free_trajectory = dummy; out.candidate_trajectories = dummy; out.r_padding_mask = nullptr;
    size_t need = 0;
    for (int mode = 0; mode < 2; ++mode) {          // the tensor-core and the exact-fp32 schedules allocate differently
        Ctx c; c.dry = true; c.save = true; c.simt = (mode == 1);
        if (e->forward(*shape, out, c) != 0) return 0;
        if (e->grads) {
            c.simt = (mode == 1);
            if (e->backward(*shape, nullptr, c) != 0) return 0;
        }
        if (c.off > need) need = c.off;
    }
    return need + 4096;
}

int rift_b200_forward(rift_b200_engine* e, const rift_b200_batch* batch, const rift_b200_outputs* out, void* workspace,
                      size_t workspace_bytes, int flags, void* stream) {
    RIFT_REQUIRE(e && batch && out && workspace, "forward: null argument");
    RIFT_REQUIRE(e->bound, "forward: bind_arena first");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "forward: workspace must be 256-byte aligned");
    Ctx c; c.st = S(stream); c.base = static_cast<char*>(workspace); c.cap = workspace_bytes;
    c.save = (flags & RIFT_B200_FWD_SAVE_FOR_BACKWARD) != 0;
    c.simt = (flags & RIFT_B200_GEMM_SIMT) != 0;
    return e->forward(*batch, *out, c);
}

size_t rift_b200_weight_cache_bytes(const rift_b200_engine* e) { return (e && e->bound) ? e->weight_cache_bytes() : 0; }

int rift_b200_bind_weight_cache(rift_b200_engine* e, void* cache, size_t bytes) {
    RIFT_REQUIRE(e && e->bound, "bind_weight_cache: bind_arena first");
    return e->bind_weight_cache(cache, bytes);
}

int rift_b200_params_updated(rift_b200_engine* e, int trainable_only) {
    RIFT_REQUIRE(e != nullptr, "params_updated: null engine");
    if (trainable_only == 2) { e->dirty_train = false; e->dirty_all = false; }
    else if (trainable_only) e->dirty_train = true;
    else e->dirty_all = true;
    return 0;
}

int rift_b200_backward(rift_b200_engine* e, const rift_b200_batch* batch, const float* dlogits, void* workspace,
                       size_t workspace_bytes, int flags, void* stream) {
    RIFT_REQUIRE(e && batch && dlogits && workspace, "backward: null argument");
    RIFT_REQUIRE(e->bound, "backward: bind_arena first");
    Ctx c; c.st = S(stream); c.base = static_cast<char*>(workspace); c.cap = workspace_bytes;
    c.off = e->fwd_ws_end;          // scratch goes after the activations saved by forward
    c.simt = (flags & RIFT_B200_GEMM_SIMT) != 0;
    return e->backward(*batch, dlogits, c);
}

// ---------------------------------------------------------------- objectives
size_t rift_b200_objective_scratch_bytes(int bs) { return (size_t)bs * (sizeof(double) + sizeof(int)) + 64; }

int rift_b200_group_objective(int algo, const float* logits, const float* old_logits, const float* ref_logits,
                              const double* advantage, const uint8_t* valid, const uint8_t* r_pad, int bs, int R, int Mo,
                              float clip_lo, float clip_hi, float dual_clip, float kl_weight, void* scratch, double* out3,
                              float* dlogits, int scale_by_count, void* stream) {
    RIFT_REQUIRE(logits && old_logits && advantage && valid && scratch && out3, "group_objective: null argument");
    RIFT_REQUIRE(algo == 0 || algo == 1, "group_objective: algo must be 0 (rift) or 1 (grpo)");
    double* part_sum = static_cast<double*>(scratch);
    int* part_cnt = reinterpret_cast<int*>(part_sum + bs);
    return launch_group_objective(algo, logits, old_logits, ref_logits, advantage, valid, r_pad, bs, R, Mo, clip_lo, clip_hi,
                                  dual_clip, kl_weight, part_sum, part_cnt, out3, dlogits, scale_by_count, S(stream));
}

int rift_b200_action_objective(int mode, const float* logits, const uint8_t* r_pad, const long long* action_mode,
                               const float* weight, const float* old_log_prob, int bs, int R, int Mo, float clip_epsilon,
                               float lambda_entropy, float inv_n, const float* extra_loss, float* scratch, float* loss_out,
                               float* dlogits, int* chosen, void* stream) {
    RIFT_REQUIRE(logits && r_pad && weight && scratch && loss_out, "action_objective: null argument");
    RIFT_REQUIRE(mode == 1 || (action_mode && old_log_prob), "action_objective: PPO needs action_mode and old_log_prob");
    return launch_action_objective(mode, logits, r_pad, action_mode, weight, old_log_prob, bs, R, Mo, clip_epsilon,
                                   lambda_entropy, inv_n, extra_loss, scratch, loss_out, dlogits, chosen, S(stream));
}

int rift_b200_teacher_objective(const float* logits, const uint8_t* r_pad, const float* trajectory, const float* teacher_infos,
                                int bs, int R, int Mo, int T, int frame_rate, float inv_n, float weight, float* scratch,
                                float* loss_out, float* dlogits, int accumulate, int* label_out, void* stream) {
    RIFT_REQUIRE(logits && r_pad && trajectory && teacher_infos && scratch && loss_out, "teacher_objective: null argument");
    return launch_teacher_objective(logits, r_pad, trajectory, teacher_infos, bs, R, Mo, T, frame_rate, inv_n, weight, scratch,
                                    loss_out, dlogits, accumulate, label_out, S(stream));
}

int rift_b200_smooth_l1(const float* value, const float* target, int n, float inv_n, float* loss_out, float* dvalue,
                        void* stream) {
    RIFT_REQUIRE(value && target && loss_out, "smooth_l1: null argument");
    return launch_smooth_l1(value, target, n, inv_n, loss_out, dvalue, S(stream));
}

int rift_b200_group_advantage(const double* returns, const long long* offsets, long long n_groups, int G, double* advantage,
                              void* stream) {
    RIFT_REQUIRE(returns && advantage, "group_advantage: null argument");
    return launch_group_advantage(returns, offsets, n_groups, G, advantage, S(stream));
}

int rift_b200_gae(const float* rewards, const float* undones, const float* values, const float* next_values,
                  const float* unterminated, int n, float gamma, float lambda_, float* advantage, float* reward_sum,
                  float* advantage_normalised, void* stream) {
    RIFT_REQUIRE(rewards && undones && values && next_values && unterminated && advantage, "gae: null argument");
    return launch_gae(rewards, undones, values, next_values, unterminated, n, gamma, lambda_, advantage, reward_sum,
                      advantage_normalised, S(stream));
}

int rift_b200_discounted_return(const float* rewards, const float* dones, int n, float gamma, float* returns, void* stream) {
    RIFT_REQUIRE(rewards && dones && returns, "discounted_return: null argument");
    return launch_discounted_return(rewards, dones, n, gamma, returns, S(stream));
}

size_t rift_b200_optim_scratch_bytes(void) { return (size_t)optim_scratch_doubles() * sizeof(double); }

int rift_b200_clip_adamw(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const double* count,
                         float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                         void* scratch, float* scal_out, void* stream) {
    RIFT_REQUIRE(p && g && m && v && scratch && scal_out, "clip_adamw: null argument");
    return launch_clip_adamw(p, g, m, v, n, n_decay, count, max_norm, lr, beta1, beta2, eps, weight_decay, step,
                             static_cast<double*>(scratch), scal_out, S(stream));
}

int rift_b200_clip_adamw_dev(float* p, const float* g, float* m, float* v, long long n, long long n_decay, const double* count,
                             float max_norm, float* hyper, float beta1, float beta2, float eps, float weight_decay,
                             void* scratch, float* scal_out, void* stream) {
    RIFT_REQUIRE(p && g && m && v && scratch && scal_out && hyper, "clip_adamw_dev: null argument");
    return launch_clip_adamw(p, g, m, v, n, n_decay, count, max_norm, 0.f, beta1, beta2, eps, weight_decay, 0,
                             static_cast<double*>(scratch), scal_out, S(stream), hyper);
}

int rift_b200_refresh_weights(rift_b200_engine* e, void* stream) {
    RIFT_REQUIRE(e && e->bound, "refresh_weights: bind_arena first");
    return e->refresh_weights(S(stream));
}

// ---------------------------------------------------------------- primitive operators (tests)
int rift_b200_op_linear(const float* x, int rows, int K, const float* w, const float* bias, int N, int act, const float* res,
                        float* y, int simt, void* stream) {
    GemmArgs a;
    a.A = x; a.sam = K; a.B = w; a.sbn = K; a.C = y; a.ldc = N; a.M = rows; a.N = N; a.K = K;
    a.bias = bias; a.act = act; a.res = res; a.ldres = N;
    (void)simt;
    return launch_gemm_simt(a, S(stream));
}

size_t rift_b200_op_linear_tc_scratch_bytes(int rows, int N, int K) {
    const int Kp = (K + 63) / 64 * 64;
    return 2 * (((size_t)N * Kp * 2 + 255) & ~(size_t)255) + 512 + 2 * (((size_t)rows * Kp * 2 + 255) & ~(size_t)255);
}

// tcgen05 path of op_linear: splits `w` into bf16 planes inside `scratch`, builds the TMA descriptors and runs
// the tensor-core kernel (synchronises the stream once; test / micro-benchmark use only)
int rift_b200_op_linear_tc(const float* x, int rows, int K, const float* w, const float* bias, int N, int act,
                           const float* res, float* y, void* scratch, size_t scratch_bytes, int resplit, void* stream) {
    RIFT_REQUIRE(x && w && y && scratch, "op_linear_tc: null argument");
    RIFT_REQUIRE(scratch_bytes >= rift_b200_op_linear_tc_scratch_bytes(rows, N, K), "op_linear_tc: scratch too small");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "op_linear_tc: scratch must be 256-byte aligned");
    TcWeight tw;
    tw.src = w; tw.ld_src = K; tw.N = N; tw.K = K; tw.Kp = (K + 63) / 64 * 64;
    const size_t plane = ((size_t)N * tw.Kp * 2 + 255) & ~(size_t)255;
    char* p = static_cast<char*>(scratch);
    tw.hi = p; tw.lo = p + plane;
    if (resplit & 1) {
        std::vector<char> job(split_job_bytes());
        fill_split_job(job.data(), w, K, N, K, tw.Kp, tw.hi, tw.lo, 0, 0);
        RIFT_CUDA_OK(cudaMemcpyAsync(p + 2 * plane, job.data(), job.size(), cudaMemcpyHostToDevice, S(stream)));
        RIFT_CUDA_OK(cudaStreamSynchronize(S(stream)));
        int r = launch_split_weights(p + 2 * plane, 1, split_job_units(N, tw.Kp), S(stream));
        if (r) return r;
    }
    int r = 0;
    GemmArgs a;
    a.A = x; a.sam = K; a.B = w; a.sbn = K; a.C = y; a.ldc = N; a.M = rows; a.N = N; a.K = K;
    a.bias = bias; a.act = act; a.res = res; a.ldres = N;
    char* ap = p + 2 * plane + 512;
    const size_t aplane = ((size_t)rows * tw.Kp * 2 + 255) & ~(size_t)255;
    if (!(resplit & 2)) r = launch_pack_split(x, K, rows, K, tw.Kp, ap, ap + aplane, S(stream));   // bit 1: reuse the A planes
    if (r) return r;                                                              // A planes are per-call data
    return launch_gemm_tc(a, ap, ap + aplane, tw.Kp, tw, 0, 0, S(stream));
}

// op_linear_tc with every output form of the epilogue: y (+)= beta * y, optional pre-activation copy and split-bf16 planes
int rift_b200_op_linear_tc_full(const float* x, int rows, int K, const float* w, const float* bias, int N, int act,
                                const float* res, float* y, float beta, float* preact, void* out_hi, void* out_lo, void* scratch,
                                size_t scratch_bytes, void* stream) {
    RIFT_REQUIRE(x && w && scratch && (y || (out_hi && out_lo)), "op_linear_tc_full: null argument");
    RIFT_REQUIRE(scratch_bytes >= rift_b200_op_linear_tc_scratch_bytes(rows, N, K), "op_linear_tc_full: scratch too small");
    TcWeight tw;
    tw.src = w; tw.ld_src = K; tw.N = N; tw.K = K; tw.Kp = (K + 63) / 64 * 64;
    const size_t plane = ((size_t)N * tw.Kp * 2 + 255) & ~(size_t)255;
    char* p = static_cast<char*>(scratch);
    tw.hi = p; tw.lo = p + plane;
    std::vector<char> job(split_job_bytes());
    fill_split_job(job.data(), w, K, N, K, tw.Kp, tw.hi, tw.lo, 0, 0);
    RIFT_CUDA_OK(cudaMemcpyAsync(p + 2 * plane, job.data(), job.size(), cudaMemcpyHostToDevice, S(stream)));
    RIFT_CUDA_OK(cudaStreamSynchronize(S(stream)));
    int r = launch_split_weights(p + 2 * plane, 1, split_job_units(N, tw.Kp), S(stream));
    if (r) return r;
    GemmArgs a;
    a.A = x; a.sam = K; a.B = w; a.sbn = K; a.C = y; a.ldc = N; a.M = rows; a.N = N; a.K = K;
    a.bias = bias; a.act = act; a.res = res; a.ldres = N; a.beta = beta; a.preact = preact;
    if (out_hi && out_lo) { a.out_planes.hi = static_cast<uint16_t*>(out_hi); a.out_planes.lo = static_cast<uint16_t*>(out_lo); a.out_planes.Kp = (N + 63) / 64 * 64; }
    char* ap = p + 2 * plane + 512;
    const size_t aplane = ((size_t)rows * tw.Kp * 2 + 255) & ~(size_t)255;
    r = launch_pack_split(x, K, rows, K, tw.Kp, ap, ap + aplane, S(stream));
    if (r) return r;
    return launch_gemm_tc(a, ap, ap + aplane, tw.Kp, tw, 0, 0, S(stream));
}

size_t rift_b200_op_fused_mlp_scratch_bytes(int D, int Hd) {
    const size_t p1 = ((size_t)Hd * ((D + 63) / 64 * 64) * 2 + 255) & ~(size_t)255, p2 = ((size_t)D * ((Hd + 63) / 64 * 64) * 2 + 255) & ~(size_t)255;
    return 2 * p1 + 2 * p2 + 1024;
}

// Pre-LN MLP sub-block y = x + fc2(act(fc1(LN(x)))) through the fused cluster kernel (fused_block.cu); the optional
// outputs are the tensors the backward reads.  Splits w1 / w2 into planes inside `scratch` first (test / micro-benchmark use).
int rift_b200_op_fused_mlp(const float* x, int rows, int D, int Hd, int act, const float* ln_g, const float* ln_b,
                           const float* w1, const float* b1, const float* w2, const float* b2, float* y, float* mean,
                           float* rstd, void* t2_hi, void* t2_lo, float* hpre, void* hm_hi, void* hm_lo, void* scratch,
                           size_t scratch_bytes, int resplit, void* stream) {
    RIFT_REQUIRE(x && w1 && w2 && y && scratch, "op_fused_mlp: null argument");
    RIFT_REQUIRE(scratch_bytes >= rift_b200_op_fused_mlp_scratch_bytes(D, Hd), "op_fused_mlp: scratch too small");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "op_fused_mlp: scratch must be 256-byte aligned");
    RIFT_REQUIRE(fused_mlp_shape_ok(rows < 64 ? 64 : rows, D, Hd), "op_fused_mlp: unsupported shape");
    TcWeight t1, t2;
    t1.src = w1; t1.ld_src = D; t1.N = Hd; t1.K = D; t1.Kp = (D + 63) / 64 * 64;
    t2.src = w2; t2.ld_src = Hd; t2.N = D; t2.K = Hd; t2.Kp = (Hd + 63) / 64 * 64;
    const size_t p1 = ((size_t)Hd * t1.Kp * 2 + 255) & ~(size_t)255, p2 = ((size_t)D * t2.Kp * 2 + 255) & ~(size_t)255;
    char* p = static_cast<char*>(scratch);
    t1.hi = p; t1.lo = p + p1; t2.hi = p + 2 * p1; t2.lo = p + 2 * p1 + p2;
    if (resplit) {
        std::vector<char> jobs(2 * split_job_bytes());
        fill_split_job(jobs.data(), w1, D, Hd, D, t1.Kp, t1.hi, t1.lo, 0, 0);
        fill_split_job(jobs.data() + split_job_bytes(), w2, Hd, D, Hd, t2.Kp, t2.hi, t2.lo, split_job_units(Hd, t1.Kp), 0);
        char* jd = p + 2 * p1 + 2 * p2;
        RIFT_CUDA_OK(cudaMemcpyAsync(jd, jobs.data(), jobs.size(), cudaMemcpyHostToDevice, S(stream)));
        RIFT_CUDA_OK(cudaStreamSynchronize(S(stream)));
        int r = launch_split_weights(jd, 2, split_job_units(Hd, t1.Kp) + split_job_units(D, t2.Kp), S(stream));
        if (r) return r;
    }
    FusedMlpArgs a;
    a.X = x; a.ldx = D; a.Y = y; a.ldy = D; a.rows = rows; a.D = D; a.Hd = Hd; a.act = act;
    a.ln_g = ln_g; a.ln_b = ln_b; a.b1 = b1; a.b2 = b2; a.ln_mean = mean; a.ln_rstd = rstd; a.hpre = hpre;
    if (t2_hi && t2_lo) { a.t2p.hi = static_cast<uint16_t*>(t2_hi); a.t2p.lo = static_cast<uint16_t*>(t2_lo); a.t2p.Kp = t1.Kp; }
    if (hm_hi && hm_lo) { a.hmp.hi = static_cast<uint16_t*>(hm_hi); a.hmp.lo = static_cast<uint16_t*>(hm_lo); a.hmp.Kp = t2.Kp; }
    return launch_fused_mlp(a, t1, t2, S(stream));
}

size_t rift_b200_op_wgrad_tc_scratch_bytes(int rows, int N, int K) {
    const size_t np_ = (size_t)(N + 63) / 64 * 64, kp_ = (size_t)(K + 63) / 64 * 64;
    return 2 * (((size_t)rows * np_ * 2 + 255) & ~(size_t)255) + 2 * (((size_t)rows * kp_ * 2 + 255) & ~(size_t)255);
}

// dW[N, K] += dY^T X and db[N] += colsum(dY) on the tcgen05 weight-gradient path (MN-major operands, split-K parts added
// with red.global.add, bias gradient from the second accumulator); dW / db must hold the running sums (e.g. zeros)
int rift_b200_op_wgrad_tc(const float* dY, const float* X, int rows, int N, int K, float* dW, float* db, int splits,
                          void* scratch, size_t scratch_bytes, void* stream) {
    RIFT_REQUIRE(dY && X && dW && scratch, "op_wgrad_tc: null argument");
    RIFT_REQUIRE(scratch_bytes >= rift_b200_op_wgrad_tc_scratch_bytes(rows, N, K), "op_wgrad_tc: scratch too small");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "op_wgrad_tc: scratch must be 256-byte aligned");
    const int Np = (N + 63) / 64 * 64, Kp = (K + 63) / 64 * 64, K4 = (K + 3) & ~3;
    char* p = static_cast<char*>(scratch);
    const size_t ypl = ((size_t)rows * Np * 2 + 255) & ~(size_t)255, xpl = ((size_t)rows * Kp * 2 + 255) & ~(size_t)255;
    void *yh = p, *yl = p + ypl, *xh = p + 2 * ypl, *xl = p + 2 * ypl + xpl;
    int r = launch_pack_split(dY, N, rows, N, Np, yh, yl, S(stream));
    if (r) return r;
    r = launch_pack_split(X, K, rows, K, Kp, xh, xl, S(stream));
    if (r) return r;
    GemmArgs a;
    a.C = dW; a.ldc = K; a.M = N; a.N = K4; a.K = rows;
    a.atomic_out = true; a.colsum_out = db;
    if (K4 != K) a.n_store = K;
    PlaneOp A{yh, yl, rows, Np, 0, 0};
    PlaneOp B{xh, xl, rows, Kp, 0, 0};
    return launch_gemm_tc_ex(a, A, B, true, splits, nullptr, S(stream));
}

// n <= 12 products dW_i[N_i, K_i] += dY_i^T X_i, db_i[N_i] += colsum(dY_i) in ONE grouped launch (wgrad_group_kernel); arrays of
// n host-side entries; scratch as for n calls of rift_b200_op_wgrad_tc (sum of the per-product sizes)
int rift_b200_op_wgrad_group(int n, const float* const* dY, const float* const* X, const int* rows, const int* N, const int* K,
                             float* const* dW, float* const* db, void* scratch, size_t scratch_bytes, void* stream) {
    RIFT_REQUIRE(n >= 1 && n <= WGRAD_GROUP_MAX && dY && X && rows && N && K && dW && db && scratch, "op_wgrad_group: bad argument");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "op_wgrad_group: scratch must be 256-byte aligned");
    WgradItem items[WGRAD_GROUP_MAX];
    char* p = static_cast<char*>(scratch);
    size_t used = 0;
    for (int i = 0; i < n; ++i) {
        RIFT_REQUIRE(wgrad_group_takes(N[i], K[i], rows[i], K[i], dW[i]), "op_wgrad_group: product not eligible (N >= 64, K > 64, K % 4 == 0, rows >= 64)");
        const int Np = (N[i] + 63) / 64 * 64, Kp = (K[i] + 63) / 64 * 64;
        const size_t ypl = ((size_t)rows[i] * Np * 2 + 255) & ~(size_t)255, xpl = ((size_t)rows[i] * Kp * 2 + 255) & ~(size_t)255;
        used += 2 * ypl + 2 * xpl;
        RIFT_REQUIRE(used <= scratch_bytes, "op_wgrad_group: scratch too small");
        void *yh = p, *yl = p + ypl, *xh = p + 2 * ypl, *xl = p + 2 * ypl + xpl;
        p += 2 * ypl + 2 * xpl;
        int r = launch_pack_split(dY[i], N[i], rows[i], N[i], Np, yh, yl, S(stream));
        if (r) return r;
        r = launch_pack_split(X[i], K[i], rows[i], K[i], Kp, xh, xl, S(stream));
        if (r) return r;
        items[i].A = PlaneOp{yh, yl, rows[i], Np, 0, 0};
        items[i].B = PlaneOp{xh, xl, rows[i], Kp, 0, 0};
        items[i].C = dW[i]; items[i].ldc = K[i]; items[i].colsum = db[i];
        items[i].M = N[i]; items[i].N = K[i]; items[i].K = rows[i];
    }
    return launch_wgrad_group(items, n, 3, S(stream));
}

int rift_b200_op_gemm(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk, float* C,
                      long long ldc, int M, int N, int K, float beta, int split_k, float* split_ws, int simt, void* stream) {
    GemmArgs a;
    a.A = A; a.sam = sam; a.sak = sak; a.B = B; a.sbn = sbn; a.sbk = sbk; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
    a.beta = beta; a.split_k = split_k; a.split_ws = split_ws;
    (void)simt;
    return launch_gemm_simt(a, S(stream));
}

int rift_b200_op_layernorm(const float* x, int rows, int C, const float* gamma, const float* beta, int relu, float* y,
                           float* mean, float* rstd, void* stream) {
    return launch_layernorm(x, C, rows, C, gamma, beta, y, C, relu, nullptr, 0, nullptr, mean, rstd, S(stream));
}

int rift_b200_op_layernorm_bwd(const float* x, const float* dy, int rows, int C, const float* gamma, const float* mean,
                               const float* rstd, const float* y_relu, float* dx, float* dgamma, float* dbeta, float* scratch,
                               void* stream) {
    return launch_layernorm_bwd(x, C, dy, C, rows, C, gamma, mean, rstd, y_relu, C, dx, C, 0, dgamma, dbeta, scratch, S(stream));
}

int rift_b200_op_attention(const float* qkv, int B, int Sq, int H, int hd, const uint8_t* key_padding, float* out,
                           void* stream) {
    const int D = H * hd;
    AttnArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = out;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.B = B; a.H = H; a.Sq = Sq; a.Sk = Sq; a.hd = hd;
    a.q_outer = Sq; a.k_outer = Sq;
    a.kpm = key_padding; a.scale = 1.f / sqrtf((float)hd);
    return launch_attention(a, S(stream));
}

int rift_b200_op_nat_attention(const float* qkv, int n_seq, int L, int heads, int hd, int ksize, const float* rpb, float* out,
                               void* stream) {
    return launch_nat_attention(qkv, n_seq, L, heads, hd, ksize, rpb, out, S(stream));
}

int rift_b200_op_attention_bwd(const float* qkv, const float* d_out, int B, int Sq, int H, int hd, const uint8_t* key_padding,
                               float* out, float* lse, float* dqkv, void* stream) {
    RIFT_REQUIRE(qkv && d_out && out && lse && dqkv, "op_attention_bwd: null argument");
    const int D = H * hd;
    AttnArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = out;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.B = B; a.H = H; a.Sq = Sq; a.Sk = Sq; a.hd = hd;
    a.q_outer = Sq; a.k_outer = Sq;
    a.kpm = key_padding; a.scale = 1.f / sqrtf((float)hd);
    a.lse = lse;
    int r = launch_attention(a, S(stream));
    if (r) return r;
    return launch_attention_bwd(a, d_out, D, dqkv, 3 * D, dqkv + D, dqkv + 2 * D, 3 * D, 3 * D, S(stream));
}

int rift_b200_op_nat_attention_bwd(const float* qkv, const float* d_out, int n_seq, int L, int heads, int hd, int ksize,
                                   const float* rpb, float* dqkv, float* drpb_partial, void* stream) {
    RIFT_REQUIRE(qkv && d_out && rpb && dqkv, "op_nat_attention_bwd: null argument");
    return launch_nat_attention_bwd(qkv, d_out, n_seq, L, heads, hd, ksize, rpb, dqkv, drpb_partial, S(stream));
}

int rift_b200_op_act_bwd(const float* ref, float* dy, long long n, int act, void* stream) {
    return launch_act_bwd(ref, dy, n, act, S(stream));
}

int rift_b200_op_add_inplace(float* dst, const float* src, long long n, void* stream) {
    return launch_add_inplace(dst, src, n, S(stream));
}

int rift_b200_op_colsum(const float* x, int rows, int C, float* out, int accumulate, float* scratch, void* stream) {
    return launch_colsum(x, C, rows, C, out, accumulate, scratch, S(stream));
}

int rift_b200_op_masked_maxpool(const float* x, const uint8_t* mask, int groups, int n, int C, float* out, int* argmax,
                                void* stream) {
    return launch_masked_maxpool(x, mask, groups, n, C, out, argmax, S(stream));
}

}  // extern "C"
