// Internal launcher declarations (host side).  Every launcher returns 0 on success, <0 on error
// (message via rift::get_last_error()).  All pointers are device pointers unless noted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "split.cuh"

namespace rift {

// ------------------------------------------------------------------ rl_kernels.cu
int launch_group_advantage(const double* ret, const long long* offsets, long long n_groups, int G, double* adv,
                           cudaStream_t st);
int launch_group_objective(int algo /*0 rift, 1 grpo*/, const float* logits, const float* old_logits,
                           const float* ref_logits, const double* adv, const uint8_t* valid, const uint8_t* r_pad,
                           int bs, int R, int Mo, float clip_lo, float clip_hi, float dual_clip, float kl_w,
                           double* part_sum, int* part_cnt, double* out3, float* dlogits, int scale_dlogits,
                           cudaStream_t st);
int launch_action_objective(int mode /*0 ppo, 1 reinforce*/, const float* logits, const uint8_t* r_pad,
                            const long long* action_mode, const float* weight, const float* old_log_prob, int bs, int R,
                            int Mo, float eps_clip, float lambda_entropy, float inv_n, const float* extra_loss,
                            float* part, float* loss_out, float* dlogits, int* chosen, cudaStream_t st);
// SFT / RTR teacher cross-entropy (sft_trainer.py:123-215); dlogits assigned or accumulated, loss_out likewise
int launch_teacher_objective(const float* logits, const uint8_t* r_pad, const float* traj, const float* teacher, int bs, int R,
                             int Mo, int T, int step, float inv_n, float weight, float* part, float* loss_out, float* dlogits,
                             int accumulate, int* label_out, cudaStream_t st);
int launch_smooth_l1(const float* value, const float* target, int n, float inv_n, float* loss_out, float* dvalue,
                     cudaStream_t st);
int launch_gae(const float* rewards, const float* undones, const float* values, const float* next_values,
               const float* unterminated, int n, float gamma, float lam, float* adv, float* reward_sum,
               float* adv_normalised, cudaStream_t st);
int launch_discounted_return(const float* rewards, const float* dones, int n, float gamma, float* out, cudaStream_t st);

// ------------------------------------------------------------------ optim_kernels.cu
int optim_scratch_doubles();
int launch_clip_adamw(float* p, const float* g, float* m, float* v, long long n, long long n_decay,
                      const double* count, float max_norm, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, double* scratch, float* scal, cudaStream_t st, float* hyper = nullptr);

// ------------------------------------------------------------------ gemm (gemm_simt.cu / gemm_tc.cu)
// C[m,n] = post( act( (sum_k A(m,k) B(n,k) + pre[m/pre_div, n]) * colscale[n] + bias[n] ) + res[m/res_div, n] ) + beta*C[m,n]
struct GemmArgs {
    const float* A = nullptr; long long sam = 0, sak = 1;    // A(m,k) = A[m*sam + k*sak]
    const float* B = nullptr; long long sbn = 0, sbk = 1;    // B(n,k) = B[n*sbn + k*sbk]
    float* C = nullptr; long long ldc = 0;
    int M = 0, N = 0, K = 0;
    const float* bias = nullptr;         // [N]
    const float* colscale = nullptr;     // [N]  (eval-mode BatchNorm folded into the producing Linear)
    const float* pre = nullptr; long long ldpre = 0; int pre_div = 1;   // added before scale/bias/act
    const float* res = nullptr; long long ldres = 0; int res_div = 1;   // added after act
    int res_mod = 0;                     // if >0 the residual row is (m % res_mod) instead of m / res_div
    int act = 0;                         // ActKind
    float beta = 0.f;                    // accumulate into C (gradient arenas)
    float alpha = 1.f;                   // scales the raw product
    int split_k = 1;                     // >1: partial products in workspace, reduced deterministically
    float* split_ws = nullptr;           // [split_k, M, N] when split_k > 1
    float* preact = nullptr;             // optional copy of the value before `act` (same ldc), for backward
    Planes out_planes;                   // optional split-bf16 copy of the result (tcgen05 path; C may then be null)
    // tcgen05 path, backward of an activation fused into the data-gradient product: the result is multiplied by
    // act'(dact_ref[m, n]) (ReLU: dact_ref = the activation's OUTPUT; GELU: its pre-activation input), pitch lddact
    const float* dact_ref = nullptr; long long lddact = 0; int dact = 0;
    int terms = 3;                       // tcgen05 path: 3 = split-bf16 x3 products, 1 = plain bf16 (hi planes only)
    bool atomic_out = false;             // tcgen05 path: C += raw product with red.global.add from every (tile, split-K part):
                                         // no partial slabs, no reduce kernel, summation order not fixed (weight gradients)
    float* colsum_out = nullptr;         // with atomic_out on the MN-major (dW = dY^T X) form: colsum_out[m] += sum_k A[k, m]
                                         // (the bias gradient, from a second tensor-core accumulator)
    int n_store = 0;                     // tcgen05 path: >0 = only the first n_store of the N (padded) columns are real; the
                                         // result then goes through the partial buffer and C / ldc need no alignment
};
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
// ws_pitch: row pitch of the partial slabs (0 = a.N); the reduce writes columns [0, a.N)
int launch_splitk_reduce(const float* ws, int splits, const GemmArgs& a, cudaStream_t st, int ws_pitch = 0);

// Second stage of a two-stage reduction may run on another stream: after the first stage is enqueued on the
// caller's stream, `ev` is recorded there and `st` waits for it (null st: everything stays on the caller's stream).
// The partial buffer must then stay untouched until that stream has consumed it.
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t ev = nullptr; };

// ------------------------------------------------------------------ nn_kernels.cu
int launch_layernorm(const float* x, long long ldx, int rows, int C, const float* gamma, const float* beta, float* y,
                     long long ldy, int relu, const float* add_rowmod, int rowmod, float* y2, float* mean, float* rstd,
                     cudaStream_t st, Planes yp = Planes(), Planes y2p = Planes());
int launch_layernorm_bwd(const float* x, long long ldx, const float* dy, long long lddy, int rows, int C,
                         const float* gamma, const float* mean, const float* rstd, const float* y_for_relu, long long ldy,
                         float* dx, long long lddx, int dx_accumulate, float* dgamma, float* dbeta, float* scratch,
                         cudaStream_t st, SideStream fin = SideStream(), Planes dx_planes = Planes(), const uint8_t* zero_flag = nullptr,
                         int zero_div = 1);      // dx_planes: dx also as split-bf16 planes (pitch C, C % 64 == 0); dx[row] = 0 where zero_flag[row / zero_div]
int layernorm_bwd_scratch_floats(int C);

struct AttnArgs {
    const float* q = nullptr; const float* k = nullptr; const float* v = nullptr; float* o = nullptr;
    long long ldq = 0, ldk = 0, ldv = 0, ldo = 0;    // row strides (floats); head h lives at column h*hd
    int B = 0, H = 0, Sq = 0, Sk = 0, hd = 0;
    // row(b, s) = (b / inner_n) * outer + (b % inner_n) * inner + s * seq
    int q_inner_n = 1; long long q_outer = 0, q_inner = 0, q_seq = 1;
    int k_inner_n = 1; long long k_outer = 0, k_inner = 0, k_seq = 1;
    const uint8_t* kpm = nullptr; int kpm_div = 1;   // key padding mask [B / kpm_div, Sk], 1 = masked
    int kpm_mod = 0;                                 // if >0 the mask row is ((b + kpm_off) % kpm_mod) (reference r2r quirk;
    int kpm_off = 0;                                 //  kpm_off = position of this micro-batch inside the whole batch)
    float scale = 1.f;
    // output rows default to the query rows; o_custom selects an own mapping (shared learned query)
    int o_custom = 0; int o_inner_n = 1; long long o_outer = 0, o_inner = 0, o_seq = 1;
    float* lse = nullptr;                            // [B, H, Sq] log-sum-exp of scaled logits (for backward)
    Planes o_planes;                                 // optional split-bf16 copy of the output (o may then be null)
};
int launch_attention(const AttnArgs& a, cudaStream_t st);
// optional plane destinations of dQ / dK / dV (pointers already offset to the gradient's first column inside the
// in-projection's dY matrix; an fp32 destination may then be null).  Vector kernel only: attention_bwd_planes_ok
struct AttnBwdPlanes { Planes q, k, v; };
bool attention_bwd_planes_ok(int Sq, int Sk, int hd);
int launch_attention_bwd(const AttnArgs& a, const float* d_o, long long lddo, float* dq, long long lddq, float* dk,
                         float* dv, long long lddk, long long lddv, cudaStream_t st, const AttnBwdPlanes& pl = AttnBwdPlanes());

int launch_nat_attention(const float* qkv, int n_seq, int L, int heads, int hd, int ksize, const float* rpb, float* out,
                         cudaStream_t st, Planes op = Planes());
int launch_nat_attention_bwd(const float* qkv, const float* d_out, int n_seq, int L, int heads, int hd, int ksize,
                             const float* rpb, float* dqkv, float* drpb_partial, cudaStream_t st, Planes dqkv_planes = Planes());     // planes: dqkv may be null
bool nat_attention_bwd_planes_ok(int L, int ksize);

int launch_im2col_k3(const float* x, int n_seq, int L, int C, int stride, float* out, cudaStream_t st,
                     Planes op = Planes());   // -> (n_seq*Lout, C*3)
int launch_im2col_k3_last(const float* x, int n_seq, int L, int C, float* out, cudaStream_t st,
                          Planes op = Planes());           // -> (n_seq, C*3) at t=L-1
int launch_col2im_k3(const float* dcols, int n_seq, int L, int C, int stride, float* dx, int accumulate, cudaStream_t st);
int launch_col2im_k3_last(const float* dcols, int n_seq, int L, int C, float* dx, cudaStream_t st);
int launch_fpn_upsample_add(float* dst, const float* src, int n_seq, int Ld, int Ls, int C, cudaStream_t st);
int launch_fpn_upsample_add_bwd(const float* ddst, float* dsrc, int n_seq, int Ld, int Ls, int C, cudaStream_t st);

int launch_masked_maxpool(const float* x, const uint8_t* mask, int groups, int n, int C, float* out, int* argmax,
                          cudaStream_t st, Planes op = Planes());
int launch_masked_maxpool_bwd(const float* dout, const int* argmax, int groups, int n, int C, float* dx, int accumulate,
                              cudaStream_t st);
int launch_mask_any(const uint8_t* mask, int rows, int n, uint8_t* any_out, uint8_t* none_out, cudaStream_t st);
int launch_token_masks(const uint8_t* agent_valid, int agent_T, int Th, const uint8_t* map_valid, int P, int bs, int A,
                       int Mp, uint8_t* agent_any, uint8_t* key_pad, cudaStream_t st);
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* lin_bias,
                   int n, float* scale, float* shift, cudaStream_t st);

int launch_agent_features(const float* pos, const float* heading, const float* vel, const float* shape,
                          const uint8_t* valid, int n_agents, int Th, int Tstride, float* feat, cudaStream_t st);
int launch_map_features(const float* point_position, const float* point_vector, const float* point_orientation,
                        const float* polygon_center, int n_poly, int P, float* feat, cudaStream_t st);
int launch_ref_features(const float* position, const float* vector, const float* orientation, int n_ref, int Pr,
                        float* feat, float* rpos, cudaStream_t st);
int launch_token_pos(const float* agent_pos, const float* agent_heading, const float* polygon_center, int bs, int A,
                     int Th, int Tstride, int Mp, float* pos, cudaStream_t st);
int launch_fourier_features(const float* x, int rows, int d, int dsel, const float* freqs, int nfreq, float* feat,
                            int ldf, cudaStream_t st, Planes op = Planes());
int launch_fourier_features_bwd(const float* x, int rows, int d, int dsel, const float* freqs, int nfreq,
                                const float* dfeat, int ldf, float* dfreqs_partial, cudaStream_t st);

int launch_state_tokens(const float* cur_state, int cs_stride, int bs, int n_tok, int D, const float* const* w,
                        const float* const* b, const float* pos_embed, float* toks, cudaStream_t st);
int launch_agent_assemble(const float* x_hist, const float* x_ego, const uint8_t* agent_any, const int8_t* category,
                          const float* type_emb, int bs, int A, int S, int D, float* tokens, cudaStream_t st);
int launch_map_assemble(const float* x_poly, const float* x_speed, const int8_t* ptype, const uint8_t* on_route,
                        const int8_t* tl, const uint8_t* has_speed, const float* type_emb, const float* route_emb,
                        const float* tl_emb, const float* unknown_emb, int bs, int Mp, int A, int S, int D, float* tokens,
                        cudaStream_t st);
int launch_query_init(const float* u, const float* v, int rows, int Mo, int D, float* q, cudaStream_t st);   // q[row] = u[row/Mo] + v[row%Mo]
int launch_zero_rows(float* x, const uint8_t* rowflag, int flag_div, int rows, int C, cudaStream_t st);      // x[row]=0 where flag[row/div]
int launch_interleave_heads(const float* loc, const float* yaw, const float* vel, long long rows, int T, float* out,
                            cudaStream_t st);
int launch_mask_logits(float* pi, const uint8_t* r_pad, long long rows, int Mo, float fill, cudaStream_t st);
int launch_gather_rows(const float* x, long long ldx_batch, int bs, int row0, int nrows, int C, float* out, cudaStream_t st);
int launch_colsum(const float* x, long long ldx, int rows, int C, float* out, int accumulate, float* scratch,
                  cudaStream_t st, SideStream fin = SideStream());
// column sum of a split-bf16 plane pair (hi + lo), two-stage like launch_colsum; scratch: 148 * C floats
int launch_colsum_planes(const Planes& p, int rows, int C, float* out, int accumulate, float* scratch, cudaStream_t st,
                         SideStream fin = SideStream());
// second stage alone: out[c] (+)= sum over `slabs` partial rows of pitch C
int launch_colsum_final(const float* partial, int slabs, int C, float* out, int accumulate, cudaStream_t st,
                        SideStream fin = SideStream());
int launch_act_bwd(const float* pre_or_post, float* dy, long long n, int act, cudaStream_t st, Planes flat_planes = Planes());   // dy *= act'(.) in place, or -> planes (pitch == row length) with dy left untouched
int launch_add_inplace(float* dst, const float* src, long long n, cudaStream_t st);
int launch_scale_shift_rows_bwd(float* dy, const float* colscale, long long rows, int C, cudaStream_t st);
int launch_traj_outputs(const float* trajectory, long long n_traj, int T, float* cand, cudaStream_t st);

// ------------------------------------------------------------------ nn_bwd_kernels.cu
// dW[N, K] += dY^T X for narrow first-layer inputs (K <= 32, dW contiguous); partial: wgrad_narrow_slabs(M) * N * K floats
int wgrad_narrow_slabs(int M);
int launch_wgrad_narrow(const float* dY, long long lddy, const float* X, long long ldx, int M, int N, int K, float* dW,
                        float* partial, cudaStream_t st);
int launch_groupsum(const float* x, long long ldx, int groups, int n, int C, float* out, long long ldo, int accumulate,
                    cudaStream_t st);
int launch_modsum(const float* x, long long ldx, long long rows, int C, int mod, float* out, int accumulate, cudaStream_t st);
int launch_embedding_bwd(const float* dy, long long lddy, long long row_offset, int inner, int outer_stride_rows,
                         const int8_t* idx, long long rows, int C, int n_emb, float* demb, int invert_mask, float* scratch,
                         cudaStream_t st);
int launch_masked_gather_rows(const float* src, long long lds, long long row_offset, int inner, int outer_stride_rows,
                              const uint8_t* keep, long long rows, int C, float* out, int zero_inner0, cudaStream_t st);
int launch_scale_cols(float* dy, const float* colscale, long long rows, int C, cudaStream_t st);
int launch_bn_affine_bwd(const float* dz, const float* v, long long rows, int C, const float* gamma, const float* beta,
                         float* dgamma, float* dbeta, float* scratch, cudaStream_t st);
int launch_fourier_freq_bwd(const float* x, int rows, int d, int dsel, const float* freqs, int nfreq, const float* dfeat,
                            int ldf, float* contrib, cudaStream_t st);
int launch_state_tokens_bwd(const float* cur, int cs_stride, int bs, int n_tok, int D, const float* dtok, float* dw_all,
                            float* db_all, cudaStream_t st);

}  // namespace rift
