/* rift_b200 — C ABI of the B200-native RIFT policy-update hot path.
 *
 * Drop-in boundary for the reference's fine-tuning inner loop (paths relative to the reference
 * tree, CurryChen77/RIFT @ 49d643c):
 *   PlanningModel.forward                   rift/cbv/planning/pluto/model/pluto_model.py:122-225
 *   LightningTrainer.get_{rift,grpo,ppo,reinforce}_loss
 *                                           rift/cbv/planning/fine_tuner/rlft/<algo>_pluto/<algo>_trainer.py
 *   loss.backward / clip_grad_norm_ / AdamW rift/cbv/planning/fine_tuner/rlft/config/lightning/custom_lightning.yaml:40-41,
 *                                           rift/cbv/planning/fine_tuner/rlft/rift_pluto/rift_trainer.py:279-362
 *   group-relative advantage                rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:466-470
 *   GAE / discounted return                 rift/cbv/planning/fine_tuner/rlft/ppo_pluto/ppo_datamodule.py:22-37,163
 *                                           rift/cbv/planning/fine_tuner/rlft/reinforce_pluto/reinforce_datamodule.py:19-38
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; pass tensor.data_ptr())
 *     unless the parameter is documented as host memory;
 *   - every call is asynchronous on the `stream` argument (a cudaStream_t passed as void*; use
 *     torch.cuda.current_stream().cuda_stream);
 *   - return value 0 = ok, -1 = invalid argument, -2 = CUDA error; rift_b200_last_error() gives the text;
 *   - no internal threads, no hidden allocations after rift_b200_create; a handle is not thread-safe;
 *   - bool tensors are passed as uint8 (torch.bool storage), int8 categorical tensors as int8.
 */
#ifndef RIFT_B200_H
#define RIFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rift_b200_engine rift_b200_engine;

const char* rift_b200_last_error(void);
int rift_b200_version(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py: gpu_launches) */
long long rift_b200_launch_count(void);
/* profiling aid: CTA 0 of every following tcgen05 GEMM launch writes %globaltimer stamps (ns) into dev_buf
 * (64 uint64: 0 entry, 1 set-up done, 2 first operands landed, 3 first tile's MMAs issued, 4+2i / 5+2i
 * epilogue start / end of its i-th tile, 60 last role done, 61 TMEM released); NULL switches it off */
void rift_b200_debug_gemm_trace(void* dev_buf);
/* the same for the fused sub-block kernels (16 uint64: 0 set-up done, 1 dependency met, 2 LayerNorm planes written, 3 / 4 GEMM1
 * starts / issued, 5 its accumulator complete, 6 / 7 GEMM2 starts / issued, 8 its accumulator complete, 9 partial tile written,
 * 10 cluster barrier passed, 11 output stored, 12 TMEM released) */
void rift_b200_debug_fused_trace(void* dev_buf);

/* ---- model description (PlanningModel.__init__, pluto_model.py:23-44) ---- */
typedef struct {
    int dim, num_heads, encoder_depth, decoder_depth, num_modes;
    int history_steps, future_steps, state_channel, ref_points;
    int value_hidden0, value_hidden1;      /* CriticPPO hidden widths, 0 = no value net */
} rift_b200_model_config;

/* one state-dict entry of the flat fp32 parameter arena (name = the reference's state_dict key) */
typedef struct {
    const char* name;        /* host string */
    long long offset;        /* element offset into the arena */
    long long numel;
    int trainable;           /* LightningTrainer.freeze_parameters outcome (rift_trainer.py:78-90) */
} rift_b200_param_entry;

/* A collated PlutoFeature.data (pluto_feature.py:25-96) as raw device pointers. */
typedef struct {
    int bs, A, Mp, P, R, Pr;
    int agent_T;                             /* stored time length of agent tensors (>= history_steps) */
    const float* agent_position;             /* (bs, A, agent_T, 2) */
    const float* agent_heading;              /* (bs, A, agent_T)    */
    const float* agent_velocity;             /* (bs, A, agent_T, 2) */
    const float* agent_shape;                /* (bs, A, agent_T, 2) */
    const int8_t* agent_category;            /* (bs, A)             */
    const uint8_t* agent_valid_mask;         /* (bs, A, agent_T)    */
    const float* map_point_position;         /* (bs, Mp, 3, P, 2)   */
    const float* map_point_vector;           /* (bs, Mp, 3, P, 2)   */
    const float* map_point_orientation;      /* (bs, Mp, 3, P)      */
    const float* map_polygon_center;         /* (bs, Mp, 3)         */
    const int8_t* map_polygon_type;          /* (bs, Mp)            */
    const uint8_t* map_polygon_on_route;     /* (bs, Mp)            */
    const int8_t* map_polygon_tl_status;     /* (bs, Mp)            */
    const uint8_t* map_polygon_has_speed_limit; /* (bs, Mp)         */
    const float* map_polygon_speed_limit;    /* (bs, Mp)            */
    const uint8_t* map_valid_mask;           /* (bs, Mp, P)         */
    const float* ref_position;               /* (bs, R, Pr, 2)      */
    const float* ref_vector;                 /* (bs, R, Pr, 2)      */
    const float* ref_orientation;            /* (bs, R, Pr)         */
    const uint8_t* ref_valid_mask;           /* (bs, R, Pr)         */
    const float* current_state;              /* (bs, cs_stride), first state_channel columns used */
    int cs_stride;
    /* Micro-batching: when this struct describes samples [b_offset, b_offset + bs) of a larger collated batch, give the
     * WHOLE batch's reference-line mask and size.  The reference masks r2r attention row j = b * Mo + m with
     * r_pad[j % bs] (planning_decoder.py:56-60), which depends on the position inside the whole batch; with these
     * fields a slice computes exactly what the whole batch would.  NULL / 0 = this struct is the whole batch. */
    const uint8_t* ref_valid_mask_global;    /* (bs_global, R, Pr) */
    int bs_global, b_offset;
} rift_b200_batch;

/* Outputs of PlanningModel.forward; any pointer may be NULL to skip materialising that tensor. */
typedef struct {
    float* probability;          /* (bs, R, Mo) logits, padded reference lines filled with -1e6 */
    float* trajectory;           /* (bs, R, Mo, T, 6) */
    float* prediction;           /* (bs, A-1, T, 6)   */
    float* hidden;               /* (bs, dim)         */
    float* ref_free_trajectory;  /* (bs, T, 4)        */
    float* candidate_trajectories; /* (bs, R, Mo, T, 3), needs trajectory */
    uint8_t* r_padding_mask;     /* (bs, R) 1 = padded reference line */
} rift_b200_outputs;

/* ---- engine lifecycle ---- */
int rift_b200_create(const rift_b200_model_config* cfg, const rift_b200_param_entry* entries, int n_entries,
                     rift_b200_engine** out);
void rift_b200_destroy(rift_b200_engine* e);
/* params / grads: flat fp32 arenas laid out by `entries` (grads may be NULL for inference) */
int rift_b200_bind_arena(rift_b200_engine* e, float* params, float* grads, long long numel);
/* bytes of workspace rift_b200_forward/backward need for a batch of this shape */
size_t rift_b200_workspace_bytes(const rift_b200_engine* e, const rift_b200_batch* shape);

/* tcgen05 path: the engine keeps split-bf16 copies (hi + lo planes) of every weight matrix and the TMA
 * descriptors over them in caller-owned memory; without a bound cache every GEMM runs on the exact-fp32
 * SIMT kernel.  Call rift_b200_params_updated after the parameter arena was written from outside the
 * library (optimizer step: trainable_only = 1; checkpoint load: 0) - the planes are refreshed lazily by
 * the next forward. */
size_t rift_b200_weight_cache_bytes(const rift_b200_engine* e);
int rift_b200_bind_weight_cache(rift_b200_engine* e, void* cache, size_t bytes);
/* trainable_only: 1 = the trainable entries changed, 0 = everything changed, 2 = mark this engine's planes CLEAN (they live
 * in a weight cache shared with another engine of the same layout, which refreshes them) */
int rift_b200_params_updated(rift_b200_engine* e, int trainable_only);
/* re-split whatever rift_b200_params_updated marked stale NOW, on `stream` (a captured CUDA graph replays device
 * work only: after a checkpoint load the frozen planes must be refreshed eagerly, not by the next eager forward) */
int rift_b200_refresh_weights(rift_b200_engine* e, void* stream);

/* flags */
#define RIFT_B200_FWD_SAVE_FOR_BACKWARD 1   /* keep activations needed by rift_b200_backward in the workspace */
#define RIFT_B200_GEMM_SIMT 2               /* force the exact-fp32 SIMT GEMM (validation path) */

int rift_b200_forward(rift_b200_engine* e, const rift_b200_batch* batch, const rift_b200_outputs* out,
                      void* workspace, size_t workspace_bytes, int flags, void* stream);
/* d(loss)/d(probability) -> gradient arena (accumulated with `=` semantics after an internal zero of the
 * trainable range).  Must follow a forward with RIFT_B200_FWD_SAVE_FOR_BACKWARD on the same workspace. */
int rift_b200_backward(rift_b200_engine* e, const rift_b200_batch* batch, const float* dlogits,
                       void* workspace, size_t workspace_bytes, int flags, void* stream);

/* ---- RL objectives (standalone operators; usable without an engine) ---- */
/* RIFT (algo 0: clip [lo,hi] + dual clip) and GRPO (algo 1: clip + kl_weight * KL(ref || new)).
 * logits/old/ref (bs,R,Mo) f32, advantage (bs,R,Mo) f64, valid (bs,R,Mo) u8, r_pad (bs,R) u8 or NULL.
 * scratch: bs doubles + bs ints (use rift_b200_objective_scratch_bytes).
 * out3 (device, 3 doubles) = { loss, sum of objectives over valid, valid count }.
 * dlogits (may be NULL): d loss / d logits; if scale_by_count == 0 it is left multiplied by the valid
 * count (data-parallel use: all-reduce gradients and counts first, rift_b200_clip_adamw divides). */
size_t rift_b200_objective_scratch_bytes(int bs);
int rift_b200_group_objective(int algo, const float* logits, const float* old_logits, const float* ref_logits,
                              const double* advantage, const uint8_t* valid, const uint8_t* r_pad,
                              int bs, int R, int Mo, float clip_lo, float clip_hi, float dual_clip, float kl_weight,
                              void* scratch, double* out3, float* dlogits, int scale_by_count, void* stream);
/* PPO (mode 0) / REINFORCE (mode 1).  inv_n = 1 / global batch size.  extra_loss (device float, may be
 * NULL) is added to the scalar (PPO value loss).  loss_out: device float.  scratch: bs floats. */
int rift_b200_action_objective(int mode, const float* logits, const uint8_t* r_pad, const long long* action_mode,
                               const float* weight, const float* old_log_prob, int bs, int R, int Mo,
                               float clip_epsilon, float lambda_entropy, float inv_n, const float* extra_loss,
                               float* scratch, float* loss_out, float* dlogits, int* chosen, void* stream);
/* SFT / RTR teacher objective (fine_tuner/sft/sft_trainer.py:123-215, rtr_pluto/rtr_trainer.py:130-195): cross-entropy of the
 * candidate logits against the one-hot label (model's best reference line, mode whose PID target speed is closest to the
 * teacher's).  trajectory (bs, R, Mo, T, 6); teacher_infos (bs, 5) = [target speed, origin x, origin y, heading, speed].
 * loss_out (device float) = weight * inv_n * sum_b CE_b, assigned or (accumulate != 0) added; dlogits likewise (may be NULL);
 * label_out (bs ints, may be NULL) = flat label index r * Mo + m.  scratch: bs floats. */
int rift_b200_teacher_objective(const float* logits, const uint8_t* r_pad, const float* trajectory, const float* teacher_infos,
                                int bs, int R, int Mo, int T, int frame_rate, float inv_n, float weight, float* scratch,
                                float* loss_out, float* dlogits, int accumulate, int* label_out, void* stream);
int rift_b200_smooth_l1(const float* value, const float* target, int n, float inv_n, float* loss_out, float* dvalue,
                        void* stream);

/* ---- advantage / buffer passes ---- */
/* (ret - mean) / (std + 1e-5) per group in float64, bit-identical to numpy (population std).
 * offsets == NULL: n_groups contiguous groups of size G; else offsets[n_groups + 1] (int64 element offsets). */
int rift_b200_group_advantage(const double* returns, const long long* offsets, long long n_groups, int G,
                              double* advantage, void* stream);
int rift_b200_gae(const float* rewards, const float* undones, const float* values, const float* next_values,
                  const float* unterminated, int n, float gamma, float lambda_, float* advantage, float* reward_sum,
                  float* advantage_normalised, void* stream);
int rift_b200_discounted_return(const float* rewards, const float* dones, int n, float gamma, float* returns,
                                void* stream);

/* ---- optimizer: clip_grad_norm_(max_norm) + AdamW over the flat trainable range ---- */
/* p/g/m/v: n elements each, 16-byte aligned; the first n_decay elements get weight_decay.
 * count (device double, may be NULL): gradients are divided by it first (see group_objective).
 * scratch: rift_b200_optim_scratch_bytes() bytes.  scal_out (device, 2 floats) = { grad norm, applied scale }. */
size_t rift_b200_optim_scratch_bytes(void);
int rift_b200_clip_adamw(float* p, const float* g, float* m, float* v, long long n, long long n_decay,
                         const double* count, float max_norm, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int step, void* scratch, float* scal_out, void* stream);
/* Graph-capturable form of the same update: the learning rate and the number of updates applied so far live on the
 * device.  hyper (device, 2 floats): [0] learning rate (the host rewrites it when WarmupCosLR steps, warmup_cos_lr.py:39-54),
 * [1] updates applied so far (the kernel increments it).  A step whose `count` is <= 0 (no valid objective term: the
 * reference's loss is then a constant and Lightning takes no optimizer step) leaves parameters, moments and the
 * counter untouched.  scal_out (device, 8 floats) = { grad norm, applied scale, applied?, lr / bias1, sqrt(bias2),
 * 1 - lr * wd, -, - }. */
int rift_b200_clip_adamw_dev(float* p, const float* g, float* m, float* v, long long n, long long n_decay,
                             const double* count, float max_norm, float* hyper, float beta1, float beta2, float eps,
                             float weight_decay, void* scratch, float* scal_out, void* stream);

/* ---- device-resident replay buffer: GPU collate (cbv_rollout_buffer.py:16-138 + pluto_feature.py:25-96 + rift_datamodule.py:20-51)
 * One field = one tensor of the buffer arena laid out [capacity][item_stride_bytes]; the gather copies the first copy_bytes
 * bytes of slot idx[b] to dst + b * dst_stride_bytes for every b < bs (zero padding beyond an item's own extent is already in
 * the slot).  `fields` is a HOST array (passed to the kernel by value), idx a device array of bs slot indices. */
#define RIFT_B200_GATHER_MAX_FIELDS 40
typedef struct {
    const void* src; void* dst;
    long long item_stride_bytes, copy_bytes, dst_stride_bytes;
} rift_b200_gather_field;
int rift_b200_gather_fields(const rift_b200_gather_field* fields, int n_fields, const long long* idx, int bs, void* stream);

/* ---- candidate-rollout evaluator (traj_evaluator.py:115-475 TrajEvaluator.get_grpo_advantage, track_propogate.py:160-780
 * TrackPropagate, reward_model.py:34-50 DenseRewardModel, kinematic_bicycle_model.py:33-61).  All arrays are DEVICE pointers
 * unless named *_host.  G = R * M candidates; a candidate is scored over its first 40 frames and rolled out for 80.
 *   ref_line_info   trajectory [R, M, T_in >= 40, 6] (x, y, cos, sin, vx, vy); reference lines packed ref_pos [sum n_r, 2],
 *                   ref_angle [sum n_r], ref_offsets [R + 1]; out delta_dis / delta_angle [G, 40]
 *   center_rollout  state6_host = {origin x, y, heading (rad, right-handed), speed, width, length}; pid_buf [2, pid_slots, 20],
 *                   pid_ptr / pid_len [2, pid_slots] int32 persist between calls (zero-initialise once); out center [G, 80, 2],
 *                   angle / speed (smoothed) / acc / yaw_rate / yaw_acc [G, 80], vertices [G, 80, 4, 2]
 *   other_rollout   float64 like the reference: location [N, 3], heading_deg [N], speed [N], control [N, 3] = steer, throttle,
 *                   brake; extent [N, 2]; out vertices [N, n_frames, 4, 2] in the right-handed frame
 *   returns         off_road_mask [H, W] uint8 (1 = off road; NULL = no raster) with map_pose4_host = {origin x, y, angle,
 *                   metres per pixel}; out returns [G] float64; collision_out / offroad_out [G, 80] uint8 (may be NULL) */
int rift_b200_eval_ref_line_info(const float* trajectory, int R, int M, int T_in, const float* ref_pos, const float* ref_angle,
                                 const int* ref_offsets, float* delta_dis, float* delta_angle, void* stream);
int rift_b200_eval_center_rollout(const float* trajectory, int G, int T_in, const float* state6_host, float dt, float* pid_buf,
                                  int* pid_ptr, int* pid_len, int pid_slots, float* center, float* angle, float* speed, float* acc,
                                  float* yaw_rate, float* yaw_acc, float* vertices, void* stream);
int rift_b200_eval_other_rollout(const double* location, const double* heading_deg, const double* speed, const double* control,
                                 const double* extent, int N, int n_frames, int near_lane_change, double inflation, double* vertices,
                                 void* stream);
int rift_b200_eval_returns(const float* delta_dis, const float* delta_angle, const float* speed, const float* acc, const float* yaw_rate,
                           const float* yaw_acc, const float* center, const float* vertices, const double* other_vertices, int N,
                           int other_frames, const uint8_t* off_road_mask, int H, int W, const double* map_pose4_host, int G,
                           double gamma, double* returns, uint8_t* collision_out, uint8_t* offroad_out, void* stream);

/* ---- primitive operators exported for the kernel-level parity tests (tests/test_ops_gpu.py) ---- */
int rift_b200_op_add_inplace(float* dst, const float* src, long long n, void* stream);      /* dst += src */
int rift_b200_op_linear(const float* x, int rows, int K, const float* w, const float* bias, int N, int act,
                        const float* res, float* y, int simt, void* stream);
size_t rift_b200_op_linear_tc_scratch_bytes(int rows, int N, int K);
int rift_b200_op_linear_tc(const float* x, int rows, int K, const float* w, const float* bias, int N, int act,
                           const float* res, float* y, void* scratch, size_t scratch_bytes, int resplit, void* stream);
/* every output form of the tensor-core epilogue: y = act(x w^T + bias) + res + beta * y (y may be NULL when planes are
 * requested), preact (may be NULL) = value before `act`, out_hi / out_lo (may be NULL) = split-bf16 planes [rows, ceil64(N)] */
int rift_b200_op_linear_tc_full(const float* x, int rows, int K, const float* w, const float* bias, int N, int act,
                                const float* res, float* y, float beta, float* preact, void* out_hi, void* out_lo, void* scratch,
                                size_t scratch_bytes, void* stream);
/* fused pre-LN MLP sub-block (one cluster kernel: LayerNorm -> fc1 + act -> fc2 + residual; layers/transformer.py:83-94):
 * y = x + fc2(act(fc1(LN(x)))); optional outputs (may be NULL): LayerNorm mean / rstd [rows], LN(x) planes, fc1
 * pre-activation [rows, Hd], hidden planes.  w1 [Hd, D], w2 [D, Hd]; resplit != 0 re-splits the weights into scratch */
size_t rift_b200_op_fused_mlp_scratch_bytes(int D, int Hd);
int rift_b200_op_fused_mlp(const float* x, int rows, int D, int Hd, int act, const float* ln_g, const float* ln_b,
                           const float* w1, const float* b1, const float* w2, const float* b2, float* y, float* mean,
                           float* rstd, void* t2_hi, void* t2_lo, float* hpre, void* hm_hi, void* hm_lo, void* scratch,
                           size_t scratch_bytes, int resplit, void* stream);
/* weight-gradient path of the tcgen05 GEMM: dW[N, K] += dY[rows, N]^T X[rows, K], db[N] += column sums of dY (db may be NULL);
 * split-K parts are added with red.global.add, the bias gradient comes from a second tensor-core accumulator */
size_t rift_b200_op_wgrad_tc_scratch_bytes(int rows, int N, int K);
int rift_b200_op_wgrad_tc(const float* dY, const float* X, int rows, int N, int K, float* dW, float* db, int splits,
                          void* scratch, size_t scratch_bytes, void* stream);
/* n <= 12 weight-gradient products (each as rift_b200_op_wgrad_tc) in ONE grouped persistent launch; host arrays of n entries;
 * every product needs N >= 64, K > 64, K % 4 == 0, rows >= 64; db[i] may be NULL */
int rift_b200_op_wgrad_group(int n, const float* const* dY, const float* const* X, const int* rows, const int* N, const int* K,
                             float* const* dW, float* const* db, void* scratch, size_t scratch_bytes, void* stream);
int rift_b200_op_gemm(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk,
                      float* C, long long ldc, int M, int N, int K, float beta, int split_k, float* split_ws,
                      int simt, void* stream);
int rift_b200_op_layernorm(const float* x, int rows, int C, const float* gamma, const float* beta, int relu, float* y,
                           float* mean, float* rstd, void* stream);
/* scratch: 592 * 2 * C floats (needed when dgamma / dbeta are requested) */
int rift_b200_op_layernorm_bwd(const float* x, const float* dy, int rows, int C, const float* gamma, const float* mean,
                               const float* rstd, const float* y_relu, float* dx, float* dgamma, float* dbeta,
                               float* scratch, void* stream);
int rift_b200_op_attention(const float* qkv, int B, int S, int H, int hd, const uint8_t* key_padding, float* out,
                           void* stream);
int rift_b200_op_nat_attention(const float* qkv, int n_seq, int L, int heads, int hd, int ksize, const float* rpb,
                               float* out, void* stream);
/* forward (writes out [B,S,H*hd] and lse [B,H,S]) followed by the backward: dqkv [B,S,3*H*hd] = d(sum(out * d_out)) / d qkv */
int rift_b200_op_attention_bwd(const float* qkv, const float* d_out, int B, int S, int H, int hd, const uint8_t* key_padding,
                               float* out, float* lse, float* dqkv, void* stream);
/* drpb_partial (may be NULL): [n_seq * heads][2 * ksize - 1] per-(sequence, head) bias gradients (column-sum them) */
int rift_b200_op_nat_attention_bwd(const float* qkv, const float* d_out, int n_seq, int L, int heads, int hd, int ksize,
                                   const float* rpb, float* dqkv, float* drpb_partial, void* stream);
/* dy *= act'(ref) in place (act 1: ReLU, ref = output; act 2: GELU, ref = pre-activation) */
int rift_b200_op_act_bwd(const float* ref, float* dy, long long n, int act, void* stream);
/* out[c] (+)= sum_r x[r, c]; scratch: 148 * C floats */
int rift_b200_op_colsum(const float* x, int rows, int C, float* out, int accumulate, float* scratch, void* stream);
int rift_b200_op_masked_maxpool(const float* x, const uint8_t* mask, int groups, int n, int C, float* out, int* argmax,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIFT_B200_H */
