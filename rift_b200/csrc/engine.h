// Host-side executor of the Pluto policy: parameter views over the flat arena, the forward
// schedule, the saved-activation tape and the backward schedule.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rift_b200.h"
#include "common.cuh"
#include "gemm_tc.h"
#include "ops.h"

namespace rift {

struct Lin {            // nn.Linear / flattened Conv1d: W [N, K] row-major, b [N]
    const float* W = nullptr; const float* b = nullptr;
    float* dW = nullptr; float* db = nullptr;
    int N = 0, K = 0;
    long long ldw = 0;      // row pitch of W (== K unless this is a column slice of a wider matrix)
    bool train = false;
    int tc = -1, tc_n0 = 0, tc_k0 = 0;      // index into the engine's pre-split weight planes (+ slice origin)
    int tcT = -1;                           // index of the transposed planes (W^T) used by the backward
};
struct Norm {           // LayerNorm
    const float* g = nullptr; const float* b = nullptr;
    float* dg = nullptr; float* db = nullptr;
    int C = 0;
    bool train = false;
};
struct BNorm {          // BatchNorm1d (eval: running statistics)
    Norm affine;
    const float* mean = nullptr; const float* var = nullptr;
};
struct Vec {            // bare nn.Parameter / nn.Embedding table
    const float* p = nullptr; float* d = nullptr; long long n = 0; bool train = false;
};
struct MLPLayerP { Lin l0; Norm n; Lin l3; };
struct FourierP { Vec freqs; int d = 0; std::vector<MLPLayerP> mlps; Norm out_n; Lin out_l; };
struct PointsEncP { Lin f0; BNorm fbn; Lin f3; Lin s0; BNorm sbn; Lin s3; };
struct MHAP { Lin in; Lin out; };                 // in = packed in_proj (3D x D)
struct NatBlockP { Norm n1; Vec rpb; Lin qkv, proj; Norm n2; Lin fc1, fc2; };
struct NatLevelP { std::vector<NatBlockP> blocks; Lin down; Norm down_n; bool has_down = false; int dim = 0, heads = 0, ksize = 0; };
struct NatEncP { Lin embed; std::vector<NatLevelP> levels; Norm norms[3]; Lin lateral[3]; Lin fpn; };
struct StateAttnP { Vec pos_embed, query; Lin lin[8]; MHAP attn; };
struct EncBlockP { Norm n1; MHAP attn; Norm n2; Lin fc1, fc2; };
struct DecBlockP { MHAP r2r, m2m, cross; Lin ffn0, ffn3; Norm n1, n2, n3, n4; };

struct Model {
    FourierP pos_emb;
    NatEncP hist;
    StateAttnP ego;
    Vec agent_type_emb;
    PointsEncP poly_enc;
    FourierP speed_emb;
    Vec map_type_emb, map_route_emb, map_tl_emb, map_unknown_emb;
    std::vector<EncBlockP> enc;
    Norm final_norm;
    MLPLayerP pred_loc, pred_yaw, pred_vel;
    Vec m_emb, m_pos;
    std::vector<DecBlockP> dec;
    FourierP r_pos_emb;
    PointsEncP r_enc;
    Lin q_proj, cat_x_proj;
    MLPLayerP loc_head, yaw_head, vel_head, pi_head;
    Lin hidden0, hidden2;
    MLPLayerP ref_free;
    bool full = false;      // some trainable parameter lies outside planning_decoder.pi_head
};

// An activation [rows, C]: fp32 (row pitch ld) and / or split-bf16 planes for the tcgen05 GEMM
struct Act {
    float* f = nullptr; long long ld = 0;
    Planes p;
    int rows = 0, C = 0;
};

struct Ctx {            // per-call state: stream + bump allocator over the caller's workspace
    cudaStream_t st = nullptr;
    char* base = nullptr;
    size_t cap = 0, off = 0;
    bool dry = false;      // measure only: advance the allocator, launch nothing
    bool simt = true;      // force the exact-fp32 GEMM everywhere
    bool save = false;     // keep what backward needs
    bool full = false;     // save && trainable parameters outside pi_head: keep every activation (fp32)
    bool tc_bwd = false;   // backward GEMMs on the tcgen05 path (false: exact-fp32 SIMT)
    bool used_fused = false;   // forward: some sub-block ran as a fused kernel (its tape holds operand planes only, no fp32 copies)
    const std::vector<TcWeight>* tcw = nullptr;
    // fork / join streams (null: everything stays on `st`).  `side` carries parameter-gradient work of the
    // backward (weight-gradient GEMMs, bias / LayerNorm-parameter reductions), `br` whole independent branches
    // (map / reference-line encoders in the forward, parameter-only sub-graphs in the backward).
    cudaStream_t side = nullptr, br = nullptr;
    cudaEvent_t* events = nullptr; int n_events = 0; int ev_next = 0;
    cudaEvent_t next_event() { cudaEvent_t e = events[ev_next]; ev_next = (ev_next + 1) % n_events; return e; }
    template <class T> T* alloc(size_t n) {
        const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        const size_t at = off;
        off += bytes;
        if (dry) return reinterpret_cast<T*>((char*)nullptr + 256 + at);   // non-null placeholder
        if (off > cap) return nullptr;
        return reinterpret_cast<T*>(base + at);
    }
    bool planes_for(int C) const { return !simt && C >= 32; }
    // weight-gradient products waiting for a grouped launch (engine_ops.h::flush_wgrads); `pending_stream` produced their operands
    // (two queues: products whose operands were produced on the caller's stream / on the branch stream - a queue is ordered
    //  after ONE stream, and alternating between the two would cut the groups short)
    WgradItem pending[2][WGRAD_GROUP_MAX]; int n_pending[2] = {0, 0}; cudaStream_t pending_stream[2] = {nullptr, nullptr};
    // fp32 gradient buffers that the main chain keeps updating in place (residual-stream gradients): kernels moved to
    // another stream must not read them.  Everything else the backward allocates is written once (bump allocator).
    const void* inplace_bufs[12] = {}; int n_inplace = 0;
    void mark_inplace(const void* p) { if (p && n_inplace < 12) inplace_bufs[n_inplace++] = p; }
    bool updated_inplace(const void* p) const {
        if (n_inplace >= 12) return true;                 // registry overflow: assume the worst
        for (int i = 0; i < n_inplace; ++i) if (inplace_bufs[i] == p) return true;
        return false;
    }
};

// ---------------------------------------------------------------- saved activations (forward -> backward)
struct LNSave { const float* x = nullptr; float* mean = nullptr; float* rstd = nullptr; int rows = 0; };
struct MlpTape { Act x; Act h; LNSave ln; Act a; };                       // MLPLayer: x -> l0 -> h -> LN+ReLU -> a -> l3
struct FourierTape {
    const float* x = nullptr; int rows = 0;
    struct Dim { Act feat; Act h; LNSave ln; Act hn; };
    std::vector<Dim> dims;
    float* acc = nullptr; LNSave ln_o; Act on;
};
struct PointsTape {
    Act F; int groups = 0, n = 0, Cin = 0; const uint8_t* mask = nullptr;
    float *sc1 = nullptr, *sh1 = nullptr, *sc2 = nullptr, *sh2 = nullptr;
    float* h1pre = nullptr; Act h1; Act f; Act pooled; int* arg1 = nullptr; float* gp = nullptr;
    float* h2pre = nullptr; Act h2; Act o; int* arg2 = nullptr;
};
struct NatBlockTape { float* x = nullptr; LNSave ln1; Act t1; float* qkv = nullptr; Act att; float* x1 = nullptr; LNSave ln2; Act t2;
                      float* hpre = nullptr; Act hm; };
struct NatTape {
    int NA = 0; int Ls[3] = {0, 0, 0};
    Act col0;
    std::vector<NatBlockTape> blocks;       // 6, in forward order
    float* xlev[3] = {nullptr, nullptr, nullptr};   // level output (input of norm_i and of the downsample)
    LNSave ln_lev[3]; Act colL[3];
    Act colD[2]; float* xd[2] = {nullptr, nullptr}; LNSave ln_down[2];
    Act colF;
};
struct EgoTape { float* toks = nullptr; Act toks_a; float* kv = nullptr; float* qv = nullptr; float* lse = nullptr; Act eo; };
struct EncBlockTape { float* X = nullptr; LNSave ln1; Act t1; float* qkv = nullptr; float* lse = nullptr; Act att; float* X1 = nullptr;
                      LNSave ln2; Act t2; float* hpre = nullptr; Act hm; };
struct DecBlockTape {
    float* q = nullptr; LNSave ln1; Act t1; float* qkv1 = nullptr; float* lse1 = nullptr; Act a1; float* q1 = nullptr;
    LNSave ln2; Act t2; Act t2p; float* qkv2 = nullptr; float* lse2 = nullptr; Act a2; float* q2 = nullptr;
    LNSave ln3; Act t3; float* qc = nullptr; float* kvc = nullptr; float* lse3 = nullptr; Act a3; float* q3 = nullptr;
    LNSave ln4; Act t4; Act hm; float* hpre4 = nullptr;     // hpre4: fc1 pre-activation (fused forward), else null
};
struct Tape {
    bool valid = false, full = false, fused = false;
    int bs = 0, A = 0, Mp = 0, P = 0, R = 0, Pr = 0, S = 0;
    uint8_t *agent_any = nullptr, *key_pad = nullptr, *r_pad = nullptr;
    const uint8_t* r_pad_r2r = nullptr; int r2r_mod = 0, r2r_off = 0;     // r2r key padding: mask rows of the WHOLE batch (quirk)
    NatTape nat; EgoTape ego;
    PointsTape poly; FourierTape speed;
    float* pos = nullptr; FourierTape pos_emb;
    std::vector<EncBlockTape> enc;
    float* Xlast = nullptr; LNSave ln_final; Act Xn;
    float* rpos = nullptr; PointsTape renc; FourierTape rpos_emb; Act r_emb;
    std::vector<DecBlockTape> dec;
    float* qlast = nullptr; Act qlast_a;
    Act xego_rows;            // Xn[:, 0] viewed (bs rows, pitch S*D)
    Act qf; MlpTape pi;
};

struct ParamRef { long long offset, numel; bool trainable; };

}  // namespace rift

struct rift_b200_engine {
    rift_b200_model_config cfg;
    std::unordered_map<std::string, rift::ParamRef> table;
    float* params = nullptr; float* grads = nullptr; long long numel = 0;
    long long train_lo = 0, train_hi = 0;     // element span covering every trainable entry
    rift::Model m;
    bool bound = false;
    rift::Tape tape;
    // tcgen05 path: pre-split bf16 weight planes (caller-owned memory) + TMA descriptors
    std::vector<rift::TcWeight> tcw;
    void* wcache = nullptr; size_t wcache_bytes = 0;
    void* jobs_all = nullptr; void* jobs_train = nullptr;
    int n_jobs_all = 0, n_jobs_train = 0; long long total_all = 0, total_train = 0;
    bool dirty_all = true, dirty_train = true;
    size_t weight_cache_bytes() const;
    int bind_weight_cache(void* cache, size_t bytes);
    int refresh_weights(cudaStream_t st);
    size_t fwd_ws_end = 0;                    // workspace offset where backward scratch may start
    // fork / join streams of the schedules (created on first use; RIFT_B200_STREAMS=0 keeps one stream)
    cudaStream_t s_side = nullptr, s_br = nullptr;
    std::vector<cudaEvent_t> events;
    int streams_state = 0;                    // 0 = not initialised, 1 = on, -1 = off
    int attach_streams(rift::Ctx& c);
    ~rift_b200_engine();

    int build_model();
    int forward(const rift_b200_batch& bt, const rift_b200_outputs& out, rift::Ctx& c);
    int backward(const rift_b200_batch& bt, const float* dlogits, rift::Ctx& c);
    int backward_impl(const rift_b200_batch& bt, const float* dlogits, rift::Ctx& c);
};
