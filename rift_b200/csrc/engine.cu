// Forward / backward schedule of the Pluto trajectory policy on the rift_b200 kernels.
//
// Mirrors PlanningModel.forward (rift/cbv/planning/pluto/model/pluto_model.py:122-225) in the
// deterministic parity mode of SURVEY 8(c): dropout / drop-path identity, BatchNorm running
// statistics, no state-token dropping.  Ragged structure (x[mask] gathers in the reference) is
// handled densely: every padded row is computed and then masked, which is equivalent because no
// op in the model mixes rows of different agents / polylines except through masked attention and
// masked max-pooling.
#include "engine.h"

#include <math.h>

namespace rift {

#define TRY(x)                   \
    do {                         \
        int _r = (x);            \
        if (_r != 0) return _r;  \
    } while (0)

#define ALLOC(var, T, n)                                                           \
    T* var = c.alloc<T>((size_t)(n));                                              \
    if (!var) {                                                                    \
        set_last_error("workspace too small (see rift_b200_workspace_bytes)");     \
        return -1;                                                                 \
    }

static const int NAT_HEADS_[3] = {2, 4, 8};
static const int NAT_K_[3] = {3, 3, 5};
static const int NFREQ = 64, FIN = 129, FLD = 132;    // Fourier feature width, padded row stride
static const int PE_H1 = 128, PE_H2 = 256;

// ---------------------------------------------------------------------------------- ops
static int gemm(Ctx& c, const GemmArgs& a) {
    if (c.dry) return 0;
    return launch_gemm_simt(a, c.st);
}

// GEMM against a (possibly sliced) weight matrix: tcgen05 path when the shape qualifies and the
// weight has pre-split planes, exact-fp32 SIMT otherwise (tiny K / N == 1 / misaligned views).
static int gemm_w(Ctx& c, GemmArgs& a, const Lin& L, long long ldw) {
    a.B = L.W; a.sbn = ldw; a.sbk = 1; a.N = L.N; a.K = L.K;
    // shape-only decision so that the dry (sizing) pass and the real pass allocate identically
    const bool tc = !c.simt && L.tc >= 0 && c.tcw && a.sak == 1 && gemm_tc_shape_ok(a.M, a.N, a.K) && (a.ldc % 4) == 0 &&
                    (a.ldres % 4) == 0 && (a.ldpre % 4) == 0;
    if (tc) {
        const int Kp = tc_pitch(a.K);
        ALLOC(a_hi, uint16_t, (size_t)a.M * Kp);
        ALLOC(a_lo, uint16_t, (size_t)a.M * Kp);
        if (c.dry) return 0;
        if (gemm_tc_eligible(a)) {
            TRY(launch_pack_split(a.A, a.sam, a.M, a.K, Kp, a_hi, a_lo, c.st));
            return launch_gemm_tc(a, a_hi, a_lo, Kp, (*c.tcw)[L.tc], L.tc_n0, L.tc_k0, c.st);
        }
    }
    if (c.dry) return 0;
    return launch_gemm_simt(a, c.st);
}

// Y = act((X W^T [+pre]) [*colscale] + b) [+ res]
static int linear_ld(Ctx& c, const float* X, long long ldx, int M, const Lin& L, long long ldw, float* Y, long long ldy,
                     int act = ACT_NONE, const float* res = nullptr, long long ldres = 0, int res_div = 1,
                     float* preact = nullptr) {
    GemmArgs a;
    a.A = X; a.sam = ldx; a.sak = 1;
    a.C = Y; a.ldc = ldy; a.M = M;
    a.bias = L.b; a.act = act; a.res = res; a.ldres = ldres; a.res_div = res_div; a.preact = preact;
    return gemm_w(c, a, L, ldw);
}
static int linear(Ctx& c, const float* X, long long ldx, int M, const Lin& L, float* Y, long long ldy, int act = ACT_NONE,
                  const float* res = nullptr, long long ldres = 0, int res_div = 1, float* preact = nullptr) {
    return linear_ld(c, X, ldx, M, L, L.K, Y, ldy, act, res, ldres, res_div, preact);
}

// slice of a Linear: output rows [n0, n0+n), input columns [k0, k0+k).  NOTE: a slice keeps the parent's
// row pitch; callers pass it as `ldw` to linear_ld.
static Lin slice(const Lin& L, int n0, int n, int k0, int k, bool with_bias) {
    Lin s;
    s.W = L.W + (long long)n0 * L.K + k0;
    s.b = (with_bias && L.b) ? L.b + n0 : nullptr;
    s.dW = L.dW ? L.dW + (long long)n0 * L.K + k0 : nullptr;
    s.db = (with_bias && L.db) ? L.db + n0 : nullptr;
    s.N = n; s.K = k; s.train = L.train;
    s.tc = L.tc; s.tc_n0 = L.tc_n0 + n0; s.tc_k0 = L.tc_k0 + k0;
    return s;
}

static int layernorm(Ctx& c, const float* x, int rows, const Norm& n, float* y, int relu = 0, float* mean = nullptr,
                     float* rstd = nullptr, const float* add_rowmod = nullptr, int rowmod = 0, float* y2 = nullptr) {
    if (c.dry) return 0;
    return launch_layernorm(x, n.C, rows, n.C, n.g, n.b, y, n.C, relu, add_rowmod, rowmod, y2, mean, rstd, c.st);
}

// MLPLayer: Linear -> LayerNorm -> ReLU -> Linear  (layers/mlp_layer.py:4-16)
static int mlp_layer(Ctx& c, const float* X, long long ldx, int M, const MLPLayerP& p, float* Y, long long ldy,
                     float** h_out = nullptr, float** a_out = nullptr, float** mean_out = nullptr, float** rstd_out = nullptr) {
    ALLOC(h, float, (size_t)M * p.l0.N);
    ALLOC(a, float, (size_t)M * p.l0.N);
    float* mean = nullptr; float* rstd = nullptr;
    if (mean_out) {
        mean = c.alloc<float>(M); rstd = c.alloc<float>(M);
        if (!mean || !rstd) { set_last_error("workspace too small"); return -1; }
    }
    TRY(linear(c, X, ldx, M, p.l0, h, p.l0.N));
    TRY(layernorm(c, h, M, p.n, a, 1, mean, rstd));
    TRY(linear(c, a, p.l0.N, M, p.l3, Y, ldy));
    if (h_out) *h_out = h;
    if (a_out) *a_out = a;
    if (mean_out) { *mean_out = mean; *rstd_out = rstd; }
    return 0;
}

// FourierEmbedding (layers/fourier_embedding.py:45-55): out = to_out(sum_i mlp_i(feat_i)) [+ res]
static int fourier(Ctx& c, const float* x, int rows, const FourierP& p, float* out, const float* res) {
    const int D = p.out_l.N;
    ALLOC(acc, float, (size_t)rows * D);
    for (int i = 0; i < p.d; ++i) {
        ALLOC(feat, float, (size_t)rows * FLD);
        ALLOC(h, float, (size_t)rows * D);
        ALLOC(hn, float, (size_t)rows * D);
        if (!c.dry) TRY(launch_fourier_features(x, rows, p.d, i, p.freqs.p, NFREQ, feat, FLD, c.st));
        TRY(linear(c, feat, FLD, rows, p.mlps[i].l0, h, D));
        TRY(layernorm(c, h, rows, p.mlps[i].n, hn, 1));
        TRY(linear(c, hn, D, rows, p.mlps[i].l3, acc, D, ACT_NONE, i > 0 ? acc : nullptr, D, 1));
    }
    ALLOC(on, float, (size_t)rows * D);
    TRY(layernorm(c, acc, rows, p.out_n, on, 1));
    TRY(linear(c, on, D, rows, p.out_l, out, D, ACT_NONE, res, D, 1));
    return 0;
}

// PointsEncoder (layers/embedding.py:271-296) over `groups` polylines of `n` points, C_in channels
static int points_encoder(Ctx& c, const float* F, int groups, int n, int Cin, const uint8_t* mask, const PointsEncP& p,
                          float* out) {
    const int rows = groups * n;
    const int Cout = p.s3.N;
    ALLOC(sc1, float, PE_H1); ALLOC(sh1, float, PE_H1); ALLOC(sc2, float, PE_H2); ALLOC(sh2, float, PE_H2);
    ALLOC(h1, float, (size_t)rows * PE_H1);
    ALLOC(f, float, (size_t)rows * PE_H2);
    ALLOC(pooled, float, (size_t)groups * PE_H2);
    ALLOC(arg1, int, (size_t)groups * PE_H2);
    ALLOC(gp, float, (size_t)groups * PE_H2);
    ALLOC(h2, float, (size_t)rows * PE_H2);
    ALLOC(o, float, (size_t)rows * Cout);
    ALLOC(arg2, int, (size_t)groups * Cout);
    if (!c.dry) {
        TRY(launch_bn_fold(p.fbn.affine.g, p.fbn.affine.b, p.fbn.mean, p.fbn.var, p.f0.b, PE_H1, sc1, sh1, c.st));
        TRY(launch_bn_fold(p.sbn.affine.g, p.sbn.affine.b, p.sbn.mean, p.sbn.var, p.s0.b, PE_H2, sc2, sh2, c.st));
    }
    {   // first_mlp: Linear(C,128) + BN + ReLU + Linear(128,256)
        GemmArgs a;
        a.A = F; a.sam = Cin; a.C = h1; a.ldc = PE_H1; a.M = rows;
        a.colscale = sc1; a.bias = sh1; a.act = ACT_RELU;
        TRY(gemm_w(c, a, p.f0, Cin));
        TRY(linear(c, h1, PE_H1, rows, p.f3, f, PE_H2));
    }
    if (!c.dry) TRY(launch_masked_maxpool(f, mask, groups, n, PE_H2, pooled, arg1, c.st));
    {   // second_mlp on cat[feat, pooled]: the pooled half of the weight acts once per polyline
        Lin wb = slice(p.s0, 0, PE_H2, PE_H2, PE_H2, false);
        TRY(linear_ld(c, pooled, PE_H2, groups, wb, 2 * PE_H2, gp, PE_H2));
        GemmArgs a;
        a.A = f; a.sam = PE_H2; a.C = h2; a.ldc = PE_H2; a.M = rows;
        a.pre = gp; a.ldpre = PE_H2; a.pre_div = n; a.colscale = sc2; a.bias = sh2; a.act = ACT_RELU;
        TRY(gemm_w(c, a, slice(p.s0, 0, PE_H2, 0, PE_H2, false), 2 * PE_H2));
        TRY(linear(c, h2, PE_H2, rows, p.s3, o, Cout));
    }
    if (!c.dry) TRY(launch_masked_maxpool(o, mask, groups, n, Cout, out, arg2, c.st));
    return 0;
}

// nn.MultiheadAttention self-attention over rows laid out (B, S, D) contiguous
static int mha_self(Ctx& c, const float* x, int B, int S, int D, int H, const MHAP& p, const uint8_t* kpm, float* out,
                    const float* res) {
    const int rows = B * S;
    ALLOC(qkv, float, (size_t)rows * 3 * D);
    ALLOC(att, float, (size_t)rows * D);
    TRY(linear(c, x, D, rows, p.in, qkv, 3 * D));
    if (!c.dry) {
        AttnArgs a;
        a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = att;
        a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
        a.B = B; a.H = H; a.Sq = S; a.Sk = S; a.hd = D / H;
        a.q_outer = S; a.k_outer = S;
        a.kpm = kpm; a.kpm_div = 1; a.scale = 1.f / sqrtf((float)(D / H));
        TRY(launch_attention(a, c.st));
    }
    TRY(linear(c, att, D, rows, p.out, out, D, ACT_NONE, res, D, 1));
    return 0;
}

}  // namespace rift

using namespace rift;

// =====================================================================================
// parameter views
// =====================================================================================
namespace {
struct Binder {
    rift_b200_engine* e;
    int err = 0;
    const ParamRef* find(const std::string& n) {
        auto it = e->table.find(n);
        if (it == e->table.end()) {
            if (!err) set_last_error("parameter missing from the arena table: " + n);
            err = -1;
            return nullptr;
        }
        return &it->second;
    }
    Vec vec(const std::string& n) {
        Vec v;
        if (auto r = find(n)) {
            v.p = e->params + r->offset; v.n = r->numel; v.train = r->trainable && e->grads;
            v.d = v.train ? e->grads + r->offset : nullptr;
        }
        return v;
    }
    Lin lin(const std::string& p, int N, int K, bool bias = true, const char* wname = ".weight", const char* bname = ".bias") {
        Lin l; l.N = N; l.K = K;
        if (auto r = find(p + wname)) {
            if (r->numel != (long long)N * K) { set_last_error("shape mismatch for " + p + wname); err = -1; }
            l.W = e->params + r->offset; l.train = r->trainable && e->grads;
            l.dW = l.train ? e->grads + r->offset : nullptr;
            if (N >= 16 && K >= 32) {          // candidate for the tcgen05 path: gets pre-split bf16 planes
                TcWeight w;
                w.src = l.W; w.ld_src = K; w.N = N; w.K = K; w.Kp = (K + 63) / 64 * 64; w.trainable = r->trainable;
                l.tc = (int)e->tcw.size();
                e->tcw.push_back(w);
            }
        }
        if (bias) if (auto r = find(p + bname)) {
            l.b = e->params + r->offset;
            l.db = (r->trainable && e->grads) ? e->grads + r->offset : nullptr;
        }
        return l;
    }
    Norm norm(const std::string& p, int C) {
        Norm n; n.C = C;
        if (auto r = find(p + ".weight")) {
            if (r->numel != C) { set_last_error("shape mismatch for " + p + ".weight"); err = -1; }
            n.g = e->params + r->offset; n.train = r->trainable && e->grads;
            n.dg = n.train ? e->grads + r->offset : nullptr;
        }
        if (auto r = find(p + ".bias")) { n.b = e->params + r->offset; n.db = n.train ? e->grads + r->offset : nullptr; }
        return n;
    }
    BNorm bnorm(const std::string& p, int C) {
        BNorm b; b.affine = norm(p, C);
        if (auto r = find(p + ".running_mean")) b.mean = e->params + r->offset;
        if (auto r = find(p + ".running_var")) b.var = e->params + r->offset;
        return b;
    }
    MLPLayerP mlp(const std::string& p, int cin, int hid, int cout) {
        MLPLayerP m; m.l0 = lin(p + ".mlp.0", hid, cin); m.n = norm(p + ".mlp.1", hid); m.l3 = lin(p + ".mlp.3", cout, hid);
        return m;
    }
    FourierP fourier(const std::string& p, int d, int D) {
        FourierP f; f.d = d; f.freqs = vec(p + ".freqs.weight");
        for (int i = 0; i < d; ++i) {
            MLPLayerP m;
            const std::string q = p + ".mlps." + std::to_string(i);
            m.l0 = lin(q + ".0", D, FIN); m.n = norm(q + ".1", D); m.l3 = lin(q + ".3", D, D);
            f.mlps.push_back(m);
        }
        f.out_n = norm(p + ".to_out.0", D); f.out_l = lin(p + ".to_out.2", D, D);
        return f;
    }
    PointsEncP points(const std::string& p, int cin, int cout) {
        PointsEncP e2;
        e2.f0 = lin(p + ".first_mlp.0", PE_H1, cin); e2.fbn = bnorm(p + ".first_mlp.1", PE_H1);
        e2.f3 = lin(p + ".first_mlp.3", PE_H2, PE_H1);
        e2.s0 = lin(p + ".second_mlp.0", PE_H2, 2 * PE_H2); e2.sbn = bnorm(p + ".second_mlp.1", PE_H2);
        e2.s3 = lin(p + ".second_mlp.3", cout, PE_H2);
        return e2;
    }
    MHAP mha(const std::string& p, int D) {
        MHAP m; m.in = lin(p, 3 * D, D, true, ".in_proj_weight", ".in_proj_bias"); m.out = lin(p + ".out_proj", D, D);
        return m;
    }
};
}  // namespace

// ---- pre-split weight planes for the tcgen05 path -------------------------------------------
// cache layout: [plane hi | plane lo] per weight (256 B aligned), then the two split-job tables.
size_t rift_b200_engine::weight_cache_bytes() const {
    size_t off = 0;
    for (const TcWeight& w : tcw) off += 2 * (((size_t)w.N * w.Kp * 2 + 255) & ~(size_t)255);
    off += 2 * ((tcw.size() * split_job_bytes() + 255) & ~(size_t)255);
    return off + 256;
}

int rift_b200_engine::bind_weight_cache(void* cache, size_t bytes) {
    RIFT_REQUIRE(cache != nullptr && bytes >= weight_cache_bytes(), "bind_weight_cache: buffer too small");
    RIFT_REQUIRE((reinterpret_cast<uintptr_t>(cache) & 255) == 0, "bind_weight_cache: buffer must be 256-byte aligned");
    wcache = cache; wcache_bytes = bytes;
    char* p = static_cast<char*>(cache);
    std::vector<char> all(tcw.size() * split_job_bytes()), tr(tcw.size() * split_job_bytes());
    n_jobs_all = n_jobs_train = 0; total_all = total_train = 0;
    for (TcWeight& w : tcw) {
        const size_t plane = ((size_t)w.N * w.Kp * 2 + 255) & ~(size_t)255;
        w.hi = p; w.lo = p + plane; p += 2 * plane;
        TRY(make_weight_tensor_map(w.tm_hi, w.hi, w.N, w.Kp));
        TRY(make_weight_tensor_map(w.tm_lo, w.lo, w.N, w.Kp));
        fill_split_job(all.data() + n_jobs_all * split_job_bytes(), w.src, w.ld_src, w.N, w.K, w.Kp, w.hi, w.lo, total_all);
        ++n_jobs_all; total_all += (long long)w.N * w.Kp;
        if (w.trainable) {
            fill_split_job(tr.data() + n_jobs_train * split_job_bytes(), w.src, w.ld_src, w.N, w.K, w.Kp, w.hi, w.lo, total_train);
            ++n_jobs_train; total_train += (long long)w.N * w.Kp;
        }
    }
    const size_t tbl = (tcw.size() * split_job_bytes() + 255) & ~(size_t)255;
    jobs_all = p; jobs_train = p + tbl;
    if (n_jobs_all) RIFT_CUDA_OK(cudaMemcpy(jobs_all, all.data(), n_jobs_all * split_job_bytes(), cudaMemcpyHostToDevice));
    if (n_jobs_train) RIFT_CUDA_OK(cudaMemcpy(jobs_train, tr.data(), n_jobs_train * split_job_bytes(), cudaMemcpyHostToDevice));
    dirty_all = true; dirty_train = true;
    return 0;
}

int rift_b200_engine::refresh_weights(cudaStream_t st) {
    if (!wcache) return 0;
    if (dirty_all) TRY(launch_split_weights(jobs_all, n_jobs_all, total_all, st));
    else if (dirty_train) TRY(launch_split_weights(jobs_train, n_jobs_train, total_train, st));
    dirty_all = dirty_train = false;
    return 0;
}

int rift_b200_engine::build_model() {
    Binder b{this};
    const int D = cfg.dim, T = cfg.future_steps;
    m = Model();
    tcw.clear();
    wcache = nullptr;
    m.pos_emb = b.fourier("pos_emb", 3, D);
    {   // NATSequenceEncoder (layers/embedding.py:8-87)
        const std::string p = "agent_encoder.history_encoder";
        const int dims[3] = {D / 4, D / 2, D};
        m.hist.embed = b.lin(p + ".embed.proj", dims[0], 9 * 3);
        for (int i = 0; i < 3; ++i) {
            NatLevelP lv; lv.dim = dims[i]; lv.heads = NAT_HEADS_[i]; lv.ksize = NAT_K_[i];
            for (int j = 0; j < 2; ++j) {
                const std::string q = p + ".levels." + std::to_string(i) + ".blocks." + std::to_string(j);
                NatBlockP nb;
                nb.n1 = b.norm(q + ".norm1", dims[i]); nb.rpb = b.vec(q + ".attn.rpb");
                nb.qkv = b.lin(q + ".attn.qkv", 3 * dims[i], dims[i]); nb.proj = b.lin(q + ".attn.proj", dims[i], dims[i]);
                nb.n2 = b.norm(q + ".norm2", dims[i]);
                nb.fc1 = b.lin(q + ".mlp.fc1", 3 * dims[i], dims[i]); nb.fc2 = b.lin(q + ".mlp.fc2", dims[i], 3 * dims[i]);
                lv.blocks.push_back(nb);
            }
            if (i < 2) {
                const std::string q = p + ".levels." + std::to_string(i) + ".downsample";
                lv.has_down = true;
                lv.down = b.lin(q + ".reduction", 2 * dims[i], dims[i] * 3, false);
                lv.down_n = b.norm(q + ".norm", 2 * dims[i]);
            }
            m.hist.levels.push_back(lv);
            m.hist.norms[i] = b.norm(p + ".norm" + std::to_string(i), dims[i]);
            m.hist.lateral[i] = b.lin(p + ".lateral_convs." + std::to_string(i), D, dims[i] * 3);
        }
        m.hist.fpn = b.lin(p + ".fpn_conv", D, D * 3);
    }
    {
        const std::string p = "agent_encoder.ego_state_emb";
        m.ego.pos_embed = b.vec(p + ".pos_embed"); m.ego.query = b.vec(p + ".query");
        for (int i = 0; i < cfg.state_channel; ++i) m.ego.lin[i] = b.lin(p + ".linears." + std::to_string(i), D, 1);
        m.ego.attn = b.mha(p + ".attn", D);
    }
    m.agent_type_emb = b.vec("agent_encoder.type_emb.weight");
    m.poly_enc = b.points("map_encoder.polygon_encoder", 10, D);
    m.speed_emb = b.fourier("map_encoder.speed_limit_emb", 1, D);
    m.map_type_emb = b.vec("map_encoder.type_emb.weight");
    m.map_route_emb = b.vec("map_encoder.on_route_emb.weight");
    m.map_tl_emb = b.vec("map_encoder.traffic_light_emb.weight");
    m.map_unknown_emb = b.vec("map_encoder.unknown_speed_emb.weight");
    for (int i = 0; i < cfg.encoder_depth; ++i) {
        const std::string p = "encoder_blocks." + std::to_string(i);
        EncBlockP eb;
        eb.n1 = b.norm(p + ".norm1", D); eb.attn = b.mha(p + ".attn", D); eb.n2 = b.norm(p + ".norm2", D);
        eb.fc1 = b.lin(p + ".mlp.fc1", 4 * D, D); eb.fc2 = b.lin(p + ".mlp.fc2", D, 4 * D);
        m.enc.push_back(eb);
    }
    m.final_norm = b.norm("norm", D);
    m.pred_loc = b.mlp("agent_predictor.loc_predictor", D, 2 * D, 2 * T);
    m.pred_yaw = b.mlp("agent_predictor.yaw_predictor", D, 2 * D, 2 * T);
    m.pred_vel = b.mlp("agent_predictor.vel_predictor", D, 2 * D, 2 * T);
    m.m_emb = b.vec("planning_decoder.m_emb"); m.m_pos = b.vec("planning_decoder.m_pos");
    for (int i = 0; i < cfg.decoder_depth; ++i) {
        const std::string p = "planning_decoder.decoder_blocks." + std::to_string(i);
        DecBlockP db;
        db.r2r = b.mha(p + ".r2r_attn", D); db.m2m = b.mha(p + ".m2m_attn", D); db.cross = b.mha(p + ".cross_attn", D);
        db.ffn0 = b.lin(p + ".ffn.0", 4 * D, D); db.ffn3 = b.lin(p + ".ffn.3", D, 4 * D);
        db.n1 = b.norm(p + ".norm1", D); db.n2 = b.norm(p + ".norm2", D); db.n3 = b.norm(p + ".norm3", D); db.n4 = b.norm(p + ".norm4", D);
        m.dec.push_back(db);
    }
    m.r_pos_emb = b.fourier("planning_decoder.r_pos_emb", 3, D);
    m.r_enc = b.points("planning_decoder.r_encoder", 6, D);
    m.q_proj = b.lin("planning_decoder.q_proj", D, 2 * D);
    m.cat_x_proj = b.lin("planning_decoder.cat_x_proj", D, 2 * D);
    m.loc_head = b.mlp("planning_decoder.loc_head", D, 2 * D, 2 * T);
    m.yaw_head = b.mlp("planning_decoder.yaw_head", D, 2 * D, 2 * T);
    m.vel_head = b.mlp("planning_decoder.vel_head", D, 2 * D, 2 * T);
    m.pi_head = b.mlp("planning_decoder.pi_head", D, D, 1);
    m.hidden0 = b.lin("hidden_proj.0", D, D); m.hidden2 = b.lin("hidden_proj.2", D, D);
    m.ref_free = b.mlp("ref_free_decoder", D, 2 * D, 4 * T);
    if (b.err) return b.err;
    m.any_trainable_outside_pi_head = false;
    for (auto& kv : table)
        if (kv.second.trainable && kv.first.rfind("planning_decoder.pi_head.", 0) != 0 && kv.first.rfind("value_net.", 0) != 0)
            m.any_trainable_outside_pi_head = true;
    return 0;
}

// =====================================================================================
// forward
// =====================================================================================
int rift_b200_engine::forward(const rift_b200_batch& bt, const rift_b200_outputs& out, Ctx& c) {
    const int bs = bt.bs, A = bt.A, Mp = bt.Mp, P = bt.P, R = bt.R, Pr = bt.Pr;
    const int D = cfg.dim, H = cfg.num_heads, Mo = cfg.num_modes, T = cfg.future_steps, Th = cfg.history_steps;
    const int S = A + Mp;
    RIFT_REQUIRE(bs > 0 && A > 0 && R > 0 && Mp >= 0, "forward: empty batch");
    RIFT_REQUIRE(bt.agent_T >= Th, "forward: agent tensors shorter than history_steps");
    RIFT_REQUIRE(D % H == 0 && D / H == 32, "forward: encoder head_dim must be 32");
    pi_tape.valid = false;
    if (!wcache) c.simt = true;            // no pre-split planes bound: exact-fp32 SIMT GEMMs only
    c.tcw = &tcw;
    if (!c.dry && !c.simt) TRY(refresh_weights(c.st));

    // ---------------- masks
    ALLOC(agent_any, uint8_t, (size_t)bs * A);
    ALLOC(key_pad, uint8_t, (size_t)bs * S);
    ALLOC(r_any, uint8_t, (size_t)bs * R);
    ALLOC(r_pad, uint8_t, (size_t)bs * R);
    if (!c.dry) {
        TRY(launch_token_masks(bt.agent_valid_mask, bt.agent_T, Th, bt.map_valid_mask, P, bs, A, Mp, agent_any, key_pad, c.st));
        TRY(launch_mask_any(bt.ref_valid_mask, bs * R, Pr, r_any, r_pad, c.st));
        if (out.r_padding_mask)
            RIFT_CUDA_OK(cudaMemcpyAsync(out.r_padding_mask, r_pad, (size_t)bs * R, cudaMemcpyDeviceToDevice, c.st));
    }
    ALLOC(tokens, float, (size_t)bs * S * D);

    // ---------------- AgentEncoder (modules/agent_encoder.py:54-94)
    {
        const int NA = bs * A;
        const int Ls[3] = {Th - 1, (Th - 1 + 2 - 3) / 2 + 1, ((Th - 1 + 2 - 3) / 2 + 1 + 2 - 3) / 2 + 1};
        ALLOC(F0, float, (size_t)NA * Ls[0] * 9);
        ALLOC(col0, float, (size_t)NA * Ls[0] * 27);
        if (!c.dry) {
            TRY(launch_agent_features(bt.agent_position, bt.agent_heading, bt.agent_velocity, bt.agent_shape,
                                      bt.agent_valid_mask, NA, Th, bt.agent_T, F0, c.st));
            TRY(launch_im2col_k3(F0, NA, Ls[0], 9, 1, col0, c.st));
        }
        float* x = c.alloc<float>((size_t)NA * Ls[0] * m.hist.embed.N);
        if (!x) { set_last_error("workspace too small"); return -1; }
        TRY(linear(c, col0, 27, NA * Ls[0], m.hist.embed, x, m.hist.embed.N));
        float* lat[3] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < 3; ++i) {
            const NatLevelP& lv = m.hist.levels[i];
            const int d = lv.dim, L = Ls[i], rows = NA * L;
            for (const NatBlockP& nb : lv.blocks) {
                ALLOC(t1, float, (size_t)rows * d);
                ALLOC(qkv, float, (size_t)rows * 3 * d);
                ALLOC(att, float, (size_t)rows * d);
                ALLOC(x1, float, (size_t)rows * d);
                ALLOC(t2, float, (size_t)rows * d);
                ALLOC(hm, float, (size_t)rows * 3 * d);
                ALLOC(x2, float, (size_t)rows * d);
                TRY(layernorm(c, x, rows, nb.n1, t1));
                TRY(linear(c, t1, d, rows, nb.qkv, qkv, 3 * d));
                if (!c.dry) TRY(launch_nat_attention(qkv, NA, L, lv.heads, d / lv.heads, lv.ksize, nb.rpb.p, att, c.st));
                TRY(linear(c, att, d, rows, nb.proj, x1, d, ACT_NONE, x, d, 1));
                TRY(layernorm(c, x1, rows, nb.n2, t2));
                TRY(linear(c, t2, d, rows, nb.fc1, hm, 3 * d, ACT_GELU));
                TRY(linear(c, hm, 3 * d, rows, nb.fc2, x2, d, ACT_NONE, x1, d, 1));
                x = x2;
            }
            // per-level output -> LayerNorm -> lateral Conv1d(k3) to D channels
            ALLOC(o, float, (size_t)rows * d);
            ALLOC(colL, float, (size_t)rows * 3 * d);
            lat[i] = c.alloc<float>((size_t)rows * D);
            if (!lat[i]) { set_last_error("workspace too small"); return -1; }
            TRY(layernorm(c, x, rows, m.hist.norms[i], o));
            if (!c.dry) TRY(launch_im2col_k3(o, NA, L, d, 1, colL, c.st));
            TRY(linear(c, colL, 3 * d, rows, m.hist.lateral[i], lat[i], D));
            if (lv.has_down) {
                const int Ln = Ls[i + 1];
                ALLOC(colD, float, (size_t)NA * Ln * 3 * d);
                ALLOC(xd, float, (size_t)NA * Ln * 2 * d);
                ALLOC(xn, float, (size_t)NA * Ln * 2 * d);
                if (!c.dry) TRY(launch_im2col_k3(x, NA, L, d, 2, colD, c.st));
                TRY(linear(c, colD, 3 * d, NA * Ln, lv.down, xd, 2 * d));
                TRY(layernorm(c, xd, NA * Ln, lv.down_n, xn));
                x = xn;
            }
        }
        if (!c.dry) {
            TRY(launch_fpn_upsample_add(lat[1], lat[2], NA, Ls[1], Ls[2], D, c.st));
            TRY(launch_fpn_upsample_add(lat[0], lat[1], NA, Ls[0], Ls[1], D, c.st));
        }
        ALLOC(colF, float, (size_t)NA * 3 * D);
        ALLOC(x_hist, float, (size_t)NA * D);
        if (!c.dry) TRY(launch_im2col_k3_last(lat[0], NA, Ls[0], D, colF, c.st));
        TRY(linear(c, colF, 3 * D, NA, m.hist.fpn, x_hist, D));

        // StateAttentionEncoder (modules/agent_encoder.py:97-140), 4 heads hard-coded (:104)
        const int nt = cfg.state_channel, eh = 4;
        ALLOC(toks, float, (size_t)bs * nt * D);
        ALLOC(kv, float, (size_t)bs * nt * 2 * D);
        ALLOC(qv, float, (size_t)D);
        ALLOC(eo, float, (size_t)bs * D);
        ALLOC(x_ego, float, (size_t)bs * D);
        if (!c.dry) {
            const float* w[8]; const float* bb[8];
            for (int i = 0; i < nt; ++i) { w[i] = m.ego.lin[i].W; bb[i] = m.ego.lin[i].b; }
            TRY(launch_state_tokens(bt.current_state, bt.cs_stride, bs, nt, D, w, bb, m.ego.pos_embed.p, toks, c.st));
        }
        TRY(linear(c, m.ego.query.p, D, 1, slice(m.ego.attn.in, 0, D, 0, D, true), qv, D));
        TRY(linear(c, toks, D, bs * nt, slice(m.ego.attn.in, D, 2 * D, 0, D, true), kv, 2 * D));
        if (!c.dry) {
            AttnArgs a;
            a.q = qv; a.k = kv; a.v = kv + D; a.o = eo;
            a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
            a.B = bs; a.H = eh; a.Sq = 1; a.Sk = nt; a.hd = D / eh;
            a.q_outer = 0; a.q_seq = 0;                    // one learned query shared by every sample
            a.k_outer = nt;
            a.o_custom = 1; a.o_outer = 1; a.o_seq = 0;    // ... but one output row per sample
            a.scale = 1.f / sqrtf((float)(D / eh));
            TRY(launch_attention(a, c.st));
        }
        TRY(linear(c, eo, D, bs, m.ego.attn.out, x_ego, D));
        if (!c.dry)
            TRY(launch_agent_assemble(x_hist, x_ego, agent_any, bt.agent_category, m.agent_type_emb.p, bs, A, S, D, tokens, c.st));
    }

    // ---------------- MapEncoder (modules/map_encoder.py:31-93)
    if (Mp > 0) {
        const int NP = bs * Mp;
        ALLOC(Fm, float, (size_t)NP * P * 10);
        ALLOC(x_poly, float, (size_t)NP * D);
        ALLOC(x_speed, float, (size_t)NP * D);
        if (!c.dry)
            TRY(launch_map_features(bt.map_point_position, bt.map_point_vector, bt.map_point_orientation, bt.map_polygon_center,
                                    NP, P, Fm, c.st));
        TRY(points_encoder(c, Fm, NP, P, 10, bt.map_valid_mask, m.poly_enc, x_poly));
        TRY(fourier(c, bt.map_polygon_speed_limit, NP, m.speed_emb, x_speed, nullptr));
        if (!c.dry)
            TRY(launch_map_assemble(x_poly, x_speed, bt.map_polygon_type, bt.map_polygon_on_route, bt.map_polygon_tl_status,
                                    bt.map_polygon_has_speed_limit, m.map_type_emb.p, m.map_route_emb.p, m.map_tl_emb.p,
                                    m.map_unknown_emb.p, bs, Mp, A, S, D, tokens, c.st));
    }

    // ---------------- + pos_emb, encoder blocks, final norm (pluto_model.py:144-154)
    ALLOC(pos, float, (size_t)bs * S * 3);
    float* X = c.alloc<float>((size_t)bs * S * D);
    if (!X) { set_last_error("workspace too small"); return -1; }
    if (!c.dry)
        TRY(launch_token_pos(bt.agent_position, bt.agent_heading, bt.map_polygon_center, bs, A, Th, bt.agent_T, Mp, pos, c.st));
    TRY(fourier(c, pos, bs * S, m.pos_emb, X, tokens));
    const int rowsE = bs * S;
    for (const EncBlockP& eb : m.enc) {
        ALLOC(t1, float, (size_t)rowsE * D);
        ALLOC(X1, float, (size_t)rowsE * D);
        ALLOC(t2, float, (size_t)rowsE * D);
        ALLOC(hm, float, (size_t)rowsE * 4 * D);
        ALLOC(X2, float, (size_t)rowsE * D);
        TRY(layernorm(c, X, rowsE, eb.n1, t1));
        TRY(mha_self(c, t1, bs, S, D, H, eb.attn, key_pad, X1, X));
        TRY(layernorm(c, X1, rowsE, eb.n2, t2));
        TRY(linear(c, t2, D, rowsE, eb.fc1, hm, 4 * D, ACT_GELU));
        TRY(linear(c, hm, 4 * D, rowsE, eb.fc2, X2, D, ACT_NONE, X1, D, 1));
        X = X2;
    }
    ALLOC(Xn, float, (size_t)rowsE * D);
    TRY(layernorm(c, X, rowsE, m.final_norm, Xn));

    // ---------------- AgentPredictor (modules/agent_predictor.py:17-29) — only when asked for
    if (out.prediction && A > 1) {
        const int rows = bs * (A - 1);
        ALLOC(xa, float, (size_t)rows * D);
        ALLOC(pl, float, (size_t)rows * 2 * T);
        ALLOC(py, float, (size_t)rows * 2 * T);
        ALLOC(pv, float, (size_t)rows * 2 * T);
        if (!c.dry) TRY(launch_gather_rows(Xn, (long long)S * D, bs, 1, A - 1, D, xa, c.st));
        TRY(mlp_layer(c, xa, D, rows, m.pred_loc, pl, 2 * T));
        TRY(mlp_layer(c, xa, D, rows, m.pred_yaw, py, 2 * T));
        TRY(mlp_layer(c, xa, D, rows, m.pred_vel, pv, 2 * T));
        if (!c.dry) TRY(launch_interleave_heads(pl, py, pv, rows, T, out.prediction, c.st));
    }

    // ---------------- PlanningDecoder (modules/planning_decoder.py:135-188)
    const int NR = bs * R, rowsQ = bs * R * Mo;
    float* q = nullptr;
    {
        ALLOC(Fr, float, (size_t)NR * Pr * 6);
        ALLOC(rpos, float, (size_t)NR * 3);
        ALLOC(r_enc, float, (size_t)NR * D);
        ALLOC(r_emb, float, (size_t)NR * D);
        ALLOC(u, float, (size_t)NR * D);
        ALLOC(v, float, (size_t)Mo * D);
        q = c.alloc<float>((size_t)rowsQ * D);
        if (!q) { set_last_error("workspace too small"); return -1; }
        if (!c.dry) TRY(launch_ref_features(bt.ref_position, bt.ref_vector, bt.ref_orientation, NR, Pr, Fr, rpos, c.st));
        TRY(points_encoder(c, Fr, NR, Pr, 6, bt.ref_valid_mask, m.r_enc, r_enc));
        TRY(fourier(c, rpos, NR, m.r_pos_emb, r_emb, r_enc));
        // q = q_proj(cat[r_emb (per line), m_emb (per mode)]) split into its two column blocks
        TRY(linear_ld(c, r_emb, D, NR, slice(m.q_proj, 0, D, 0, D, false), 2 * D, u, D));
        TRY(linear_ld(c, m.m_emb.p, D, Mo, slice(m.q_proj, 0, D, D, D, true), 2 * D, v, D));
        if (!c.dry) TRY(launch_query_init(u, v, rowsQ, Mo, D, q, c.st));
    }
    const float att_scale = 1.f / sqrtf((float)(D / H));
    for (const DecBlockP& db : m.dec) {
        // (i) r2r self-attention over reference lines, per mode
        ALLOC(t1, float, (size_t)rowsQ * D);
        ALLOC(qkv1, float, (size_t)rowsQ * 3 * D);
        ALLOC(a1, float, (size_t)rowsQ * D);
        ALLOC(q1, float, (size_t)rowsQ * D);
        TRY(layernorm(c, q, rowsQ, db.n1, t1));
        TRY(linear(c, t1, D, rowsQ, db.r2r.in, qkv1, 3 * D));
        if (!c.dry) {
            AttnArgs a;
            a.q = qkv1; a.k = qkv1 + D; a.v = qkv1 + 2 * D; a.o = a1;
            a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
            a.B = bs * Mo; a.H = H; a.Sq = R; a.Sk = R; a.hd = D / H;
            a.q_inner_n = Mo; a.q_outer = (long long)R * Mo; a.q_inner = 1; a.q_seq = Mo;
            a.k_inner_n = Mo; a.k_outer = (long long)R * Mo; a.k_inner = 1; a.k_seq = Mo;
            // The reference passes key_padding_mask = r_pad.repeat(Mo, 1) for a batch laid out (b, m)
            // (planning_decoder.py:56-60): batch row j = b*Mo + m is masked with r_pad[j % bs], not with
            // r_pad[b].  Reproduced as is - parity is defined by what the reference computes.
            a.kpm = r_pad; a.kpm_mod = bs; a.scale = att_scale;
            TRY(launch_attention(a, c.st));
        }
        TRY(linear(c, a1, D, rowsQ, db.r2r.out, q1, D, ACT_NONE, q, D, 1));
        // (ii) m2m self-attention over modes on valid reference lines; padded lines -> 0
        ALLOC(t2, float, (size_t)rowsQ * D);
        ALLOC(t2p, float, (size_t)rowsQ * D);
        ALLOC(qkv2, float, (size_t)rowsQ * 3 * D);
        ALLOC(a2, float, (size_t)rowsQ * D);
        ALLOC(q2, float, (size_t)rowsQ * D);
        TRY(layernorm(c, q1, rowsQ, db.n2, t2, 0, nullptr, nullptr, m.m_pos.p, Mo, t2p));
        TRY(linear(c, t2p, D, rowsQ, slice(db.m2m.in, 0, 2 * D, 0, D, true), qkv2, 3 * D));
        TRY(linear(c, t2, D, rowsQ, slice(db.m2m.in, 2 * D, D, 0, D, true), qkv2 + 2 * D, 3 * D));
        if (!c.dry) {
            AttnArgs a;
            a.q = qkv2; a.k = qkv2 + D; a.v = qkv2 + 2 * D; a.o = a2;
            a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
            a.B = NR; a.H = H; a.Sq = Mo; a.Sk = Mo; a.hd = D / H;
            a.q_outer = Mo; a.k_outer = Mo; a.scale = att_scale;
            TRY(launch_attention(a, c.st));
        }
        TRY(linear(c, a2, D, rowsQ, db.m2m.out, q2, D, ACT_NONE, q1, D, 1));
        if (!c.dry) TRY(launch_zero_rows(q2, r_pad, Mo, rowsQ, D, c.st));
        // (iii) cross-attention to the scene encoding
        ALLOC(t3, float, (size_t)rowsQ * D);
        ALLOC(qc, float, (size_t)rowsQ * D);
        ALLOC(kvc, float, (size_t)rowsE * 2 * D);
        ALLOC(a3, float, (size_t)rowsQ * D);
        ALLOC(q3, float, (size_t)rowsQ * D);
        TRY(layernorm(c, q2, rowsQ, db.n3, t3));
        TRY(linear(c, t3, D, rowsQ, slice(db.cross.in, 0, D, 0, D, true), qc, D));
        TRY(linear(c, Xn, D, rowsE, slice(db.cross.in, D, 2 * D, 0, D, true), kvc, 2 * D));
        if (!c.dry) {
            AttnArgs a;
            a.q = qc; a.k = kvc; a.v = kvc + D; a.o = a3;
            a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
            a.B = bs; a.H = H; a.Sq = R * Mo; a.Sk = S; a.hd = D / H;
            a.q_outer = (long long)R * Mo; a.k_outer = S;
            a.kpm = key_pad; a.kpm_div = 1; a.scale = att_scale;
            TRY(launch_attention(a, c.st));
        }
        TRY(linear(c, a3, D, rowsQ, db.cross.out, q3, D, ACT_NONE, q2, D, 1));
        // (iv) ReLU FFN
        ALLOC(t4, float, (size_t)rowsQ * D);
        ALLOC(hm, float, (size_t)rowsQ * 4 * D);
        ALLOC(q4, float, (size_t)rowsQ * D);
        TRY(layernorm(c, q3, rowsQ, db.n4, t4));
        TRY(linear(c, t4, D, rowsQ, db.ffn0, hm, 4 * D, ACT_RELU));
        TRY(linear(c, hm, 4 * D, rowsQ, db.ffn3, q4, D, ACT_NONE, q3, D, 1));
        q = q4;
    }
    // cat_x_proj(cat[q, x_ego]) : the ego half of the weight acts once per sample
    ALLOC(eg, float, (size_t)bs * D);
    ALLOC(qf, float, (size_t)rowsQ * D);
    TRY(linear_ld(c, Xn, (long long)S * D, bs, slice(m.cat_x_proj, 0, D, D, D, true), 2 * D, eg, D));
    TRY(linear_ld(c, q, D, rowsQ, slice(m.cat_x_proj, 0, D, 0, D, false), 2 * D, qf, D, ACT_NONE, eg, D, R * Mo));

    if (out.trajectory) {
        ALLOC(tl, float, (size_t)rowsQ * 2 * T);
        ALLOC(ty, float, (size_t)rowsQ * 2 * T);
        ALLOC(tv, float, (size_t)rowsQ * 2 * T);
        TRY(mlp_layer(c, qf, D, rowsQ, m.loc_head, tl, 2 * T));
        TRY(mlp_layer(c, qf, D, rowsQ, m.yaw_head, ty, 2 * T));
        TRY(mlp_layer(c, qf, D, rowsQ, m.vel_head, tv, 2 * T));
        if (!c.dry) {
            TRY(launch_interleave_heads(tl, ty, tv, rowsQ, T, out.trajectory, c.st));
            if (out.candidate_trajectories) TRY(launch_traj_outputs(out.trajectory, rowsQ, T, out.candidate_trajectories, c.st));
        }
    }
    {
        ALLOC(pi, float, (size_t)rowsQ);
        float *h = nullptr, *a = nullptr, *mean = nullptr, *rstd = nullptr;
        TRY(mlp_layer(c, qf, D, rowsQ, m.pi_head, pi, 1, &h, &a, &mean, &rstd));
        if (!c.dry) {
            TRY(launch_mask_logits(pi, r_pad, NR, Mo, -1e6f, c.st));
            if (out.probability)
                RIFT_CUDA_OK(cudaMemcpyAsync(out.probability, pi, (size_t)rowsQ * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
            pi_tape.q = qf; pi_tape.h = h; pi_tape.a = a; pi_tape.mean = mean; pi_tape.rstd = rstd; pi_tape.rows = rowsQ;
            pi_tape.valid = c.save;
        }
    }
    if (out.hidden) {
        ALLOC(hh, float, (size_t)bs * D);
        TRY(linear(c, Xn, (long long)S * D, bs, m.hidden0, hh, D, ACT_RELU));
        TRY(linear(c, hh, D, bs, m.hidden2, out.hidden, D));
    }
    if (out.ref_free_trajectory) TRY(mlp_layer(c, Xn, (long long)S * D, bs, m.ref_free, out.ref_free_trajectory, 4 * T));
    fwd_ws_end = c.off;
    return 0;
}

// =====================================================================================
// backward (gradient of the objective wrt the trainable parameters)
// =====================================================================================
namespace rift {
// dW += dY^T X ; db += colsum(dY) ; dX = dY W   (each optional)
static int linear_bwd(Ctx& c, const float* X, long long ldx, const float* dY, long long lddy, int M, const Lin& L,
                      long long ldw, float* dX, long long lddx, float dx_beta) {
    if (L.train && L.dW) {
        const int splits = M >= 2048 ? std::min(64, M / 512) : 1;
        float* ws = nullptr;
        if (splits > 1) { ws = c.alloc<float>((size_t)splits * L.N * L.K); if (!ws) { set_last_error("workspace too small"); return -1; } }
        GemmArgs a;
        a.A = dY; a.sam = 1; a.sak = lddy;             // A(m = out feature, k = row)
        a.B = X; a.sbn = 1; a.sbk = ldx;               // B(n = in feature,  k = row)
        a.C = L.dW; a.ldc = ldw; a.M = L.N; a.N = L.K; a.K = M; a.beta = 1.f;
        a.split_k = splits; a.split_ws = ws;
        TRY(gemm(c, a));
    }
    if (L.train && L.db) {
        ALLOC(sc, float, (size_t)148 * L.N);
        if (!c.dry) TRY(launch_colsum(dY, lddy, M, L.N, L.db, 1, sc, c.st));
    }
    if (dX) {
        GemmArgs a;
        a.A = dY; a.sam = lddy; a.sak = 1;             // A(m = row, k = out feature)
        a.B = L.W; a.sbn = 1; a.sbk = ldw;             // B(n = in feature, k = out feature)
        a.C = dX; a.ldc = lddx; a.M = M; a.N = L.K; a.K = L.N; a.beta = dx_beta;
        TRY(gemm(c, a));
    }
    return 0;
}
}  // namespace rift

int rift_b200_engine::backward(const rift_b200_batch& bt, const float* dlogits, Ctx& c) {
    RIFT_REQUIRE(grads != nullptr, "backward: no gradient arena bound");
    RIFT_REQUIRE(c.dry || pi_tape.valid, "backward: run forward with RIFT_B200_FWD_SAVE_FOR_BACKWARD first");
    RIFT_REQUIRE(!m.any_trainable_outside_pi_head,
                 "backward: trainable layers outside planning_decoder.pi_head are not supported by this build");
    const int D = cfg.dim;
    const long long rows = c.dry ? (long long)bt.bs * bt.R * cfg.num_modes : pi_tape.rows;
    if (!c.dry && train_hi > train_lo)
        RIFT_CUDA_OK(cudaMemsetAsync(grads + train_lo, 0, (size_t)(train_hi - train_lo) * sizeof(float), c.st));
    const MLPLayerP& p = m.pi_head;
    if (!p.l0.train) return 0;
    // pi = l3(a) ; a = relu(LN(h)) ; h = l0(q)
    ALLOC(da, float, (size_t)rows * D);
    ALLOC(dh, float, (size_t)rows * D);
    ALLOC(lnsc, float, (size_t)layernorm_bwd_scratch_floats(D));
    TRY(linear_bwd(c, pi_tape.a, D, dlogits, 1, (int)rows, p.l3, D, da, D, 0.f));
    if (!c.dry)
        TRY(launch_layernorm_bwd(pi_tape.h, D, da, D, (int)rows, D, p.n.g, pi_tape.mean, pi_tape.rstd, pi_tape.a, D, dh, D, 0,
                                 p.n.dg, p.n.db, lnsc, c.st));
    TRY(linear_bwd(c, pi_tape.q, D, dh, D, (int)rows, p.l0, D, nullptr, 0, 0.f));
    return 0;
}
