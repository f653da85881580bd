// Parameter binding and the forward schedule of the Pluto trajectory policy on the rift_b200 kernels.
//
// Mirrors PlanningModel.forward (rift/cbv/planning/pluto/model/pluto_model.py:122-225) in the
// deterministic parity mode of SURVEY 8(c): dropout / drop-path identity, BatchNorm running
// statistics, no state-token dropping.  Ragged structure (x[mask] gathers in the reference) is
// handled densely: every padded row is computed and then masked, which is equivalent because no
// op in the model mixes rows of different agents / polylines except through masked attention and
// masked max-pooling.
//
// Activations that feed a GEMM travel as split-bf16 planes written directly by their producer
// (LayerNorm, attention, GEMM epilogues, im2col, pooling); fp32 copies are kept only where another
// kernel reads them (residual stream, attention q/k/v, pooling inputs) or when the backward needs them.
#include <stdlib.h>

#include "engine_ops.h"
#include "fused_block.h"

using namespace rift;

static const int NAT_HEADS_[3] = {2, 4, 8};
static const int NAT_K_[3] = {3, 3, 5};

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
        Lin l; l.N = N; l.K = K; l.ldw = K;
        if (auto r = find(p + wname)) {
            if (r->numel != (long long)N * K) { set_last_error("shape mismatch for " + p + wname); err = -1; }
            l.W = e->params + r->offset; l.train = r->trainable && e->grads;
            l.dW = l.train ? e->grads + r->offset : nullptr;
            if (N >= 16 && K >= 32) {          // candidate for the tcgen05 path: gets pre-split bf16 planes
                TcWeight w;
                w.src = l.W; w.ld_src = K; w.N = N; w.K = K; w.Kp = tc_pitch(K); w.trainable = r->trainable;
                l.tc = (int)e->tcw.size();
                e->tcw.push_back(w);
                if (e->grads) {                // W^T planes [K, Np] for the data-gradient product dX = dY W
                    TcWeight t;
                    t.src = l.W; t.ld_src = K; t.N = K; t.K = N; t.Kp = tc_pitch(N); t.trainable = r->trainable; t.transpose = true;
                    l.tcT = (int)e->tcw.size();
                    e->tcw.push_back(t);
                }
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
        fill_split_job(all.data() + n_jobs_all * split_job_bytes(), w.src, w.ld_src, w.N, w.K, w.Kp, w.hi, w.lo, total_all,
                       w.transpose ? 1 : 0);
        ++n_jobs_all; total_all += split_job_units(w.N, w.Kp);
        if (w.trainable) {
            fill_split_job(tr.data() + n_jobs_train * split_job_bytes(), w.src, w.ld_src, w.N, w.K, w.Kp, w.hi, w.lo, total_train,
                           w.transpose ? 1 : 0);
            ++n_jobs_train; total_train += split_job_units(w.N, w.Kp);
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

// ---- fork / join streams ----------------------------------------------------------------------
int rift_b200_engine::attach_streams(Ctx& c) {
    if (c.dry) return 0;
    if (streams_state == 0) {
        const char* env = getenv("RIFT_B200_STREAMS");
        if (env && atoi(env) == 0) streams_state = -1;
        else {
            // the side stream (parameter gradients: nothing on the step's critical path waits for them before the final
            // join) gets the LOWEST priority, the branch stream the highest: when an SM frees up, CTAs of the main chain and
            // of the branches it will join go first.  RIFT_B200_STREAM_PRIO=1 enables.
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = numerically largest = least urgent
            const char* pe = getenv("RIFT_B200_STREAM_PRIO");
            const bool prio = pe && atoi(pe) != 0;            // measured neutral to slightly worse (8.47 vs 8.39 ms): off by default
            RIFT_CUDA_OK(cudaStreamCreateWithPriority(&s_side, cudaStreamNonBlocking, prio ? lo : 0));
            RIFT_CUDA_OK(cudaStreamCreateWithPriority(&s_br, cudaStreamNonBlocking, prio ? hi : 0));
            events.resize(64);
            for (auto& ev : events) RIFT_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            streams_state = 1;
        }
    }
    if (streams_state == 1) { c.side = s_side; c.br = s_br; c.events = events.data(); c.n_events = (int)events.size(); }
    return 0;
}

rift_b200_engine::~rift_b200_engine() {
    for (auto& ev : events) cudaEventDestroy(ev);
    if (s_side) cudaStreamDestroy(s_side);
    if (s_br) cudaStreamDestroy(s_br);
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
    m.full = false;
    for (auto& kv : table)
        if (kv.second.trainable && kv.first.rfind("planning_decoder.pi_head.", 0) != 0 && kv.first.rfind("value_net.", 0) != 0)
            m.full = true;
    if (!grads) m.full = false;
    return 0;
}

// =====================================================================================
// forward building blocks
// =====================================================================================
namespace rift {

// MLPLayer: Linear -> LayerNorm -> ReLU -> Linear  (layers/mlp_layer.py:4-16); Y fp32 into dst (pitch ldy)
static int mlp_layer(Ctx& c, Act& X, const MLPLayerP& p, float* dst, long long ldy, MlpTape* tape) {
    Act h, a;
    LNSave ln;
    TRY(linear_new(c, X, p.l0, Epi(), W_F, &h));
    // `a` feeds l3; the backward of the fused LN+ReLU reads it in fp32
    TRY(layernorm_new(c, h.f, X.rows, p.n, 1, want_in(c, X.rows, {&p.l3}, tape != nullptr), &a, tape ? &ln : nullptr));
    TRY(linear_into(c, a, p.l3, Epi(), dst, ldy));
    if (tape) { tape->x = X; tape->h = h; tape->ln = ln; tape->a = a; }
    return 0;
}

// FourierEmbedding (layers/fourier_embedding.py:45-55): out = to_out(sum_i mlp_i(feat_i)) [+ res]
static int fourier(Ctx& c, const float* x, int rows, const FourierP& p, float* out, const float* res, FourierTape* tape) {
    const int D = p.out_l.N;
    ALLOC(acc, float, (size_t)rows * D);
    if (tape) { tape->x = x; tape->rows = rows; tape->dims.clear(); tape->acc = acc; }
    for (int i = 0; i < p.d; ++i) {
        Act feat, h, hn;
        LNSave ln;
        const MLPLayerP& ml = p.mlps[i];
        TRY(new_act(c, rows, FIN, want_in(c, rows, {&ml.l0}), &feat));
        if (feat.f) feat.ld = FIN;
        if (!c.dry) TRY(launch_fourier_features(x, rows, p.d, i, p.freqs.p, NFREQ, feat.f, FIN, c.st, feat.p));
        TRY(linear_new(c, feat, ml.l0, Epi(), W_F, &h));
        TRY(layernorm_new(c, h.f, rows, ml.n, 1, want_in(c, rows, {&ml.l3}, tape != nullptr), &hn, tape ? &ln : nullptr));
        Epi e;
        if (i > 0) { e.res = acc; e.ldres = D; }
        TRY(linear_into(c, hn, ml.l3, e, acc, D));
        if (tape) tape->dims.push_back(FourierTape::Dim{feat, h, ln, hn});
    }
    Act on;
    LNSave lno;
    TRY(layernorm_new(c, acc, rows, p.out_n, 1, want_in(c, rows, {&p.out_l}, tape != nullptr), &on, tape ? &lno : nullptr));
    Epi e;
    e.res = res; e.ldres = D;
    TRY(linear_into(c, on, p.out_l, e, out, D));
    if (tape) { tape->ln_o = lno; tape->on = on; }
    return 0;
}

// PointsEncoder (layers/embedding.py:271-296) over `groups` polylines of `n` points, C_in channels
static int points_encoder(Ctx& c, Act& F, int groups, int n, const uint8_t* mask, const PointsEncP& p, Act* out, int out_want,
                          PointsTape* tape) {
    const int rows = groups * n, Cin = F.C, Cout = p.s3.N;
    ALLOC(sc1, float, PE_H1); ALLOC(sh1, float, PE_H1); ALLOC(sc2, float, PE_H2); ALLOC(sh2, float, PE_H2);
    ALLOC(arg1, int, (size_t)groups * PE_H2);
    ALLOC(arg2, int, (size_t)groups * Cout);
    float *h1pre = nullptr, *h2pre = nullptr;
    if (tape) {
        h1pre = c.alloc<float>((size_t)rows * PE_H1); h2pre = c.alloc<float>((size_t)rows * PE_H2);
        if (!h1pre || !h2pre) { set_last_error("workspace too small"); return -1; }
    }
    if (!c.dry) {
        TRY(launch_bn_fold(p.fbn.affine.g, p.fbn.affine.b, p.fbn.mean, p.fbn.var, p.f0.b, PE_H1, sc1, sh1, c.st));
        TRY(launch_bn_fold(p.sbn.affine.g, p.sbn.affine.b, p.sbn.mean, p.sbn.var, p.s0.b, PE_H2, sc2, sh2, c.st));
    }
    const Lin s0a = slice(p.s0, 0, PE_H2, 0, PE_H2, false), s0b = slice(p.s0, 0, PE_H2, PE_H2, PE_H2, false);
    Act h1, f, pooled, gp, h2, o;
    {   // first_mlp: Linear(C,128) + BN + ReLU + Linear(128,256)
        Epi e; e.colscale = sc1; e.shift = sh1; e.act = ACT_RELU; e.preact = h1pre;
        TRY(linear_new(c, F, p.f0, e, want_in(c, rows, {&p.f3}, tape != nullptr), &h1));
        TRY(linear_new(c, h1, p.f3, Epi(), want_in(c, rows, {&s0a}, true), &f));       // fp32: pooled over points
    }
    TRY(new_act(c, groups, PE_H2, want_in(c, groups, {&s0b}), &pooled));
    if (!c.dry) TRY(launch_masked_maxpool(f.f, mask, groups, n, PE_H2, pooled.f, arg1, c.st, pooled.p));
    {   // second_mlp on cat[feat, pooled]: the pooled half of the weight acts once per polyline
        TRY(linear_new(c, pooled, s0b, Epi(), W_F, &gp));
        Epi e; e.pre = gp.f; e.ldpre = PE_H2; e.pre_div = n; e.colscale = sc2; e.shift = sh2; e.act = ACT_RELU; e.preact = h2pre;
        TRY(linear_new(c, f, s0a, e, want_in(c, rows, {&p.s3}, tape != nullptr), &h2));
        TRY(linear_new(c, h2, p.s3, Epi(), W_F, &o));
    }
    TRY(new_act(c, groups, Cout, out_want, out));
    if (!c.dry) TRY(launch_masked_maxpool(o.f, mask, groups, n, Cout, out->f, arg2, c.st, out->p));
    if (tape) {
        tape->F = F; tape->groups = groups; tape->n = n; tape->Cin = Cin; tape->mask = mask;
        tape->sc1 = sc1; tape->sh1 = sh1; tape->sc2 = sc2; tape->sh2 = sh2;
        tape->h1pre = h1pre; tape->h1 = h1; tape->f = f; tape->pooled = pooled; tape->arg1 = arg1; tape->gp = gp.f;
        tape->h2pre = h2pre; tape->h2 = h2; tape->o = o; tape->arg2 = arg2;
    }
    return 0;
}

// Pre-LN MLP sub-block X2 = X1 + fc2(act(fc1(LN(X1))))  (layers/transformer.py:83-94, planning_decoder.py:80-86, NATLayer MLP).
// One fused cluster kernel (fused_block.cu) when the shape allows, else LayerNorm + two GEMMs.  With `tape` the tensors the
// backward reads are kept: LayerNorm statistics, LN(X1) planes, the fc1 pre-activation and the hidden planes.
struct MlpBlockTape { LNSave* ln; Act* t2; float** hpre; Act* hm; };
static int mlp_block(Ctx& c, float* X1, int rows, const Norm& n2, const Lin& fc1, const Lin& fc2, int act, float* X2,
                     bool keep_pre, const MlpBlockTape* tape) {
    const int D = n2.C, Hd = fc1.N;
    const bool fused = !c.simt && c.tcw && fc1.tc >= 0 && fc2.tc >= 0 && fc1.tc_n0 == 0 && fc1.tc_k0 == 0 && fc2.tc_n0 == 0 &&
                       fc2.tc_k0 == 0 && fc1.b && fc2.b && fused_blocks_enabled() && fused_mlp_shape_ok(rows, D, Hd);
    if (fused) {
        FusedMlpArgs a;
        a.X = X1; a.ldx = D; a.Y = X2; a.ldy = D; a.rows = rows; a.D = D; a.Hd = Hd; a.act = act;
        a.ln_g = n2.g; a.ln_b = n2.b; a.b1 = fc1.b; a.b2 = fc2.b;
        if (tape) {
            Act t2, hm;
            TRY(new_act(c, rows, D, W_P, &t2));
            TRY(new_act(c, rows, Hd, W_P, &hm));
            ALLOC(mean, float, rows);
            ALLOC(rstd, float, rows);
            ALLOC(hpre, float, (size_t)rows * Hd);
            a.t2p = t2.p; a.hmp = hm.p; a.ln_mean = mean; a.ln_rstd = rstd; a.hpre = hpre;
            tape->ln->x = X1; tape->ln->mean = mean; tape->ln->rstd = rstd; tape->ln->rows = rows;
            *tape->t2 = t2; *tape->hm = hm; *tape->hpre = hpre;
            c.used_fused = true;
        }
        if (c.dry) return 0;
        return launch_fused_mlp(a, (*c.tcw)[fc1.tc], (*c.tcw)[fc2.tc], c.st);
    }
    float* hpre = nullptr;
    if (tape && keep_pre) { hpre = c.alloc<float>((size_t)rows * Hd); if (!hpre) { set_last_error("workspace too small"); return -1; } }
    Act t2, hm;
    TRY(layernorm_new(c, X1, rows, n2, 0, want_in(c, rows, {&fc1}), &t2, tape ? tape->ln : nullptr));
    { Epi e; e.act = act; e.preact = hpre; TRY(linear_new(c, t2, fc1, e, want_in(c, rows, {&fc2}), &hm)); }
    { Epi e; e.res = X1; e.ldres = D; TRY(linear_into(c, hm, fc2, e, X2, D)); }
    if (tape) { *tape->t2 = t2; *tape->hm = hm; *tape->hpre = hpre; }
    return 0;
}

}  // namespace rift

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
    if (!wcache) c.simt = true;            // no pre-split planes bound: exact-fp32 SIMT GEMMs only
    c.tcw = &tcw;
    c.full = c.save && m.full;
    const bool full = c.full;
    if (!c.dry && !c.simt) TRY(refresh_weights(c.st));
    Tape& tp = tape;
    if (!c.dry) {
        tp = Tape();
        tp.bs = bs; tp.A = A; tp.Mp = Mp; tp.P = P; tp.R = R; tp.Pr = Pr; tp.S = S;
    }
    Tape scratch_tape;                      // the sizing pass records into a throw-away tape
    Tape& T_ = c.dry ? scratch_tape : tp;

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
    T_.agent_any = agent_any; T_.key_pad = key_pad; T_.r_pad = r_pad;
    // r2r key padding of a micro-batch: the mask rows of the whole batch, indexed with this slice's offset
    const uint8_t* r_pad_r2r = r_pad;
    int r2r_mod = bs, r2r_off = 0;
    if (bt.ref_valid_mask_global && bt.bs_global > bs) {
        ALLOC(rg_any, uint8_t, (size_t)bt.bs_global * R);
        ALLOC(rg_pad, uint8_t, (size_t)bt.bs_global * R);
        if (!c.dry) TRY(launch_mask_any(bt.ref_valid_mask_global, bt.bs_global * R, Pr, rg_any, rg_pad, c.st));
        r_pad_r2r = rg_pad; r2r_mod = bt.bs_global; r2r_off = bt.b_offset * cfg.num_modes;
    }
    T_.r_pad_r2r = r_pad_r2r; T_.r2r_mod = r2r_mod; T_.r2r_off = r2r_off;
    ALLOC(tokens, float, (size_t)bs * S * D);
    // The map encoder and the reference-line encoder / query initialisation depend only on the inputs: they run
    // on the branch stream next to the agent encoder (joined before pos_emb and before the decoder respectively).
    TRY(attach_streams(c));
    TRY(fork_to(c, c.br));

    // ---------------- AgentEncoder (modules/agent_encoder.py:54-94)
    {
        NatTape& nt = T_.nat;
        const int NA = bs * A;
        const int Ls[3] = {Th - 1, (Th - 1 + 2 - 3) / 2 + 1, ((Th - 1 + 2 - 3) / 2 + 1 + 2 - 3) / 2 + 1};
        nt.NA = NA; nt.Ls[0] = Ls[0]; nt.Ls[1] = Ls[1]; nt.Ls[2] = Ls[2];
        nt.blocks.clear();
        ALLOC(F0, float, (size_t)NA * Ls[0] * 9);
        Act col0;
        TRY(new_act(c, NA * Ls[0], 27, W_F, &col0));
        if (!c.dry) {
            TRY(launch_agent_features(bt.agent_position, bt.agent_heading, bt.agent_velocity, bt.agent_shape,
                                      bt.agent_valid_mask, NA, Th, bt.agent_T, F0, c.st));
            TRY(launch_im2col_k3(F0, NA, Ls[0], 9, 1, col0.f, c.st));
        }
        nt.col0 = col0;
        Act xa;
        TRY(linear_new(c, col0, m.hist.embed, Epi(), W_F, &xa));
        float* x = xa.f;
        float* lat[3] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < 3; ++i) {
            const NatLevelP& lv = m.hist.levels[i];
            const int d = lv.dim, L = Ls[i], rows = NA * L;
            for (const NatBlockP& nb : lv.blocks) {
                NatBlockTape bt_;
                bt_.x = x;
                ALLOC(qkv, float, (size_t)rows * 3 * d);
                ALLOC(x1, float, (size_t)rows * d);
                ALLOC(x2, float, (size_t)rows * d);
                Act t1, att;
                TRY(layernorm_new(c, x, rows, nb.n1, 0, want_in(c, rows, {&nb.qkv}), &t1, full ? &bt_.ln1 : nullptr));
                TRY(linear_into(c, t1, nb.qkv, Epi(), qkv, 3 * d));
                TRY(new_act(c, rows, d, want_in(c, rows, {&nb.proj}), &att));
                if (!c.dry) TRY(launch_nat_attention(qkv, NA, L, lv.heads, d / lv.heads, lv.ksize, nb.rpb.p, att.f, c.st, att.p));
                { Epi e; e.res = x; e.ldres = d; TRY(linear_into(c, att, nb.proj, e, x1, d)); }
                {
                    MlpBlockTape mt{&bt_.ln2, &bt_.t2, &bt_.hpre, &bt_.hm};
                    TRY(mlp_block(c, x1, rows, nb.n2, nb.fc1, nb.fc2, ACT_GELU, x2, true, full ? &mt : nullptr));
                }
                bt_.t1 = t1; bt_.qkv = qkv; bt_.att = att; bt_.x1 = x1;
                nt.blocks.push_back(bt_);
                x = x2;
            }
            nt.xlev[i] = x;
            // per-level output -> LayerNorm -> lateral Conv1d(k3) to D channels
            Act o, colL;
            TRY(layernorm_new(c, x, rows, m.hist.norms[i], 0, W_F, &o, full ? &nt.ln_lev[i] : nullptr));
            TRY(new_act(c, rows, 3 * d, want_in(c, rows, {&m.hist.lateral[i]}), &colL));
            if (!c.dry) TRY(launch_im2col_k3(o.f, NA, L, d, 1, colL.f, c.st, colL.p));
            lat[i] = c.alloc<float>((size_t)rows * D);
            if (!lat[i]) { set_last_error("workspace too small"); return -1; }
            TRY(linear_into(c, colL, m.hist.lateral[i], Epi(), lat[i], D));
            nt.colL[i] = colL;
            if (lv.has_down) {
                const int Ln = Ls[i + 1];
                Act colD, xdn;
                ALLOC(xd, float, (size_t)NA * Ln * 2 * d);
                TRY(new_act(c, NA * Ln, 3 * d, want_in(c, NA * Ln, {&lv.down}), &colD));
                if (!c.dry) TRY(launch_im2col_k3(x, NA, L, d, 2, colD.f, c.st, colD.p));
                TRY(linear_into(c, colD, lv.down, Epi(), xd, 2 * d));
                TRY(layernorm_new(c, xd, NA * Ln, lv.down_n, 0, W_F, &xdn, full ? &nt.ln_down[i] : nullptr));
                nt.colD[i] = colD; nt.xd[i] = xd;
                x = xdn.f;
            }
        }
        if (!c.dry) {
            TRY(launch_fpn_upsample_add(lat[1], lat[2], NA, Ls[1], Ls[2], D, c.st));
            TRY(launch_fpn_upsample_add(lat[0], lat[1], NA, Ls[0], Ls[1], D, c.st));
        }
        Act colF, x_hist;
        TRY(new_act(c, NA, 3 * D, want_in(c, NA, {&m.hist.fpn}), &colF));
        if (!c.dry) TRY(launch_im2col_k3_last(lat[0], NA, Ls[0], D, colF.f, c.st, colF.p));
        TRY(linear_new(c, colF, m.hist.fpn, Epi(), W_F, &x_hist));
        nt.colF = colF;

        // StateAttentionEncoder (modules/agent_encoder.py:97-140), 4 heads hard-coded (:104)
        EgoTape& et = T_.ego;
        const int ntok = cfg.state_channel, eh = 4;
        const Lin in_q = slice(m.ego.attn.in, 0, D, 0, D, true), in_kv = slice(m.ego.attn.in, D, 2 * D, 0, D, true);
        Act toks, eo, x_ego;
        cudaEvent_t ego_done = nullptr;
        {   // independent of the history encoder above: branch stream (first in its queue)
        OnStream on_br(c, c.br);
        ALLOC(kv, float, (size_t)bs * ntok * 2 * D);
        ALLOC(qv, float, (size_t)D);
        ALLOC(elsp, float, (size_t)bs * eh);
        TRY(new_act(c, bs * ntok, D, W_F, &toks));
        if (!c.dry) {
            const float* w[8]; const float* bb[8];
            for (int i = 0; i < ntok; ++i) { w[i] = m.ego.lin[i].W; bb[i] = m.ego.lin[i].b; }
            TRY(launch_state_tokens(bt.current_state, bt.cs_stride, bs, ntok, D, w, bb, m.ego.pos_embed.p, toks.f, c.st));
        }
        Act qa = act_f32(const_cast<float*>(m.ego.query.p), D, 1, D);
        TRY(linear_into(c, qa, in_q, Epi(), qv, D));
        TRY(linear_into(c, toks, in_kv, Epi(), kv, 2 * D));
        TRY(new_act(c, bs, D, want_in(c, bs, {&m.ego.attn.out}), &eo));
        if (!c.dry) {
            AttnArgs a;
            a.q = qv; a.k = kv; a.v = kv + D; a.o = eo.f; a.o_planes = eo.p;
            a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
            a.B = bs; a.H = eh; a.Sq = 1; a.Sk = ntok; a.hd = D / eh;
            a.q_outer = 0; a.q_seq = 0;                    // one learned query shared by every sample
            a.k_outer = ntok;
            a.o_custom = 1; a.o_outer = 1; a.o_seq = 0;    // ... but one output row per sample
            a.scale = 1.f / sqrtf((float)(D / eh));
            a.lse = full ? elsp : nullptr;
            TRY(launch_attention(a, c.st));
        }
        TRY(linear_new(c, eo, m.ego.attn.out, Epi(), W_F, &x_ego));
        et.toks = toks.f; et.toks_a = toks; et.kv = kv; et.qv = qv; et.lse = elsp; et.eo = eo;
        if (!c.dry && c.br) { ego_done = c.next_event(); RIFT_CUDA_OK(cudaEventRecord(ego_done, c.br)); }
        }
        if (ego_done) RIFT_CUDA_OK(cudaStreamWaitEvent(c.st, ego_done, 0));
        if (!c.dry)
            TRY(launch_agent_assemble(x_hist.f, x_ego.f, agent_any, bt.agent_category, m.agent_type_emb.p, bs, A, S, D, tokens, c.st));
    }

    // ---------------- MapEncoder (modules/map_encoder.py:31-93)
    if (Mp > 0) {
        OnStream on_br(c, c.br);
        const int NP = bs * Mp;
        Act Fm, x_poly;
        TRY(new_act(c, NP * P, 10, W_F, &Fm));
        ALLOC(x_speed, float, (size_t)NP * D);
        if (!c.dry)
            TRY(launch_map_features(bt.map_point_position, bt.map_point_vector, bt.map_point_orientation, bt.map_polygon_center,
                                    NP, P, Fm.f, c.st));
        TRY(points_encoder(c, Fm, NP, P, bt.map_valid_mask, m.poly_enc, &x_poly, W_F, full ? &T_.poly : nullptr));
        TRY(fourier(c, bt.map_polygon_speed_limit, NP, m.speed_emb, x_speed, nullptr, full ? &T_.speed : nullptr));
        if (!c.dry)
            TRY(launch_map_assemble(x_poly.f, x_speed, bt.map_polygon_type, bt.map_polygon_on_route, bt.map_polygon_tl_status,
                                    bt.map_polygon_has_speed_limit, m.map_type_emb.p, m.map_route_emb.p, m.map_tl_emb.p,
                                    m.map_unknown_emb.p, bs, Mp, A, S, D, tokens, c.st));
    }

    // recorded here, waited for by the main stream right before pos_emb: covers the map tokens only (the branch
    // keeps going with the reference lines / queries)
    cudaEvent_t map_done = nullptr;
    if (!c.dry && c.br) { map_done = c.next_event(); RIFT_CUDA_OK(cudaEventRecord(map_done, c.br)); }
    const int NR = bs * R, rowsQ = bs * R * Mo;
    float* q = nullptr;
    {   // reference-line encoder + query initialisation (planning_decoder.py:135-160), on the branch stream
        OnStream on_br(c, c.br);
        Act Fr, r_enc, r_emb, u, v;
        const Lin qa = slice(m.q_proj, 0, D, 0, D, false), qb = slice(m.q_proj, 0, D, D, D, true);
        TRY(new_act(c, NR * Pr, 6, W_F, &Fr));
        ALLOC(rpos, float, (size_t)NR * 3);
        if (!c.dry) TRY(launch_ref_features(bt.ref_position, bt.ref_vector, bt.ref_orientation, NR, Pr, Fr.f, rpos, c.st));
        T_.rpos = rpos;
        TRY(points_encoder(c, Fr, NR, Pr, bt.ref_valid_mask, m.r_enc, &r_enc, W_F, full ? &T_.renc : nullptr));
        TRY(new_act(c, NR, D, W_F, &r_emb));
        TRY(fourier(c, rpos, NR, m.r_pos_emb, r_emb.f, r_enc.f, full ? &T_.rpos_emb : nullptr));
        T_.r_emb = r_emb;
        // q = q_proj(cat[r_emb (per line), m_emb (per mode)]) split into its two column blocks
        TRY(linear_new(c, r_emb, qa, Epi(), W_F, &u));
        Act me = act_f32(const_cast<float*>(m.m_emb.p), D, Mo, D);
        TRY(linear_new(c, me, qb, Epi(), W_F, &v));
        q = c.alloc<float>((size_t)rowsQ * D);
        if (!q) { set_last_error("workspace too small"); return -1; }
        if (!c.dry) TRY(launch_query_init(u.f, v.f, rowsQ, Mo, D, q, c.st));
    }
    // ---------------- + pos_emb, encoder blocks, final norm (pluto_model.py:144-154)
    if (map_done) RIFT_CUDA_OK(cudaStreamWaitEvent(c.st, map_done, 0));
    ALLOC(pos, float, (size_t)bs * S * 3);
    float* X = c.alloc<float>((size_t)bs * S * D);
    if (!X) { set_last_error("workspace too small"); return -1; }
    if (!c.dry)
        TRY(launch_token_pos(bt.agent_position, bt.agent_heading, bt.map_polygon_center, bs, A, Th, bt.agent_T, Mp, pos, c.st));
    T_.pos = pos;
    TRY(fourier(c, pos, bs * S, m.pos_emb, X, tokens, full ? &T_.pos_emb : nullptr));
    const int rowsE = bs * S;
    const float att_scale = 1.f / sqrtf((float)(D / H));
    T_.enc.clear();
    for (const EncBlockP& eb : m.enc) {
        EncBlockTape et;
        et.X = X;
        ALLOC(qkv, float, (size_t)rowsE * 3 * D);
        ALLOC(X1, float, (size_t)rowsE * D);
        ALLOC(X2, float, (size_t)rowsE * D);
        float* lse = nullptr;
        if (full) {
            lse = c.alloc<float>((size_t)bs * H * S);
            if (!lse) { set_last_error("workspace too small"); return -1; }
        }
        Act t1, att;
        TRY(layernorm_new(c, X, rowsE, eb.n1, 0, want_in(c, rowsE, {&eb.attn.in}), &t1, full ? &et.ln1 : nullptr));
        TRY(linear_into(c, t1, eb.attn.in, Epi(), qkv, 3 * D));
        TRY(new_act(c, rowsE, D, want_in(c, rowsE, {&eb.attn.out}), &att));
        if (!c.dry) {
            AttnArgs a;
            a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D; a.o = att.f; a.o_planes = att.p;
            a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
            a.B = bs; a.H = H; a.Sq = S; a.Sk = S; a.hd = D / H;
            a.q_outer = S; a.k_outer = S;
            a.kpm = key_pad; a.kpm_div = 1; a.scale = att_scale; a.lse = lse;
            TRY(launch_attention(a, c.st));
        }
        { Epi e; e.res = X; e.ldres = D; TRY(linear_into(c, att, eb.attn.out, e, X1, D)); }
        {
            MlpBlockTape mt{&et.ln2, &et.t2, &et.hpre, &et.hm};
            TRY(mlp_block(c, X1, rowsE, eb.n2, eb.fc1, eb.fc2, ACT_GELU, X2, true, full ? &mt : nullptr));
        }
        et.t1 = t1; et.qkv = qkv; et.lse = lse; et.att = att; et.X1 = X1;
        T_.enc.push_back(et);
        X = X2;
    }
    T_.Xlast = X;
    Act Xn;
    {
        // consumers: cross-attention K/V projections (every decoder layer), cat_x_proj, hidden / ref-free heads
        const Lin kv0 = slice(m.dec.empty() ? m.cat_x_proj : m.dec[0].cross.in, D, 2 * D, 0, D, true);
        TRY(layernorm_new(c, X, rowsE, m.final_norm, 0, want_in(c, rowsE, {&kv0}, true), &Xn, full ? &T_.ln_final : nullptr));
    }
    T_.Xn = Xn;
    Act xego_rows = act_f32(Xn.f, (long long)S * D, bs, D);       // x[:, 0]

    // ---------------- AgentPredictor (modules/agent_predictor.py:17-29) — only when asked for
    if (out.prediction && A > 1) {
        const int rows = bs * (A - 1);
        Act xa;
        TRY(new_act(c, rows, D, W_F, &xa));
        ALLOC(pl, float, (size_t)rows * 2 * T);
        ALLOC(py, float, (size_t)rows * 2 * T);
        ALLOC(pv, float, (size_t)rows * 2 * T);
        if (!c.dry) TRY(launch_gather_rows(Xn.f, (long long)S * D, bs, 1, A - 1, D, xa.f, c.st));
        TRY(mlp_layer(c, xa, m.pred_loc, pl, 2 * T, nullptr));
        TRY(mlp_layer(c, xa, m.pred_yaw, py, 2 * T, nullptr));
        TRY(mlp_layer(c, xa, m.pred_vel, pv, 2 * T, nullptr));
        if (!c.dry) TRY(launch_interleave_heads(pl, py, pv, rows, T, out.prediction, c.st));
    }

    // ---------------- PlanningDecoder (modules/planning_decoder.py:135-188)
    // cross-attention K / V projections of the scene encoding do not depend on the query chain: all layers' on the
    // branch stream, behind the query initialisation
    std::vector<float*> kvc_all(m.dec.size(), nullptr);
    TRY(fork_to(c, c.br));
    {
        OnStream on_br(c, c.br);
        for (size_t l = 0; l < m.dec.size(); ++l) {
            const Lin cr_kv = slice(m.dec[l].cross.in, D, 2 * D, 0, D, true);
            kvc_all[l] = c.alloc<float>((size_t)rowsE * 2 * D);
            if (!kvc_all[l]) { set_last_error("workspace too small"); return -1; }
            TRY(linear_into(c, Xn, cr_kv, Epi(), kvc_all[l], 2 * D));
        }
    }
    TRY(join_from(c, c.br));                 // query initialisation + K / V projections (branch stream) are complete
    T_.dec.clear();
    size_t dec_l = 0;
    for (const DecBlockP& db : m.dec) {
        DecBlockTape dt;
        dt.q = q;
        const Lin m2m_qk = slice(db.m2m.in, 0, 2 * D, 0, D, true), m2m_v = slice(db.m2m.in, 2 * D, D, 0, D, true);
        const Lin cr_q = slice(db.cross.in, 0, D, 0, D, true), cr_kv = slice(db.cross.in, D, 2 * D, 0, D, true);
        float *lse1 = nullptr, *lse2 = nullptr, *lse3 = nullptr;
        if (full) {
            lse1 = c.alloc<float>((size_t)bs * Mo * H * R); lse2 = c.alloc<float>((size_t)NR * H * Mo);
            lse3 = c.alloc<float>((size_t)bs * H * R * Mo);
            if (!lse1 || !lse2 || !lse3) { set_last_error("workspace too small"); return -1; }
        }
        // (i) r2r self-attention over reference lines, per mode
        ALLOC(qkv1, float, (size_t)rowsQ * 3 * D);
        ALLOC(q1, float, (size_t)rowsQ * D);
        Act t1, a1;
        TRY(layernorm_new(c, q, rowsQ, db.n1, 0, want_in(c, rowsQ, {&db.r2r.in}), &t1, full ? &dt.ln1 : nullptr));
        TRY(linear_into(c, t1, db.r2r.in, Epi(), qkv1, 3 * D));
        TRY(new_act(c, rowsQ, D, want_in(c, rowsQ, {&db.r2r.out}), &a1));
        if (!c.dry) {
            AttnArgs a;
            a.q = qkv1; a.k = qkv1 + D; a.v = qkv1 + 2 * D; a.o = a1.f; a.o_planes = a1.p;
            a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
            a.B = bs * Mo; a.H = H; a.Sq = R; a.Sk = R; a.hd = D / H;
            a.q_inner_n = Mo; a.q_outer = (long long)R * Mo; a.q_inner = 1; a.q_seq = Mo;
            a.k_inner_n = Mo; a.k_outer = (long long)R * Mo; a.k_inner = 1; a.k_seq = Mo;
            // The reference passes key_padding_mask = r_pad.repeat(Mo, 1) for a batch laid out (b, m)
            // (planning_decoder.py:56-60): batch row j = b*Mo + m is masked with r_pad[j % bs], not with
            // r_pad[b].  Reproduced as is - parity is defined by what the reference computes.
            a.kpm = r_pad_r2r; a.kpm_mod = r2r_mod; a.kpm_off = r2r_off; a.scale = att_scale; a.lse = lse1;
            TRY(launch_attention(a, c.st));
        }
        { Epi e; e.res = q; e.ldres = D; TRY(linear_into(c, a1, db.r2r.out, e, q1, D)); }
        // (ii) m2m self-attention over modes on valid reference lines; padded lines -> 0
        ALLOC(qkv2, float, (size_t)rowsQ * 3 * D);
        ALLOC(q2, float, (size_t)rowsQ * D);
        Act t2, t2p, a2;
        TRY(layernorm_new(c, q1, rowsQ, db.n2, 0, want_in(c, rowsQ, {&m2m_v}), &t2, full ? &dt.ln2 : nullptr, m.m_pos.p, Mo,
                          want_in(c, rowsQ, {&m2m_qk}), &t2p));
        TRY(linear_into(c, t2p, m2m_qk, Epi(), qkv2, 3 * D));
        TRY(linear_into(c, t2, m2m_v, Epi(), qkv2 + 2 * D, 3 * D));
        TRY(new_act(c, rowsQ, D, want_in(c, rowsQ, {&db.m2m.out}), &a2));
        if (!c.dry) {
            AttnArgs a;
            a.q = qkv2; a.k = qkv2 + D; a.v = qkv2 + 2 * D; a.o = a2.f; a.o_planes = a2.p;
            a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
            a.B = NR; a.H = H; a.Sq = Mo; a.Sk = Mo; a.hd = D / H;
            a.q_outer = Mo; a.k_outer = Mo; a.scale = att_scale; a.lse = lse2;
            TRY(launch_attention(a, c.st));
        }
        { Epi e; e.res = q1; e.ldres = D; TRY(linear_into(c, a2, db.m2m.out, e, q2, D)); }
        if (!c.dry) TRY(launch_zero_rows(q2, r_pad, Mo, rowsQ, D, c.st));
        // (iii) cross-attention to the scene encoding
        ALLOC(qc, float, (size_t)rowsQ * D);
        float* kvc = kvc_all[dec_l++];
        ALLOC(q3, float, (size_t)rowsQ * D);
        Act t3, a3;
        TRY(layernorm_new(c, q2, rowsQ, db.n3, 0, want_in(c, rowsQ, {&cr_q}), &t3, full ? &dt.ln3 : nullptr));
        TRY(linear_into(c, t3, cr_q, Epi(), qc, D));
        (void)cr_kv;
        TRY(new_act(c, rowsQ, D, want_in(c, rowsQ, {&db.cross.out}), &a3));
        if (!c.dry) {
            AttnArgs a;
            a.q = qc; a.k = kvc; a.v = kvc + D; a.o = a3.f; a.o_planes = a3.p;
            a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
            a.B = bs; a.H = H; a.Sq = R * Mo; a.Sk = S; a.hd = D / H;
            a.q_outer = (long long)R * Mo; a.k_outer = S;
            a.kpm = key_pad; a.kpm_div = 1; a.scale = att_scale; a.lse = lse3;
            TRY(launch_attention(a, c.st));
        }
        { Epi e; e.res = q2; e.ldres = D; TRY(linear_into(c, a3, db.cross.out, e, q3, D)); }
        // (iv) ReLU FFN
        ALLOC(q4, float, (size_t)rowsQ * D);
        {
            MlpBlockTape mt{&dt.ln4, &dt.t4, &dt.hpre4, &dt.hm};
            TRY(mlp_block(c, q3, rowsQ, db.n4, db.ffn0, db.ffn3, ACT_RELU, q4, false, full ? &mt : nullptr));
        }
        dt.t1 = t1; dt.qkv1 = qkv1; dt.lse1 = lse1; dt.a1 = a1; dt.q1 = q1;
        dt.t2 = t2; dt.t2p = t2p; dt.qkv2 = qkv2; dt.lse2 = lse2; dt.a2 = a2; dt.q2 = q2;
        dt.t3 = t3; dt.qc = qc; dt.kvc = kvc; dt.lse3 = lse3; dt.a3 = a3; dt.q3 = q3;
        T_.dec.push_back(dt);
        q = q4;
    }
    T_.qlast = q;
    // cat_x_proj(cat[q, x_ego]) : the ego half of the weight acts once per sample
    Act eg, qf;
    {
        const Lin ca = slice(m.cat_x_proj, 0, D, 0, D, false), cb = slice(m.cat_x_proj, 0, D, D, D, true);
        Act qa = act_f32(q, D, rowsQ, D);
        TRY(linear_new(c, xego_rows, cb, Epi(), W_F, &eg));
        Epi e; e.res = eg.f; e.ldres = D; e.res_div = R * Mo;
        const bool heads = out.trajectory != nullptr;
        TRY(linear_new(c, qa, ca, e,
                       heads ? want_in(c, rowsQ, {&m.pi_head.l0, &m.loc_head.l0}, c.save) : want_in(c, rowsQ, {&m.pi_head.l0}, c.save),
                       &qf));
        T_.qlast_a = qa; T_.xego_rows = xego_rows;
    }
    T_.qf = qf;

    if (out.trajectory) {
        ALLOC(tl, float, (size_t)rowsQ * 2 * T);
        ALLOC(ty, float, (size_t)rowsQ * 2 * T);
        ALLOC(tv, float, (size_t)rowsQ * 2 * T);
        TRY(mlp_layer(c, qf, m.loc_head, tl, 2 * T, nullptr));
        TRY(mlp_layer(c, qf, m.yaw_head, ty, 2 * T, nullptr));
        TRY(mlp_layer(c, qf, m.vel_head, tv, 2 * T, nullptr));
        if (!c.dry) {
            TRY(launch_interleave_heads(tl, ty, tv, rowsQ, T, out.trajectory, c.st));
            if (out.candidate_trajectories) TRY(launch_traj_outputs(out.trajectory, rowsQ, T, out.candidate_trajectories, c.st));
        }
    }
    {
        ALLOC(pi, float, (size_t)rowsQ);
        TRY(mlp_layer(c, qf, m.pi_head, pi, 1, c.save ? &T_.pi : nullptr));
        if (!c.dry) {
            TRY(launch_mask_logits(pi, r_pad, NR, Mo, -1e6f, c.st));
            if (out.probability)
                RIFT_CUDA_OK(cudaMemcpyAsync(out.probability, pi, (size_t)rowsQ * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
        }
    }
    if (out.hidden) {
        Act hh;
        { Epi e; e.act = ACT_RELU; TRY(linear_new(c, xego_rows, m.hidden0, e, want_in(c, bs, {&m.hidden2}), &hh)); }
        TRY(linear_into(c, hh, m.hidden2, Epi(), out.hidden, D));
    }
    if (out.ref_free_trajectory) TRY(mlp_layer(c, xego_rows, m.ref_free, out.ref_free_trajectory, 4 * T, nullptr));
    if (!c.dry) { tp.valid = c.save; tp.full = full; tp.fused = c.used_fused; }
    fwd_ws_end = c.off;
    return 0;
}
