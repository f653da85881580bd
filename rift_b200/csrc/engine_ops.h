// Operator wrappers shared by the forward (engine.cu) and backward (engine_bwd.cu) schedules.
#pragma once
#include <math.h>
#include <stdlib.h>

#include <initializer_list>

#include "engine.h"

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

// ---------------------------------------------------------------- fork / join between the schedule's streams
// `to` starts after everything enqueued on c.st so far
inline int fork_to(Ctx& c, cudaStream_t to) {
    if (c.dry || !to || to == c.st) return 0;
    cudaEvent_t e = c.next_event();
    RIFT_CUDA_OK(cudaEventRecord(e, c.st));
    RIFT_CUDA_OK(cudaStreamWaitEvent(to, e, 0));
    return 0;
}
// c.st continues after everything enqueued on `from` so far
inline int join_from(Ctx& c, cudaStream_t from) {
    if (c.dry || !from || from == c.st) return 0;
    cudaEvent_t e = c.next_event();
    RIFT_CUDA_OK(cudaEventRecord(e, from));
    RIFT_CUDA_OK(cudaStreamWaitEvent(c.st, e, 0));
    return 0;
}
// Grouped weight gradients (RIFT_B200_WGRAD_GROUP=0 switches them off): lin_bwd queues eligible products instead of launching
// them; a group of up to 12 goes out as ONE persistent kernel on the side stream, ordered after everything the stream that
// produced their operand planes has enqueued so far.  The operands are planes nobody writes again (bump allocator), so the
// delay between queueing and launching is harmless.
inline bool wgrad_group_on() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_WGRAD_GROUP"); return !(e && atoi(e) == 0); }();
    return on;
}
// products per grouped launch (RIFT_B200_WGRAD_GROUP_SIZE, default and maximum 12)
inline int wgrad_group_size() {
    static const int n = [] { const char* e = getenv("RIFT_B200_WGRAD_GROUP_SIZE"); const int v = e ? atoi(e) : 0; return (v > 0 && v <= WGRAD_GROUP_MAX) ? v : WGRAD_GROUP_MAX; }();
    return n;
}
inline int bwd_terms();
inline int flush_wgrads(Ctx& c, int q) {
    if (c.n_pending[q] == 0) return 0;
    const int n = c.n_pending[q];
    c.n_pending[q] = 0;
    if (c.dry) return 0;
    // MEASUREMENT ONLY (results are wrong): drop the grouped weight gradients to see what the step costs without them
    static const bool skip = [] { const char* e = getenv("RIFT_B200_DEBUG_SKIP_WGRAD"); return e && atoi(e) != 0; }();
    if (skip) return 0;
    cudaStream_t from = c.pending_stream[q];
    cudaStream_t to = c.side ? c.side : from;
    if (to != from) {
        cudaEvent_t e = c.next_event();
        RIFT_CUDA_OK(cudaEventRecord(e, from));
        RIFT_CUDA_OK(cudaStreamWaitEvent(to, e, 0));
    }
    return launch_wgrad_group(c.pending[q], n, bwd_terms(), to);
}
inline int flush_wgrads(Ctx& c) {
    TRY(flush_wgrads(c, 0));
    return flush_wgrads(c, 1);
}
// launches inside the scope go to `s` (when non-null)
struct OnStream {
    Ctx& c; cudaStream_t saved;
    OnStream(Ctx& c_, cudaStream_t s) : c(c_), saved(c_.st) { if (s) c.st = s; }
    ~OnStream() { c.st = saved; }
};

constexpr int W_F = 1, W_P = 2;             // "want" bits for an activation: fp32 / split planes
constexpr int NFREQ = 64, FIN = 129;
constexpr int PE_H1 = 128, PE_H2 = 256;

// shape-only routing decision (identical in the sizing pass and the real pass)
inline bool route_tc(const Ctx& c, const Lin& L, int M, long long ldc) {
    return !c.simt && L.tc >= 0 && c.tcw && gemm_tc_shape_ok(M, L.N, L.K) && (ldc % 4) == 0;
}

// which representations a GEMM-input activation needs: planes when some consuming Linear runs on the
// tcgen05 path, fp32 when some consumer runs on the SIMT kernel, another kernel reads it, or the
// full-backward mode keeps every activation
inline int want_in(const Ctx& c, int rows, std::initializer_list<const Lin*> consumers, bool f32_reader = false) {
    bool p = false, f = f32_reader || c.full || c.simt;
    for (const Lin* L : consumers) {
        if (route_tc(c, *L, rows, 4)) p = true; else f = true;
    }
    return (p ? W_P : 0) | (f ? W_F : 0);
}

inline int new_act(Ctx& c, int rows, int C, int want, Act* a) {
    *a = Act();
    a->rows = rows; a->C = C;
    if (want & W_F) {
        a->f = c.alloc<float>((size_t)rows * C); a->ld = C;
        if (!a->f) { set_last_error("workspace too small"); return -1; }
    }
    if (want & W_P) {
        a->p.Kp = tc_pitch(C);
        a->p.hi = c.alloc<uint16_t>((size_t)rows * a->p.Kp);
        a->p.lo = c.alloc<uint16_t>((size_t)rows * a->p.Kp);
        if (!a->p.hi || !a->p.lo) { set_last_error("workspace too small"); return -1; }
    }
    return 0;
}
inline Act act_f32(float* f, long long ld, int rows, int C) { Act a; a.f = f; a.ld = ld; a.rows = rows; a.C = C; return a; }

// RIFT_B200_PACK_TRACE=1: one stderr line per standalone pack (site, rows, features) - which activations still reach a
// tensor-core GEMM as fp32 and pay a pack kernel on the way
inline void pack_trace(const char* site, int rows, int C) {
    static const bool on = getenv("RIFT_B200_PACK_TRACE") != nullptr;
    if (on) fprintf(stderr, "PACK %s %d %d\n", site, rows, C);
}
inline int ensure_planes(Ctx& c, Act& x) {
    if (x.p.on()) return 0;
    x.p.Kp = tc_pitch(x.C);
    x.p.hi = c.alloc<uint16_t>((size_t)x.rows * x.p.Kp);
    x.p.lo = c.alloc<uint16_t>((size_t)x.rows * x.p.Kp);
    if (!x.p.hi || !x.p.lo) { set_last_error("workspace too small"); return -1; }
    if (c.dry) return 0;
    if (!x.f) { set_last_error("internal: activation has neither fp32 nor planes"); return -1; }
    pack_trace("operand", x.rows, x.C);
    return launch_pack_split(x.f, x.ld, x.rows, x.C, x.p.Kp, x.p.hi, x.p.lo, c.st);
}

struct Epi {
    int act = ACT_NONE;
    const float* res = nullptr; long long ldres = 0; int res_div = 1, res_mod = 0;
    const float* pre = nullptr; long long ldpre = 0; int pre_div = 1;
    const float* colscale = nullptr; const float* shift = nullptr;     // eval-BatchNorm fold: replaces the bias
    float* preact = nullptr;
};

// dst (fp32, pitch ldc; may be null on the tensor-core path when only planes are wanted) and / or planes
inline int gemm_fwd(Ctx& c, Act& X, const Lin& L, const Epi& e, float* dst, long long ldc, Planes outp) {
    GemmArgs a;
    a.A = X.f; a.sam = X.ld; a.sak = 1;
    a.B = L.W; a.sbn = L.ldw; a.sbk = 1;
    a.C = dst; a.ldc = ldc; a.M = X.rows; a.N = L.N; a.K = L.K;
    a.bias = e.colscale ? e.shift : L.b; a.colscale = e.colscale;
    a.act = e.act; a.res = e.res; a.ldres = e.ldres; a.res_div = e.res_div; a.res_mod = e.res_mod;
    a.pre = e.pre; a.ldpre = e.ldpre; a.pre_div = e.pre_div; a.preact = e.preact;
    a.out_planes = outp;
    if (route_tc(c, L, X.rows, ldc)) {
        TRY(ensure_planes(c, X));
        if (c.dry) return 0;
        if (!gemm_tc_eligible(a)) { set_last_error("internal: tensor-core GEMM arguments misaligned"); return -1; }
        return launch_gemm_tc(a, X.p.hi, X.p.lo, X.p.Kp, (*c.tcw)[L.tc], L.tc_n0, L.tc_k0, c.st);
    }
    if (c.dry) return 0;
    if (!X.f || !dst) { set_last_error("internal: SIMT GEMM needs fp32 operands"); return -1; }
    a.out_planes = Planes();
    TRY(launch_gemm_simt(a, c.st));
    if (outp.on()) TRY(launch_pack_split(dst, ldc, X.rows, L.N, outp.Kp, outp.hi, outp.lo, c.st));
    return 0;
}

// Y = epilogue(X W^T): allocates Y in the requested representations
inline int linear_new(Ctx& c, Act& X, const Lin& L, const Epi& e, int want, Act* Y) {
    if (!route_tc(c, L, X.rows, L.N)) want |= W_F;          // the SIMT kernel only writes fp32
    TRY(new_act(c, X.rows, L.N, want, Y));
    return gemm_fwd(c, X, L, e, Y->f, L.N, Y->p);
}
// fp32 result into caller-provided storage (column block of a wider buffer)
inline int linear_into(Ctx& c, Act& X, const Lin& L, const Epi& e, float* dst, long long ldc) {
    return gemm_fwd(c, X, L, e, dst, ldc, Planes());
}

// slice of a Linear: output rows [n0, n0+n), input columns [k0, k0+k) (keeps the parent's row pitch)
inline Lin slice(const Lin& L, int n0, int n, int k0, int k, bool with_bias) {
    Lin s;
    s.W = L.W + (long long)n0 * L.ldw + k0;
    s.b = (with_bias && L.b) ? L.b + n0 : nullptr;
    s.dW = L.dW ? L.dW + (long long)n0 * L.ldw + k0 : nullptr;
    s.db = (with_bias && L.db) ? L.db + n0 : nullptr;
    s.N = n; s.K = k; s.ldw = L.ldw; s.train = L.train;
    s.tc = L.tc; s.tcT = L.tcT; s.tc_n0 = L.tc_n0 + n0; s.tc_k0 = L.tc_k0 + k0;
    return s;
}

// y = LN(x) [ReLU] in the requested representations; optional y2 = y + add[row % rowmod]
inline int layernorm_new(Ctx& c, const float* x, int rows, const Norm& n, int relu, int want, Act* y, LNSave* save,
                         const float* add_rowmod = nullptr, int rowmod = 0, int want2 = 0, Act* y2 = nullptr) {
    TRY(new_act(c, rows, n.C, want, y));
    if (y2) TRY(new_act(c, rows, n.C, want2, y2));
    float* mean = nullptr; float* rstd = nullptr;
    if (save) {
        mean = c.alloc<float>(rows); rstd = c.alloc<float>(rows);
        if (!mean || !rstd) { set_last_error("workspace too small"); return -1; }
        save->x = x; save->mean = mean; save->rstd = rstd; save->rows = rows;
    }
    if (c.dry) return 0;
    return launch_layernorm(x, n.C, rows, n.C, n.g, n.b, y->f, n.C, relu, add_rowmod, rowmod, y2 ? y2->f : nullptr, mean, rstd,
                            c.st, y->p, y2 ? y2->p : Planes());
}

// ---------------------------------------------------------------- attention argument builders (fwd + bwd)
inline AttnArgs attn_self(const float* qkv, int B, int S, int D, int H, const uint8_t* kpm, float scale) {
    AttnArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.B = B; a.H = H; a.Sq = S; a.Sk = S; a.hd = D / H;
    a.q_outer = S; a.k_outer = S;
    a.kpm = kpm; a.kpm_div = 1; a.scale = scale;
    return a;
}
// r2r: batch (b, m), sequence over reference lines.  The reference passes key_padding_mask =
// r_pad.repeat(Mo, 1) for a batch laid out (b, m) (planning_decoder.py:56-60): batch row j = b*Mo + m is
// masked with r_pad[j % bs], not with r_pad[b].  Reproduced as is - parity is what the reference computes.
inline AttnArgs attn_r2r(const float* qkv, int bs, int R, int Mo, int D, int H, const uint8_t* r_pad, float scale) {
    AttnArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.B = bs * Mo; a.H = H; a.Sq = R; a.Sk = R; a.hd = D / H;
    a.q_inner_n = Mo; a.q_outer = (long long)R * Mo; a.q_inner = 1; a.q_seq = Mo;
    a.k_inner_n = Mo; a.k_outer = (long long)R * Mo; a.k_inner = 1; a.k_seq = Mo;
    a.kpm = r_pad; a.kpm_mod = bs; a.scale = scale;
    return a;
}
inline AttnArgs attn_m2m(const float* qkv, int NR, int Mo, int D, int H, float scale) {
    AttnArgs a;
    a.q = qkv; a.k = qkv + D; a.v = qkv + 2 * D;
    a.ldq = a.ldk = a.ldv = 3 * D; a.ldo = D;
    a.B = NR; a.H = H; a.Sq = Mo; a.Sk = Mo; a.hd = D / H;
    a.q_outer = Mo; a.k_outer = Mo; a.scale = scale;
    return a;
}
inline AttnArgs attn_cross(const float* qc, const float* kvc, int bs, int RM, int S, int D, int H, const uint8_t* key_pad,
                           float scale) {
    AttnArgs a;
    a.q = qc; a.k = kvc; a.v = kvc + D;
    a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
    a.B = bs; a.H = H; a.Sq = RM; a.Sk = S; a.hd = D / H;
    a.q_outer = RM; a.k_outer = S;
    a.kpm = key_pad; a.kpm_div = 1; a.scale = scale;
    return a;
}
// StateAttentionEncoder: one learned query shared by every sample, one output row per sample
inline AttnArgs attn_ego(const float* qv, const float* kv, int bs, int ntok, int D, int eh) {
    AttnArgs a;
    a.q = qv; a.k = kv; a.v = kv + D;
    a.ldq = D; a.ldk = a.ldv = 2 * D; a.ldo = D;
    a.B = bs; a.H = eh; a.Sq = 1; a.Sk = ntok; a.hd = D / eh;
    a.q_outer = 0; a.q_seq = 0;
    a.k_outer = ntok;
    a.o_custom = 1; a.o_outer = 1; a.o_seq = 0;
    a.scale = 1.f / sqrtf((float)(D / eh));
    return a;
}

// ---------------------------------------------------------------- backward helpers (exact-fp32 SIMT GEMMs)
inline int simt(Ctx& c, const GemmArgs& a) {
    if (c.dry) return 0;
    return launch_gemm_simt(a, c.st);
}
// dW += dY^T X ; db += colsum(dY) ; dX (=|+=) dY W        (X: fp32 [M, K] pitch ldx, optional planes Xp)
// Tensor-core route (c.tc_bwd): dY is packed once into split-bf16 planes [M, Np];
//   dX = dYp x (W^T planes)            K-major tcgen05 GEMM
//   dW += dYp^T x Xp                   MN-major tcgen05 GEMM over the same [rows, features] planes, split-K over rows
// Optional fusions (tensor-core route only):
//   dYp_in  : dY already exists as split-bf16 planes (written by the producing GEMM) - no pack; the bias gradient is
//             then summed from the planes, entirely on the side stream, and the fp32 dY may be null
//   dact_ref: the data gradient is multiplied by act'(dact_ref) in the GEMM epilogue and, with dX_planes, leaves as
//             planes (dX may then be null) - the activation backward and the next layer's pack disappear
struct LinBwdFuse {
    const Planes* dYp_in = nullptr;
    const float* dact_ref = nullptr; long long lddact = 0; int dact = 0;
    Planes* dX_planes = nullptr;
};
// shape-only: will lin_bwd run both of this layer's products on the tensor-core route?
inline bool lin_bwd_all_tc(const Ctx& c, int M, const Lin& L, bool need_dx) {
    const bool tc_on = c.tc_bwd && c.tcw && L.tcT >= 0;
    const bool want_w = L.train && L.dW;
    const bool tc_d = gemm_tc_shape_ok(M, L.K, L.N);
    const bool tc_w = gemm_tc_shape_ok(L.N, L.K, M) && L.N >= 64 && (L.K % 4) == 0 && (L.ldw % 4) == 0;
    return tc_on && (!need_dx || tc_d) && (!want_w || tc_w);
}
// RIFT_B200_BWD_TERMS=1: the gradient products (dX = dY W, dW = dY^T X) run as plain bf16 (hi planes only) instead of
// split-bf16 x3.  The 1e-3 tolerance of BASELINE.json is on logits and loss (forward: always x3); gradients are checked
// by their norms (2e-3) - see DESIGN.md section 2 for the measured error of both forms.
inline int bwd_terms() {
    static const int t = [] { const char* e = getenv("RIFT_B200_BWD_TERMS"); return (e && atoi(e) == 1) ? 1 : 3; }();
    return t;
}
inline int lin_bwd(Ctx& c, const float* X, long long ldx, const float* dY, long long lddy, int M, const Lin& L, float* dX,
                   long long lddx, float dx_beta, bool bias_grad = true, const Planes* Xp = nullptr,
                   const LinBwdFuse* fz = nullptr) {
    const bool want_w = L.train && L.dW;
    const bool tc_on = c.tc_bwd && c.tcw && L.tcT >= 0;
    // data gradient dX[M, K]: when the caller's dX pitch leaves room, K is padded to a multiple of 4 (the extra
    // columns read rows of the W^T planes that do not exist -> TMA zero fill -> zeros in dX's pad columns)
    const int Kd = (dX && lddx % 4 == 0 && lddx >= ((L.K + 3) & ~3)) ? ((L.K + 3) & ~3) : L.K;
    const bool dx_wanted = dX != nullptr || (fz && fz->dX_planes);
    const bool tc_d = tc_on && dx_wanted && gemm_tc_shape_ok(M, Kd, L.N) && (lddx % 4) == 0;
    // weight gradient dW[N, K] = dY^T X: K that is not a multiple of 4 (FourierEmbedding: 129) runs with the column
    // count padded to 4 (the operand planes are zero there) and stores only the real columns through the reduce
    const int Kpad4 = (L.K + 3) & ~3;
    const bool w_padded = (L.K % 4) != 0 || (L.ldw % 4) != 0;
    // (the weight-gradient product reads activation planes only, so it does not need the layer's W^T planes: first layers
    // with 16 <= K < 32 - the history encoder's 27-wide im2col - take it too)
    const bool tc_w = c.tc_bwd && c.tcw && want_w && gemm_tc_shape_ok(L.N, Kpad4, M) && L.N >= 64 && (Xp && Xp->on() ? Xp->Kp >= Kpad4 : true);
    // Parameter gradients go to the side stream: they read only dYp / the saved operand planes (never the fp32 dY, which
    // the main stream may update in place later) and nothing downstream on the main stream depends on them before the
    // final join of backward().  On the tensor-core route the weight-gradient GEMM adds every (tile, split-K part)
    // straight into dW with red.global.add (the arena was zeroed when backward started) and produces the bias gradient
    // from a second accumulator (dY^T x ones): no partial slabs, no reduce / column-sum kernels.
    const bool want_b = bias_grad && L.train && L.db;
    // RIFT_B200_WGRAD_ATOMIC=0: the previous form (partial slabs + fixed-order reduce kernel, separate column sums)
    static const bool atomic_w = [] { const char* e = getenv("RIFT_B200_WGRAD_ATOMIC"); return !(e && atoi(e) == 0); }();
    const bool b_in_wgrad = want_b && tc_w && atomic_w;
    float* sc = nullptr;
    if (want_b && !b_in_wgrad) {
        const size_t slabs_max = (size_t)(pack_colsum_slabs(M, tc_pitch(L.N)) > 148 ? pack_colsum_slabs(M, tc_pitch(L.N)) : 148);
        sc = c.alloc<float>(slabs_max * L.N);
        if (!sc) { set_last_error("workspace too small"); return -1; }
    }
    Planes dYp;
    bool forked = false, bias_done = b_in_wgrad;
    if (fz && fz->dYp_in && fz->dYp_in->on()) {
        if (!(tc_d || !dx_wanted) || (want_w && !tc_w)) { set_last_error("internal: pre-packed dY needs the tensor-core route"); return -1; }
        dYp = *fz->dYp_in;
        if (want_b && !bias_done && !c.dry) {            // bias gradient from the planes, both stages on the side stream
            TRY(fork_to(c, c.side));
            forked = true;
            OnStream on(c, c.side);
            TRY(launch_colsum_planes(dYp, M, L.N, L.db, 1, sc, c.st));
        }
        bias_done = true;
    } else if (tc_d || tc_w) {
        dYp.Kp = tc_pitch(L.N);
        dYp.hi = c.alloc<uint16_t>((size_t)M * dYp.Kp);
        dYp.lo = c.alloc<uint16_t>((size_t)M * dYp.Kp);
        if (!dYp.hi || !dYp.lo) { set_last_error("workspace too small"); return -1; }
        if (!c.dry) {
            int slabs = 0;
            pack_trace(want_b && !bias_done ? "bwd_dY_colsum" : "bwd_dY", M, L.N);
            if (want_b && !bias_done) TRY(launch_pack_split_colsum(dY, lddy, M, L.N, dYp.Kp, dYp.hi, dYp.lo, sc, &slabs, c.st));   // one pass over dY
            if (slabs > 0) {
                SideStream fin;
                if (c.side) { fin.st = c.side; fin.ev = c.next_event(); forked = true; }
                TRY(launch_colsum_final(sc, slabs, L.N, L.db, 1, c.st, fin));
                bias_done = true;
            } else {
                TRY(launch_pack_split(dY, lddy, M, L.N, dYp.Kp, dYp.hi, dYp.lo, c.st));
            }
        }
    }
    if (want_b && !bias_done && !c.dry) {
        SideStream fin;
        if (c.side) { fin.st = c.side; fin.ev = c.next_event(); forked = true; }
        TRY(launch_colsum(dY, lddy, M, L.N, L.db, 1, sc, c.st, fin));
    }
    // weight-gradient kernels that read the fp32 dY (narrow first layers, shapes the tensor cores do not take) also leave
    // the main chain, unless dY is one of the buffers the chain updates in place later.  The side stream must then see
    // everything the main stream has done so far (an earlier fork in this call predates the colsum launches above).
    static const bool small_side = [] { const char* e = getenv("RIFT_B200_SMALL_WGRAD_SIDE"); return !(e && atoi(e) == 0); }();
    const bool off_chain = small_side && c.side != nullptr && !c.updated_inplace(dY) && !(fz && fz->dYp_in);
    if (!tc_w) forked = false;
    if (tc_w) {
        Planes xp;
        if (Xp && Xp->on()) xp = *Xp;
        else {
            xp.Kp = tc_pitch(L.K);
            xp.hi = c.alloc<uint16_t>((size_t)M * xp.Kp);
            xp.lo = c.alloc<uint16_t>((size_t)M * xp.Kp);
            if (!xp.hi || !xp.lo) { set_last_error("workspace too small"); return -1; }
            if (!c.dry) { pack_trace("bwd_X", M, L.K); TRY(launch_pack_split(X, ldx, M, L.K, xp.Kp, xp.hi, xp.lo, c.st)); }
            forked = false;                      // the side stream has not seen this pack yet
        }
        const int tiles = ((L.N + 127) / 128) * ((L.K + (L.K <= 64 ? 63 : 127)) / (L.K <= 64 ? 64 : 128));
        const int num_kb = (M + 63) / 64;
        // one wave of (tile, split) work items; RIFT_B200_WGRAD_CTAS caps the wave below the SM count, which leaves SMs to
        // the data-gradient chain running next to these side-stream products
        static const int wave = [] { const char* e = getenv("RIFT_B200_WGRAD_CTAS"); const int v = e ? atoi(e) : 148; return v > 0 ? v : 148; }();
        // split count: at most RIFT_B200_WGRAD_SPLIT_CAP (24) parts per tile.  More parts shorten the kernel but take SMs
        // from the data-gradient chain running beside it: measured 8.86 ms per step with 148, 8.85 with 64, 8.66 with 24.
        static const int split_cap = [] { const char* e = getenv("RIFT_B200_WGRAD_SPLIT_CAP"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 24; }();
        int splits = wave / tiles;
        int cap = atomic_w ? split_cap : 24;
        // very tall activations (the history encoder's 20480 / 40960-row levels: 320 / 640 k-blocks) sit at the end of the
        // backward, where the data-gradient chain has finished and nothing competes for the SMs: one part per 8 k-blocks
        if (atomic_w && num_kb >= 256 && num_kb / 8 > cap) cap = num_kb / 8;
        if (splits > cap) splits = cap;
        if (splits > num_kb) splits = num_kb;
        if (splits < 1) splits = 1;
        float* ws = nullptr;
        if (!atomic_w && (splits > 1 || w_padded)) { ws = c.alloc<float>((size_t)splits * L.N * Kpad4); if (!ws) { set_last_error("workspace too small"); return -1; } }
        if (!c.dry && wgrad_group_on() && atomic_w && !w_padded && wgrad_group_takes(L.N, L.K, M, L.ldw, L.dW)) {
            const int q = (c.br != nullptr && c.st == c.br) ? 1 : 0;
            if (c.n_pending[q] && c.pending_stream[q] != c.st) TRY(flush_wgrads(c, q));
            c.pending_stream[q] = c.st;
            WgradItem& w = c.pending[q][c.n_pending[q]++];
            w.A = PlaneOp{dYp.hi, dYp.lo, M, dYp.Kp, 0, 0};
            w.B = PlaneOp{xp.hi, xp.lo, M, xp.Kp, 0, 0};
            w.C = L.dW; w.ldc = L.ldw; w.colsum = b_in_wgrad ? L.db : nullptr;
            w.M = L.N; w.N = L.K; w.K = M;
            if (c.n_pending[q] == wgrad_group_size()) TRY(flush_wgrads(c, q));
        } else if (!c.dry) {
            if (!forked) TRY(fork_to(c, c.side));
            OnStream on(c, c.side);
            GemmArgs a;
            a.C = L.dW; a.ldc = L.ldw; a.M = L.N; a.N = Kpad4; a.K = M;
            if (atomic_w) { a.atomic_out = true; if (b_in_wgrad) a.colsum_out = L.db; }
            else a.beta = 1.f;
            a.terms = bwd_terms();
            if (w_padded) a.n_store = L.K;
            PlaneOp A{dYp.hi, dYp.lo, M, dYp.Kp, 0, 0};
            PlaneOp B{xp.hi, xp.lo, M, xp.Kp, 0, 0};
            TRY(launch_gemm_tc_ex(a, A, B, true, splits, ws, c.st));
        }
    } else if (want_w && L.K <= 32 && L.ldw == L.K && M >= 1024) {
        // narrow first-layer input (raw features / im2col of 9 channels): dedicated kernel instead of the generic SIMT GEMM
        ALLOC(wp, float, (size_t)wgrad_narrow_slabs(M) * L.N * L.K);
        if (!c.dry) {
            if (off_chain && !forked) { TRY(fork_to(c, c.side)); forked = true; }
            OnStream on(c, off_chain ? c.side : nullptr);
            TRY(launch_wgrad_narrow(dY, lddy, X, ldx, M, L.N, L.K, L.dW, wp, c.st));
        }
    } else if (want_w && L.N == 1 && L.K >= 32 && M >= 1024) {
        // single-output Linear (probability head): dW[1, K] = sum_r dY[r] X[r, :] = the narrow kernel with the roles of
        // dY and X swapped ([K, 1] and [1, K] are the same memory)
        ALLOC(wp, float, (size_t)wgrad_narrow_slabs(M) * L.K);
        if (!c.dry) {
            if (off_chain && !forked) { TRY(fork_to(c, c.side)); forked = true; }
            OnStream on(c, off_chain ? c.side : nullptr);
            TRY(launch_wgrad_narrow(X, ldx, dY, lddy, M, L.K, 1, L.dW, wp, c.st));
        }
    } else if (want_w) {
        const int splits = M >= 2048 ? (M / 512 < 64 ? M / 512 : 64) : 1;
        float* ws = nullptr;
        if (splits > 1) { ws = c.alloc<float>((size_t)splits * L.N * L.K); if (!ws) { set_last_error("workspace too small"); return -1; } }
        GemmArgs a;
        a.A = dY; a.sam = 1; a.sak = lddy;             // A(m = out feature, k = row)
        a.B = X; a.sbn = 1; a.sbk = ldx;               // B(n = in feature,  k = row)
        a.C = L.dW; a.ldc = L.ldw; a.M = L.N; a.N = L.K; a.K = M; a.beta = 1.f;
        a.split_k = splits; a.split_ws = ws;
        if (!c.dry && off_chain && !forked) { TRY(fork_to(c, c.side)); forked = true; }
        OnStream on(c, (!c.dry && off_chain) ? c.side : nullptr);
        TRY(simt(c, a));
    }
    if (tc_d) {
        if (!c.dry) {
            const TcWeight& wt = (*c.tcw)[L.tcT];          // planes [K_full, Np_full] of W^T; the slice origin swaps roles
            GemmArgs a;
            a.C = dX; a.ldc = lddx; a.M = M; a.N = Kd; a.K = L.N; a.beta = dx_beta;
            a.terms = bwd_terms();
            if (fz) {
                a.dact_ref = fz->dact_ref; a.lddact = fz->lddact; a.dact = fz->dact;
                if (fz->dX_planes) a.out_planes = *fz->dX_planes;
            }
            PlaneOp A{dYp.hi, dYp.lo, M, dYp.Kp, 0, 0};
            PlaneOp B{wt.hi, wt.lo, wt.N, wt.Kp, L.tc_k0, L.tc_n0};
            TRY(launch_gemm_tc_ex(a, A, B, false, 1, nullptr, c.st));
        }
    } else if (dX) {
        GemmArgs a;
        a.A = dY; a.sam = lddy; a.sak = 1;             // A(m = row, k = out feature)
        a.B = L.W; a.sbn = 1; a.sbk = L.ldw;           // B(n = in feature, k = out feature)
        a.C = dX; a.ldc = lddx; a.M = M; a.N = L.K; a.K = L.N; a.beta = dx_beta;
        TRY(simt(c, a));
    }
    return 0;
}
// LayerNorm backward: dx (=|+=) ... ; parameter gradients accumulate
// dx_planes (optional, tensor-core backward only): the updated dx also leaves as split-bf16 planes for the GEMM that takes
// it as dY next (LinBwdFuse::dYp_in) - the standalone pack between two sub-blocks disappears; left off() when the shape
// does not allow it, and the consumer packs as before.  zero_flag: dx rows to clear (replaces a zero_rows launch).
// RIFT_B200_FUSE_LN_PLANES=0 switches both off.
inline bool ln_bwd_fuse_on() {
    static const bool on = [] { const char* e = getenv("RIFT_B200_FUSE_LN_PLANES"); return !(e && atoi(e) == 0); }();
    return on;
}
// planes [rows, tc_pitch(N)] that a backward kernel fills directly (attention backward -> in-projection's dY)
inline int new_planes(Ctx& c, int rows, int N, Planes* p) {
    p->Kp = tc_pitch(N);
    p->hi = c.alloc<uint16_t>((size_t)rows * p->Kp);
    p->lo = c.alloc<uint16_t>((size_t)rows * p->Kp);
    if (!p->hi || !p->lo) { set_last_error("workspace too small"); return -1; }
    return 0;
}
inline Planes plane_cols(const Planes& p, int col0) { Planes q = p; q.hi += col0; q.lo += col0; return q; }
// RIFT_B200_FUSE_ATTN_PLANES=0: attention backward writes fp32 and every in-projection packs its dY again
inline bool attn_planes_on(const Ctx& c) {
    static const bool on = [] { const char* e = getenv("RIFT_B200_FUSE_ATTN_PLANES"); return !(e && atoi(e) == 0); }();
    return on && c.tc_bwd;
}
inline int ln_bwd(Ctx& c, const LNSave& s, const Norm& n, const float* dy, const float* y_relu, float* dx, int accumulate,
                  Planes* dx_planes = nullptr, int plane_rows = 0, const uint8_t* zero_flag = nullptr, int zero_div = 1) {
    ALLOC(sc, float, (size_t)layernorm_bwd_scratch_floats(n.C));
    const bool fusable = ln_bwd_fuse_on() && c.tc_bwd && (n.C % 64) == 0 && dx != nullptr;
    Planes out;
    if (dx_planes) *dx_planes = Planes();
    if (dx_planes && fusable) {                          // (plane_rows: the dry pass has no tape to read the row count from)
        out.Kp = n.C;
        out.hi = c.alloc<uint16_t>((size_t)plane_rows * out.Kp);
        out.lo = c.alloc<uint16_t>((size_t)plane_rows * out.Kp);
        if (!out.hi || !out.lo) { set_last_error("workspace too small"); return -1; }
        *dx_planes = out;
    }
    if (c.dry) return 0;
    if (out.on() && plane_rows != s.rows) { set_last_error("internal: ln_bwd plane rows"); return -1; }
    const bool pg = n.train && n.dg;
    // The parameter gradients (column sums over all rows, atomics into 2 C addresses) run as a second, dx-less launch of the
    // same kernel on the side stream: the main chain keeps only the row-wise part.  dy must not be a buffer the chain
    // updates in place.  Off by default (RIFT_B200_LN_PG_SIDE=1 enables): measured 9.62 vs 9.26 ms per step - the step is
    // bound by the sum of kernel time over all streams, and the second launch re-reads x and dy.
    static const bool pg_side = [] { const char* e = getenv("RIFT_B200_LN_PG_SIDE"); return e && atoi(e) != 0; }();
    if (pg && pg_side && c.side && dx && !c.updated_inplace(dy)) {
        TRY(fork_to(c, c.side));
        {
            OnStream on(c, c.side);
            TRY(launch_layernorm_bwd(s.x, n.C, dy, n.C, s.rows, n.C, n.g, s.mean, s.rstd, y_relu, n.C, nullptr, n.C, 0, n.dg, n.db, sc, c.st));
        }
        if (zero_flag && !fusable) {
            TRY(launch_layernorm_bwd(s.x, n.C, dy, n.C, s.rows, n.C, n.g, s.mean, s.rstd, y_relu, n.C, dx, n.C, accumulate, nullptr, nullptr,
                                     sc, c.st));
            return launch_zero_rows(dx, zero_flag, zero_div, s.rows, n.C, c.st);
        }
        return launch_layernorm_bwd(s.x, n.C, dy, n.C, s.rows, n.C, n.g, s.mean, s.rstd, y_relu, n.C, dx, n.C, accumulate, nullptr, nullptr,
                                    sc, c.st, SideStream(), out, zero_flag, zero_div);
    }
    SideStream fin;
    if (pg && c.side) { fin.st = c.side; fin.ev = c.next_event(); }       // dgamma / dbeta reduction off the main chain
    if (zero_flag && !fusable) {                         // generic shapes: the separate kernel, after the update
        TRY(launch_layernorm_bwd(s.x, n.C, dy, n.C, s.rows, n.C, n.g, s.mean, s.rstd, y_relu, n.C, dx, n.C, accumulate,
                                 pg ? n.dg : nullptr, pg ? n.db : nullptr, sc, c.st, fin));
        return launch_zero_rows(dx, zero_flag, zero_div, s.rows, n.C, c.st);
    }
    return launch_layernorm_bwd(s.x, n.C, dy, n.C, s.rows, n.C, n.g, s.mean, s.rstd, y_relu, n.C, dx, n.C, accumulate,
                                pg ? n.dg : nullptr, pg ? n.db : nullptr, sc, c.st, fin, out, zero_flag, zero_div);
}

}  // namespace rift
