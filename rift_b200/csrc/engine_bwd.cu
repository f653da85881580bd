// Backward schedule: d(loss)/d(probability) -> gradient arena, for any trainable set
// (LightningTrainer.freeze_parameters, rift_trainer.py:78-90).  With the reference's default set
// (planning_decoder.pi_head) only the head is differentiated; otherwise the whole graph that reaches the
// logits is walked in reverse: pi_head <- cat_x_proj <- decoder blocks <- {query init, reference-line
// encoder, scene encoding} <- encoder blocks <- {pos_emb, agent encoder (NAT + ego state attention), map
// encoder}.  The trajectory / prediction / hidden / ref-free heads receive no gradient from any RL
// objective (their outputs are not part of the loss), exactly like autograd in the reference.
#include <stdlib.h>

#include "engine_ops.h"

using namespace rift;

namespace rift {

#define ZALLOC(var, n)                                                                                 \
    ALLOC(var, float, n);                                                                              \
    if (!c.dry) RIFT_CUDA_OK(cudaMemsetAsync(var, 0, (size_t)(n) * sizeof(float), c.st));

static int act_bwd(Ctx& c, const float* ref, float* dy, long long n, int act, Planes flat_planes = Planes()) {
    if (c.dry) return 0;
    return launch_act_bwd(ref, dy, n, act, c.st, flat_planes);
}
static int add_into(Ctx& c, float* dst, const float* src, long long n) {
    if (c.dry || !dst) return 0;
    return launch_add_inplace(dst, src, n, c.st);
}

// MLPLayer backward; dX may be null
static int mlp_bwd(Ctx& c, const MlpTape& t, const MLPLayerP& p, const float* dY, long long lddy, float* dX) {
    const int rows = t.x.rows, Hd = p.l0.N;
    ALLOC(da, float, (size_t)rows * Hd);
    ALLOC(dh, float, (size_t)rows * Hd);
    TRY(lin_bwd(c, t.a.f, Hd, dY, lddy, rows, p.l3, da, Hd, 0.f, true, &t.a.p));
    TRY(ln_bwd(c, t.ln, p.n, da, t.a.f, dh, 0));
    TRY(lin_bwd(c, t.x.f, t.x.ld, dh, Hd, rows, p.l0, dX, p.l0.K, 0.f, true, &t.x.p));
    return 0;
}

// FourierEmbedding backward (parameters only: its inputs are raw coordinates)
static int fourier_bwd(Ctx& c, const FourierTape& t, const FourierP& p, const float* dOut) {
    const int rows = t.rows, D = p.out_l.N;
    ALLOC(d_on, float, (size_t)rows * D);
    ALLOC(d_acc, float, (size_t)rows * D);
    TRY(lin_bwd(c, t.on.f, D, dOut, D, rows, p.out_l, d_on, D, 0.f, true, &t.on.p));
    TRY(ln_bwd(c, t.ln_o, p.out_n, d_on, t.on.f, d_acc, 0));
    for (int i = 0; i < p.d; ++i) {
        const MLPLayerP& ml = p.mlps[i];
        const FourierTape::Dim& dm = t.dims[i];
        ALLOC(d_hn, float, (size_t)rows * D);
        ALLOC(d_h, float, (size_t)rows * D);
        TRY(lin_bwd(c, dm.hn.f, D, d_acc, D, rows, ml.l3, d_hn, D, 0.f, true, &dm.hn.p));
        TRY(ln_bwd(c, dm.ln, ml.n, d_hn, dm.hn.f, d_h, 0));
        float* d_feat = nullptr;
        constexpr int FINP = (FIN + 3) & ~3;             // pitch of d_feat: lets the data gradient run on the tcgen05 path
        if (p.freqs.train) { d_feat = c.alloc<float>((size_t)rows * FINP); if (!d_feat) { set_last_error("workspace too small"); return -1; } }
        TRY(lin_bwd(c, dm.feat.f, FIN, d_h, D, rows, ml.l0, d_feat, FINP, 0.f, true, &dm.feat.p));
        if (p.freqs.train) {
            ALLOC(contrib, float, (size_t)rows * NFREQ);
            ALLOC(sc, float, (size_t)148 * NFREQ);
            if (!c.dry) {
                TRY(launch_fourier_freq_bwd(t.x, rows, p.d, i, p.freqs.p, NFREQ, d_feat, FINP, contrib, c.st));
                TRY(launch_colsum(contrib, NFREQ, rows, NFREQ, p.freqs.d + (long long)i * NFREQ, 1, sc, c.st));
            }
        }
    }
    return 0;
}

// PointsEncoder backward (parameters only: its inputs are raw geometry)
static int points_bwd(Ctx& c, const PointsTape& t, const PointsEncP& p, const float* dOut) {
    const int groups = t.groups, n = t.n, rows = groups * n, Cout = p.s3.N;
    const Lin s0a = slice(p.s0, 0, PE_H2, 0, PE_H2, false), s0b = slice(p.s0, 0, PE_H2, PE_H2, PE_H2, false);
    ALLOC(d_o, float, (size_t)rows * Cout);
    ALLOC(d_h2, float, (size_t)rows * PE_H2);
    ALLOC(df, float, (size_t)rows * PE_H2);
    ALLOC(d_gp, float, (size_t)groups * PE_H2);
    ALLOC(d_pooled, float, (size_t)groups * PE_H2);
    ALLOC(d_h1, float, (size_t)rows * PE_H1);
    ALLOC(sc, float, (size_t)148 * 2 * PE_H2);
    if (!c.dry) TRY(launch_masked_maxpool_bwd(dOut, t.arg2, groups, n, Cout, d_o, 0, c.st));
    TRY(lin_bwd(c, t.h2.f, PE_H2, d_o, Cout, rows, p.s3, d_h2, PE_H2, 0.f, true, &t.h2.p));
    TRY(act_bwd(c, t.h2.f, d_h2, (long long)rows * PE_H2, ACT_RELU));
    if (!c.dry) {
        if (p.sbn.affine.train)
            TRY(launch_bn_affine_bwd(d_h2, t.h2pre, rows, PE_H2, p.sbn.affine.g, p.sbn.affine.b, p.sbn.affine.dg, p.sbn.affine.db, sc, c.st));
        TRY(launch_scale_cols(d_h2, t.sc2, rows, PE_H2, c.st));                      // through the folded BatchNorm scale
        if (p.s0.train && p.s0.db) TRY(launch_colsum(d_h2, PE_H2, rows, PE_H2, p.s0.db, 1, sc, c.st));
    }
    TRY(lin_bwd(c, t.f.f, PE_H2, d_h2, PE_H2, rows, s0a, df, PE_H2, 0.f, false, &t.f.p));
    if (!c.dry) TRY(launch_groupsum(d_h2, PE_H2, groups, n, PE_H2, d_gp, PE_H2, 0, c.st));
    TRY(lin_bwd(c, t.pooled.f, PE_H2, d_gp, PE_H2, groups, s0b, d_pooled, PE_H2, 0.f, false, &t.pooled.p));
    if (!c.dry) TRY(launch_masked_maxpool_bwd(d_pooled, t.arg1, groups, n, PE_H2, df, 1, c.st));
    TRY(lin_bwd(c, t.h1.f, PE_H1, df, PE_H2, rows, p.f3, d_h1, PE_H1, 0.f, true, &t.h1.p));
    TRY(act_bwd(c, t.h1.f, d_h1, (long long)rows * PE_H1, ACT_RELU));
    if (!c.dry) {
        if (p.fbn.affine.train)
            TRY(launch_bn_affine_bwd(d_h1, t.h1pre, rows, PE_H1, p.fbn.affine.g, p.fbn.affine.b, p.fbn.affine.dg, p.fbn.affine.db, sc, c.st));
        TRY(launch_scale_cols(d_h1, t.sc1, rows, PE_H1, c.st));
        if (p.f0.train && p.f0.db) TRY(launch_colsum(d_h1, PE_H1, rows, PE_H1, p.f0.db, 1, sc, c.st));
    }
    TRY(lin_bwd(c, t.F.f, t.F.ld, d_h1, PE_H1, rows, p.f0, nullptr, 0, 0.f, false, &t.F.p));
    return 0;
}

// generic pre-LN attention + MLP block tail: X2 = X1 + fc2(act(fc1(LN2(X1))));  dX (rows x D) is updated in place
// dXp (optional): on entry the split-bf16 planes of dX when its producer wrote them (else off()), on return the planes of
// the updated dX (or off()) - the residual-stream gradient then never needs a pack kernel between sub-blocks
static int mlp_tail_bwd(Ctx& c, int rows, int D, int Hd, const Act& hm, const float* act_ref, int act, const Act& t2, const LNSave& ln2,
                        const Norm& n2, const Lin& fc1, const Lin& fc2, float* dX, Planes* dXp = nullptr) {
    ALLOC(d_t2, float, (size_t)rows * D);
    const Planes dY_in = (dXp && dXp->on() && lin_bwd_all_tc(c, rows, fc2, true)) ? *dXp : Planes();
    // (ReLU only: GELU's derivative - erf + exp per element - costs more inside the GEMM epilogue, where it is
    // instruction-latency bound, than the bandwidth-bound activation-backward kernel it would replace; measured)
    static const bool fuse_gelu = [] { const char* e = getenv("RIFT_B200_FUSE_GELU_BWD"); return e && atoi(e) != 0; }();
    if ((act == ACT_RELU || fuse_gelu) && lin_bwd_all_tc(c, rows, fc2, true) && lin_bwd_all_tc(c, rows, fc1, true) && (Hd % 4) == 0) {
        // tensor-core route: fc2's data-gradient GEMM applies act'(.) in its epilogue and writes d_hm as split-bf16
        // planes only; fc1's backward consumes them directly (no activation-backward kernel, no pack, bias gradient
        // summed from the planes on the side stream)
        Planes dhp;
        dhp.Kp = tc_pitch(Hd);
        dhp.hi = c.alloc<uint16_t>((size_t)rows * dhp.Kp);
        dhp.lo = c.alloc<uint16_t>((size_t)rows * dhp.Kp);
        if (!dhp.hi || !dhp.lo) { set_last_error("workspace too small"); return -1; }
        LinBwdFuse f2; f2.dact_ref = act_ref; f2.lddact = Hd; f2.dact = act; f2.dX_planes = &dhp;
        if (dY_in.on()) f2.dYp_in = &dY_in;
        TRY(lin_bwd(c, hm.f, Hd, dX, D, rows, fc2, nullptr, Hd, 0.f, true, &hm.p, &f2));
        LinBwdFuse f1; f1.dYp_in = &dhp;
        TRY(lin_bwd(c, t2.f, D, nullptr, Hd, rows, fc1, d_t2, D, 0.f, true, &t2.p, &f1));
    } else {
        ALLOC(d_hm, float, (size_t)rows * Hd);
        LinBwdFuse f2; f2.dYp_in = &dY_in;
        TRY(lin_bwd(c, hm.f, Hd, dX, D, rows, fc2, d_hm, Hd, 0.f, true, &hm.p, dY_in.on() ? &f2 : nullptr));
        // the activation backward writes fc1's dY as split-bf16 planes (and no fp32) when fc1's products take the tensor cores
        Planes dhp;
        if (attn_planes_on(c) && (Hd % 64) == 0 && lin_bwd_all_tc(c, rows, fc1, true)) TRY(new_planes(c, rows, Hd, &dhp));
        TRY(act_bwd(c, act_ref, d_hm, (long long)rows * Hd, act, dhp));
        LinBwdFuse f1; f1.dYp_in = &dhp;
        TRY(lin_bwd(c, t2.f, D, d_hm, Hd, rows, fc1, d_t2, D, 0.f, true, &t2.p, dhp.on() ? &f1 : nullptr));
    }
    TRY(ln_bwd(c, ln2, n2, d_t2, nullptr, dX, 1, dXp, rows));
    return 0;
}

}  // namespace rift

int rift_b200_engine::backward(const rift_b200_batch& bt, const float* dlogits, Ctx& c) {
    TRY(attach_streams(c));
    int r = backward_impl(bt, dlogits, c);
    if (r) { c.n_pending[0] = c.n_pending[1] = 0; return r; }
    TRY(flush_wgrads(c));                        // the last queued weight-gradient products
    // parameter-gradient work (side stream) and the parameter-only branches rejoin the caller's stream here
    TRY(join_from(c, c.br));
    TRY(join_from(c, c.side));
    return 0;
}

int rift_b200_engine::backward_impl(const rift_b200_batch& bt, const float* dlogits, Ctx& c) {
    RIFT_REQUIRE(grads != nullptr, "backward: no gradient arena bound");
    RIFT_REQUIRE(c.dry || tape.valid, "backward: run forward with RIFT_B200_FWD_SAVE_FOR_BACKWARD first");
    const int D = cfg.dim, H = cfg.num_heads, Mo = cfg.num_modes;
    const int bs = bt.bs, A = bt.A, Mp = bt.Mp, R = bt.R, S = A + Mp;
    const int NR = bs * R, rowsQ = NR * Mo, rowsE = bs * S;
    const float att_scale = 1.f / sqrtf((float)(D / H));
    const bool full = c.dry ? m.full : tape.full;
    c.tcw = &tcw;
    c.tc_bwd = !c.simt && wcache != nullptr;      // c.simt on entry = caller asked for the exact-fp32 path
    RIFT_REQUIRE(c.dry || c.tc_bwd || !tape.fused, "backward: the exact-fp32 backward needs an exact-fp32 forward (the fused forward "
                                                   "keeps operand planes only)");
    c.simt = true;                                // (forward-style routing helpers are not used below)
    c.full = full;
    Tape dry_tape;
    if (c.dry) {                    // sizing pass: a tape with the right block counts, pointers unused
        dry_tape.enc.resize(m.enc.size()); dry_tape.dec.resize(m.dec.size());
        dry_tape.nat.blocks.resize(6);
        dry_tape.nat.NA = bs * A;
        const int Th = cfg.history_steps;
        dry_tape.nat.Ls[0] = Th - 1; dry_tape.nat.Ls[1] = (Th - 1 + 2 - 3) / 2 + 1; dry_tape.nat.Ls[2] = (dry_tape.nat.Ls[1] + 2 - 3) / 2 + 1;
        for (const FourierP* fp : {&m.pos_emb, &m.speed_emb, &m.r_pos_emb}) (void)fp;
        dry_tape.pos_emb.dims.resize(3); dry_tape.pos_emb.rows = rowsE;
        dry_tape.speed.dims.resize(1); dry_tape.speed.rows = bs * Mp;
        dry_tape.rpos_emb.dims.resize(3); dry_tape.rpos_emb.rows = NR;
        dry_tape.poly.groups = bs * Mp; dry_tape.poly.n = bt.P; dry_tape.poly.F.ld = 10;
        dry_tape.renc.groups = NR; dry_tape.renc.n = bt.Pr; dry_tape.renc.F.ld = 6;
        dry_tape.pi.x.rows = rowsQ; dry_tape.pi.x.ld = D;
    }
    const Tape& tp = c.dry ? dry_tape : tape;
    if (!c.dry && train_hi > train_lo)
        RIFT_CUDA_OK(cudaMemsetAsync(grads + train_lo, 0, (size_t)(train_hi - train_lo) * sizeof(float), c.st));

    // ---------------- pi_head
    if (!full) {
        if (m.pi_head.l0.train || m.pi_head.l3.train || m.pi_head.n.train) TRY(mlp_bwd(c, tp.pi, m.pi_head, dlogits, 1, nullptr));
        return 0;
    }
    ALLOC(dqf, float, (size_t)rowsQ * D);
    TRY(mlp_bwd(c, tp.pi, m.pi_head, dlogits, 1, dqf));

    // ---------------- cat_x_proj: qf = q Wa^T + (x_ego Wb^T + b)[row / (R*Mo)]
    const Lin ca = slice(m.cat_x_proj, 0, D, 0, D, false), cb = slice(m.cat_x_proj, 0, D, D, D, true);
    ALLOC(dq, float, (size_t)rowsQ * D);            // running gradient of the decoder residual stream
    c.n_inplace = 0;
    c.mark_inplace(dq);
    ALLOC(deg, float, (size_t)bs * D);
    ZALLOC(dXn, (size_t)rowsE * D);
    c.mark_inplace(dXn);
    // dqp: split-bf16 planes of dq whenever its last producer could write them (GEMM epilogue, LayerNorm backward); each
    // consumer below falls back to packing dq itself when they are off()
    Planes dqp;
    {
        LinBwdFuse fq;
        if (ln_bwd_fuse_on() && lin_bwd_all_tc(c, rowsQ, ca, true)) {
            dqp.Kp = tc_pitch(D);
            dqp.hi = c.alloc<uint16_t>((size_t)rowsQ * dqp.Kp);
            dqp.lo = c.alloc<uint16_t>((size_t)rowsQ * dqp.Kp);
            if (!dqp.hi || !dqp.lo) { set_last_error("workspace too small"); return -1; }
            fq.dX_planes = &dqp;
        }
        TRY(lin_bwd(c, tp.qlast, D, dqf, D, rowsQ, ca, dq, D, 0.f, false, nullptr, dqp.on() ? &fq : nullptr));
    }
    // dq as pre-packed dY of `L` when both of L's gradient products take the tensor-core route
    auto with_dq = [&](LinBwdFuse& f, const Lin& L) -> const LinBwdFuse* {
        if (!dqp.on() || !lin_bwd_all_tc(c, rowsQ, L, true)) return nullptr;
        f.dYp_in = &dqp;
        return &f;
    };
    if (!c.dry) TRY(launch_groupsum(dqf, D, bs, R * Mo, D, deg, D, 0, c.st));
    TRY(lin_bwd(c, tp.Xn.f, (long long)S * D, deg, D, bs, cb, dXn, (long long)S * D, 0.f));

    // ---------------- decoder blocks, last to first
    for (int l = (int)m.dec.size() - 1; l >= 0; --l) {
        const DecBlockP& db = m.dec[l];
        const DecBlockTape& dt = tp.dec[l];
        const Lin m2m_qk = slice(db.m2m.in, 0, 2 * D, 0, D, true), m2m_v = slice(db.m2m.in, 2 * D, D, 0, D, true);
        const Lin cr_q = slice(db.cross.in, 0, D, 0, D, true), cr_kv = slice(db.cross.in, D, 2 * D, 0, D, true);
        // (iv) ReLU FFN
        // ReLU'(.) from the post-activation values, or - after the fused forward, which keeps no fp32 hidden - from the pre-activation
        TRY(mlp_tail_bwd(c, rowsQ, D, 4 * D, dt.hm, dt.hpre4 ? dt.hpre4 : dt.hm.f, ACT_RELU, dt.t4, dt.ln4, db.n4, db.ffn0, db.ffn3, dq, &dqp));
        // (iii) cross-attention
        {
            ALLOC(d_a3, float, (size_t)rowsQ * D);
            ALLOC(d_qc, float, (size_t)rowsQ * D);
            ALLOC(d_kvc, float, (size_t)rowsE * 2 * D);
            ALLOC(d_t3, float, (size_t)rowsQ * D);
            LinBwdFuse f3;
            TRY(lin_bwd(c, dt.a3.f, D, dq, D, rowsQ, db.cross.out, d_a3, D, 0.f, true, &dt.a3.p, with_dq(f3, db.cross.out)));
            // dQ / dK / dV leave the attention backward as split-bf16 planes when both projections' gradient products run on
            // the tensor cores: no fp32 copy, no pack
            const bool ap = attn_planes_on(c) && attention_bwd_planes_ok(R * Mo, S, D / H) && lin_bwd_all_tc(c, rowsQ, cr_q, true) &&
                            lin_bwd_all_tc(c, rowsE, cr_kv, true);
            Planes pqc, pkvc;
            if (ap) { TRY(new_planes(c, rowsQ, D, &pqc)); TRY(new_planes(c, rowsE, 2 * D, &pkvc)); }
            if (!c.dry) {
                AttnArgs a = attn_cross(dt.qc, dt.kvc, bs, R * Mo, S, D, H, tp.key_pad, att_scale);
                a.lse = dt.lse3;
                AttnBwdPlanes pl;
                if (ap) { pl.q = pqc; pl.k = pkvc; pl.v = plane_cols(pkvc, D); }
                TRY(launch_attention_bwd(a, d_a3, D, ap ? nullptr : d_qc, D, ap ? nullptr : d_kvc, ap ? nullptr : d_kvc + D, 2 * D, 2 * D, c.st, pl));
            }
            LinBwdFuse fqc, fkvc;
            if (ap) { fqc.dYp_in = &pqc; fkvc.dYp_in = &pkvc; }
            TRY(lin_bwd(c, dt.t3.f, D, d_qc, D, rowsQ, cr_q, d_t3, D, 0.f, true, &dt.t3.p, ap ? &fqc : nullptr));
            // the K / V projection's gradients feed dXn, which the query chain never reads: branch stream
            // (in order there, so the accumulation into dXn stays sequential)
            TRY(fork_to(c, c.br));
            {
                OnStream on_br(c, c.br);
                TRY(lin_bwd(c, tp.Xn.f, D, d_kvc, 2 * D, rowsE, cr_kv, dXn, D, 1.f, true, &tp.Xn.p, ap ? &fkvc : nullptr));
            }
            // (ii) below starts from dq with the rows of padded reference lines cleared (their forward output was overwritten
            // with 0 -> no gradient through them): done by this LayerNorm backward on its way out
            TRY(ln_bwd(c, dt.ln3, db.n3, d_t3, nullptr, dq, 1, &dqp, rowsQ, tp.r_pad, Mo));
        }
        // (ii) m2m
        {
            ALLOC(d_a2, float, (size_t)rowsQ * D);
            ALLOC(dqkv2, float, (size_t)rowsQ * 3 * D);
            ALLOC(d_t2, float, (size_t)rowsQ * D);
            ALLOC(sc, float, (size_t)Mo * D);
            LinBwdFuse f2o;
            TRY(lin_bwd(c, dt.a2.f, D, dq, D, rowsQ, db.m2m.out, d_a2, D, 0.f, true, &dt.a2.p, with_dq(f2o, db.m2m.out)));
            const bool ap = attn_planes_on(c) && attention_bwd_planes_ok(Mo, Mo, D / H) && lin_bwd_all_tc(c, rowsQ, m2m_qk, true) &&
                            lin_bwd_all_tc(c, rowsQ, m2m_v, true);
            Planes pqk2, pv2;
            if (ap) { TRY(new_planes(c, rowsQ, 2 * D, &pqk2)); TRY(new_planes(c, rowsQ, D, &pv2)); }
            if (!c.dry) {
                AttnArgs a = attn_m2m(dt.qkv2, NR, Mo, D, H, att_scale);
                a.lse = dt.lse2;
                AttnBwdPlanes pl;
                if (ap) { pl.q = pqk2; pl.k = plane_cols(pqk2, D); pl.v = pv2; }
                TRY(launch_attention_bwd(a, d_a2, D, ap ? nullptr : dqkv2, 3 * D, ap ? nullptr : dqkv2 + D, ap ? nullptr : dqkv2 + 2 * D, 3 * D,
                                         3 * D, c.st, pl));
            }
            LinBwdFuse fqk2, fv2;
            if (ap) { fqk2.dYp_in = &pqk2; fv2.dYp_in = &pv2; }
            TRY(lin_bwd(c, dt.t2p.f, D, dqkv2, 3 * D, rowsQ, m2m_qk, d_t2, D, 0.f, true, &dt.t2p.p, ap ? &fqk2 : nullptr));   // d(LN2 out + m_pos)
            if (m.m_pos.train && !c.dry) TRY(launch_modsum(d_t2, D, rowsQ, D, Mo, m.m_pos.d, 1, c.st));
            (void)sc;
            TRY(lin_bwd(c, dt.t2.f, D, dqkv2 + 2 * D, 3 * D, rowsQ, m2m_v, d_t2, D, 1.f, true, &dt.t2.p, ap ? &fv2 : nullptr));   // + value path
            TRY(ln_bwd(c, dt.ln2, db.n2, d_t2, nullptr, dq, 1, &dqp, rowsQ));
        }
        // (i) r2r
        {
            ALLOC(d_a1, float, (size_t)rowsQ * D);
            ALLOC(dqkv1, float, (size_t)rowsQ * 3 * D);
            ALLOC(d_t1, float, (size_t)rowsQ * D);
            LinBwdFuse f1o;
            TRY(lin_bwd(c, dt.a1.f, D, dq, D, rowsQ, db.r2r.out, d_a1, D, 0.f, true, &dt.a1.p, with_dq(f1o, db.r2r.out)));
            const bool ap = attn_planes_on(c) && attention_bwd_planes_ok(R, R, D / H) && lin_bwd_all_tc(c, rowsQ, db.r2r.in, true);
            Planes p1;
            if (ap) TRY(new_planes(c, rowsQ, 3 * D, &p1));
            if (!c.dry) {
                AttnArgs a = attn_r2r(dt.qkv1, bs, R, Mo, D, H, tp.r_pad_r2r, att_scale);
                a.kpm_mod = tp.r2r_mod; a.kpm_off = tp.r2r_off;
                a.lse = dt.lse1;
                AttnBwdPlanes pl;
                if (ap) { pl.q = p1; pl.k = plane_cols(p1, D); pl.v = plane_cols(p1, 2 * D); }
                TRY(launch_attention_bwd(a, d_a1, D, ap ? nullptr : dqkv1, 3 * D, ap ? nullptr : dqkv1 + D, ap ? nullptr : dqkv1 + 2 * D, 3 * D,
                                         3 * D, c.st, pl));
            }
            LinBwdFuse f1i;
            if (ap) f1i.dYp_in = &p1;
            TRY(lin_bwd(c, dt.t1.f, D, dqkv1, 3 * D, rowsQ, db.r2r.in, d_t1, D, 0.f, true, &dt.t1.p, ap ? &f1i : nullptr));
            TRY(ln_bwd(c, dt.ln1, db.n1, d_t1, nullptr, dq, 1, &dqp, rowsQ));
        }
    }

    // dXn is complete once the branch stream has run the last K / V projection backward (recorded before the
    // parameter-only branches are queued behind it)
    cudaEvent_t dxn_done = nullptr;
    if (!c.dry && c.br) { dxn_done = c.next_event(); RIFT_CUDA_OK(cudaEventRecord(dxn_done, c.br)); }

    // ---------------- query init: q0[row] = u[row / Mo] + v[row % Mo]
    {
        const Lin qa = slice(m.q_proj, 0, D, 0, D, false), qb = slice(m.q_proj, 0, D, D, D, true);
        ALLOC(du, float, (size_t)NR * D);
        ALLOC(dv, float, (size_t)Mo * D);
        ALLOC(d_remb, float, (size_t)NR * D);
        if (!c.dry) {
            TRY(launch_groupsum(dq, D, NR, Mo, D, du, D, 0, c.st));
            TRY(launch_modsum(dq, D, rowsQ, D, Mo, dv, 0, c.st));
        }
        TRY(lin_bwd(c, tp.r_emb.f, D, du, D, NR, qa, d_remb, D, 0.f, false));
        TRY(lin_bwd(c, m.m_emb.p, D, dv, D, Mo, qb, m.m_emb.train ? m.m_emb.d : nullptr, D, 1.f));
        // r_emb = r_encoder(points) + r_pos_emb(first point): parameter-only sub-graphs -> branch stream
        TRY(fork_to(c, c.br));
        {
            OnStream on_br(c, c.br);
            TRY(fourier_bwd(c, tp.rpos_emb, m.r_pos_emb, d_remb));
            TRY(points_bwd(c, tp.renc, m.r_enc, d_remb));
        }
    }

    // ---------------- scene encoding: final norm <- encoder blocks
    ALLOC(dX, float, (size_t)rowsE * D);
    c.mark_inplace(dX);
    if (dxn_done) RIFT_CUDA_OK(cudaStreamWaitEvent(c.st, dxn_done, 0));
    Planes dXp;                                      // planes of the encoder stream's gradient, same protocol as dqp
    TRY(ln_bwd(c, tp.ln_final, m.final_norm, dXn, nullptr, dX, 0, &dXp, rowsE));
    for (int l = (int)m.enc.size() - 1; l >= 0; --l) {
        const EncBlockP& eb = m.enc[l];
        const EncBlockTape& et = tp.enc[l];
        TRY(mlp_tail_bwd(c, rowsE, D, 4 * D, et.hm, et.hpre, ACT_GELU, et.t2, et.ln2, eb.n2, eb.fc1, eb.fc2, dX, &dXp));
        ALLOC(d_att, float, (size_t)rowsE * D);
        ALLOC(dqkv, float, (size_t)rowsE * 3 * D);
        ALLOC(d_t1, float, (size_t)rowsE * D);
        LinBwdFuse fo;
        if (dXp.on() && lin_bwd_all_tc(c, rowsE, eb.attn.out, true)) fo.dYp_in = &dXp;
        TRY(lin_bwd(c, et.att.f, D, dX, D, rowsE, eb.attn.out, d_att, D, 0.f, true, &et.att.p, fo.dYp_in ? &fo : nullptr));
        const bool ap = attn_planes_on(c) && attention_bwd_planes_ok(S, S, D / H) && lin_bwd_all_tc(c, rowsE, eb.attn.in, true);
        Planes pe;
        if (ap) TRY(new_planes(c, rowsE, 3 * D, &pe));
        if (!c.dry) {
            AttnArgs a = attn_self(et.qkv, bs, S, D, H, tp.key_pad, att_scale);
            a.lse = et.lse;
            AttnBwdPlanes pl;
            if (ap) { pl.q = pe; pl.k = plane_cols(pe, D); pl.v = plane_cols(pe, 2 * D); }
            TRY(launch_attention_bwd(a, d_att, D, ap ? nullptr : dqkv, 3 * D, ap ? nullptr : dqkv + D, ap ? nullptr : dqkv + 2 * D, 3 * D, 3 * D,
                                     c.st, pl));
        }
        LinBwdFuse fei;
        if (ap) fei.dYp_in = &pe;
        TRY(lin_bwd(c, et.t1.f, D, dqkv, 3 * D, rowsE, eb.attn.in, d_t1, D, 0.f, true, &et.t1.p, ap ? &fei : nullptr));
        TRY(ln_bwd(c, et.ln1, eb.n1, d_t1, nullptr, dX, 1, &dXp, rowsE));
    }

    // ---------------- tokens = [agent tokens ; map tokens] + pos_emb
    // dX is final from here on (only read below): pos_emb's parameters on the branch stream
    TRY(fork_to(c, c.br));
    {
        OnStream on_br(c, c.br);
        TRY(fourier_bwd(c, tp.pos_emb, m.pos_emb, dX));
    }
    ALLOC(esc, float, (size_t)148 * 4 * D);
    // map side
    if (Mp > 0) {
        const int NP = bs * Mp;
        ALLOC(dx_poly, float, (size_t)NP * D);
        ALLOC(dx_speed, float, (size_t)NP * D);
        if (!c.dry) {
            TRY(launch_masked_gather_rows(dX, D, A, Mp, S, nullptr, NP, D, dx_poly, 0, c.st));
            TRY(launch_masked_gather_rows(dX, D, A, Mp, S, bt.map_polygon_has_speed_limit, NP, D, dx_speed, 0, c.st));
            if (m.map_type_emb.train)
                TRY(launch_embedding_bwd(dX, D, A, Mp, S, bt.map_polygon_type, NP, D, 3, m.map_type_emb.d, 0, esc, c.st));
            if (m.map_route_emb.train)
                TRY(launch_embedding_bwd(dX, D, A, Mp, S, reinterpret_cast<const int8_t*>(bt.map_polygon_on_route), NP, D, 2,
                                         m.map_route_emb.d, 0, esc, c.st));
            if (m.map_tl_emb.train)
                TRY(launch_embedding_bwd(dX, D, A, Mp, S, bt.map_polygon_tl_status, NP, D, 4, m.map_tl_emb.d, 0, esc, c.st));
            if (m.map_unknown_emb.train)
                TRY(launch_embedding_bwd(dX, D, A, Mp, S, reinterpret_cast<const int8_t*>(bt.map_polygon_has_speed_limit), NP, D, 1,
                                         m.map_unknown_emb.d, 1, esc, c.st));
        }
        TRY(fork_to(c, c.br));
        {
            OnStream on_br(c, c.br);
            TRY(points_bwd(c, tp.poly, m.poly_enc, dx_poly));
            TRY(fourier_bwd(c, tp.speed, m.speed_emb, dx_speed));
        }
    }
    // agent side
    {
        const int NA = bs * A;
        ALLOC(dx_hist, float, (size_t)NA * D);
        ALLOC(dx_ego, float, (size_t)bs * D);
        if (!c.dry) {
            TRY(launch_masked_gather_rows(dX, D, 0, A, S, tp.agent_any, NA, D, dx_hist, 1, c.st));     // token 0 is the ego embedding
            TRY(launch_masked_gather_rows(dX, D, 0, 1, S, nullptr, bs, D, dx_ego, 0, c.st));
            if (m.agent_type_emb.train)
                TRY(launch_embedding_bwd(dX, D, 0, A, S, bt.agent_category, NA, D, 4, m.agent_type_emb.d, 0, esc, c.st));
        }
        // ---- StateAttentionEncoder: parameter-only from here (its inputs are raw states) -> branch stream
        TRY(fork_to(c, c.br));
        {
            OnStream on_br(c, c.br);
            ALLOC(esc_ego, float, (size_t)148 * D);      // private reduction scratch (esc is used by the main stream below)
            const EgoTape& et = tp.ego;
            const int ntok = cfg.state_channel, eh = 4;
            const Lin in_q = slice(m.ego.attn.in, 0, D, 0, D, true), in_kv = slice(m.ego.attn.in, D, 2 * D, 0, D, true);
            ALLOC(d_eo, float, (size_t)bs * D);
            ALLOC(dq_b, float, (size_t)bs * D);
            ALLOC(d_kv, float, (size_t)bs * ntok * 2 * D);
            ALLOC(d_qv, float, (size_t)D);
            ALLOC(d_toks, float, (size_t)bs * ntok * D);
            ALLOC(dwb, float, (size_t)2 * ntok * D);
            TRY(lin_bwd(c, et.eo.f, D, dx_ego, D, bs, m.ego.attn.out, d_eo, D, 0.f, true, &et.eo.p));
            if (!c.dry) {
                AttnArgs a = attn_ego(et.qv, et.kv, bs, ntok, D, eh);
                a.lse = et.lse;
                TRY(launch_attention_bwd(a, d_eo, D, dq_b, D, d_kv, d_kv + D, 2 * D, 2 * D, c.st));
                TRY(launch_colsum(dq_b, D, bs, D, d_qv, 0, esc_ego, c.st));
            }
            TRY(lin_bwd(c, m.ego.query.p, D, d_qv, D, 1, in_q, m.ego.query.train ? m.ego.query.d : nullptr, D, 1.f));
            TRY(lin_bwd(c, et.toks, D, d_kv, 2 * D, bs * ntok, in_kv, d_toks, D, 0.f));
            if (!c.dry) {
                if (m.ego.pos_embed.train) TRY(launch_modsum(d_toks, D, (long long)bs * ntok, D, ntok, m.ego.pos_embed.d, 1, c.st));
                TRY(launch_state_tokens_bwd(bt.current_state, bt.cs_stride, bs, ntok, D, d_toks, dwb, dwb + (size_t)ntok * D, c.st));
                for (int i = 0; i < ntok; ++i) {
                    if (m.ego.lin[i].train && m.ego.lin[i].dW) TRY(launch_add_inplace(m.ego.lin[i].dW, dwb + (size_t)i * D, D, c.st));
                    if (m.ego.lin[i].train && m.ego.lin[i].db) TRY(launch_add_inplace(m.ego.lin[i].db, dwb + (size_t)(ntok + i) * D, D, c.st));
                }
            }
        }
        // ---- NATSequenceEncoder
        {
            const NatTape& nt = tp.nat;
            const int* Ls = nt.Ls;
            ALLOC(d_colF, float, (size_t)NA * 3 * D);
            ALLOC(dlat0, float, (size_t)NA * Ls[0] * D);
            ZALLOC(dlat1, (size_t)NA * Ls[1] * D);
            ZALLOC(dlat2, (size_t)NA * Ls[2] * D);
            TRY(lin_bwd(c, nt.colF.f, 3 * D, dx_hist, D, NA, m.hist.fpn, d_colF, 3 * D, 0.f, true, &nt.colF.p));
            if (!c.dry) {
                TRY(launch_col2im_k3_last(d_colF, NA, Ls[0], D, dlat0, c.st));
                TRY(launch_fpn_upsample_add_bwd(dlat0, dlat1, NA, Ls[0], Ls[1], D, c.st));
                TRY(launch_fpn_upsample_add_bwd(dlat1, dlat2, NA, Ls[1], Ls[2], D, c.st));
            }
            float* dlat[3] = {dlat0, dlat1, dlat2};
            float* dxn = nullptr;                   // gradient wrt the input of the level above (down-sampled stream)
            for (int i = 2; i >= 0; --i) {
                const NatLevelP& lv = m.hist.levels[i];
                const int d = lv.dim, L = Ls[i], rows = NA * L;
                ALLOC(d_colL, float, (size_t)rows * 3 * d);
                ALLOC(d_o, float, (size_t)rows * d);
                ALLOC(dx, float, (size_t)rows * d);
                c.mark_inplace(dx);
                TRY(lin_bwd(c, nt.colL[i].f, 3 * d, dlat[i], D, rows, m.hist.lateral[i], d_colL, 3 * d, 0.f, true, &nt.colL[i].p));
                if (!c.dry) TRY(launch_col2im_k3(d_colL, NA, L, d, 1, d_o, 0, c.st));
                Planes dxp;                          // planes of dx (valid only while no other kernel has added into dx)
                TRY(ln_bwd(c, nt.ln_lev[i], m.hist.norms[i], d_o, nullptr, dx, 0, lv.has_down ? nullptr : &dxp, rows));
                if (lv.has_down) {
                    const int Ln = Ls[i + 1];
                    ALLOC(d_xd, float, (size_t)NA * Ln * 2 * d);
                    ALLOC(d_colD, float, (size_t)NA * Ln * 3 * d);
                    TRY(ln_bwd(c, nt.ln_down[i], lv.down_n, dxn, nullptr, d_xd, 0));
                    TRY(lin_bwd(c, nt.colD[i].f, 3 * d, d_xd, 2 * d, NA * Ln, lv.down, d_colD, 3 * d, 0.f, true, &nt.colD[i].p));
                    if (!c.dry) TRY(launch_col2im_k3(d_colD, NA, L, d, 2, dx, 1, c.st));
                }
                for (int j = 1; j >= 0; --j) {
                    const NatBlockP& nb = lv.blocks[j];
                    const NatBlockTape& bt_ = nt.blocks[i * 2 + j];
                    TRY(mlp_tail_bwd(c, rows, d, 3 * d, bt_.hm, bt_.hpre, ACT_GELU, bt_.t2, bt_.ln2, nb.n2, nb.fc1, nb.fc2, dx, &dxp));
                    ALLOC(d_att, float, (size_t)rows * d);
                    ALLOC(dqkv, float, (size_t)rows * 3 * d);
                    ALLOC(d_t1, float, (size_t)rows * d);
                    const int nrel = 2 * lv.ksize - 1;
                    ALLOC(drpb, float, (size_t)NA * lv.heads * nrel);
                    LinBwdFuse fp;
                    if (dxp.on() && lin_bwd_all_tc(c, rows, nb.proj, true)) fp.dYp_in = &dxp;
                    TRY(lin_bwd(c, bt_.att.f, d, dx, d, rows, nb.proj, d_att, d, 0.f, true, &bt_.att.p, fp.dYp_in ? &fp : nullptr));
                    const bool ap = attn_planes_on(c) && nat_attention_bwd_planes_ok(L, lv.ksize) && ((3 * d) % 64) == 0 &&
                                    lin_bwd_all_tc(c, rows, nb.qkv, true);
                    Planes pn;
                    if (ap) TRY(new_planes(c, rows, 3 * d, &pn));
                    if (!c.dry) {
                        TRY(launch_nat_attention_bwd(bt_.qkv, d_att, NA, L, lv.heads, d / lv.heads, lv.ksize, nb.rpb.p, ap ? nullptr : dqkv,
                                                     nb.rpb.train ? drpb : nullptr, c.st, pn));
                        if (nb.rpb.train) TRY(launch_colsum(drpb, lv.heads * nrel, NA, lv.heads * nrel, nb.rpb.d, 1, esc, c.st));
                    }
                    LinBwdFuse fn;
                    if (ap) fn.dYp_in = &pn;
                    TRY(lin_bwd(c, bt_.t1.f, d, dqkv, 3 * d, rows, nb.qkv, d_t1, d, 0.f, true, &bt_.t1.p, ap ? &fn : nullptr));
                    TRY(ln_bwd(c, bt_.ln1, nb.n1, d_t1, nullptr, dx, 1, &dxp, rows));
                }
                dxn = dx;
            }
            TRY(lin_bwd(c, nt.col0.f, 27, dxn, m.hist.embed.N, NA * Ls[0], m.hist.embed, nullptr, 0, 0.f, true, &nt.col0.p));
        }
    }
    return 0;
}
