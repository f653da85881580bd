"""GPU: the policy forward, the pi_head backward and whole policy-update steps through the public
API, against (a) golden vectors written by the unmodified reference and (b) the CPU oracle on
larger seeded inputs.  Tolerance from BASELINE.json north_star: logits and loss within 1e-3
relative (fp32), indices exact."""
import numpy as np
import pytest
import torch

from rift_b200.planning_model import PlanningModel
from rift_b200.trainer import TRAINERS
from rift_b200.config import MODEL_ZOO
from rift_b200.synth import synth_state_dict, synth_features, synth_rl_extras
from tests.helpers import CASES, case_inputs, golden, check_golden, oracle_losses, to_torch_tree

pytestmark = pytest.mark.gpu

RTOL = 1e-3
TRAINER_KW = dict(lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3, frame_rate=10)


def build(cfg, sd, trainable=()):
    m = PlanningModel.from_config(cfg, trainable_layers=trainable)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m


def make_batch(feats, ex):
    b = {"cur_pluto_feature_torch": to_torch_tree(feats, "cuda")}
    for k in ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits", "return"):
        b[k + "_torch"] = torch.from_numpy(ex[k].copy()).cuda()
    return b


def valid_logits(prob, feats):
    keep = torch.from_numpy(feats["reference_line"]["valid_mask"].any(-1))
    return prob[keep]


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(name):
    cfg, sd, feats, _ = case_inputs(name)
    g = golden(name)
    out = build(cfg, sd)(to_torch_tree(feats, "cuda"))
    for k in ("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory", "candidate_trajectories"):
        check_golden(g, "out_" + k, out[k].cpu().numpy(), rtol=RTOL)
    # padded reference lines carry exactly the reference's fill value
    pad = ~torch.from_numpy(feats["reference_line"]["valid_mask"].any(-1))
    assert torch.equal(out["r_padding_mask"].cpu(), pad)
    assert (out["probability"].cpu()[pad] == -1e6).all()
    best = out["probability"].reshape(out["probability"].shape[0], -1).argmax(-1).cpu().numpy()
    assert np.array_equal(best, g["out_best_index"])


@pytest.mark.parametrize("model,bs,A,Mp,R", [("small", 8, 16, 20, 6), ("medium", 4, 32, 20, 6), ("small", 2, 49, 150, 6)])
def test_forward_matches_oracle_larger(model, bs, A, Mp, R):
    cfg = MODEL_ZOO[model]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, bs, A, Mp, R, seed=3, ragged=True)
    ex = synth_rl_extras(cfg, feats, seed=4)
    with torch.no_grad():
        _, ref, _ = oracle_losses(cfg, sd, feats, ex, "rift")
    out = build(cfg, sd)(to_torch_tree(feats, "cuda"))
    for k in ("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory"):
        a, b = out[k].cpu(), ref[k]
        err = (a - b).abs().max().item()
        assert err <= RTOL * b.abs().max().item(), (k, err)
    a, b = valid_logits(out["probability"].cpu(), feats), valid_logits(ref["probability"], feats)
    assert (a - b).abs().max().item() <= RTOL * b.abs().max().item()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("algo", ["rift", "grpo", "reinforce"])
def test_training_step_loss_and_pi_head_grads_golden(name, algo):
    cfg, sd, feats, ex = case_inputs(name)
    g = golden(name)
    model = build(cfg, sd)
    tr = TRAINERS[algo](model, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    loss = tr.training_step(make_batch(feats, ex))
    ref = float(g[f"loss_{algo}"])
    assert abs(float(loss) - ref) <= RTOL * max(abs(ref), 1e-3)
    if algo == "reinforce":
        return
    count = float(tr._count) if tr._count is not None else 1.0
    for k in g.files:
        if k.startswith(f"grad_{algo}/") and not k.endswith("@stats"):
            n = k.split("/", 1)[1].split("@")[0]
            got = (model.arena.grad_view(n) / count).cpu().numpy()
            check_golden(g, f"grad_{algo}/{n}", got, rtol=RTOL, atol=2e-6)


FULL_LAYERS = ["pos_emb", "agent_encoder", "map_encoder", "static_objects_encoder", "encoder_blocks", "norm",
               "agent_predictor", "planning_decoder", "hidden_proj", "ref_free_decoder"]


# Element-wise gradient tolerance (fraction of the tensor's max |g|).  With exact-fp32 GEMMs the CUDA
# gradients match autograd to 2e-3 everywhere.  With the split-bf16 tensor-core forward (activations
# accurate to ~1e-5) a handful of ReLU / max-pool decisions sitting within 1e-5 of their threshold flip,
# which moves single gradient elements by up to ~0.5 % of the tensor maximum; tensor norms still agree to 2e-3.
GRAD_ETOL = {True: 2e-3, False: 3e-2, "bwd_only": 2e-3}


@pytest.mark.parametrize("exact", [True, False], ids=["fp32", "tcgen05"])
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("algo", ["rift", "grpo"])
def test_full_backward_matches_reference_golden(name, algo, exact):
    """Every parameter gradient of the whole policy (all modules trainable) against the reference's autograd:
    per-tensor L2 norm and sum for all 426 tensors, complete tensors for a sample of them."""
    from rift_b200.config import param_spec, is_buffer
    cfg, sd, feats, ex = case_inputs(name)
    g = golden(name)
    model = build(cfg, sd)
    model.exact_fp32 = exact
    tr = TRAINERS[algo](model, trainable_layers=FULL_LAYERS, **TRAINER_KW)
    loss = tr.training_step(make_batch(feats, ex))
    ref = float(g[f"loss_{algo}"])
    assert abs(float(loss) - ref) <= RTOL * max(abs(ref), 1e-3)
    count = float(tr._count)
    names = [n for n, _, _ in param_spec(cfg) if not is_buffer(n)]
    stats = g[f"fullgrad_{algo}_stats"]
    gmax = stats[:, 1].max()
    bad = []
    for i, n in enumerate(names):
        if n not in model.arena.trainable:
            assert stats[i, 1] == 0.0, f"{n} has a reference gradient but is not in the trainable arena"
            continue
        gv = (model.arena.grad_view(n).double() / count).cpu()
        l2 = float(gv.pow(2).sum().sqrt())
        tol = 2e-3 * stats[i, 1] + 1e-6 * gmax
        if abs(l2 - stats[i, 1]) > tol or abs(float(gv.sum()) - stats[i, 0]) > 2e-3 * stats[i, 1] * np.sqrt(gv.numel()) + 1e-6 * gmax:
            bad.append((n, l2, stats[i, 1], float(gv.sum()), stats[i, 0]))
    assert not bad, f"{len(bad)} tensors differ, first: {bad[:5]}"
    for k in g.files:
        if k.startswith(f"fullgrad_{algo}/") and not k.endswith("@stats"):
            n = k.split("/", 1)[1].split("@")[0]
            got = (model.arena.grad_view(n) / count).cpu().numpy()
            check_golden(g, f"fullgrad_{algo}/{n}", got, rtol=GRAD_ETOL[exact], atol=1e-6 * gmax)


@pytest.mark.parametrize("exact", [True, False, "bwd_only"], ids=["fp32", "tcgen05", "fp32fwd_tcgen05bwd"])
def test_full_backward_matches_oracle_autograd_cfg2_like(exact):
    """A larger ragged batch: CUDA gradients vs torch autograd through the CPU oracle, tensor by tensor."""
    cfg = MODEL_ZOO["small"]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, 6, 12, 14, 4, seed=3, ragged=True)
    ex = synth_rl_extras(cfg, feats, seed=4)
    model = build(cfg, sd)
    model.exact_fp32 = bool(exact)          # "bwd_only": exact forward (no ReLU / arg-max flips), tensor-core backward
    if exact == "bwd_only":
        model.exact_bwd = False
    tr = TRAINERS["grpo"](model, trainable_layers=FULL_LAYERS, **TRAINER_KW)
    loss = tr.training_step(make_batch(feats, ex))
    count = float(tr._count)
    prefixes = tuple(p for p in FULL_LAYERS)
    ref_loss, _, sdt = oracle_losses(cfg, sd, feats, ex, "grpo", requires_grad=prefixes)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss.detach())) <= RTOL * abs(float(ref_loss.detach()))
    gmax = max(float(t.grad.abs().max()) for t in sdt.values() if t.grad is not None)
    for n in model.arena.trainable:
        ref = sdt[n].grad
        assert ref is not None, n
        got = (model.arena.grad_view(n) / count).cpu()
        err = float((got - ref).abs().max())
        assert err <= GRAD_ETOL[exact] * float(ref.abs().max()) + 1e-6 * gmax, (n, err, float(ref.abs().max()))


def test_three_policy_updates_match_reference_parameters():
    """forward -> GRPO loss -> backward -> clip 0.5 -> AdamW, three times, vs the reference's torch loop."""
    name = "cfg1_small"
    cfg, sd, feats, ex = case_inputs(name)
    g = golden(name)
    model = build(cfg, sd)
    tr = TRAINERS["grpo"](model, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    tr.configure_optimizers()
    tr.optimizer.param_groups[0]["lr"] = 1e-4        # the golden loop holds lr at 1e-4 (no scheduler)
    batch = make_batch(feats, ex)
    for step in range(3):
        loss = tr.step(batch)
        if step > 0:
            ref = float(g[f"loss_grpo_step{step}"])
            assert abs(float(loss) - ref) <= RTOL * max(abs(ref), 1e-3)
        gn = float(g[f"gradnorm_grpo_step{step}"])
        assert abs(tr.optimizer.grad_norm() - gn) <= RTOL * gn
        if step in (0, 2):
            for n in model.arena.trainable:
                key = f"param_grpo_step{step + 1}/{n}"
                if key not in g.files:
                    continue
                # elements whose gradient is round-off sized move by noise in the reference itself
                sel = np.abs(g[f"grad_grpo/{n}"]) > 1e-6
                err = np.abs(model.arena.view(n).cpu().numpy() - g[key])[sel]
                assert err.size == 0 or err.max() <= 0.05 * 1e-4, (key, err.max())


@pytest.mark.parametrize("through", ["trainer", "model"])
def test_state_dict_load_after_graph_capture_takes_effect(through):
    """A captured step graph reads frozen weights through pre-split planes: loading a checkpoint AFTER the graph exists must
    change what the next step computes (trainer.load_state_dict drops the graphs, PlanningModel.load_state_dict re-splits every
    plane), frozen layers included."""
    cfg, sd, feats, ex = case_inputs("cfg1_small")
    model = build(cfg, sd)
    tr = TRAINERS["grpo"](model, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    tr.configure_optimizers()
    batch = make_batch(feats, ex)
    for _ in range(3):                                   # the second call captures, the third replays
        tr.step(batch)
    sd2 = {k: (v * 1.05 + 0.01).astype(v.dtype) if v.dtype.kind == "f" else v for k, v in sd.items()}
    t2 = {k: torch.from_numpy(v) for k, v in sd2.items()}
    if through == "trainer":
        tr.load_state_dict({"model." + k: v for k, v in t2.items()})
    else:
        model.load_state_dict(t2)
    got = [float(tr.step(batch)) for _ in range(2)][0]   # first step after the load: eager or replay, must use the new weights
    fresh = build(cfg, sd2)
    tr2 = TRAINERS["grpo"](fresh, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    tr2.configure_optimizers()
    want = float(tr2.step(batch))
    stale = float(TRAINERS["grpo"](build(cfg, sd), trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW).training_step(batch))
    assert abs(got - want) <= 1e-5 * max(abs(want), 1e-3), (got, want)
    assert abs(want - stale) > 1e-3 * max(abs(want), 1e-3), "the two checkpoints must differ for the test to mean anything"


def test_state_dict_roundtrip_and_prefixed_checkpoint():
    cfg, sd, _, _ = case_inputs("cfg1_small")
    model = build(cfg, sd)
    back = model.state_dict()
    for k, v in sd.items():
        assert np.array_equal(back[k].cpu().numpy(), v), k
    tr = TRAINERS["rift"](model, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    ck = tr.state_dict()
    assert all(k.startswith("model.") for k in ck)
    model2 = PlanningModel.from_config(cfg)
    model2.load_state_dict(ck)                       # training checkpoint keys (pluto.py:135-137)
    for k, v in sd.items():
        assert np.array_equal(model2.state_dict()[k].cpu().numpy(), v), k
    with pytest.raises(ValueError, match="not found in the model"):
        TRAINERS["rift"](model, trainable_layers=["planning_decoder.nope"], **TRAINER_KW)


def test_freezing_preserves_values_and_forward():
    cfg, sd, feats, _ = case_inputs("ragged_small")
    model = build(cfg, sd)
    d = to_torch_tree(feats, "cuda")
    p0 = model(d)["probability"].clone()
    model.set_trainable_layers(["planning_decoder.pi_head"])      # arena is re-laid out
    assert torch.equal(model(d)["probability"], p0)


def test_tensor_core_path_agrees_with_exact_fp32_path():
    """The tcgen05 split-bf16 GEMMs against the exact-fp32 SIMT GEMMs of the same library, whole forward."""
    cfg = MODEL_ZOO["medium"]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, 16, 32, 20, 6, seed=5, ragged=True)
    d = to_torch_tree(feats, "cuda")
    model = build(cfg, sd)
    model.exact_fp32 = False
    a = model(d)
    model.exact_fp32 = True
    b = model(d)
    keep = torch.from_numpy(feats["reference_line"]["valid_mask"].any(-1))
    for k in ("probability", "trajectory", "hidden", "prediction"):
        x, y = (a[k].cpu()[keep], b[k].cpu()[keep]) if k == "probability" else (a[k].cpu(), b[k].cpu())
        assert (x - y).abs().max().item() <= 2e-4 * y.abs().max().item(), k


def test_cfg2_shape_properties():
    """BASELINE configs[1] shape (64 x 32 agents, R=6, Pluto-medium): finite outputs, exact padding,
    softmax-gradient rows sum to zero, loss decreases over a few updates on a fixed batch."""
    cfg = MODEL_ZOO["medium"]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, 64, 32, 20, 6, seed=1)
    ex = synth_rl_extras(cfg, feats, seed=2)
    model = build(cfg, sd)
    tr = TRAINERS["grpo"](model, trainable_layers=["planning_decoder.pi_head"], **TRAINER_KW)
    tr.configure_optimizers()
    tr.optimizer.param_groups[0]["lr"] = 1e-3
    batch = make_batch(feats, ex)
    losses = [float(tr.step(batch)) for _ in range(6)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]
    g = model.arena.grad_view("planning_decoder.pi_head.mlp.3.bias")
    assert abs(float(g.sum())) < 1e-3 * max(float(model.arena.grads.abs().max()), 1e-6)


# BASELINE.json configs at their own shapes: cfg2 exactly, a ragged variant of it, and the
# per-rank shard of configs[3] (256 x 48 over 8 GPUs = 32 samples of 48 agents) - Pluto-medium, every module trainable.
BASELINE_SHAPES = {
    "cfg2": dict(bs=64, A=32, Mp=20, R=6, ragged=False, algo="grpo", clip=(0.8, 1.2)),
    "cfg2_ragged": dict(bs=64, A=32, Mp=20, R=6, ragged=True, algo="grpo", clip=(0.8, 1.2)),
    "cfg4_rank": dict(bs=32, A=48, Mp=20, R=6, ragged=False, algo="grpo", clip=(0.8, 1.2)),
}


# Norm tolerance per gradient tensor.  With the exact-fp32 forward the tensor-core backward holds 2e-3 on every tensor.
# With the split-bf16 forward (activations accurate to ~1e-5) a few ReLU / arg-max / max-pool decisions within 1e-5 of
# their threshold differ from the CPU oracle's; at 64 x 72 candidates that moves the norm of a handful of SMALL tensors
# (biases, rpb) by up to ~3e-3 - the flipped decisions are a property of the forward rounding, not of the backward
# arithmetic, which the "fp32fwd_tcgen05bwd" variant shows.
# Measured on a B200 (profiles/r2_parity_baseline_shapes.txt), worst tensor of cfg2 / ragged cfg2 / cfg4-per-rank:
#   exact forward + tensor-core backward : norm 3.9e-4, relative L2 2.9e-3, element 5.8e-3, 382 of 17.0 M elements beyond 2e-3
#   tensor-core forward + backward       : norm 4.2e-3, relative L2 2.9e-2, element 1.1e-1, 26 541 of 17.0 M elements (0.16 %)
# norm = | ||g|| - ||g_ref|| | / ||g_ref||, l2 = ||g - g_ref|| / ||g_ref||, element = max |g - g_ref| / (max |g_ref| + 1e-3 gmax).
PARITY_TOL = {"tcgen05": dict(norm=6e-3, l2=4e-2, element=0.15, over=3e-3),
              # (exact forward: the few ReLU / arg-max decisions within one ulp of their threshold still flip against the CPU
              #  oracle - which ones depends on summation order inside the forward kernels; one flipped unit of a 1024-wide FFN
              #  moves single gradient elements by ~2 % of their tensor's max.  The COUNT bound is the regression guard.)
              "fp32fwd_tcgen05bwd": dict(norm=2e-3, l2=5e-3, element=3e-2, over=1e-4)}


@pytest.mark.parametrize("mode", ["tcgen05", "fp32fwd_tcgen05bwd"])
@pytest.mark.parametrize("name", list(BASELINE_SHAPES))
def test_baseline_shape_parity_vs_oracle(name, mode):
    """The benchmarked configuration itself against the CPU oracle: logits and loss at 1e-3, the L2 norm of EVERY
    gradient tensor (NORM_TOL), element-wise errors bounded by GRAD_ETOL, and the number of elements beyond 2e-3 of
    their tensor's maximum recorded and bounded (at most 0.2 % of all gradient elements), so that a regression cannot
    hide among the flipped ReLU / arg-max decisions."""
    s = BASELINE_SHAPES[name]
    cfg = MODEL_ZOO["medium"]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, s["bs"], s["A"], s["Mp"], s["R"], seed=1, ragged=s["ragged"])
    ex = synth_rl_extras(cfg, feats, seed=2)
    model = build(cfg, sd)
    if mode == "fp32fwd_tcgen05bwd":
        model.exact_fp32, model.exact_bwd = True, False
    tr = TRAINERS[s["algo"]](model, trainable_layers=FULL_LAYERS, **TRAINER_KW)
    loss = tr.training_step(make_batch(feats, ex))
    count = float(tr._count)
    ref_loss, ref_out, sdt = oracle_losses(cfg, sd, feats, ex, s["algo"], requires_grad=tuple(FULL_LAYERS))
    ref_loss.backward()
    rl = float(ref_loss.detach())
    assert abs(float(loss) - rl) <= RTOL * max(abs(rl), 1e-3), (float(loss), rl)
    out = model(to_torch_tree(feats, "cuda"), outputs=("probability",))
    a, b = valid_logits(out["probability"].cpu(), feats), valid_logits(ref_out["probability"].detach(), feats)
    assert (a - b).abs().max().item() <= RTOL * b.abs().max().item()
    best = out["probability"].reshape(s["bs"], -1).argmax(-1).cpu()
    assert torch.equal(best, ref_out["probability"].detach().reshape(s["bs"], -1).argmax(-1))
    gmax = max(float(t.grad.abs().max()) for t in sdt.values() if t.grad is not None)
    over, total, worst_norm, worst_l2, worst_el = 0, 0, (0.0, ""), (0.0, ""), (0.0, "")
    for n in sorted(model.arena.trainable):
        ref = sdt[n].grad
        assert ref is not None, n
        got = (model.arena.grad_view(n) / count).cpu()
        rn, gn = float(ref.double().pow(2).sum().sqrt()), float(got.double().pow(2).sum().sqrt())
        floor = 1e-6 * gmax * ref.numel() ** 0.5             # tensors whose whole gradient is round-off sized are skipped
        err = (got - ref).abs()
        if rn > floor:
            worst_norm = max(worst_norm, (abs(gn - rn) / rn, n))
            worst_l2 = max(worst_l2, (float((got - ref).double().pow(2).sum().sqrt()) / rn, n))
        tmax = float(ref.abs().max())
        worst_el = max(worst_el, (float(err.max()) / (tmax + 1e-3 * gmax), n))
        over += int((err > 2e-3 * tmax + 1e-6 * gmax).sum())
        total += ref.numel()
    print(f"[{name}/{mode}] loss {float(loss):.6f} vs {rl:.6f} count {count:.0f} gmax {gmax:.3e}; norm dev {worst_norm[0]:.2e} ({worst_norm[1]}); "
          f"rel L2 err {worst_l2[0]:.2e} ({worst_l2[1]}); element err / (tensor max + 1e-3 gmax) {worst_el[0]:.2e} ({worst_el[1]}); "
          f"{over} of {total} elements beyond 2e-3 of their tensor max")
    tol = PARITY_TOL[mode]
    assert worst_norm[0] <= tol["norm"], worst_norm
    assert worst_l2[0] <= tol["l2"], worst_l2
    assert worst_el[0] <= tol["element"], worst_el
    assert over <= tol["over"] * total, (over, total)


@pytest.mark.parametrize("K", [2, 4])
def test_micro_batched_step_equals_whole_batch_step(K):
    """K concurrent micro-batches (own engine / stream / gradient arena each) against the same batch run whole: same loss,
    same valid count, same gradients up to summation order - on a RAGGED batch, where the reference's r2r key-padding
    indexing (r_pad[j % bs] over the whole batch) makes a sample's output depend on its position in the batch."""
    cfg = MODEL_ZOO["small"]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, 32, 10, 12, 4, seed=3, ragged=True)
    ex = synth_rl_extras(cfg, feats, seed=4)
    out = {}
    for k in (1, K):
        model = build(cfg, sd)
        tr = TRAINERS["grpo"](model, trainable_layers=FULL_LAYERS, **TRAINER_KW)
        tr.micro_batches = k
        tr.use_cuda_graph = False
        loss = tr.training_step(make_batch(feats, ex))
        out[k] = (float(loss), float(tr._count), model.arena.grads[:model.arena.n_train].clone())
    assert out[1][1] == out[K][1]
    assert abs(out[1][0] - out[K][0]) <= 2e-6 * max(1.0, abs(out[1][0]))      # (8-sample slices route some GEMMs to the exact kernel)
    g1, gk = out[1][2], out[K][2]
    # K = 2: same kernels on both sides, only the summation order differs.  K = 4: the 8-sample slices send more (small) GEMMs
    # to the exact-fp32 kernel than the whole batch does, and the ~1e-5 activation differences flip a few ReLU / arg-max decisions
    assert float((g1 - gk).abs().max()) <= (2e-5 if K == 2 else 1e-2) * float(g1.abs().max())
    # and through the CUDA graph (whole step incl. optimizer): parameters after two steps agree
    params = {}
    for k in (1, K):
        model = build(cfg, sd)
        tr = TRAINERS["grpo"](model, trainable_layers=FULL_LAYERS, **TRAINER_KW)
        tr.micro_batches = k
        tr.configure_optimizers()
        b = make_batch(feats, ex)
        for _ in range(3):
            tr.step(b)
        params[k] = model.arena.params[:model.arena.n_train].clone()
    # (AdamW turns round-off sized gradient differences into +-lr steps on elements whose gradient is ~0: compare in lr units)
    d = (params[1] - params[K]).abs()
    assert float((d > 1.5e-4).float().mean()) < (1e-3 if K == 2 else 2e-2)
