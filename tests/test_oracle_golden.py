"""CPU: the oracle restatement against golden vectors produced by the unmodified reference
(oracle/make_golden.py), i.e. the pin that makes the oracle trustworthy on the GPU box."""
import json

import numpy as np
import pytest
import torch

from tests.helpers import CASES, case_inputs, golden, check_golden, oracle_losses
from rift_b200.config import MODEL_ZOO, param_spec, is_buffer
from oracle import loss_oracle as lo


def test_param_table_matches_reference_state_dict():
    spec = json.load(open(golden.__globals__["GOLDEN"] + "/state_dict_spec.json"))
    for mname in ("small", "medium"):
        ours = [[n, list(s), "torch.float32" if k == "f32" else "torch.int64"]
                for n, s, k in param_spec(MODEL_ZOO[mname]())]
        assert ours == spec[mname]


@pytest.mark.parametrize("name", list(CASES))
def test_forward_matches_reference_golden(name):
    cfg, sd, feats, extras = case_inputs(name)
    g = golden(name)
    with torch.no_grad():
        _, out, _ = oracle_losses(cfg, sd, feats, extras, "rift")
    for k in ("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory"):
        check_golden(g, "out_" + k, out[k].numpy(), rtol=2e-5)
    assert np.array_equal(g["out_best_index"], out["best_index"].numpy())


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("algo", ["rift", "grpo", "ppo", "reinforce"])
def test_losses_and_pi_head_grads_match_reference_golden(name, algo):
    cfg, sd, feats, extras = case_inputs(name, ppo=(algo == "ppo"))
    g = golden(name)
    loss, out, sdt = oracle_losses(cfg, sd, feats, extras, algo,
                                   requires_grad=("planning_decoder.pi_head", "value_net"))
    ref = float(g[f"loss_{algo}"])
    assert abs(float(loss.detach()) - ref) <= 1e-5 * max(abs(ref), 1e-3)
    if algo == "reinforce":
        return
    loss.backward()
    for k in g.files:
        if k.startswith(f"grad_{algo}/"):
            n = k.split("/", 1)[1].split("@")[0]
            if k.endswith("@stats"):
                continue
            check_golden(g, f"grad_{algo}/{n}", sdt[n].grad.numpy(), rtol=2e-4, atol=5e-7)  # mlp.3.bias grad is an exact-zero sum: pure round-off


def test_group_advantage_bit_exact():
    g = golden("advantage")
    for k in g.files:
        if not k.startswith("ret_"):
            continue
        G = int(k[4:])
        ret, adv = g[k], g[f"adv_{G}"]
        for i in range(ret.shape[0]):
            assert np.array_equal(lo.group_advantage(ret[i]), adv[i])
            assert np.array_equal(lo.group_advantage_explicit(ret[i]), adv[i]), (G, i)


def test_buffer_passes():
    g = golden("buffer_pass")
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    adv = lo.gae(t["rewards"], 1 - t["dones"], t["values"], t["next_values"], 1 - t["terminated"])
    assert torch.equal(adv, t["gae"])
    assert torch.allclose(lo.ppo_normalise(adv), t["gae_normalised"], rtol=1e-6, atol=1e-7)
    assert torch.equal(lo.discounted_return(t["rewards"], t["dones"]), t["discounted_return"])


def test_optimizer_grouping_and_steps():
    name = "cfg1_small"
    cfg, sd, feats, extras = case_inputs(name)
    g = golden(name)
    groups = json.loads(str(g["optim_groups"]))
    names = [(n, s) for n, s, _ in param_spec(cfg) if n.startswith("planning_decoder.pi_head")]
    decay, no_decay = lo.decay_partition(names)
    assert [decay, no_decay] == groups
    # three clip+AdamW steps on pi_head with the oracle loss reproduce the reference parameters
    params = {k: torch.from_numpy(v.copy()) for k, v in sd.items()}
    state = {}
    for step in range(3):
        sd_np = {k: v.numpy() for k, v in params.items()}
        loss, _, sdt = oracle_losses(cfg, sd_np, feats, extras, "grpo", requires_grad=("planning_decoder.pi_head",))
        loss.backward()
        grads = {n: t.grad for n, t in sdt.items() if t.grad is not None}
        tn = lo.clip_adamw_step(params, grads, state, lr=1e-4)
        assert abs(tn - float(g[f"gradnorm_grpo_step{step}"])) <= 1e-4 * tn
        if step in (0, 2):
            for n in grads:
                # Adam's first steps are lr*g/(|g|+eps): where |g| is round-off-sized (exact-zero sums
                # such as d/d(LN bias) of an always-active unit) the update is noise in the reference
                # itself, so only elements with |g| > 1e-6 are compared, to 2% of one lr-sized step
                key = f"param_grpo_step{step + 1}/{n}"
                if key not in g.files:
                    continue
                sel = np.abs(g[f"grad_grpo/{n}"]) > 1e-6
                err = np.abs(params[n].numpy() - g[key])[sel]
                assert err.size == 0 or err.max() <= 0.02 * 1e-4, (key, err.max())


def test_full_model_decay_partition_counts():
    # 137 decay tensors / 289 no-decay tensors for Pluto-small
    cfg = MODEL_ZOO["small"]()
    names = [(n, s) for n, s, _ in param_spec(cfg) if not is_buffer(n)]
    decay, no_decay = lo.decay_partition(names)
    assert (len(decay), len(no_decay)) == (137, 289)


@pytest.mark.parametrize("kind", ["sft", "rtr", "rs"])
def test_sft_family_objectives_match_reference_golden(kind):
    """fine_tuner/sft: SFT teacher cross-entropy, RTR = 5 x PPO + teacher, RS = REINFORCE - loss, d loss / d logits and
    the teacher label against the reference trainers' own outputs (tests/golden/sft_objectives.npz)."""
    from oracle import pluto_oracle as po
    from tests.helpers import to_torch_tree
    g = golden("sft_objectives")
    cfg, sd_np, feats, extras = case_inputs("ragged_small", ppo=(kind == "rtr"))
    sd = {k: torch.from_numpy(v.copy()) for k, v in sd_np.items()}
    data, ex = to_torch_tree(feats), to_torch_tree(extras)
    with torch.no_grad():
        out = po.planning_model_forward(data, sd, cfg)
    prob = out["probability"].clone().requires_grad_(True)
    r_pad = ~data["reference_line"]["valid_mask"].any(-1)
    ti = torch.from_numpy(g["teacher_infos"])
    if kind == "sft":
        loss = lo.sft_loss(out["trajectory"], prob, r_pad, ti)
    elif kind == "rtr":
        value = po.critic_ppo(ex["state"], sd)
        loss = lo.rtr_loss(out["trajectory"], prob, r_pad, ti, ex["action_mode"], value, ex["advantage"], ex["reward_sum"],
                           ex["old_log_prob"])
    else:
        loss = lo.reinforce_loss(prob, r_pad, ex["return"])
    ref = float(g[f"loss_{kind}"])
    assert abs(float(loss.detach()) - ref) <= 1e-5 * max(abs(ref), 1e-3)
    loss.backward()
    assert np.abs(prob.grad.numpy() - g[f"dlogits_{kind}"]).max() <= 1e-5 * np.abs(g[f"dlogits_{kind}"]).max()
    if kind != "rs":
        label, _ = lo.teacher_label(out["trajectory"], prob.detach(), r_pad, ti)
        assert np.array_equal(label.numpy(), g[f"label_{kind}"])
