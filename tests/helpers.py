"""Shared fixtures for the parity tests: seeded cases, golden loading, oracle evaluation."""
import os

import numpy as np
import torch

from rift_b200.config import MODEL_ZOO
from rift_b200.synth import synth_state_dict, synth_features, synth_rl_extras

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must mirror oracle/make_golden.py:CASES
CASES = {
    "cfg1_small": dict(model="small", kw=dict(future_steps=40), bs=1, A=8, Mp=20, R=1, ragged=False),
    "ragged_small": dict(model="small", kw={}, bs=3, A=9, Mp=11, R=3, ragged=True),
    "medium_tiny": dict(model="medium", kw={}, bs=2, A=6, Mp=7, R=2, ragged=True),
}


def case_cfg(name, ppo=False):
    c = CASES[name]
    kw = dict(c["kw"])
    if ppo:
        kw["value_hidden"] = (256, 256)
    return MODEL_ZOO[c["model"]](**kw)


def case_inputs(name, ppo=False):
    c = CASES[name]
    cfg = case_cfg(name, ppo)
    feats = synth_features(cfg, c["bs"], c["A"], c["Mp"], c["R"], seed=1, ragged=c["ragged"])
    extras = synth_rl_extras(cfg, feats, seed=2)
    sd = synth_state_dict(cfg, seed=7)
    return cfg, sd, feats, extras


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def check_golden(g, key, arr, rtol, atol=0.0):
    """Compare `arr` with golden `key`, which may be stored whole or as an every-7th sample + stats."""
    arr = np.asarray(arr)
    if key in g.files:
        ref = g[key]
        assert ref.shape == arr.shape, (key, ref.shape, arr.shape)
        scale = max(float(np.abs(ref).max()), 1e-30)
        err = float(np.abs(ref.astype(np.float64) - arr.astype(np.float64)).max())
        assert err <= rtol * scale + atol, f"{key}: max err {err:.3e} vs scale {scale:.3e}"
    else:
        ref = g[key + "@s7"]
        got = arr.reshape(-1)[::7]
        scale = max(float(np.abs(ref).max()), 1e-30)
        err = float(np.abs(ref.astype(np.float64) - got.astype(np.float64)).max())
        assert err <= rtol * scale + atol, f"{key}@s7: max err {err:.3e} vs scale {scale:.3e}"
        st = g[key + "@stats"]
        l2 = np.sqrt((arr.astype(np.float64) ** 2).sum())
        assert abs(l2 - st[1]) <= max(rtol, 1e-5) * max(st[1], 1e-30) + atol, f"{key}@l2: {l2} vs {st[1]}"


def to_torch_tree(tree, device="cpu"):
    if isinstance(tree, dict):
        return {k: to_torch_tree(v, device) for k, v in tree.items()}
    if isinstance(tree, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(tree)).to(device)
    return tree


def oracle_losses(cfg, sd_np, feats, extras, algo, requires_grad=()):
    """Oracle forward + objective.  Returns (loss, outputs, sd tensors)."""
    from oracle import pluto_oracle as po, loss_oracle as lo
    sd = {k: torch.from_numpy(v.copy()) for k, v in sd_np.items()}
    for n in sd:
        if any(n.startswith(p) for p in requires_grad) and sd[n].dtype.is_floating_point:
            sd[n].requires_grad_(True)
    data = to_torch_tree(feats)
    out = po.planning_model_forward(data, sd, cfg)
    prob = out["probability"]
    r_pad = ~data["reference_line"]["valid_mask"].any(-1)
    ex = to_torch_tree(extras)
    if algo == "rift":
        loss = lo.rift_loss(prob, ex["old_group_logits"], ex["group_advantage"], ex["group_advantage_mask"], r_pad)
    elif algo == "grpo":
        loss = lo.grpo_loss(prob, ex["old_group_logits"], ex["ref_group_logits"], ex["group_advantage"],
                            ex["group_advantage_mask"], r_pad)
    elif algo == "ppo":
        value = po.critic_ppo(ex["state"], sd)
        loss = lo.ppo_loss(prob, r_pad, ex["action_mode"], value, ex["advantage"], ex["reward_sum"],
                           ex["old_log_prob"])
    elif algo == "reinforce":
        loss = lo.reinforce_loss(prob, r_pad, ex["return"])
    else:
        raise ValueError(algo)
    return loss, out, sd
