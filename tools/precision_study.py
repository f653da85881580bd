"""Numerics study (not a test): end-to-end logit / loss error of the forward under emulated operand
precisions of the GEMMs.  RIFT_B200_EMULATE=0|1|2|3 python tools/precision_study.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200.config import MODEL_ZOO
from rift_b200.planning_model import PlanningModel
from rift_b200.synth import synth_state_dict, synth_features, synth_rl_extras
from tests.helpers import oracle_losses, to_torch_tree
for model, bs, A, Mp, R in (("small", 8, 16, 20, 6), ("medium", 4, 32, 20, 6), ("medium", 64, 32, 20, 6)):
    cfg = MODEL_ZOO[model]()
    sd = synth_state_dict(cfg, seed=7)
    feats = synth_features(cfg, bs, A, Mp, R, seed=3, ragged=True)
    ex = synth_rl_extras(cfg, feats, seed=4)
    with torch.no_grad():
        loss, ref, _ = oracle_losses(cfg, sd, feats, ex, "grpo")
    m = PlanningModel.from_config(cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    out = m(to_torch_tree(feats, "cuda"))
    keep = torch.from_numpy(feats["reference_line"]["valid_mask"].any(-1))
    res = {}
    for k in ("probability", "trajectory", "hidden", "ref_free_trajectory", "prediction"):
        a, b = out[k].cpu(), ref[k]
        if k == "probability":
            a, b = a[keep], b[keep]
        res[k] = ((a - b).abs().max() / b.abs().max()).item()
    from rift_b200 import functional as F
    t = {k: torch.from_numpy(v).cuda() for k, v in ex.items()}
    l2, _, _ = F.group_objective("grpo", out["probability"], t["old_group_logits"], t["group_advantage"],
                                 t["group_advantage_mask"], out["r_padding_mask"], t["ref_group_logits"])
    res["loss_rel"] = abs(float(l2) - float(loss)) / abs(float(loss))
    print(os.environ.get("RIFT_B200_EMULATE", "0"), model, bs, {k: f"{v:.2e}" for k, v in res.items()}, flush=True)
