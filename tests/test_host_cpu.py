"""CPU: host-side mirrors (feature collate, datamodule collates, LR schedule) and the data-parallel reduce
under a 2-process gloo group."""
import os
import socket

import numpy as np
import pytest
import torch

from rift_b200.feature import PlutoFeature, pad_first_dim
from rift_b200.datamodule import RIFTCollate, GRPOCollate, PPOCollate, ReinforceCollate
from rift_b200.config import pluto_small
from rift_b200.synth import synth_features, synth_rl_extras
from oracle import loss_oracle as lo


def _samples(n=4, seed=0):
    """Per-sample (un-collated) PlutoFeature dicts with ragged agent / polygon / reference-line counts."""
    cfg = pluto_small()
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for i in range(n):
        A, Mp, R = int(rng.integers(2, 7)), int(rng.integers(1, 6)), int(rng.integers(1, 4))
        f = synth_features(cfg, 1, A, Mp, R, seed=seed * 100 + i)
        d = {k: ({kk: vv[0] for kk, vv in v.items()} if isinstance(v, dict) else v[0]) for k, v in f.items()}
        out.append((d, A, Mp, R))
    return out


def test_collate_pads_with_zeros_like_pad_sequence():
    samples = _samples()
    feats = [PlutoFeature(data=d) for d, *_ in samples]
    b = PlutoFeature.collate(feats).data
    A, Mp, R = (max(s[i] for s in samples) for i in (1, 2, 3))
    assert b["agent"]["position"].shape == (4, A, 21, 2) and b["agent"]["position"].dtype == torch.float32
    assert b["map"]["point_position"].shape[:2] == (4, Mp) and b["reference_line"]["position"].shape[:2] == (4, R)
    assert b["agent"]["valid_mask"].dtype == torch.bool and b["agent"]["category"].dtype == torch.int8
    for i, (d, a, m, r) in enumerate(samples):
        assert torch.equal(b["agent"]["heading"][i, :a], torch.from_numpy(d["agent"]["heading"]))
        assert not b["agent"]["valid_mask"][i, a:].any() and float(b["agent"]["position"][i, a:].abs().sum()) == 0
        assert not b["reference_line"]["valid_mask"][i, r:].any()
    assert b["current_state"].shape == (4, 7)
    ref = torch.nn.utils.rnn.pad_sequence([torch.from_numpy(d["map"]["polygon_center"]) for d, *_ in samples], batch_first=True)
    assert torch.equal(b["map"]["polygon_center"], ref)


@pytest.mark.reference
def test_collate_matches_reference_collate():
    from oracle import ref_shim
    RefFeature = ref_shim.pluto_feature_cls()
    samples = _samples(5, seed=3)
    ours = PlutoFeature.collate([PlutoFeature(data=d) for d, *_ in samples]).data
    ref = RefFeature.collate([RefFeature(data=d).to_feature_tensor() for d, *_ in samples]).data
    for k, v in ref.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                assert torch.equal(ours[k][kk], vv), (k, kk)
        else:
            assert torch.equal(ours[k], v), k


def test_collate_callables_produce_the_reference_keys():
    samples = _samples(3, seed=5)
    cfg = pluto_small()
    items = []
    for d, A, Mp, R in samples:
        adv = np.random.default_rng(0).normal(size=(R, 12))
        vm = np.ones((R, 12), bool)
        lg = np.random.default_rng(1).normal(size=(R, 12)).astype(np.float32)
        items.append({"CBVs_obs": {"raw_pluto_feature": PlutoFeature(data=d)},
                      "CBVs_group_advantage": {"advantage": adv, "valid_mask": vm},
                      "CBVs_actions_old_group_logits": {"logits": lg, "valid_mask": vm},
                      "CBVs_actions_ref_group_logits": {"logits": lg, "valid_mask": vm},
                      "CBVs_state": torch.zeros(128), "CBVs_advantage": torch.tensor(0.5), "CBVs_reward_sum": torch.tensor(1.0),
                      "CBVs_old_log_prob": torch.tensor(-2.0), "CBVs_action_mode": torch.tensor([0, 3]),
                      "CBVs_return": torch.tensor(-4.0)})
    R = max(s[3] for s in samples)
    b = RIFTCollate()(items)
    assert set(b) == {"cur_pluto_feature_torch", "group_advantage_torch", "group_advantage_mask_torch",
                      "old_group_logits_torch", "old_group_logits_mask_torch"}
    assert b["group_advantage_torch"].shape == (3, R, 12) and b["group_advantage_torch"].dtype == torch.float64
    g = GRPOCollate()(items)
    assert "ref_group_logits_torch" in g and g["ref_group_logits_torch"].dtype == torch.float32
    p = PPOCollate()(items)
    assert p["action_mode_torch"].shape == (3, 2) and p["state_torch"].shape == (3, 128)
    assert ReinforceCollate()(items)["return_torch"].shape == (3,)


@pytest.mark.reference
def test_collate_callables_match_reference_collates_by_value():
    """a2: RIFT / GRPO / PPO / REINFORCE collates against the reference's own callables executed from
    rift_datamodule.py:20-51, grpo_datamodule.py:20-57, ppo_datamodule.py:40-70, reinforce_datamodule.py:41-64
    on ragged samples: every tensor torch.equal, same dtypes."""
    from oracle import ref_shim
    RefFeature = ref_shim.pluto_feature_cls()
    base = "rift/cbv/planning/fine_tuner/rlft/"
    ref_cls = {
        "rift": ref_shim.ref_class(base + "rift_pluto/rift_datamodule.py", "RIFTCollate"),
        "grpo": ref_shim.ref_class(base + "grpo_pluto/grpo_datamodule.py", "GRPOCollate"),
        "ppo": ref_shim.ref_class(base + "ppo_pluto/ppo_datamodule.py", "PPOCollate"),
        "reinforce": ref_shim.ref_class(base + "reinforce_pluto/reinforce_datamodule.py", "ReinforceCollate"),
    }
    ours_cls = {"rift": RIFTCollate, "grpo": GRPOCollate, "ppo": PPOCollate, "reinforce": ReinforceCollate}
    samples = _samples(6, seed=9)
    rng = np.random.default_rng(4)

    def items(feature_cls, tensorise):
        out = []
        r2 = np.random.default_rng(4)
        for d, A, Mp, R in samples:
            vm = np.ones((R, 12), bool)
            vm[R - 1, 6:] = R > 1            # a partly valid last line
            f = feature_cls(data=d)
            out.append({"CBVs_obs": {"raw_pluto_feature": f.to_feature_tensor() if tensorise else f},
                        "CBVs_group_advantage": {"advantage": r2.normal(size=(R, 12)), "valid_mask": vm},
                        "CBVs_actions_old_group_logits": {"logits": r2.normal(size=(R, 12)).astype(np.float32), "valid_mask": vm},
                        "CBVs_actions_ref_group_logits": {"logits": r2.normal(size=(R, 12)).astype(np.float32), "valid_mask": vm},
                        "CBVs_state": torch.from_numpy(r2.normal(size=128).astype(np.float32)),
                        "CBVs_advantage": torch.tensor(float(r2.normal())), "CBVs_reward_sum": torch.tensor(float(r2.normal())),
                        "CBVs_old_log_prob": torch.tensor(float(-r2.uniform(1, 5))),
                        "CBVs_action_mode": torch.tensor([int(r2.integers(0, R)), int(r2.integers(0, 12))]),
                        "CBVs_return": torch.tensor(float(r2.normal()))})
        return out

    def same(a, b, path):
        if isinstance(b, dict):
            assert set(a) == set(b), (path, set(a) ^ set(b))
            for k in b:
                same(a[k], b[k], path + "/" + k)
        elif hasattr(b, "data") and isinstance(b.data, dict):
            same(a.data, b.data, path + ".data")
        else:
            assert a.dtype == b.dtype and torch.equal(a, b), path

    for algo in ref_cls:
        ref = ref_cls[algo]()(items(RefFeature, True))
        ours = ours_cls[algo]()(items(PlutoFeature, False))
        same(ours, ref, algo)


def test_warmup_cos_lr_matches_reference_formula():
    from rift_b200.trainer import WarmupCosLR

    class Opt:
        param_groups = [{"lr": 0.0}]
    s = WarmupCosLR(Opt(), lr=1e-4, min_lr=0.9e-4, epochs=16, warmup_epochs=3)
    got = []
    for e in range(16):
        got.append(Opt.param_groups[0]["lr"])
        s.step()
    want = [lo.warmup_cos_lr(e, 1e-4, 0.9e-4, 16, 3) for e in range(16)]
    assert np.allclose(got, want, rtol=0, atol=1e-15)
    # SURVEY App. D.6 (values executed by the reference code)
    assert np.allclose(got[:6], [3.33e-5, 6.67e-5, 1.00e-4, 1.00e-4, 9.99e-5, 9.94e-5], rtol=2e-3)


# ------------------------------------------------------------------ data-parallel reduce (gloo, world_size 2)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_data_parallel_reduce_reproduces_global_masked_mean():
    import subprocess
    import sys
    import tempfile
    port = _free_port()
    outdir = tempfile.mkdtemp(prefix="rift_dp_")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dp_worker.py")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(port), os.path.join(outdir, f"r{r}.pt")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    res = [torch.load(os.path.join(outdir, f"r{r}.pt")) for r in range(2)]
    # single-process reference on the whole batch
    gen = torch.Generator().manual_seed(0)
    bs, R, Mo = 8, 3, 12
    z = torch.randn(bs, R, Mo, generator=gen)
    old = z + 0.3 * torch.randn(bs, R, Mo, generator=gen)
    adv = torch.randn(bs, R, Mo, generator=gen, dtype=torch.float64)
    nr = torch.tensor([3, 1, 2, 3, 1, 1, 2, 3])
    vm = (torch.arange(R)[None, :] < nr[:, None])[..., None].expand(bs, R, Mo).contiguous()
    ref = float(lo.rift_loss(z, old, adv * vm, vm, ~vm.any(-1)))
    for rank, loss, count, _ in res:
        assert count == float(vm.sum())
        assert abs(loss - ref) < 1e-12 * max(1.0, abs(ref))
    assert torch.equal(res[0][3], res[1][3])                       # every rank holds the same reduced gradients
