"""Device-resident rollout buffer: bookkeeping against the reference's CBVRolloutBuffer (CPU), GPU collate against the
host collates value for value (GPU), and the plugin's train(e_i) loop on a full 4096-slot buffer (GPU)."""
import glob
import re

import numpy as np
import pytest
import torch

from rift_b200.buffer import DeviceRolloutBuffer
from rift_b200.config import pluto_small
from rift_b200.datamodule import COLLATES
from rift_b200.feature import PlutoFeature
from rift_b200.synth import synth_features, synth_state_dict

KEYS = ["CBVs_actions", "CBVs_actions_old_group_logits", "CBVs_actions_ref_group_logits", "CBVs_group_advantage", "CBVs_obs",
        "CBVs_next_obs", "CBVs_reward", "CBVs_terminated", "CBVs_done"]


def _feature(rng, i):
    cfg = pluto_small()
    A, Mp, R = int(rng.integers(2, 9)), int(rng.integers(1, 8)), int(rng.integers(1, 5))
    f = synth_features(cfg, 1, A, Mp, R, seed=5000 + i)
    d = {k: ({kk: vv[0] for kk, vv in v.items()} if isinstance(v, dict) else v[0]) for k, v in f.items()}
    return PlutoFeature(data=d), R


def _transition(rng, i, done):
    f, R = _feature(rng, i)
    ret = rng.normal(-5, 20, R * 12)
    vm = np.ones((R, 12), bool)
    return {"CBVs_actions": (0.1, 0.0, False), "CBVs_obs": {"raw_pluto_feature": f}, "CBVs_next_obs": {"raw_pluto_feature": f},
            "CBVs_group_advantage": {"advantage": ((ret - ret.mean()) / (ret.std() + 1e-5)).reshape(R, 12), "valid_mask": vm},
            "CBVs_actions_old_group_logits": {"logits": rng.normal(0, 1, (R, 12)).astype(np.float32), "valid_mask": vm},
            "CBVs_actions_ref_group_logits": {"logits": rng.normal(0, 1, (R, 12)).astype(np.float32), "valid_mask": vm},
            "CBVs_reward": np.float32(rng.normal(-0.5, 2)), "CBVs_terminated": np.float32(0.0), "CBVs_done": bool(done)}


def _episodes(n_total, seed=0, ids=(7, 8, 9)):
    """data_dict's as carla_runner.py:204-213 hands them to buffer.store: per key a list over steps of {CBV id: value}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    made, i = 0, 0
    while made < n_total:
        steps = int(rng.integers(7, 12))
        dd = {k: [] for k in KEYS}
        dd["CBV_ids"] = []
        lens = {c: int(rng.integers(3, steps + 1)) for c in ids}        # some trajectories are too short to keep alone
        for t in range(steps):
            alive = [c for c in ids if t < lens[c]]
            dd["CBV_ids"].append(alive)
            tr = {c: _transition(rng, i + j, done=(t == lens[c] - 1)) for j, c in enumerate(alive)}
            i += len(alive)
            for k in KEYS:
                dd[k].append({c: tr[c][k] for c in alive})
        made += sum(lens.values())
        yield dd


def _fill(buf, seed=0):
    for dd in _episodes(10 ** 9, seed):
        buf.store(dd)
        if buf.buffer_full:
            return buf


@pytest.mark.reference
def test_bookkeeping_matches_reference_buffer():
    """Same store() calls -> same contents, order, buffer_pos and buffer_full as cbv_rollout_buffer.py:16-138."""
    from oracle import ref_shim
    from collections import defaultdict, deque

    class BaseBuffer:
        def __init__(self, num_scenario, mode, logger=None):
            self.num_scenario, self.mode, self.logger = num_scenario, mode, logger
    Ref = ref_shim.ref_class("rift/gym_carla/buffer/cbv_rollout_buffer.py", "CBVRolloutBuffer",
                             {"BaseBuffer": BaseBuffer, "defaultdict": defaultdict, "deque": deque})
    cfg = {"buffer_capacity": 150, "data_keys": KEYS}
    ref, ours = Ref(1, "train_cbv", cfg), DeviceRolloutBuffer(1, "train_cbv", cfg, device="cpu")
    for dd in _episodes(400, seed=3):
        ref.store(dd)
        ours.store(dd)
        assert (ref.buffer_pos, ref.buffer_full) == (ours.buffer_pos, ours.buffer_full)
        if ref.buffer_full:
            break
    assert ref.buffer_full and ref.buffer_pos == 150
    for k in KEYS:
        assert len(ref.buffer_data[k]) == len(ours.buffer_data[k]) == 150
        assert all(a is b for a, b in zip(ref.buffer_data[k], ours.buffer_data[k])), k      # the very same objects, same order
    s_ref, s_ours = ref.sample([3, 77, 149]), ours.sample([3, 77, 149])
    assert all(a is b for k in KEYS for a, b in zip(s_ref[k], s_ours[k]))
    assert ref.sample(5)["CBVs_reward"] is ours.sample(5)["CBVs_reward"]


def test_short_trajectories_are_dropped_and_capacity_is_respected():
    buf = DeviceRolloutBuffer(1, "train_cbv", {"buffer_capacity": 40, "data_keys": KEYS}, device="cpu")
    _fill(buf, seed=1)
    assert buf.buffer_full and buf.buffer_pos == 40 and len(buf) == 40
    assert all(len(d) == 40 for d in buf.buffer_data.values())
    # the device mirror holds exactly the stored items, zero padded
    for i in (0, 17, 39):
        f = buf.buffer_data["CBVs_obs"][i]["raw_pluto_feature"].data
        A = f["agent"]["heading"].shape[0]
        assert buf._extent["A"][i] == A
        assert np.array_equal(buf._arena["agent_heading"][i, :A].numpy(), np.asarray(f["agent"]["heading"], np.float32))
        assert float(buf._arena["agent_heading"][i, A:].abs().sum()) == 0.0
    buf.reset_buffer()
    assert buf.buffer_pos == 0 and not buf.buffer_full and not buf._arena


@pytest.mark.gpu
@pytest.mark.parametrize("algo", ["rift", "grpo", "ppo", "reinforce"])
def test_gpu_collate_equals_host_collate(algo):
    """collate_device(indices) == {RIFT,GRPO,PPO,Reinforce}Collate()([buffer.sample(i) ...]) for every tensor the trainer
    reads (the host collates are themselves pinned to the reference's, tests/test_host_cpu.py)."""
    from rift_b200.planning_model import PackedBatch
    buf = _fill(DeviceRolloutBuffer(1, "train_cbv", {"buffer_capacity": 96, "data_keys": KEYS}), seed=2)
    rng = np.random.default_rng(0)
    if algo == "ppo":
        buf.add_extra_data({"CBVs_state": torch.randn(96, 128), "CBVs_advantage": torch.randn(96), "CBVs_reward_sum": torch.randn(96),
                            "CBVs_old_log_prob": -torch.rand(96), "CBVs_action_mode": torch.randint(0, 12, (96, 2))})
    if algo == "reinforce":
        buf.add_extra_data({"CBVs_return": torch.randn(96)})
    for trial in range(3):
        idx = rng.permutation(96)[: int(rng.integers(1, 33))].tolist()
        dev = buf.collate_device(idx, algo)
        host = COLLATES[algo]()([buf.sample(i) for i in idx])
        ref = PackedBatch(host["cur_pluto_feature_torch"].data, "cuda")
        got = dev["cur_pluto_feature_torch"]
        assert got.shape == ref.shape
        for name, t in ref.keep.items():
            assert got.keep[name].dtype == t.dtype and torch.equal(got.keep[name], t), name
        for k, v in host.items():
            if k == "cur_pluto_feature_torch" or k.endswith("logits_mask_torch"):
                continue
            assert dev[k].dtype == v.dtype and torch.equal(dev[k].cpu(), v), k


@pytest.mark.gpu
def test_plugin_train_loop_on_a_full_4096_slot_device_buffer(tmp_path):
    """rlft_pluto.py:206-247 on the reference's buffer capacity (rift_pluto.yaml:19) with the device-resident buffer:
    every mini-batch comes from the gather kernel; checkpoint naming / what moves / buffer reset as in the reference."""
    from rift_b200.policy import CBV_POLICY_LIST
    cfg = pluto_small()
    pre = tmp_path / "pretrained.ckpt"
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}}, pre)
    config = {"ckpt_path": str(pre), "ROOT_DIR": str(tmp_path), "model_path": "models", "load_agent_info": "rift_pluto",
              "obs": {"radius": 120}, "frame_rate": 10, "rlft": {"epochs": 2, "warmup_epochs": 1, "train_batch_size": 256}}
    pol = CBV_POLICY_LIST["rift_pluto"](config)
    pol.load_model(resume=True)
    buf = DeviceRolloutBuffer(1, "train_cbv", {"buffer_capacity": 4096, "data_keys": [k for k in KEYS if "ref_group" not in k]})
    # 4096 transitions from a pool of 64 distinct scenes (synthesising 4096 scenes would only test the generator)
    rng = np.random.Generator(np.random.PCG64(4))
    pool = [_transition(rng, i, done=False) for i in range(64)]
    while not buf.buffer_full:
        n = int(rng.integers(8, 30))
        picks = [pool[int(rng.integers(0, 64))] for _ in range(n)]
        dd = {k: [{1: dict(p, CBVs_done=(t == n - 1))[k]} for t, p in enumerate(picks)] for k in buf.data_keys}
        dd["CBV_ids"] = [[1]] * n
        buf.store(dd)
    assert buf.buffer_pos == 4096
    pol.set_buffer(buf, total_routes=1)
    pol.set_mode("train")
    before = {k: v.clone().cpu() for k, v in pol.pluto_model.state_dict().items()}
    pol.train(e_i=1)
    files = glob.glob(str(tmp_path / "models" / "rift_pluto" / "*.ckpt"))
    assert len(files) == 1 and re.search(r"carla_episode=1-epoch=\d+-val_loss=-?[\d.]+\.ckpt$", files[0]), files
    after = pol.pluto_model.state_dict()
    moved = {k for k in before if before[k].dtype.is_floating_point and not torch.equal(before[k], after[k].cpu())}
    assert moved and all(k.startswith("planning_decoder.pi_head.") for k in moved)
    assert buf.buffer_pos == 0 and not buf.buffer_full              # reset_buffer() after the fit (rlft_pluto.py:246)
