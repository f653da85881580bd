"""GPU: PPO trainer (value net) against the reference goldens, and the policy plugin's train(e_i) loop on a
small in-memory rollout buffer (checkpoint naming, LR decay, what moves and what stays frozen)."""
import glob
import os
import re

import numpy as np
import pytest
import torch

from rift_b200.feature import PlutoFeature
from rift_b200.policy import CBV_POLICY_LIST
from rift_b200.trainer import TRAINERS, PPOPlutoModel
from rift_b200.config import pluto_small
from rift_b200.synth import synth_features, synth_rl_extras, synth_state_dict
from tests.helpers import CASES, case_inputs, golden, check_golden, to_torch_tree

pytestmark = pytest.mark.gpu

TRAINER_KW = dict(lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3, frame_rate=10)


@pytest.mark.parametrize("name", list(CASES))
def test_ppo_training_step_matches_reference_golden(name):
    cfg, sd, feats, ex = case_inputs(name, ppo=True)
    g = golden(name)
    model = PPOPlutoModel(cfg.radius, hidden_dim=(256, 256), dim=cfg.dim, num_heads=cfg.num_heads,
                          encoder_depth=cfg.encoder_depth, decoder_depth=cfg.decoder_depth, future_steps=cfg.future_steps)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tr = TRAINERS["ppo"](model, trainable_layers=["planning_decoder.pi_head", "value_net"], **TRAINER_KW)
    batch = {"cur_pluto_feature_torch": to_torch_tree(feats, "cuda")}
    for k in ("state", "advantage", "reward_sum", "old_log_prob", "action_mode"):
        batch[k + "_torch"] = torch.from_numpy(ex[k].copy()).cuda()
    loss = tr.training_step(batch)
    ref = float(g["loss_ppo"])
    assert abs(float(loss) - ref) <= 1e-3 * max(abs(ref), 1e-3)
    for k in g.files:
        if k.startswith("grad_ppo/") and not k.endswith("@stats"):
            n = k.split("/", 1)[1].split("@")[0]
            check_golden(g, f"grad_ppo/{n}", model.arena.grad_view(n).cpu().numpy(), rtol=1e-3, atol=2e-6)
    # three updates run and move the value net and the head, nothing else
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tr.configure_optimizers()
    for _ in range(3):
        tr.step(batch)
    after = model.state_dict()
    moved = {k for k in before if before[k].dtype.is_floating_point and not torch.equal(before[k].cpu(), after[k].cpu())}
    assert moved and all(k.startswith(("planning_decoder.pi_head.", "value_net.")) for k in moved)


class _Buffer:
    """Minimal stand-in with the read-only interface of CBVRolloutBuffer (cbv_rollout_buffer.py:77-138)."""

    def __init__(self, items, extra):
        self.items, self.extra = items, extra
        self.buffer_capacity = len(items)
        self.buffer_full = True
        self.was_reset = False

    def __len__(self):
        return len(self.items)

    def sample(self, idx):
        d = dict(self.items[idx])
        d.update({k: v[idx] for k, v in self.extra.items()})
        return d

    def get_key_data(self, key):
        return [it[key] for it in self.items] if key in self.items[0] else self.extra[key]

    def add_extra_data(self, data):
        assert all(len(v) == self.buffer_capacity for v in data.values())
        self.extra.update(data)

    def reset_buffer(self):
        self.was_reset = True


def _make_buffer(n=40, seed=0):
    cfg = pluto_small()
    rng = np.random.Generator(np.random.PCG64(seed))
    items = []
    for i in range(n):
        A, Mp, R = int(rng.integers(3, 8)), int(rng.integers(2, 7)), int(rng.integers(1, 4))
        f = synth_features(cfg, 1, A, Mp, R, seed=1000 + i)
        d = {k: ({kk: vv[0] for kk, vv in v.items()} if isinstance(v, dict) else v[0]) for k, v in f.items()}
        ret = rng.normal(-5, 20, R * 12)
        adv = ((ret - ret.mean()) / (ret.std() + 1e-5)).reshape(R, 12)
        vm = np.ones((R, 12), bool)
        items.append({
            "CBVs_obs": {"raw_pluto_feature": PlutoFeature(data=d)}, "CBVs_next_obs": {"raw_pluto_feature": PlutoFeature(data=d)},
            "CBVs_group_advantage": {"advantage": adv, "valid_mask": vm},
            "CBVs_actions_old_group_logits": {"logits": rng.normal(0, 1, (R, 12)).astype(np.float32), "valid_mask": vm},
            "CBVs_actions_ref_group_logits": {"logits": rng.normal(0, 1, (R, 12)).astype(np.float32), "valid_mask": vm},
            "CBVs_reward": np.float32(rng.normal(-0.5, 2)), "CBVs_done": np.float32(rng.uniform() < 0.1),
            "CBVs_terminated": np.float32(0.0), "CBVs_actions_old_log_prob": np.float32(-rng.uniform(1, 4)),
            "CBVs_actions_mode": np.array([0, int(rng.integers(0, 12))], np.int64),
        })
    return _Buffer(items, {})


@pytest.mark.parametrize("policy_name", ["rift_pluto", "grpo_pluto", "ppo_pluto", "reinforce_pluto"])
def test_plugin_train_loop(policy_name, tmp_path):
    cfg = pluto_small()
    pre = tmp_path / "pretrained.ckpt"
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}}, pre)
    config = {"ckpt_path": str(pre), "ROOT_DIR": str(tmp_path), "model_path": "models", "load_agent_info": policy_name,
              "obs": {"radius": 120}, "frame_rate": 10, "ppo": {"hidden_dim": [256, 256], "clip_epsilon": 0.2, "lambda_entropy": 0.01},
              "rlft": {"epochs": 2, "warmup_epochs": 1, "train_batch_size": 16}}
    pol = CBV_POLICY_LIST[policy_name](config)
    pol.load_model(resume=True)
    assert pol.continue_episode == 0 and pol.current_epoch == 0
    buf = _make_buffer()
    pol.set_buffer(buf, total_routes=1)
    pol.set_mode("train")
    before = {k: v.clone().cpu() for k, v in pol.pluto_model.state_dict().items()}
    pol.train(e_i=3)
    files = glob.glob(str(tmp_path / "models" / policy_name / "*.ckpt"))
    assert len(files) == 1 and re.search(r"carla_episode=3-epoch=\d+-val_loss=-?[\d.]+\.ckpt$", files[0]), files
    assert pol.current_epoch == 1 and buf.was_reset
    ck = torch.load(files[0], weights_only=False)["state_dict"]
    assert all(k.startswith("model.") for k in ck)
    after = pol.pluto_model.state_dict()
    moved = {k for k in before if before[k].dtype.is_floating_point and not torch.equal(before[k], after[k].cpu())}
    allowed = ("planning_decoder.pi_head.", "value_net.") if policy_name == "ppo_pluto" else ("planning_decoder.pi_head.",)
    assert moved and all(k.startswith(allowed) for k in moved), sorted(moved)[:5]
    # resume picks the newest episode and decays the closed-loop learning rate
    pol2 = CBV_POLICY_LIST[policy_name](config)
    pol2.load_model(resume=True)
    assert pol2.continue_episode == 3 and pol2.current_epoch == 1
    for k in moved:
        assert torch.equal(pol2.pluto_model.state_dict()[k].cpu(), after[k].cpu())


def test_policy_outputs_advantage_is_numpy_exact():
    cfg = pluto_small()
    config = {"obs": {"radius": 120}}
    pol = CBV_POLICY_LIST["rift_pluto"](config)
    pol.pluto_model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()})
    buf = _make_buffer(3, seed=5)
    feats = [it["CBVs_obs"]["raw_pluto_feature"] for it in buf.items]
    rng = np.random.Generator(np.random.PCG64(1))
    rets = [rng.normal(-5, 20, f.data["reference_line"]["position"].shape[0] * 12) for f in feats]
    out = pol.policy_outputs(feats, rets)
    for b, r in enumerate(rets):
        ref = (r - np.mean(r)) / (np.std(r) + 1e-5)
        got = out["group_advantage"][b].reshape(-1)[: len(r)].cpu().numpy()
        assert np.array_equal(got, ref)
        assert int(out["group_advantage_mask"][b].sum()) == len(r)
    assert out["candidate_trajectories"].shape[-1] == 3


@pytest.mark.parametrize("name", list(CASES))
def test_ppo_full_backward_matches_reference_golden(name):
    """PPO with every module trainable against the reference's autograd: per-tensor L2 norm
    and sum of all gradient tensors, value net included."""
    from rift_b200.config import param_spec, is_buffer
    from tests.test_model_gpu import FULL_LAYERS
    cfg, sd, feats, ex = case_inputs(name, ppo=True)
    g = golden(name)
    model = PPOPlutoModel(cfg.radius, hidden_dim=(256, 256), dim=cfg.dim, num_heads=cfg.num_heads,
                          encoder_depth=cfg.encoder_depth, decoder_depth=cfg.decoder_depth, future_steps=cfg.future_steps)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    tr = TRAINERS["ppo"](model, trainable_layers=FULL_LAYERS + ["value_net"], **TRAINER_KW)
    batch = {"cur_pluto_feature_torch": to_torch_tree(feats, "cuda")}
    for k in ("state", "advantage", "reward_sum", "old_log_prob", "action_mode"):
        batch[k + "_torch"] = torch.from_numpy(ex[k].copy()).cuda()
    loss = tr.training_step(batch)
    ref = float(g["loss_ppo"])
    assert abs(float(loss) - ref) <= 1e-3 * max(abs(ref), 1e-3)
    names = [n for n, _, _ in param_spec(cfg) if not is_buffer(n)]
    stats = g["fullgrad_ppo_stats"]
    assert stats.shape[0] == len(names), (stats.shape, len(names))
    gmax = stats[:, 1].max()
    bad = []
    for i, n in enumerate(names):
        if n not in model.arena.trainable:
            assert stats[i, 1] == 0.0, f"{n} has a reference gradient but is not in the trainable arena"
            continue
        gv = model.arena.grad_view(n).double().cpu()
        l2 = float(gv.pow(2).sum().sqrt())
        if abs(l2 - stats[i, 1]) > 2e-3 * stats[i, 1] + 1e-6 * gmax or \
                abs(float(gv.sum()) - stats[i, 0]) > 2e-3 * stats[i, 1] * np.sqrt(gv.numel()) + 1e-6 * gmax:
            bad.append((n, l2, stats[i, 1], float(gv.sum()), stats[i, 0]))
    assert not bad, f"{len(bad)} tensors differ, first: {bad[:5]}"


# ------------------------------------------------------------------ get_action (rollout side of the plugin)
class _P:
    pass


def _agent_state(x, y, heading, speed):
    """Duck-typed CarlaAgentState: what get_action reads (rear_axle.array / .heading, centre speed)."""
    s = _P()
    s.rear_axle = _P(); s.rear_axle.array = np.array([x, y], np.float64); s.rear_axle.heading = float(heading)
    s.dynamic_car_state = _P(); s.dynamic_car_state.center_velocity_2d = _P()
    s.dynamic_car_state.center_velocity_2d.magnitude = lambda v=float(speed): v
    s.dynamic_car_state.speed = float(speed)
    s.car_footprint = _P(); s.car_footprint.width, s.car_footprint.length = 2.0, 4.6
    s.center = _P(); s.center.array = np.array([x + 1.4 * np.cos(heading), y + 1.4 * np.sin(heading)]); s.center.heading = float(heading)
    return s


class _World:
    """Recorded world with the CarlaDataProvider calls get_action makes (carla_data_provider.py:192,619,665,1031)."""

    def __init__(self, ids, seed=0):
        rng = np.random.default_rng(seed)
        self.hist = {i: [_agent_state(*rng.normal(0, 20, 2), rng.normal(0, 1), rng.uniform(0, 8)) for _ in range(3)] for i in ids}

    def get_actor_by_id(self, i):
        a = _P(); a.id = i
        return a

    def get_history_state(self, actor):
        return self.hist[actor.id]

    def get_ego_vehicle_by_env_id(self, env_id):
        e = _P(); e.id = 1000 + env_id
        return e

    def get_CBV_nearby_agents(self, ego_id, cbv_id):
        return []


class _Evaluator:
    """Stand-in roll-out evaluator: deterministic pseudo-returns per candidate (the real one is the reference's
    TrajEvaluator or the CUDA evaluator); the plugin normalises them with the bit-exact advantage kernel."""

    def get_rollout_returns(self, history, raw_trajectories, ref_pos, ref_ang, nearby):
        t = raw_trajectories.detach().double().cpu().numpy()
        return -np.abs(t[..., :40, :2]).sum((-1, -2)).reshape(-1) - 3.0 * np.arange(t.shape[0] * t.shape[1])


@pytest.mark.parametrize("policy_name", ["rift_pluto", "grpo_pluto", "ppo_pluto"])
def test_get_action_keys_values_and_graph_replay(policy_name, tmp_path):
    """get_action(CBVs_obs_list, infos) -> the reference's return keys (rift_pluto.py:66-70, grpo_pluto.py:76-81,
    rlft_pluto.py:131-135); old / reference logits are the model's logits of the valid reference lines, the group
    advantage is numpy-exact, and the CUDA-graph replay over padded shape buckets returns what the eager forward does."""
    cfg = pluto_small()
    pre = tmp_path / "pretrained.ckpt"
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}}, pre)
    config = {"ckpt_path": str(pre), "ROOT_DIR": str(tmp_path), "model_path": "models", "load_agent_info": policy_name,
              "obs": {"radius": 120}, "frame_rate": 10, "num_scenario": 2, "topk": 10,
              "ppo": {"hidden_dim": [256, 256], "clip_epsilon": 0.2, "lambda_entropy": 0.01}}
    pol = CBV_POLICY_LIST[policy_name](config)
    pol.load_model(resume=True)
    pol.set_mode("train")
    buf = _make_buffer(5, seed=9)
    obs0 = {101 + i: {"raw_pluto_feature": buf.items[i]["CBVs_obs"]["raw_pluto_feature"]} for i in range(3)}
    obs1 = {201 + i: {"raw_pluto_feature": buf.items[3 + i]["CBVs_obs"]["raw_pluto_feature"]} for i in range(2)}
    infos = [{"env_id": 0}, {"env_id": 1}]
    pol.set_world(_World(list(obs0) + list(obs1)))
    pol.set_traj_evaluator(_Evaluator())
    results = {}
    for graph in (False, True, True):                 # second graphed call replays the captured graph
        pol.use_rollout_graph = graph
        pol.controllers.clear()
        results[graph] = pol.get_action([obs0, obs1], infos, deterministic=True)
    ref_out = pol.pluto_model.forward(PlutoFeature.collate([o["raw_pluto_feature"] for o in obs0.values()]).data, outputs=())
    for graph, data in results.items():
        assert len(data["CBVs_actions"]) == 2 and set(data["CBVs_actions"][0]) == set(obs0) and set(data["CBVs_actions"][1]) == set(obs1)
        for env, obs in ((0, obs0), (1, obs1)):
            for cid in obs:
                thr, steer, brake = data["CBVs_actions"][env][cid]
                assert 0.0 <= float(thr) <= 1.0 and -1.0 <= float(steer) <= 1.0 and bool(brake) in (True, False)
        if policy_name == "ppo_pluto":
            assert set(data) == {"CBVs_actions", "CBVs_actions_old_log_prob", "CBVs_actions_mode"}
            for cid in obs0:
                r, m = data["CBVs_actions_mode"][0][cid]
                # (-1, 11) = the reference-free trajectory won: its original index is -1 (pluto.py:222, rlft_pluto.py:160-162)
                assert ((0 <= r < 3 and 0 <= m < 12) or (r, m) == (-1, 11)) and float(data["CBVs_actions_old_log_prob"][0][cid]) <= 0.0
            continue
        want = {"CBVs_actions", "CBVs_actions_old_group_logits", "CBVs_group_advantage"}
        if policy_name == "grpo_pluto":
            want.add("CBVs_actions_ref_group_logits")
        assert set(data) == want
        for index, cid in enumerate(obs0):
            f = obs0[cid]["raw_pluto_feature"].data
            R = f["reference_line"]["position"].shape[0]
            old = data["CBVs_actions_old_group_logits"][0][cid]
            assert old["logits"].shape == (R, 12) and old["valid_mask"].all()
            got = torch.from_numpy(old["logits"])
            assert (got - ref_out["probability"][index, :R].cpu()).abs().max() <= 1e-4 * ref_out["probability"][index, :R].abs().max().cpu()
            if policy_name == "grpo_pluto":      # the frozen reference policy = the pretrained checkpoint = same weights here
                assert np.allclose(data["CBVs_actions_ref_group_logits"][0][cid]["logits"], old["logits"], rtol=0, atol=1e-5)
            adv = data["CBVs_group_advantage"][0][cid]
            assert adv["advantage"].shape == (R, 12) and adv["advantage"].dtype == np.float64
            assert abs(adv["advantage"].mean()) < 1e-9 and abs(adv["advantage"].std() - 1.0) < 1e-3
    # eager and graphed calls agree (same kernels, same inputs; padded rows are masked)
    a, b = results[False], results[True]
    for env in (0, 1):
        for cid in a["CBVs_actions"][env]:
            assert np.allclose(np.array(a["CBVs_actions"][env][cid], float), np.array(b["CBVs_actions"][env][cid], float), atol=1e-4)
    # CBVs that left the scene lose their PID state (pluto.py:114-125)
    pol.get_action([{101: obs0[101]}, {}], infos, deterministic=True)
    assert set(pol.controllers[0]) == {101} and 1 not in pol.controllers


def test_get_action_with_the_cuda_candidate_evaluator(tmp_path):
    """Train-mode get_action with the CUDA TrajEvaluator on the reference's hook (rift_pluto.py:113-145 ->
    traj_evaluator.py:422-475): (R, 12) float64 advantages, one group per CBV, finite and standardised."""
    from rift_b200.evaluator import TrajEvaluator
    cfg = pluto_small()
    pre = tmp_path / "pretrained.ckpt"
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    torch.save({"state_dict": {"model." + k: v for k, v in sd.items()}}, pre)
    config = {"ckpt_path": str(pre), "ROOT_DIR": str(tmp_path), "model_path": "models", "load_agent_info": "rift_pluto",
              "obs": {"radius": 120}, "frame_rate": 10, "num_scenario": 1, "topk": 10}
    pol = CBV_POLICY_LIST["rift_pluto"](config)
    pol.load_model(resume=True)
    pol.set_mode("train")
    buf = _make_buffer(3, seed=9)
    obs0 = {101 + i: {"raw_pluto_feature": buf.items[i]["CBVs_obs"]["raw_pluto_feature"]} for i in range(3)}
    pol.set_world(_World(list(obs0)))
    pol.set_traj_evaluator(TrajEvaluator())
    data = pol.get_action([obs0], [{"env_id": 0}], deterministic=True)
    for cid in obs0:
        R = obs0[cid]["raw_pluto_feature"].data["reference_line"]["position"].shape[0]
        adv = data["CBVs_group_advantage"][0][cid]
        assert adv["advantage"].shape == (R, 12) and adv["advantage"].dtype == np.float64 and adv["valid_mask"].all()
        assert np.isfinite(adv["advantage"]).all() and abs(adv["advantage"].mean()) < 1e-9
        assert adv["advantage"].std() == 0.0 or abs(adv["advantage"].std() - 1.0) < 1e-2


@pytest.mark.parametrize("kind", ["sft", "rtr", "rs"])
def test_sft_family_training_step_matches_reference_golden(kind):
    """SFT / RTR / RS trainers (fine_tuner/sft/sft_trainer.py:123-215, rtr_trainer.py:130-195, rs_trainer.py:154-170) on the
    CUDA path against the reference trainers' loss, teacher label and pi_head / value-net gradients."""
    g = golden("sft_objectives")
    cfg, sd, feats, ex = case_inputs("ragged_small", ppo=(kind == "rtr"))
    if kind == "rtr":
        model = PPOPlutoModel(cfg.radius, hidden_dim=(256, 256), dim=cfg.dim, num_heads=cfg.num_heads,
                              encoder_depth=cfg.encoder_depth, decoder_depth=cfg.decoder_depth, future_steps=cfg.future_steps)
    else:
        from rift_b200.planning_model import PlanningModel
        model = PlanningModel.from_config(cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    layers = ["planning_decoder.pi_head"] + (["value_net"] if kind == "rtr" else [])
    tr = TRAINERS[kind](model, trainable_layers=layers, **TRAINER_KW)
    batch = {"cur_pluto_feature_torch": to_torch_tree(feats, "cuda"), "teacher_infos": torch.from_numpy(g["teacher_infos"]).cuda()}
    for k in ("state", "advantage", "reward_sum", "old_log_prob", "action_mode", "return"):
        batch[k + "_torch"] = torch.from_numpy(ex[k].copy()).cuda()
    loss = tr.training_step(batch)
    ref = float(g[f"loss_{kind}"])
    assert abs(float(loss) - ref) <= 1e-3 * max(abs(ref), 1e-3), (float(loss), ref)
    if kind != "rs":
        assert np.array_equal(tr._label.cpu().numpy(), g[f"label_{kind}"])          # index-exact teacher label
    for k in g.files:
        if k.startswith(f"grad_{kind}/"):
            n = k.split("/", 1)[1]
            got = model.arena.grad_view(n).cpu().numpy()
            scale = max(float(np.abs(g[k]).max()), 1e-30)
            assert float(np.abs(got - g[k]).max()) <= 1e-3 * scale + 2e-6, (n, float(np.abs(got - g[k]).max()), scale)
    # validation path and a few updates run
    assert np.isfinite(float(tr.validation_loss(batch)))
    tr.configure_optimizers()
    for _ in range(3):
        tr.step(batch)
