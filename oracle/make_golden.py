"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

Every golden is produced by the reference's own modules (imported through oracle/ref_shim.py):
`PlanningModel.forward`, the four `LightningTrainer._compute_objectives`, the reference
`configure_optimizers()` parameter grouping + torch AdamW + clip_grad_norm_(0.5) (Lightning's
step order), numpy's mean/std for the group advantage, and the source text of
`get_advantages_GAE` / `compute_return` exec'd from the reference files.  Inputs and parameters
are NOT stored: they are regenerated from seeds by rift_b200/synth.py (numpy PCG64, portable).

Deterministic parity mode: trunk in eval() (dropout/droppath identity, BatchNorm
running statistics), parameters randomised by synth_state_dict (seed 7).
"""
import ast
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim  # noqa: E402
from oracle.pluto_oracle import to_torch  # noqa: E402
from rift_b200.config import MODEL_ZOO, param_spec, is_buffer  # noqa: E402
from rift_b200.synth import synth_state_dict, synth_features, synth_rl_extras  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: model, model kwargs, batch shape, ragged
    "cfg1_small": dict(model="small", kw=dict(future_steps=40), bs=1, A=8, Mp=20, R=1, ragged=False),
    "ragged_small": dict(model="small", kw={}, bs=3, A=9, Mp=11, R=3, ragged=True),
    "medium_tiny": dict(model="medium", kw={}, bs=2, A=6, Mp=7, R=2, ragged=True),
}
TRAINER_KW = dict(lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3, frame_rate=10)


def _ref_function(relpath: str, name: str, extra_globals=None):
    """exec one top-level function of a reference file without importing the file."""
    src = open(os.path.join(ref_shim.REF, relpath)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            g = {"torch": torch, "np": np}
            g.update(extra_globals or {})
            exec(compile(ast.Module([node], []), relpath, "exec"), g)
            return g[name]
    raise KeyError(name)


def build_batch(cfg, case, algo):
    PlutoFeature = ref_shim.pluto_feature_cls()
    feats = synth_features(cfg, case["bs"], case["A"], case["Mp"], case["R"], seed=1, ragged=case["ragged"])
    ex = synth_rl_extras(cfg, feats, seed=2)
    batch = {"cur_pluto_feature_torch": PlutoFeature(data=to_torch(feats))}
    for k in ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits",
              "state", "advantage", "reward_sum", "old_log_prob", "action_mode", "return"):
        batch[k + "_torch"] = torch.from_numpy(ex[k].copy())
    return batch


def build_model(cfg, algo):
    if algo == "ppo":
        Model = ref_shim.ppo_pluto_model_cls()
        m = Model(radius=cfg.radius, state_dim=cfg.dim, action_dim=1, hidden_dim=list(cfg.value_hidden),
                  dim=cfg.dim, num_heads=cfg.num_heads, encoder_depth=cfg.encoder_depth,
                  decoder_depth=cfg.decoder_depth, future_steps=cfg.future_steps)
    else:
        m = ref_shim.planning_model_cls()(
            radius=cfg.radius, dim=cfg.dim, num_heads=cfg.num_heads, encoder_depth=cfg.encoder_depth,
            decoder_depth=cfg.decoder_depth, future_steps=cfg.future_steps)
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    m.load_state_dict(sd, strict=True)
    return m


class _Store(dict):
    """Large tensors are stored as an every-7th-element sample plus (sum, l2) so the fixtures stay small."""

    def put(self, key, arr):
        arr = np.asarray(arr)
        if arr.size > 20000:
            flat = arr.reshape(-1)
            self[key + "@s7"] = flat[::7].copy()
            self[key + "@stats"] = np.array([flat.astype(np.float64).sum(),
                                             np.sqrt((flat.astype(np.float64) ** 2).sum())])
        else:
            self[key] = arr.copy()


def run_case(name, case):
    out = _Store()
    for algo in ("rift", "grpo", "ppo", "reinforce"):
        kw = dict(case["kw"])
        if algo == "ppo":
            kw["value_hidden"] = (256, 256)
        cfg = MODEL_ZOO[case["model"]](**kw)
        for mode in ("pi_head", "full"):
            if mode == "full" and algo in ("reinforce",):
                continue
            model = build_model(cfg, algo)
            layers = ["planning_decoder.pi_head"] if mode == "pi_head" else \
                [n for n, _ in model.named_children()]
            if algo == "ppo" and mode == "pi_head":
                layers = layers + ["value_net"]
            Trainer = ref_shim.trainer_cls(algo)
            tr = Trainer(model=model, trainable_layers=layers, **TRAINER_KW)
            tr.train()
            tr.model.eval()                       # deterministic parity mode
            batch = build_batch(cfg, case, algo)
            data = batch["cur_pluto_feature_torch"].data
            res = tr.forward(data)
            if algo == "rift" and mode == "pi_head":
                for k in ("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory",
                          "output_trajectory", "candidate_trajectories"):
                    out.put("out_" + k, res[k].detach().numpy())
                out["out_best_index"] = res["probability"].detach().reshape(case["bs"], -1).argmax(-1).numpy()
            prob = res["probability"]
            prob.retain_grad()
            losses = tr._compute_objectives(res, data, batch)
            loss = losses["loss"]
            out[f"loss_{algo}"] = np.asarray(loss.detach().double().numpy())
            loss.backward()
            if mode == "pi_head":
                out[f"dlogits_{algo}"] = prob.grad.numpy().copy()
            named = dict(tr.model.named_parameters())
            if mode == "pi_head" and algo in ("rift", "grpo", "ppo"):
                for n, p in named.items():
                    if p.grad is not None and n.startswith("planning_decoder.pi_head"):
                        out.put(f"grad_{algo}/{n}", p.grad.numpy())
                if algo == "ppo":
                    for n, p in named.items():
                        if n.startswith("value_net") and p.grad is not None:
                            out.put(f"grad_{algo}/{n}", p.grad.numpy())
            if mode == "full" and algo in ("rift", "grpo", "ppo"):
                names = [n for n, _, _ in param_spec(cfg) if not is_buffer(n)]
                stats = np.zeros((len(names), 3), np.float64)
                for i, n in enumerate(names):
                    g = named[n].grad
                    if g is not None:
                        g = g.double()
                        stats[i] = (g.sum(), g.pow(2).sum().sqrt(), g.flatten()[g.numel() // 2])
                out[f"fullgrad_{algo}_stats"] = stats
                # a few complete tensors from different corners of the graph
                for n in ("pos_emb.freqs.weight", "agent_encoder.history_encoder.embed.proj.weight",
                          "agent_encoder.history_encoder.levels.1.blocks.0.attn.rpb",
                          "agent_encoder.ego_state_emb.query", "agent_encoder.type_emb.weight",
                          "map_encoder.polygon_encoder.first_mlp.0.weight",
                          "map_encoder.speed_limit_emb.mlps.0.3.bias", "encoder_blocks.0.attn.in_proj_bias",
                          "planning_decoder.m_pos", "planning_decoder.m_emb",
                          "planning_decoder.r_encoder.second_mlp.1.weight",
                          "planning_decoder.decoder_blocks.1.norm2.weight",
                          "planning_decoder.cat_x_proj.bias"):
                    g = named[n].grad
                    out.put(f"fullgrad_{algo}/{n}", (torch.zeros_like(named[n]) if g is None else g).numpy())
            # optimizer: the reference grouping, torch AdamW, Lightning's clip -> step order
            if algo in ("grpo", "ppo") and mode == "pi_head":
                mod = sys.modules[Trainer.__module__]
                saved = mod.WarmupCosLR
                mod.WarmupCosLR = lambda **k: None      # ctor is incompatible with torch>=2.2
                try:
                    (opt,), _ = tr.configure_optimizers()
                finally:
                    mod.WarmupCosLR = saved
                if algo == "grpo":
                    pd = dict(tr.named_parameters())
                    inv = {id(p): n for n, p in pd.items()}
                    out["optim_groups"] = np.array(json.dumps(
                        [sorted(inv[id(p)][len("model."):] for p in g["params"]) for g in opt.param_groups]))
                for step in range(3):
                    if step > 0:
                        tr.zero_grad()
                        b2 = build_batch(cfg, case, algo)
                        r2 = tr.forward(b2["cur_pluto_feature_torch"].data)
                        l2 = tr._compute_objectives(r2, b2["cur_pluto_feature_torch"].data, b2)["loss"]
                        l2.backward()
                        out[f"loss_{algo}_step{step}"] = np.asarray(l2.detach().double().numpy())
                    tn = torch.nn.utils.clip_grad_norm_(
                        [p for p in tr.parameters() if p.requires_grad], 0.5)
                    out[f"gradnorm_{algo}_step{step}"] = np.asarray(tn.double().numpy())
                    opt.step()
                    if step in (0, 2):
                        for n, p in named.items():
                            if p.requires_grad:
                                out.put(f"param_{algo}_step{step + 1}/{n}", p.detach().numpy())
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, len(out), "arrays", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def advantage_goldens():
    """traj_evaluator.py:466-469 evaluated literally by numpy, per group."""
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    for G in (2, 7, 8, 12, 24, 72, 96, 128, 129, 144, 300, 1000):
        ret = rng.normal(-5.0, 20.0, (6, G))
        ret[1] = ret[1, 0]                      # constant group -> std 0 -> divides by 1e-5
        ret[2, : G // 2] = -200.0               # collision-dominated group
        adv = np.empty_like(ret)
        for i in range(ret.shape[0]):
            returns = ret[i]
            mean_return = np.mean(returns)
            std_return = np.std(returns) + 1e-5
            adv[i] = (returns - mean_return) / std_return
        out[f"ret_{G}"] = ret
        out[f"adv_{G}"] = adv
    np.savez_compressed(os.path.join(GOLDEN, "advantage.npz"), **out)
    print("advantage", len(out))


def buffer_pass_goldens():
    gae = _ref_function("rift/cbv/planning/fine_tuner/rlft/ppo_pluto/ppo_datamodule.py", "get_advantages_GAE")
    ret = _ref_function("rift/cbv/planning/fine_tuner/rlft/reinforce_pluto/reinforce_datamodule.py", "compute_return")
    rng = np.random.Generator(np.random.PCG64(12))
    n = 4096
    rewards = torch.from_numpy(rng.normal(-0.5, 2.0, n).astype(np.float32))
    dones = torch.from_numpy((rng.uniform(size=n) < 0.02).astype(np.float32))
    term = torch.from_numpy(((rng.uniform(size=n) < 0.5) & (dones.numpy() > 0)).astype(np.float32))
    values = torch.from_numpy(rng.normal(0, 3.0, n).astype(np.float32))
    next_values = torch.from_numpy(rng.normal(0, 3.0, n).astype(np.float32))
    adv = gae(rewards, 1 - dones, values, next_values, 1 - term)
    reward_sum = adv + values
    adv_n = (adv - adv.mean()) / (adv.std(dim=0) + 1e-5)      # ppo_datamodule.py:163
    out = dict(rewards=rewards.numpy(), dones=dones.numpy(), terminated=term.numpy(), values=values.numpy(),
               next_values=next_values.numpy(), gae=adv.numpy(), reward_sum=reward_sum.numpy(),
               gae_normalised=adv_n.numpy())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # its trailing "normalise" takes std of a 0-d tensor
        r = ret(rewards.clone(), dones.clone())
    out["discounted_return"] = r.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "buffer_pass.npz"), **out)
    print("buffer_pass ok")


class _State:
    """Duck-typed CarlaAgentState: rear_axle.array / .heading, dynamic_car_state.center_velocity_2d.magnitude()."""

    def __init__(self, x, y, heading, speed):
        class _P: pass
        self.rear_axle = _P(); self.rear_axle.array = np.array([x, y], np.float64); self.rear_axle.heading = float(heading)
        v = _P(); v.magnitude = lambda s=float(speed): s
        self.dynamic_car_state = _P(); self.dynamic_car_state.center_velocity_2d = v


def get_action_goldens():
    """The host half of get_action executed by the reference's own code: PLUTO._trim_candidates / _global_to_local
    (pluto.py:196-278) and PIDController.control_pid (pid_controller.py:57-107) over a sequence of ticks."""
    from scipy.special import softmax
    trim = ref_shim.ref_method("rift/cbv/planning/pluto/pluto.py", "PLUTO", "_trim_candidates", {"softmax": softmax})
    g2l = ref_shim.ref_method("rift/cbv/planning/pluto/pluto.py", "PLUTO", "_global_to_local")
    PID = ref_shim.pid_controller_cls()
    rng = np.random.Generator(np.random.PCG64(21))
    out = {}
    class _Self: _topk = 10
    for case, (R, nvalid, with_free) in enumerate([(6, 6, True), (6, 2, True), (3, 1, False), (1, 1, True)]):
        ctrl = PID(sample_interval=10)
        for tick in range(4):
            # forward-moving candidates: arc length grows along T so the PID sees plausible speeds
            speed_c = rng.uniform(0.0, 12.0, (R, 12, 1))
            s = np.cumsum(np.full((R, 12, 80), 0.1) * speed_c, -1)
            curv = rng.normal(0, 0.02, (R, 12, 1))
            th = curv * s
            cand = np.stack([s * np.cos(th), s * np.sin(th), th], -1).astype(np.float32)
            prob = rng.normal(0, 2.0, (R, 12)).astype(np.float32)
            prob[nvalid:] = -1e6
            free = None
            if with_free:
                sf = np.cumsum(np.full(80, 0.1) * rng.uniform(0, 10.0))
                free = np.stack([sf, 0.05 * sf, np.full(80, 0.05)], -1).astype(np.float32)
            st = _State(rng.normal(0, 50), rng.normal(0, 50), rng.normal(0, 1.5), rng.uniform(0, 10) if tick else 0.0)
            traj, score, orig, n_ref, n_mode = trim(_Self, cand.astype(np.float64), prob, st,
                                                    None if free is None else free.astype(np.float64))
            best = int(score.argmax())
            local = g2l(None, traj[best, 1:], st)
            thr, steer, brake = ctrl.control_pid(local[:, :2], st.dynamic_car_state.center_velocity_2d.magnitude())
            k = f"c{case}t{tick}_"
            out[k + "cand"], out[k + "prob"] = cand, prob
            if free is not None:
                out[k + "free"] = free
            out[k + "state"] = np.array([*st.rear_axle.array, st.rear_axle.heading, st.dynamic_car_state.center_velocity_2d.magnitude()])
            out[k + "traj"], out[k + "score"], out[k + "orig"] = traj, score, orig
            out[k + "local"] = local
            out[k + "control"] = np.array([float(thr), float(steer), float(bool(brake))])
    np.savez_compressed(os.path.join(GOLDEN, "get_action.npz"), **out)
    print("get_action", len(out))


def sft_goldens():
    """SFT / RTR / RS objectives evaluated by the reference trainers (fine_tuner/sft/*) on the ragged small case:
    loss, d loss / d logits, the teacher label, pi_head (and value-net) gradients."""
    name, case = "ragged_small", CASES["ragged_small"]
    out = {}
    rng = np.random.Generator(np.random.PCG64(31))
    teacher = np.stack([rng.uniform(0, 9, case["bs"]), rng.normal(0, 2, case["bs"]), rng.normal(0, 2, case["bs"]),
                        rng.normal(0, 0.5, case["bs"]), rng.uniform(0, 9, case["bs"])], -1).astype(np.float32)
    out["teacher_infos"] = teacher
    for kind in ("sft", "rtr", "rs"):
        kw = dict(case["kw"])
        if kind == "rtr":
            kw["value_hidden"] = (256, 256)
        cfg = MODEL_ZOO[case["model"]](**kw)
        model = build_model(cfg, "ppo" if kind == "rtr" else kind)
        layers = ["planning_decoder.pi_head"] + (["value_net"] if kind == "rtr" else [])
        Trainer = ref_shim.sft_trainer_cls(kind)
        tr = Trainer(model=model, trainable_layers=layers, **TRAINER_KW)
        tr.train()
        tr.model.eval()
        batch = build_batch(cfg, case, kind)
        batch["teacher_infos"] = torch.from_numpy(teacher)
        data = batch["cur_pluto_feature_torch"].data
        res = tr.forward(data)
        prob = res["probability"]
        prob.retain_grad()
        if kind in ("sft", "rtr"):
            bs, R, M = prob.shape
            pm = prob.detach().clone().masked_fill(~data["reference_line"]["valid_mask"].any(-1).unsqueeze(-1), -1e8)
            mx = torch.argmax(pm.view(bs, -1), dim=1)
            lab, _ = tr.generate_target_label(res["trajectory"].detach(), pm, batch["teacher_infos"], mx // M, mx % M)
            out[f"label_{kind}"] = lab.argmax(-1).numpy()
        loss = tr._compute_objectives(res, data, batch)["loss"]
        out[f"loss_{kind}"] = np.asarray(loss.detach().double().numpy())
        loss.backward()
        out[f"dlogits_{kind}"] = prob.grad.numpy().copy()
        for n, p in tr.model.named_parameters():
            if p.grad is not None and (n.startswith("planning_decoder.pi_head") or n.startswith("value_net")):
                out[f"grad_{kind}/{n}"] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN, "sft_objectives.npz"), **out)
    print("sft_objectives", len(out))


def evaluator_goldens():
    """The candidate-rollout evaluator executed by the reference's own code where it runs without CARLA / shapely:
    TrajEvaluator.get_ref_line_info / get_center_rollout / get_rollout_return (traj_evaluator.py:115-158,333-420),
    TrackPropagate.propagate + derive_kinematics (track_propogate.py:500-699), DenseRewardModel, KinematicBicycleModel's
    other-vehicle forecast (pdm_lite/kinematic_bicycle_model.py:33-61).  Two consecutive calls on the SAME TrackPropagate
    object, so the PID buffers that the reference never resets carry over exactly as they do in a rollout."""
    tp = ref_shim.track_propagate_module()
    Reward = ref_shim.dense_reward_model_cls()
    rel = "rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py"
    ref_line_info = ref_shim.ref_method(rel, "TrajEvaluator", "get_ref_line_info")
    center_rollout = ref_shim.ref_method(rel, "TrajEvaluator", "get_center_rollout")
    rollout_return = ref_shim.ref_method(rel, "TrajEvaluator", "get_rollout_return")

    class _Self:
        pass
    me = _Self()
    me.center_rollout_model = tp.TrackPropagate(virtual_time_step=0.1)
    me.reward_model = Reward()
    rng = np.random.Generator(np.random.PCG64(41))
    out = {}
    for call, (R, speed0) in enumerate([(3, 4.0), (2, 0.0), (6, 9.0)]):
        M, T = 12, 80
        # raw model-like trajectories: arc with per-candidate speed and curvature, (x, y, cos, sin, vx, vy)
        spd = rng.uniform(0.0, 11.0, (R, M, 1))
        s_ = np.cumsum(np.full((R, M, T), 0.1) * spd, -1)
        th = rng.normal(0, 0.015, (R, M, 1)) * s_ + rng.normal(0, 0.05, (R, M, 1))
        traj = np.stack([s_ * np.cos(th), s_ * np.sin(th) + rng.normal(0, 0.3, (R, M, 1)), np.cos(th), np.sin(th),
                         spd * np.cos(th), spd * np.sin(th)], -1).astype(np.float32)
        n_r = rng.integers(30, 121, R)
        ref_pos = [np.stack([np.arange(n) * 1.0, 0.02 * np.arange(n) ** 1.3 * rng.normal()], -1).astype(np.float32) for n in n_r]
        ref_ang = [np.arctan2(np.gradient(p[:, 1]), np.gradient(p[:, 0])).astype(np.float32) for p in ref_pos]
        x0, y0, h0 = rng.normal(0, 40), rng.normal(0, 40), rng.normal(0, 1.5)

        class _S:
            pass
        st = _S(); st.rear_axle = _S(); st.rear_axle.array = np.array([x0, y0]); st.rear_axle.heading = h0
        st.dynamic_car_state = _S(); st.dynamic_car_state.speed = speed0
        st.car_footprint = _S(); st.car_footprint.width, st.car_footprint.length = 2.0, 4.6
        tt = torch.from_numpy(traj)[:, :, :40]
        dd, da = ref_line_info(me, tt, [torch.from_numpy(p) for p in ref_pos], [torch.from_numpy(a) for a in ref_ang])
        c, a, v, acc, yr, ya, vert = center_rollout(me, tt.clone(), [st])
        col = rng.uniform(size=(R * M, 80)) < 0.01
        off = rng.uniform(size=(R * M, 80)) < 0.03
        ret = rollout_return(me, dd, da, v, acc, yr, ya, col, off)
        k = f"call{call}_"
        out[k + "traj"] = traj
        for i, (p, an) in enumerate(zip(ref_pos, ref_ang)):
            out[k + f"refpos{i}"], out[k + f"refang{i}"] = p, an
        out[k + "state"] = np.array([x0, y0, h0, speed0, 2.0, 4.6])
        out[k + "delta_dis"], out[k + "delta_angle"] = dd, da
        out[k + "center"], out[k + "angle"], out[k + "speed"], out[k + "acc"] = c, a, v, acc
        out[k + "yaw_rate"], out[k + "yaw_acc"], out[k + "vertices"] = yr, ya, vert
        out[k + "collision"], out[k + "off_road"], out[k + "return"] = col, off, ret
    # other-vehicle forecast (numpy, pdm_lite model)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_kbm", os.path.join(ref_shim.REF, "rift/ego/pdm_lite/kinematic_bicycle_model.py"))
    kbm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kbm)

    class _Cfg:
        time_step, front_wheel_base, rear_wheel_base, steering_gain = 0.1, -0.090769015, 1.4178275, 0.36848336
        brake_acceleration, throttle_acceleration = -4.952399, 0.5633837
        throttle_values = brake_values = None
        throttle_threshold_during_forecasting = 0.3
    model = kbm.KinematicBicycleModel(_Cfg())
    N = 5
    loc = rng.normal(0, 30, (N, 3)); hd = rng.normal(0, 1.5, N); sp = rng.uniform(0, 12, N)
    act = np.stack([rng.uniform(-0.5, 0.5, N), rng.uniform(0, 1, N), (rng.uniform(size=N) < 0.3).astype(float)], -1)
    out["other_loc"], out["other_heading_rad"], out["other_speed"], out["other_action"] = loc, hd, sp, act
    fl, fh, fv = [], [], []
    l, h, v = loc.copy(), hd.copy(), sp.copy()
    for i in range(40):
        l, h, v = model.forecast_other_vehicles(l, h, v, act)
        fl.append(l.copy()); fh.append(h.copy()); fv.append(v.copy())
    out["other_future_loc"], out["other_future_heading"], out["other_future_speed"] = np.stack(fl), np.stack(fh), np.stack(fv)
    np.savez_compressed(os.path.join(GOLDEN, "evaluator.npz"), **out)
    print("evaluator", len(out))


def state_dict_spec():
    spec = {}
    for mname, kw in (("small", {}), ("medium", {})):
        cfg = MODEL_ZOO[mname](**kw)
        m = ref_shim.planning_model_cls()(radius=cfg.radius, dim=cfg.dim, num_heads=cfg.num_heads,
                                          encoder_depth=cfg.encoder_depth, decoder_depth=cfg.decoder_depth)
        spec[mname] = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    json.dump(spec, open(os.path.join(GOLDEN, "state_dict_spec.json"), "w"))
    print("state_dict_spec", {k: len(v) for k, v in spec.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]
    if not only or "spec" in only:
        state_dict_spec()
    if not only or "adv" in only:
        advantage_goldens()
    if not only or "buf" in only:
        buffer_pass_goldens()
    if not only or "act" in only:
        get_action_goldens()
    if not only or "sft" in only:
        sft_goldens()
    if not only or "eval" in only:
        evaluator_goldens()
    for name, case in CASES.items():
        if not only or name in only:
            run_case(name, case)
