"""CBV policy plugins for the fine-tuner: the ``rift/cbv/planning`` registry surface
(``CBV_POLICY_LIST[name](config, logger)``; rift/cbv/planning/__init__.py:21-34) restricted to what
the policy-update hot path owns.

``RLFTPluto`` mirrors rift/cbv/planning/fine_tuner/rlft/rlft_pluto.py:32-300 for everything that does not
need a live CARLA world: ``set_buffer`` / ``set_mode`` / ``load_model`` / ``train(e_i)`` / ``save_model`` /
``finish`` with the same checkpoint naming (``carla_episode={e}-epoch={n}-val_loss={v}.ckpt``, newest file
wins, ``current_epoch`` = number of checkpoints drives the closed-loop LR decay), and ``get_action`` with the
reference's return keys (rlft_pluto.py:94-142, rift_pluto.py:28-72, grpo_pluto.py:35-83): one batched policy (and
reference-policy) forward per environment on the GPU, replayed from a CUDA graph per padded shape bucket, ONE
device -> host copy of what the control half needs, then the reference's host arithmetic (top-k trim, PID) from
``rift_b200/controller.py``.  The live CARLA world is reached through ``self.world`` - by default the reference's own
``CarlaDataProvider`` when it is importable, else whatever ``set_world`` was given (tests use a recorded world).
"""
import glob
import os
import re
from collections import defaultdict
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import functional as F
from .controller import PIDController, global_to_local, trim_candidates
from .datamodule import DataModule
from .feature import PlutoFeature
from .planning_model import PlanningModel
from .trainer import TRAINERS, PPOPlutoModel

_DEFAULT_CFG = dict(lr=1e-4, cl_lr_decay=0.9, min_lr=1e-6, weight_decay=1e-5, epochs=16, warmup_epochs=3,
                    trainable_layers=["planning_decoder.pi_head"], train_batch_size=256, train_ratio=0.9, gamma=0.98,
                    lambda_gae_adv=0.98)


class _PrintLogger:
    def log(self, msg, color=None):
        print(msg)


class RLFTPluto:
    name = "rlft_pluto"
    type = "learnable"
    algo = "rift"

    def __init__(self, config: dict, logger=None):
        self.config, self.logger = config, logger or _PrintLogger()
        self.cfg = dict(_DEFAULT_CFG)
        self.cfg.update(config.get("rlft", {}))
        if self.algo == "ppo" and "value_net" not in self.cfg["trainable_layers"]:
            self.cfg["trainable_layers"] = list(self.cfg["trainable_layers"]) + ["value_net"]     # ppo_training.yaml:26-28
        self.initial_lr = self.cfg["lr"]
        self.frame_rate = config.get("frame_rate", 10)
        self.mode = "train"
        self.buffer = None
        self.continue_episode = 0
        self.current_epoch = 0
        self.checkpoint = config.get("ckpt_path")
        root = config.get("ROOT_DIR", ".")
        self.model_dir = os.path.join(root, config.get("model_path", "log/rift_b200"), config.get("load_agent_info", self.name))
        mk = dict(config.get("model", {}))
        radius = config.get("obs", {}).get("radius", 120)
        if self.algo == "ppo":
            p = config.get("ppo", {})
            self.pluto_model = PPOPlutoModel(radius, hidden_dim=tuple(p.get("hidden_dim", (256, 256))),
                                             clip_epsilon=p.get("clip_epsilon", 0.2), lambda_entropy=p.get("lambda_entropy", 0.01),
                                             **mk)
        else:
            self.pluto_model = PlanningModel(radius, **mk)
        self.ref_model: Optional[PlanningModel] = None        # GRPO keeps a frozen copy of the pretrained policy
        self.trainer = None
        # rollout side (pluto.py:32-52)
        self.num_scenario = config.get("num_scenario", 1)
        self._topk = config.get("topk", 10)
        self._use_prediction = config.get("use_prediction", False)
        self._render = config.get("need_video_render", False)
        self._step_interval = 1.0 / self.frame_rate
        self.controllers = defaultdict(lambda: defaultdict(lambda: PIDController(sample_interval=self.frame_rate)))
        self.world = None                      # CarlaDataProvider-like object (set_world / lazy import of the reference's)
        self.traj_evaluator = None             # candidate roll-out evaluator (set_traj_evaluator)
        self._rollout_graphs: Dict[tuple, dict] = {}
        self.use_rollout_graph = bool(int(os.environ.get("RIFT_B200_ROLLOUT_GRAPH", "1")))
        if self._render:
            self.reset_render_data()

    # ------------------------------------------------------------------ CBVBasePolicy surface (base_policy.py:9-52)
    def set_buffer(self, buffer, total_routes=None):
        self.buffer = buffer

    def set_route_planner(self, route_planner):
        self.route_planner = route_planner

    def set_mode(self, mode):
        self.mode = mode

    def get_render_data(self, env_id):
        return self._render_data[env_id] if self._render else None

    def reset_render_data(self):
        keys = ("route_ids_list", "reference_lines_list", "route_waypoints_list", "interaction_wp_list",
                "planning_trajectory_list", "candidate_trajectories_list", "candidate_index_list", "predictions_list")
        self._render_data = {env_id: dict({"ego_states": {}, "nearby_agents_states": {}, "CBV_states": {}},
                                          **{k: [] for k in keys}) for env_id in range(self.num_scenario)}

    def set_world(self, provider):
        """The object get_action asks for actor states: get_actor_by_id, get_history_state, get_ego_vehicle_by_env_id,
        get_CBV_nearby_agents (the reference's CarlaDataProvider static interface, carla_data_provider.py:192,619,665,1031)."""
        self.world = provider

    def set_traj_evaluator(self, evaluator):
        """Candidate roll-out evaluator used in 'train' mode: either the reference's TrajEvaluator (``get_grpo_advantage``
        returns the normalised advantages, traj_evaluator.py:422-475) or an object with ``get_rollout_returns`` (raw
        returns per candidate; normalised here by the bit-exact GPU kernel)."""
        self.traj_evaluator = evaluator

    def _world(self):
        if self.world is None:
            try:
                from rift.scenario.tools.carla_data_provider import CarlaDataProvider      # inside the reference tree
            except Exception as e:
                raise RuntimeError("get_action needs a world: call set_world(provider) with a CarlaDataProvider-like "
                                   "object (the reference's is not importable here)") from e
            self.world = CarlaDataProvider
        return self.world

    def clean_up(self):
        pass

    def log_episode_reward(self, *a, **k):
        pass

    def save_model(self, episode):
        pass        # checkpoints are written by train(), like the reference (rlft_pluto.py:295-296)

    def finish(self):
        pass

    # ------------------------------------------------------------------ checkpoints (rlft_pluto.py:249-293)
    def _episode_files(self):
        pat = re.compile(r"carla_episode=(\d+)")
        files = [f for f in glob.glob(os.path.join(self.model_dir, "*.ckpt")) if pat.search(os.path.basename(f))]
        return sorted(files, key=lambda f: int(pat.search(os.path.basename(f)).group(1)), reverse=True)

    @staticmethod
    def _read_ckpt(path):
        ck = torch.load(path, map_location="cpu", weights_only=False)
        return ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck

    def load_model(self, resume=True):
        files = self._episode_files()
        if resume and files:
            self.checkpoint = files[0]
            self.continue_episode = int(re.search(r"carla_episode=(\d+)", os.path.basename(files[0])).group(1))
            self.current_epoch = len(files)
        else:
            if not resume:
                for f in files:
                    os.unlink(f)
            self.checkpoint = self.config.get("ckpt_path")
            self.continue_episode, self.current_epoch = 0, 0
        if self.checkpoint:
            sd = self._read_ckpt(self.checkpoint)
            if self.algo == "ppo":      # pretrained Pluto checkpoints carry no value net (ppo_pluto.py:112-122)
                have = {k[len("model."):] if k.startswith("model.") else k for k in sd}
                if not any(k.startswith("value_net.") for k in have):
                    self.pluto_model.load_state_dict(sd, strict=False)
                    return
            self.pluto_model.load_state_dict(sd)

    # ------------------------------------------------------------------ the policy update (rlft_pluto.py:206-247)
    def train(self, e_i: int):
        self.logger.log(">> Starting fine-tuning...", color="yellow")
        cfg = self.cfg
        lr = max(self.initial_lr * (cfg["cl_lr_decay"] ** self.current_epoch), cfg["min_lr"])
        tr = TRAINERS[self.algo](self.pluto_model, lr=lr, cl_lr_decay=cfg["cl_lr_decay"], weight_decay=cfg["weight_decay"],
                                 epochs=cfg["epochs"], warmup_epochs=cfg["warmup_epochs"], frame_rate=self.frame_rate,
                                 trainable_layers=cfg["trainable_layers"])
        tr.configure_optimizers()
        dm = DataModule(self.algo, self.buffer, cfg["train_batch_size"], cfg["train_ratio"], gamma=cfg["gamma"],
                        lambda_gae_adv=cfg["lambda_gae_adv"], seed=e_i, device=self.pluto_model.device)
        dm.preprocess_buffer(self.pluto_model, getattr(tr, "value_net", None))
        dm.setup()
        best = None
        os.makedirs(self.model_dir, exist_ok=True)
        for epoch in range(cfg["epochs"]):
            for batch in dm.train_batches():
                tr.step(batch)
            tr.on_train_epoch_end()
            vals = [float(tr.validation_loss(b)) for b in dm.val_batches()]
            val = float(np.mean(vals)) if vals else float("nan")
            # ModelCheckpoint(save_top_k=1, monitor='loss/val_loss', save_weights_only=True)  (training_builder.py:131-140)
            if best is None or val < best[0]:
                if best is not None and os.path.exists(best[1]):
                    os.unlink(best[1])
                path = os.path.join(self.model_dir, f"carla_episode={e_i}-epoch={epoch}-val_loss={val:.4f}.ckpt")
                torch.save({"state_dict": {k: v.cpu() for k, v in tr.state_dict().items()}}, path)
                best = (val, path)
        files = self._episode_files()
        self.current_epoch = len(files)
        if files:
            self.checkpoint = files[0]
            self.pluto_model.load_state_dict(self._read_ckpt(self.checkpoint))
        self.trainer = tr
        if hasattr(self.buffer, "reset_buffer"):
            self.buffer.reset_buffer()
        self.logger.log(">> Finishing fine-tuning...", color="yellow")


    # ------------------------------------------------------------------ rollout: get_action (rlft_pluto.py:94-175)
    _BUCKETS = {"A": 8, "Mp": 16, "R": 2}

    def _rollout_forward(self, feats: List[PlutoFeature], with_ref: bool):
        """One batched forward of the policy (and of the frozen reference policy) for the CBVs of one environment.
        The collated batch is zero-padded up to shape buckets (padded agents / polygons / reference lines are invalid
        rows, exactly what PlutoFeature.collate produces for a shorter sample next to a longer one) and the forward is
        replayed from a CUDA graph per (batch size, bucket) signature."""
        batch = PlutoFeature.collate(feats).data
        model = self.pluto_model
        dev = model.device
        outputs = ("trajectory", "candidate_trajectories", "ref_free_trajectory") + (("prediction",) if self._use_prediction else ())
        if not self.use_rollout_graph:
            out = model.forward(batch, outputs=outputs)
            ref = self.ref_model.forward(batch, outputs=())["probability"] if with_ref and self.ref_model is not None else None
            return batch, out, ref
        pad = _pad_to_buckets(batch, self._BUCKETS)
        key = tuple((k, kk, tuple(v.shape)) for k, d in pad.items() if isinstance(d, dict) for kk, v in d.items()) + \
            (tuple(pad["current_state"].shape), bool(with_ref and self.ref_model is not None))
        g = self._rollout_graphs.get(key)
        if g is not None and (g["ws_gen"] != model.ws_generation or
                              (g["ref_gen"] is not None and g["ref_gen"] != self.ref_model.ws_generation)):
            g = None
        if g is None:
            static = {k: ({kk: torch.as_tensor(v).to(dev, copy=True) for kk, v in d.items()} if isinstance(d, dict)
                          else torch.as_tensor(d).to(dev, copy=True)) for k, d in pad.items()}
            pb = model.pack(static)
            model.forward(pb, outputs=outputs)                      # eager once: sizes the workspace, refreshes weight planes
            if key[-1]:
                self.ref_model.forward(pb, outputs=())
            graph = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap):
                with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):
                    out = model.forward(pb, outputs=outputs)
                    ref = self.ref_model.forward(pb, outputs=())["probability"] if key[-1] else None
            torch.cuda.current_stream(dev).wait_stream(cap)
            if len(self._rollout_graphs) >= 32:
                self._rollout_graphs.clear()
            g = {"graph": graph, "pb": pb, "out": out, "ref": ref, "ws_gen": model.ws_generation,
                 "ref_gen": self.ref_model.ws_generation if key[-1] else None, "w_gen": None}
            self._rollout_graphs[key] = g
        # refresh the static inputs, replay
        flat_src = _flat_features(pad)
        for name, t in g["pb"].keep.items():
            src = flat_src[name]
            t.copy_(src.view(torch.uint8) if src.dtype == torch.bool else src, non_blocking=True)
        g["graph"].replay()
        bs, A, Mp, R = (batch["agent"]["heading"].shape[0], batch["agent"]["heading"].shape[1],
                        batch["map"]["valid_mask"].shape[1], batch["reference_line"]["valid_mask"].shape[1])
        out = dict(g["out"])
        for k in ("probability", "trajectory", "candidate_trajectories", "r_padding_mask"):
            if k in out:
                out[k] = out[k][:, :R]
        if "prediction" in out:
            out["prediction"] = out["prediction"][:, :max(A - 1, 0)]
        ref = g["ref"][:, :R] if g["ref"] is not None else None
        return batch, out, ref

    def _host_outputs(self, out, ref_prob):
        """ONE device -> host transfer per environment of what the control half reads (the reference does four .cpu()
        calls per CBV: rift_pluto.py:79-93)."""
        host = {"probability": out["probability"].float().cpu().numpy(),
                "candidate_trajectories": out["candidate_trajectories"].cpu().numpy(),
                "ref_free_trajectory": _output_ref_free(out["ref_free_trajectory"]).cpu().numpy()}
        if self._use_prediction and "prediction" in out:
            host["prediction"] = out["prediction"].cpu().numpy()
        if ref_prob is not None:
            host["ref_probability"] = ref_prob.float().cpu().numpy()
        return host

    def _control(self, env_id, CBV_id, CBV_obs, CBV_state, cand, prob, ref_free, predictions):
        """Top-k trim, best candidate, PID control (pluto.py:150-194).  cand (R', Mo, T, 3), prob (R', Mo)."""
        origin, angle = CBV_state.rear_axle.array, CBV_state.rear_axle.heading
        traj, score, orig, n_ref, n_mode = trim_candidates(cand.astype(np.float64), prob, origin, angle, self._topk,
                                                           None if ref_free is None else ref_free.astype(np.float64))
        best = int(score.argmax())
        trajectory = traj[best, 1:]
        local = global_to_local(trajectory, origin, angle)
        speed = CBV_state.dynamic_car_state.center_velocity_2d.magnitude()
        throttle, steer, brake = self.controllers[env_id][CBV_id].control_pid(local[:, :2], speed)
        if self._render:
            rd = self._render_data[env_id]
            rd["CBV_states"][CBV_id] = CBV_state
            for k, src in (("route_ids_list", "route_ids"), ("reference_lines_list", "reference_lines"),
                           ("route_waypoints_list", "route_waypoints"), ("interaction_wp_list", "interaction_wp")):
                rd[k].append(CBV_obs.get(src))
            rd["planning_trajectory_list"].append(trajectory)
            rd["candidate_trajectories_list"].append(traj)
            rd["candidate_index_list"].append(best)
            rd["predictions_list"].append(predictions)
        return (throttle, steer, brake), score, orig, best, n_mode

    def _group_terms(self, env_id, CBV_id, history, batch, out, host, index):
        """Train mode: group-relative advantages of the valid reference lines' candidates + the logits that produced
        them (rift_pluto.py:113-145)."""
        world = self._world()
        rl = batch["reference_line"]
        valid_pts = torch.as_tensor(rl["valid_mask"][index])
        r_valid = valid_pts.any(-1)
        rv = r_valid.cpu().numpy()
        if self.traj_evaluator is None:
            raise RuntimeError("train-mode get_action needs a candidate roll-out evaluator (set_traj_evaluator)")
        ego_id = world.get_ego_vehicle_by_env_id(env_id).id
        nearby = world.get_CBV_nearby_agents(ego_id, CBV_id)
        raw = out["trajectory"][index][r_valid.to(out["trajectory"].device)]
        pos = torch.as_tensor(rl["position"][index])[r_valid]
        ang = torch.as_tensor(rl["orientation"][index])[r_valid]
        vm = valid_pts[r_valid]
        ref_pos = [p[m] for p, m in zip(pos, vm)]
        ref_ang = [a[m] for a, m in zip(ang, vm)]
        if hasattr(self.traj_evaluator, "get_rollout_returns"):
            ret = np.asarray(self.traj_evaluator.get_rollout_returns(history, raw, ref_pos, ref_ang, nearby), np.float64)
            adv = F.group_advantage(torch.from_numpy(ret.reshape(1, -1)).to(self.pluto_model.device)).cpu().numpy()
            group_advantage = {"advantage": adv.reshape(int(rv.sum()), -1), "valid_mask": np.ones((int(rv.sum()), adv.size // max(int(rv.sum()), 1)), np.bool_)}
        else:
            group_advantage = self.traj_evaluator.get_grpo_advantage(history, raw, ref_pos, ref_ang, nearby)
        logits = host["probability"][index][rv]
        old = {"logits": logits, "valid_mask": np.ones_like(logits, dtype=np.bool_)}
        refl = None
        if "ref_probability" in host:
            rl_ = host["ref_probability"][index][rv]
            refl = {"logits": rl_, "valid_mask": np.ones_like(rl_, dtype=np.bool_)}
        return old, refl, group_advantage

    def _clean_CBVs(self, infos, CBVs_obs_list):
        """pluto.py:114-125 - drop the PID state of CBVs that left the scene."""
        for info, CBVs_obs in zip(infos, CBVs_obs_list):
            env_id = info["env_id"]
            if env_id in self.controllers:
                for cbv_id in list(self.controllers[env_id].keys()):
                    if cbv_id not in CBVs_obs:
                        del self.controllers[env_id][cbv_id]
                if not self.controllers[env_id]:
                    del self.controllers[env_id]

    GROUP_KEYS = False      # RIFT / GRPO: old (and reference) group logits + group advantage; else PPO-style log-prob + mode
    WITH_REF = False

    def get_action(self, CBVs_obs_list, infos, deterministic=False) -> Dict[str, List[Dict[Any, Any]]]:
        n = self.num_scenario
        actions = [{} for _ in range(n)]
        a_old, a_ref, a_adv = [{} for _ in range(n)], [{} for _ in range(n)], [{} for _ in range(n)]
        a_lp, a_mode = [{} for _ in range(n)], [{} for _ in range(n)]
        world = self._world()
        with torch.no_grad():
            for info, CBVs_obs in zip(infos, CBVs_obs_list):
                if not CBVs_obs:
                    continue
                env_id = info["env_id"]
                batch, out, ref_prob = self._rollout_forward([o["raw_pluto_feature"] for o in CBVs_obs.values()], self.WITH_REF)
                host = self._host_outputs(out, ref_prob)
                for index, (CBV_id, CBV_obs) in enumerate(CBVs_obs.items()):
                    history = world.get_history_state(world.get_actor_by_id(CBV_id))
                    state = history[-1]
                    cand, prob = host["candidate_trajectories"][index], host["probability"][index]
                    pred = host.get("prediction", [None] * (index + 1))[index] if self._use_prediction else None
                    if self.GROUP_KEYS:
                        act, _, _, _, _ = self._control(env_id, CBV_id, CBV_obs, state, cand, prob, host["ref_free_trajectory"][index], pred)
                        actions[env_id][CBV_id] = act
                        if self.mode == "train":
                            a_old[env_id][CBV_id], a_ref[env_id][CBV_id], a_adv[env_id][CBV_id] = \
                                self._group_terms(env_id, CBV_id, history, batch, out, host, index)
                        else:
                            a_old[env_id][CBV_id] = a_ref[env_id][CBV_id] = a_adv[env_id][CBV_id] = None
                    else:       # rlft_pluto.py:144-175: valid reference lines only, log-prob and (r, m) of the chosen candidate
                        rv = np.asarray(torch.as_tensor(batch["reference_line"]["valid_mask"][index]).any(-1).cpu().numpy())
                        act, score, orig, best, n_mode = self._control(env_id, CBV_id, CBV_obs, state, cand[rv], prob[rv],
                                                                       host["ref_free_trajectory"][index], pred)
                        actions[env_id][CBV_id] = act
                        a_lp[env_id][CBV_id] = np.log(score[best] + 1e-12)
                        oi = int(orig[best])
                        a_mode[env_id][CBV_id] = (oi // n_mode, oi % n_mode)
        if self._render:
            for info in infos:
                env_id = info["env_id"]
                ego = world.get_ego_vehicle_by_env_id(env_id)
                nearby = world.get_ego_nearby_agents(ego.id)
                self._render_data[env_id].update({"ego_states": {ego.id: world.get_current_state(ego)},
                                                  "nearby_agents_states": {a.id: world.get_current_state(a) for a in nearby}})
        self._clean_CBVs(infos, CBVs_obs_list)
        data = {"CBVs_actions": actions}
        if self.GROUP_KEYS:
            data["CBVs_actions_old_group_logits"] = a_old
            if self.WITH_REF:
                data["CBVs_actions_ref_group_logits"] = a_ref
            data["CBVs_group_advantage"] = a_adv
        else:
            data["CBVs_actions_old_log_prob"] = a_lp
            data["CBVs_actions_mode"] = a_mode
        return data

    # ------------------------------------------------------------------ tensor half of get_action
    @torch.no_grad()
    def policy_outputs(self, features: List[PlutoFeature], returns: Optional[List[np.ndarray]] = None) -> Dict:
        """Batched forward for the CBVs of one env (rift_pluto.py:28-72): logits, best (r, m) per CBV, candidate
        trajectories; with `returns` (one (R_valid * Mo,) float64 array per CBV, the roll-out returns of
        TrajEvaluator) also the group-relative advantages (traj_evaluator.py:466-470) padded to (R, Mo)."""
        batch = PlutoFeature.collate(features)
        out = self.pluto_model.forward(batch.data, outputs=("trajectory", "candidate_trajectories"))
        prob = out["probability"]
        bs, R, Mo = prob.shape
        best = prob.reshape(bs, -1).argmax(-1)
        res = {"probability": prob, "best_r": best // Mo, "best_m": best % Mo, "r_padding_mask": out["r_padding_mask"],
               "candidate_trajectories": out["candidate_trajectories"], "trajectory": out["trajectory"]}
        if self.ref_model is not None:
            res["ref_probability"] = self.ref_model.forward(batch.data, outputs=())["probability"]
        if returns is not None:
            sizes = [len(r) for r in returns]
            offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=prob.device)
            flat = torch.from_numpy(np.concatenate(returns).astype(np.float64)).to(prob.device)
            adv = F.group_advantage(flat, offs)
            padded = torch.zeros(bs, R * Mo, dtype=torch.float64, device=prob.device)
            mask = torch.zeros(bs, R * Mo, dtype=torch.bool, device=prob.device)
            for b, n in enumerate(sizes):
                padded[b, :n] = adv[int(offs[b]):int(offs[b]) + n]
                mask[b, :n] = True
            res["group_advantage"] = padded.view(bs, R, Mo)
            res["group_advantage_mask"] = mask.view(bs, R, Mo)
        return res


class RIFTPluto(RLFTPluto):
    name, algo = "rift_pluto", "rift"
    GROUP_KEYS = True


class GRPOPluto(RLFTPluto):
    name, algo = "grpo_pluto", "grpo"
    GROUP_KEYS, WITH_REF = True, True

    def load_model(self, resume=True):
        super().load_model(resume)
        if self.config.get("ckpt_path"):       # frozen reference policy = the pretrained checkpoint (grpo_pluto.py:35-60)
            self.ref_model = PlanningModel.from_config(self.pluto_model.cfg, device=self.pluto_model.device)
            self.ref_model.load_state_dict(self._read_ckpt(self.config["ckpt_path"]))


class PPOPluto(RLFTPluto):
    name, algo = "ppo_pluto", "ppo"


class ReinforcePluto(RLFTPluto):
    name, algo = "reinforce_pluto", "reinforce"


def _output_ref_free(ref_free: torch.Tensor) -> torch.Tensor:
    """output_ref_free_trajectory (pluto_model.py:207-215): (x, y, atan2(sin, cos)) from the (bs, T, 4) head."""
    return torch.cat([ref_free[..., :2], torch.atan2(ref_free[..., 3], ref_free[..., 2]).unsqueeze(-1)], dim=-1)


def _pad_to_buckets(batch: Dict, buckets: Dict[str, int]) -> Dict:
    """Zero-pad the agent / polygon / reference-line axes of a collated PlutoFeature.data up to multiples of the bucket
    sizes (padding = invalid rows, the same thing PlutoFeature.collate's pad_sequence produces)."""
    def up(n, m):
        return max(m, (n + m - 1) // m * m)

    def pad1(t, n):
        t = torch.as_tensor(t)
        if t.shape[1] == n:
            return t
        out = torch.zeros((t.shape[0], n) + tuple(t.shape[2:]), dtype=t.dtype)
        out[:, :t.shape[1]] = t
        return out
    A = up(batch["agent"]["heading"].shape[1], buckets["A"])
    Mp = up(batch["map"]["valid_mask"].shape[1], buckets["Mp"])
    R = up(batch["reference_line"]["valid_mask"].shape[1], buckets["R"])
    out = {"agent": {k: pad1(v, A) for k, v in batch["agent"].items()},
           "map": {k: pad1(v, Mp) for k, v in batch["map"].items()},
           "reference_line": {k: pad1(v, R) for k, v in batch["reference_line"].items()},
           "current_state": torch.as_tensor(batch["current_state"])}
    if "static_objects" in batch:
        out["static_objects"] = batch["static_objects"]
    return out


def _flat_features(data: Dict) -> Dict[str, torch.Tensor]:
    """PackedBatch.keep names -> source tensors."""
    out = {}
    for grp, prefix in (("agent", "agent_"), ("map", "map_"), ("reference_line", "ref_")):
        for k, v in data[grp].items():
            out[prefix + k] = torch.as_tensor(v)
    out["current_state"] = torch.as_tensor(data["current_state"])
    return out


CBV_POLICY_LIST = {"rift_pluto": RIFTPluto, "grpo_pluto": GRPOPluto, "ppo_pluto": PPOPluto, "reinforce_pluto": ReinforcePluto}
