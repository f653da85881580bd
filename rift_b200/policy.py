"""CBV policy plugins for the fine-tuner: the ``rift/cbv/planning`` registry surface
(``CBV_POLICY_LIST[name](config, logger)``; rift/cbv/planning/__init__.py:21-34) restricted to what
the policy-update hot path owns.

``RLFTPluto`` mirrors rift/cbv/planning/fine_tuner/rlft/rlft_pluto.py:32-300 for everything that does not
need a live CARLA world: ``set_buffer`` / ``set_mode`` / ``load_model`` / ``train(e_i)`` / ``save_model`` /
``finish`` with the same checkpoint naming (``carla_episode={e}-epoch={n}-val_loss={v}.ckpt``, newest file
wins, ``current_epoch`` = number of checkpoints drives the closed-loop LR decay), plus the tensor half of
``get_action`` (``policy_outputs``: batched forward, argmax, old / reference logits and the group-relative
advantage normalisation on the GPU).  The CARLA halves of ``get_action`` (feature building from the world,
PID control, candidate roll-outs) are out of scope for this path and stay with the reference; INTEGRATION.md
shows the three-line patch that makes the reference's plugin delegate to this one.
"""
import glob
import os
import re
from typing import Dict, List, Optional

import numpy as np
import torch

from . import functional as F
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

    # ------------------------------------------------------------------ CBVBasePolicy surface (base_policy.py:9-52)
    def set_buffer(self, buffer, total_routes=None):
        self.buffer = buffer

    def set_route_planner(self, route_planner):
        self.route_planner = route_planner

    def set_mode(self, mode):
        self.mode = mode

    def get_render_data(self, env_id):
        return None

    def reset_render_data(self):
        pass

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


class GRPOPluto(RLFTPluto):
    name, algo = "grpo_pluto", "grpo"

    def load_model(self, resume=True):
        super().load_model(resume)
        if self.config.get("ckpt_path"):       # frozen reference policy = the pretrained checkpoint (grpo_pluto.py:35-60)
            self.ref_model = PlanningModel.from_config(self.pluto_model.cfg, device=self.pluto_model.device)
            self.ref_model.load_state_dict(self._read_ckpt(self.config["ckpt_path"]))


class PPOPluto(RLFTPluto):
    name, algo = "ppo_pluto", "ppo"


class ReinforcePluto(RLFTPluto):
    name, algo = "reinforce_pluto", "reinforce"


CBV_POLICY_LIST = {"rift_pluto": RIFTPluto, "grpo_pluto": GRPOPluto, "ppo_pluto": PPOPluto, "reinforce_pluto": ReinforcePluto}
