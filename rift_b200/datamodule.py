"""Buffer -> batch plumbing of the fine-tuner (rift/cbv/planning/fine_tuner/rlft/*/*_datamodule.py).

``*Collate`` reproduce the reference collate callables key for key; ``DataModule`` is the Lightning-free
equivalent of ``RIFTDataModule`` & co: 90/10 random split, shuffled mini-batches of ``train_batch_size``,
``preprocess_buffer`` (no-op for RIFT / GRPO, GAE pass for PPO, discounted returns for REINFORCE).
The buffer is used through the reference's own read-only interface (``get_key_data``, ``sample``,
``add_extra_data``, ``buffer_capacity`` / ``__len__``; rift/gym_carla/buffer/cbv_rollout_buffer.py:77-138).
"""
from typing import Dict, List

import numpy as np
import torch

from . import functional as F
from .feature import PlutoFeature, pad_first_dim


def _t(x):
    return x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))


class RIFTCollate:
    """rift_pluto/rift_datamodule.py:20-51"""
    extra = ()

    def __call__(self, batch: List[Dict]) -> Dict:
        assert len(batch) > 0, "Batch size has to be greater than 0!"
        out = {
            "cur_pluto_feature_torch": PlutoFeature.collate([d["CBVs_obs"]["raw_pluto_feature"] for d in batch]),
            "group_advantage_torch": pad_first_dim([_t(d["CBVs_group_advantage"]["advantage"]) for d in batch]),
            "group_advantage_mask_torch": pad_first_dim([_t(d["CBVs_group_advantage"]["valid_mask"]) for d in batch]),
            "old_group_logits_torch": pad_first_dim([_t(d["CBVs_actions_old_group_logits"]["logits"]) for d in batch]),
            "old_group_logits_mask_torch": pad_first_dim([_t(d["CBVs_actions_old_group_logits"]["valid_mask"]) for d in batch]),
        }
        return out


class GRPOCollate(RIFTCollate):
    """grpo_pluto/grpo_datamodule.py:20-57 (adds the frozen reference model's logits)"""

    def __call__(self, batch):
        out = super().__call__(batch)
        out["ref_group_logits_torch"] = pad_first_dim([_t(d["CBVs_actions_ref_group_logits"]["logits"]) for d in batch])
        out["ref_group_logits_mask_torch"] = pad_first_dim([_t(d["CBVs_actions_ref_group_logits"]["valid_mask"]) for d in batch])
        return out


class PPOCollate:
    """ppo_pluto/ppo_datamodule.py:40-70"""

    def __call__(self, batch):
        assert len(batch) > 0, "Batch size has to be greater than 0!"
        out = {"cur_pluto_feature_torch": PlutoFeature.collate([d["CBVs_obs"]["raw_pluto_feature"] for d in batch])}
        for key, src in (("state_torch", "CBVs_state"), ("advantage_torch", "CBVs_advantage"),
                         ("reward_sum_torch", "CBVs_reward_sum"), ("old_log_prob_torch", "CBVs_old_log_prob"),
                         ("action_mode_torch", "CBVs_action_mode")):
            out[key] = torch.stack([_t(d[src]) for d in batch], dim=0)
        return out


class ReinforceCollate:
    """reinforce_pluto/reinforce_datamodule.py:41-64"""

    def __call__(self, batch):
        assert len(batch) > 0, "Batch size has to be greater than 0!"
        return {"cur_pluto_feature_torch": PlutoFeature.collate([d["CBVs_obs"]["raw_pluto_feature"] for d in batch]),
                "return_torch": torch.stack([_t(d["CBVs_return"]) for d in batch], dim=0)}


class SFTCollate:
    """fine_tuner/sft/sft_datamodule.py:18-41"""

    def __call__(self, batch):
        assert len(batch) > 0, "Batch size has to be greater than 0!"
        return {"cur_pluto_feature_torch": PlutoFeature.collate([d["CBVs_obs"]["raw_pluto_feature"] for d in batch]),
                "teacher_infos": torch.stack([_t(d["CBVs_teacher_infos"]) for d in batch], dim=0)}


class RTRCollate(PPOCollate):
    """fine_tuner/sft/rtr_pluto/rtr_datamodule.py:40-75 (PPO terms + the teacher infos)"""

    def __call__(self, batch):
        out = super().__call__(batch)
        out["teacher_infos"] = torch.stack([_t(d["CBVs_teacher_infos"]) for d in batch], dim=0)
        return out


COLLATES = {"rift": RIFTCollate, "grpo": GRPOCollate, "ppo": PPOCollate, "reinforce": ReinforceCollate,
            "sft": SFTCollate, "rtr": RTRCollate, "rs": ReinforceCollate}


class DataModule:
    def __init__(self, algo: str, buffer, train_batch_size=256, train_ratio=0.9, shuffle=True, gamma=0.98,
                 lambda_gae_adv=0.98, seed=0, device="cuda"):
        self.algo, self.buffer, self.collate = algo, buffer, COLLATES[algo]()
        self.train_batch_size, self.train_ratio, self.shuffle = train_batch_size, train_ratio, shuffle
        self.gamma, self.lambda_gae_adv = gamma, lambda_gae_adv
        self.device = device
        self.gen = torch.Generator().manual_seed(seed)
        self.train_idx = self.val_idx = None

    def __len__(self):
        return getattr(self.buffer, "buffer_capacity", None) or len(self.buffer)

    # ------------------------------------------------------------------ buffer passes
    def preprocess_buffer(self, model=None, value_net=None):
        """RIFT / GRPO: nothing (advantages were computed at rollout time, rift_datamodule.py:97-98).
        PPO: two no-grad sweeps for hidden / value, GAE scan, reward_sum, normalisation (ppo_datamodule.py:117-174).
        REINFORCE: discounted returns (reinforce_datamodule.py:19-38)."""
        b = self.buffer
        if self.algo == "reinforce":
            rewards = torch.from_numpy(np.stack(b.get_key_data("CBVs_reward"), 0)).float().to(self.device)
            dones = torch.from_numpy(np.stack(b.get_key_data("CBVs_done"), 0)).float().to(self.device)
            b.add_extra_data({"CBVs_return": F.discounted_return(rewards, dones, self.gamma).cpu()})
        elif self.algo == "ppo":
            rewards = torch.from_numpy(np.stack(b.get_key_data("CBVs_reward"), 0)).float().to(self.device)
            undones = 1.0 - torch.from_numpy(np.stack(b.get_key_data("CBVs_done"), 0)).float().to(self.device)
            unterm = 1.0 - torch.from_numpy(np.stack(b.get_key_data("CBVs_terminated"), 0)).float().to(self.device)
            old_log_prob = torch.from_numpy(np.stack(b.get_key_data("CBVs_actions_old_log_prob"), 0))
            action_mode = torch.from_numpy(np.stack(b.get_key_data("CBVs_actions_mode"), 0))

            def sweep(key):
                feats = [o["raw_pluto_feature"] for o in b.get_key_data(key)]
                hs, vs = [], []
                for i in range(0, len(feats), self.train_batch_size):
                    chunk = PlutoFeature.collate(feats[i:i + self.train_batch_size])
                    out = model.forward(chunk.data, outputs=("hidden",))
                    hs.append(out["hidden"])
                    vs.append(value_net.forward(out["hidden"]))
                return torch.cat(hs, 0), torch.cat(vs, 0)

            state, value = sweep("CBVs_obs")
            _, next_value = sweep("CBVs_next_obs")
            adv, reward_sum, adv_n = F.gae(rewards, undones, value, next_value, unterm, self.gamma, self.lambda_gae_adv)
            b.add_extra_data({"CBVs_state": state.cpu(), "CBVs_advantage": adv_n.cpu(), "CBVs_reward_sum": reward_sum.cpu(),
                              "CBVs_old_log_prob": old_log_prob, "CBVs_action_mode": action_mode})

    # ------------------------------------------------------------------ split + iteration
    def setup(self):
        n = len(self)
        perm = torch.randperm(n, generator=self.gen).tolist()
        n_train = int(round(n * self.train_ratio))
        self.train_idx, self.val_idx = perm[:n_train], perm[n_train:]

    def _batches(self, idx, shuffle):
        if shuffle:
            order = torch.randperm(len(idx), generator=self.gen).tolist()
            idx = [idx[i] for i in order]
        on_device = hasattr(self.buffer, "collate_device")       # DeviceRolloutBuffer: one gather kernel, no host collate
        for i in range(0, len(idx), self.train_batch_size):
            chunk = idx[i:i + self.train_batch_size]
            yield self.buffer.collate_device(chunk, self.algo) if on_device else self.collate([self.buffer.sample(j) for j in chunk])

    def train_batches(self):
        if self.train_idx is None:
            self.setup()
        return self._batches(self.train_idx, self.shuffle)

    def val_batches(self):
        if self.val_idx is None:
            self.setup()
        return self._batches(self.val_idx, False)
