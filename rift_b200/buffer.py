"""Device-resident rollout buffer with GPU collate.

``DeviceRolloutBuffer`` keeps the interface and the trajectory bookkeeping of the reference's ``CBVRolloutBuffer``
(rift/gym_carla/buffer/cbv_rollout_buffer.py:16-138: ``store`` / ``process_data_dict`` / ``add_extra_data`` /
``get_key_data`` / ``sample`` / ``reset_buffer``, ``buffer_capacity`` / ``buffer_pos`` / ``buffer_full``) and mirrors what
the policy update reads into packed device arenas while the rollout runs:

* every tensor of ``CBVs_obs['raw_pluto_feature']`` that the policy reads (PackedBatch fields), zero padded to the
  arena's capacity along its ragged first dimension (agents / polygons / reference lines),
* ``CBVs_group_advantage``, ``CBVs_actions_old_group_logits`` (+ ``..._ref_group_logits``) padded over reference lines,
* whatever ``add_extra_data`` brings (PPO: state, advantage, reward_sum, old_log_prob, action_mode; REINFORCE: return).

``collate_device(indices, algo)`` then builds the mini-batch with ONE gather kernel (``rift_b200_gather_fields``):
``PlutoFeature.collate`` + ``RIFTCollate`` & co. zero-pad a mini-batch to its longest item, which on the padded slots is
"copy the first n_batch rows of each selected slot".  The result is a ``PackedBatch`` plus device tensors - the step's
host-to-device traffic is the 8-byte-per-sample index vector.
"""
from collections import defaultdict, deque
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .planning_model import PackedBatch

# (group, key, PackedBatch name, trailing shape after the ragged dim, stored dtype); ragged dim: A / Mp / R
_FEATURE_FIELDS = (
    ("agent", "position", "agent_position", "A", torch.float32), ("agent", "heading", "agent_heading", "A", torch.float32),
    ("agent", "velocity", "agent_velocity", "A", torch.float32), ("agent", "shape", "agent_shape", "A", torch.float32),
    ("agent", "category", "agent_category", "A", torch.int8), ("agent", "valid_mask", "agent_valid_mask", "A", torch.uint8),
    ("map", "point_position", "map_point_position", "Mp", torch.float32), ("map", "point_vector", "map_point_vector", "Mp", torch.float32),
    ("map", "point_orientation", "map_point_orientation", "Mp", torch.float32), ("map", "polygon_center", "map_polygon_center", "Mp", torch.float32),
    ("map", "polygon_type", "map_polygon_type", "Mp", torch.int8), ("map", "polygon_on_route", "map_polygon_on_route", "Mp", torch.uint8),
    ("map", "polygon_tl_status", "map_polygon_tl_status", "Mp", torch.int8),
    ("map", "polygon_has_speed_limit", "map_polygon_has_speed_limit", "Mp", torch.uint8),
    ("map", "polygon_speed_limit", "map_polygon_speed_limit", "Mp", torch.float32), ("map", "valid_mask", "map_valid_mask", "Mp", torch.uint8),
    ("reference_line", "position", "ref_position", "R", torch.float32), ("reference_line", "vector", "ref_vector", "R", torch.float32),
    ("reference_line", "orientation", "ref_orientation", "R", torch.float32), ("reference_line", "valid_mask", "ref_valid_mask", "R", torch.uint8),
)
# per-algorithm padded group terms: buffer key -> ((sub key, batch key, dtype), ...)
_GROUP_TERMS = {
    "CBVs_group_advantage": (("advantage", "group_advantage_torch", torch.float64), ("valid_mask", "group_advantage_mask_torch", torch.uint8)),
    "CBVs_actions_old_group_logits": (("logits", "old_group_logits_torch", torch.float32), ("valid_mask", "old_group_logits_mask_torch", torch.uint8)),
    "CBVs_actions_ref_group_logits": (("logits", "ref_group_logits_torch", torch.float32), ("valid_mask", "ref_group_logits_mask_torch", torch.uint8)),
}
_ALGO_GROUP_KEYS = {"rift": ("CBVs_group_advantage", "CBVs_actions_old_group_logits"),
                    "grpo": ("CBVs_group_advantage", "CBVs_actions_old_group_logits", "CBVs_actions_ref_group_logits"),
                    "ppo": (), "reinforce": ()}
_ALGO_EXTRA = {"ppo": (("CBVs_state", "state_torch"), ("CBVs_advantage", "advantage_torch"), ("CBVs_reward_sum", "reward_sum_torch"),
                       ("CBVs_old_log_prob", "old_log_prob_torch"), ("CBVs_action_mode", "action_mode_torch")),
               "reinforce": (("CBVs_return", "return_torch"),), "rift": (), "grpo": ()}


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


class DeviceRolloutBuffer:
    name = "DeviceRolloutBuffer"

    def __init__(self, num_scenario=1, mode="train_cbv", cbv_config: Optional[dict] = None, logger=None, device="cuda",
                 caps: Optional[Dict[str, int]] = None):
        assert mode == "train_cbv", f"Only initialize {self.name} when training the rl-based onpolicy cbv agent"
        cbv_config = cbv_config or {}
        self.num_scenario, self.mode, self.logger = num_scenario, mode, logger
        self.buffer_capacity = int(cbv_config.get("buffer_capacity", 4096))
        self.data_keys = list(cbv_config.get("data_keys", ("CBVs_actions", "CBVs_actions_old_group_logits", "CBVs_group_advantage",
                                                           "CBVs_obs", "CBVs_next_obs", "CBVs_reward", "CBVs_terminated", "CBVs_done")))
        self.device = torch.device(device)
        self.caps = dict(A=16, Mp=32, R=4)           # grown on demand (doubling) when an item exceeds them
        self.caps.update(caps or {})
        self.reset_buffer()

    # ------------------------------------------------------------------ reference bookkeeping (cbv_rollout_buffer.py:36-103)
    def reset_buffer(self):
        self.buffer_pos = 0
        self.buffer_full = False
        self.buffer_data = {key: deque(maxlen=self.buffer_capacity) for key in self.data_keys}
        self.temp_buffer = {key: defaultdict(list) for key in self.buffer_data}
        self._arena: Dict[str, torch.Tensor] = {}     # device mirror, allocated at the first store
        self._extent = {k: np.zeros(self.buffer_capacity, np.int32) for k in ("A", "Mp", "R")}
        self._extra_dev: Dict[str, torch.Tensor] = {}
        self._collate_cache: Dict[tuple, dict] = {}   # staged outputs of collate_device per (algo, bs, extents)
        self._arena_gen = 0                           # bumped whenever an arena tensor is (re)allocated: cached field tables hold its pointers

    def _arena_generation(self):
        return self._arena_gen

    def __len__(self):
        return self.buffer_capacity if self.buffer_full else self.buffer_pos

    def process_data_dict(self, data_dict):
        """Per-CBV trajectories accumulate in the temporary buffer and move on, whole, when the CBV is done."""
        processed = {key: [] for key in self.buffer_data}
        lengths = set(len(data) for key, data in data_dict.items() if key in self.buffer_data)
        assert len(lengths) == 1, "all the data in the data dict should have same length"
        n = lengths.pop()
        for i in range(n):
            for CBV_id in data_dict["CBV_ids"][i]:
                for key, value in self.temp_buffer.items():
                    value[CBV_id].append(data_dict[key][i][CBV_id])
                if data_dict["CBVs_done"][i][CBV_id]:
                    for key, value in processed.items():
                        value.extend(self.temp_buffer[key].pop(CBV_id))
        lens = set(len(v) for v in processed.values())
        assert len(lens) == 1, "the data in the processed data dict should have same length"
        return processed, lens.pop()

    def store(self, data_dict):
        processed, n = self.process_data_dict(data_dict)
        if n <= 5:                                   # too short a trajectory is ignored
            return
        take = n
        if self.buffer_pos + n >= self.buffer_capacity:
            take = min(n, self.buffer_capacity - self.buffer_pos)
            self.buffer_full = True
        first = self.buffer_pos
        for key, data in self.buffer_data.items():
            data.extend(processed[key][:take])
        self.buffer_pos += take
        if take > 0:
            self._mirror(first, {k: v[:take] for k, v in processed.items()})

    def add_extra_data(self, data_dict: dict):
        assert self.buffer_full, "only add data when the buffer is full"
        assert all(len(v) == self.buffer_capacity for v in data_dict.values())
        self.buffer_data.update(data_dict)
        for k, v in data_dict.items():               # tensors / arrays also live on the device for collate_device
            try:
                t = v if torch.is_tensor(v) else torch.as_tensor(np.stack([_np(x) for x in v], 0))
                self._extra_dev[k] = t.to(self.device)
            except Exception:
                pass

    def get_key_data(self, key: str):
        assert self.buffer_full, "only get the data when the buffer is full"
        return self.buffer_data[key]

    def sample(self, idx):
        assert self.buffer_full, "only sample the data when the buffer is full"
        indices = idx if isinstance(idx, (list, tuple)) else [idx]
        assert all(0 <= i < self.buffer_capacity for i in indices)
        return {key: [d[i] for i in indices] if len(indices) > 1 else d[indices[0]] for key, d in self.buffer_data.items()}

    def get_all_np_data(self):
        assert self.buffer_pos == self.buffer_capacity, "only get the data when the buffer is full"
        return {key: np.stack(d).reshape(self.buffer_capacity, -1) for key, d in self.buffer_data.items()}

    # ------------------------------------------------------------------ device mirror
    def _alloc(self, name, cap, trailing, dtype):
        self._arena[name] = torch.zeros((self.buffer_capacity, cap) + tuple(trailing), dtype=dtype, device=self.device)
        self._arena_gen += 1

    def _grow(self, dim, need):
        new = self.caps[dim]
        while new < need:
            new *= 2
        for name, t in list(self._arena.items()):
            if getattr(t, "_ragged", None) == dim:
                big = torch.zeros((t.shape[0], new) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
                big[:, : t.shape[1]] = t
                big._ragged = dim
                self._arena[name] = big
        self.caps[dim] = new
        self._arena_gen += 1

    def _put(self, name, dim, first, arrays, dtype):
        """arrays: one (n_i, ...) array per item -> slots [first, first + len) of arena `name`, zero padded."""
        n_max = max(a.shape[0] for a in arrays)
        if n_max > self.caps[dim]:
            self._grow(dim, n_max)
        if name not in self._arena:
            self._alloc(name, self.caps[dim], arrays[0].shape[1:], dtype)
            self._arena[name]._ragged = dim
        host = np.zeros((len(arrays), n_max) + tuple(arrays[0].shape[1:]), arrays[0].dtype)
        for i, a in enumerate(arrays):
            host[i, : a.shape[0]] = a
        t = torch.from_numpy(host)
        if t.dtype == torch.bool:
            t = t.view(torch.uint8)
        self._arena[name][first:first + len(arrays), :n_max] = t.to(self.device, non_blocking=True).to(dtype)

    def _mirror(self, first, items: Dict[str, list]):
        if "CBVs_obs" in items:
            feats = [o["raw_pluto_feature"].data for o in items["CBVs_obs"]]
            for grp, key, name, dim, dtype in _FEATURE_FIELDS:
                arrays = [_np(f[grp][key]) for f in feats]
                arrays = [a.astype(np.float32) if a.dtype == np.float64 else a for a in arrays]
                self._put(name, dim, first, arrays, dtype)
            for i, f in enumerate(feats):
                self._extent["A"][first + i] = _np(f["agent"]["heading"]).shape[0]
                self._extent["Mp"][first + i] = _np(f["map"]["valid_mask"]).shape[0]
                self._extent["R"][first + i] = _np(f["reference_line"]["valid_mask"]).shape[0]
            cs = np.stack([_np(f["current_state"]).astype(np.float32) for f in feats], 0)
            if "current_state" not in self._arena:
                self._arena["current_state"] = torch.zeros((self.buffer_capacity, 1, cs.shape[1]), dtype=torch.float32, device=self.device)
                self._arena_gen += 1
            self._arena["current_state"][first:first + len(feats), 0] = torch.from_numpy(cs).to(self.device)
        for bkey, terms in _GROUP_TERMS.items():
            if bkey in items:
                for sub, name, dtype in terms:
                    self._put(name, "R", first, [_np(d[sub]) for d in items[bkey]], dtype)

    # ------------------------------------------------------------------ GPU collate
    def collate_device(self, indices, algo: str = "rift", reuse: bool = True) -> Dict:
        """The mini-batch the trainer reads, built on the device from slot indices: equals
        ``{RIFT,GRPO,PPO,Reinforce}Collate()([buffer.sample(i) for i in indices])`` value for value.

        With ``reuse`` (default) the output tensors - and the ``PackedBatch`` object around them - are cached per
        (algo, batch size, padded extents) and OVERWRITTEN by the next call of the same shape, like a loader's staging
        buffer: the trainer's step graph then takes the batch in place as its static input (no per-step copies, one graph
        per shape).  ``reuse=False`` returns fresh tensors."""
        idx_host = np.asarray(indices, np.int64)
        bs = len(idx_host)
        n_b = {d: int(self._extent[d][idx_host].max()) for d in ("A", "Mp", "R")}
        key = (algo, bs, n_b["A"], n_b["Mp"], n_b["R"], self._arena_generation())
        slot = self._collate_cache.get(key) if reuse else None
        if slot is None:
            fields, outs = [], {}

            def add(name, n_rows):
                src = self._arena[name]
                row_bytes = src[0, 0].numel() * src.element_size()
                dst = torch.empty((bs, n_rows) + tuple(src.shape[2:]), dtype=src.dtype, device=self.device)
                outs[name] = dst
                if n_rows > 0:
                    fields.append(_lib.GatherField(src.data_ptr(), dst.data_ptr(), src.shape[1] * row_bytes, n_rows * row_bytes, n_rows * row_bytes))

            for _, _, name, dim, _ in _FEATURE_FIELDS:
                add(name, n_b[dim])
            add("current_state", 1)
            for bkey in _ALGO_GROUP_KEYS[algo]:
                for _, name, _ in _GROUP_TERMS[bkey]:
                    add(name, n_b["R"])
            data = {"agent": {}, "map": {}, "reference_line": {}, "current_state": outs["current_state"][:, 0]}
            for grp, k, name, _, _ in _FEATURE_FIELDS:
                data[grp][k] = outs[name]
            batch = {"cur_pluto_feature_torch": PackedBatch(data, self.device)}
            for bkey in _ALGO_GROUP_KEYS[algo]:
                for _, name, dtype in _GROUP_TERMS[bkey]:
                    batch[name] = outs[name].view(torch.bool) if dtype == torch.uint8 else outs[name]
            slot = {"arr": (_lib.GatherField * len(fields))(*fields), "n": len(fields), "outs": outs, "batch": batch,
                    "idx": torch.empty(bs, dtype=torch.int64, device=self.device),
                    "idx_host": torch.empty(bs, dtype=torch.int64).pin_memory() if self.device.type == "cuda" else torch.empty(bs, dtype=torch.int64)}
            if reuse:
                if len(self._collate_cache) >= 16:
                    self._collate_cache.clear()
                self._collate_cache[key] = slot
        slot["idx_host"].copy_(torch.from_numpy(idx_host))
        slot["idx"].copy_(slot["idx_host"], non_blocking=True)
        _lib.check(_lib.lib().rift_b200_gather_fields(slot["arr"], slot["n"], _lib.ptr(slot["idx"]), bs, _lib.stream_ptr()), "gather_fields")
        batch = dict(slot["batch"])
        for bkey, name in _ALGO_EXTRA[algo]:
            batch[name] = self._extra_dev[bkey].index_select(0, slot["idx"])
        return batch
