"""``PlutoFeature`` — mirror of the reference feature container
(rift/cbv/planning/pluto/feature_builder/pluto_feature.py:18-163): ``collate`` (zero-pad the ragged
first dimension of the agent / map / reference_line / static_objects sub-dicts, stack the per-sample
tensors), ``to_device``, ``to_feature_tensor``, ``is_valid``.  Only the RL fine-tuning form (anchor sample
only: no contrastive ``data_p`` / ``data_n``) is needed on this path.

``collate`` is host-side layout work; ``collate_into_pinned`` additionally packs the batch into ONE pinned
host buffer so the step's host->device traffic is a single async copy.
"""
from dataclasses import dataclass
from typing import Any, Dict, List

import numpy as np
import torch

PAD_KEYS = ("agent", "map", "reference_line", "static_objects")
STACK_KEYS = ("current_state", "origin", "angle", "cost_maps")


def _as_tensor(v):
    """pluto/utils/utils.py:12-30 — float64 -> float32, everything else keeps its dtype."""
    if torch.is_tensor(v):
        return v.float() if v.dtype == torch.float64 else v
    a = np.asarray(v)
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    if not a.flags.c_contiguous:
        a = a.copy()
    return torch.from_numpy(a) if a.ndim else torch.tensor(a.item(), dtype=torch.from_numpy(a.reshape(1)).dtype)


def pad_first_dim(tensors: List[torch.Tensor]) -> torch.Tensor:
    """torch.nn.utils.rnn.pad_sequence(batch_first=True): zero-fill to the longest first dimension."""
    n = max(t.shape[0] for t in tensors)
    out = tensors[0].new_zeros((len(tensors), n) + tuple(tensors[0].shape[1:]))
    for i, t in enumerate(tensors):
        out[i, : t.shape[0]] = t
    return out


@dataclass
class PlutoFeature:
    data: Dict[str, Any]
    data_p: Dict[str, Any] = None
    data_n: Dict[str, Any] = None
    data_n_info: Dict[str, Any] = None

    @classmethod
    def collate(cls, feature_list: List["PlutoFeature"]) -> "PlutoFeature":
        if feature_list[0].data_p is not None or feature_list[0].data_n is not None:
            raise NotImplementedError("contrastive samples (data_p / data_n) are an imitation-pretraining feature; "
                                      "the RL fine-tuner never produces them")
        first = feature_list[0].data
        batch = {}
        for key in PAD_KEYS:
            if key in first:
                batch[key] = {k: pad_first_dim([_as_tensor(f.data[key][k]) for f in feature_list]) for k in first[key]}
        for key in STACK_KEYS:
            if key in first:
                batch[key] = torch.stack([_as_tensor(f.data[key]) for f in feature_list], dim=0)
        return PlutoFeature(data=batch)

    def to_feature_tensor(self) -> "PlutoFeature":
        def conv(v):
            return {k: conv(x) for k, x in v.items()} if isinstance(v, dict) else _as_tensor(v)
        return PlutoFeature(data={k: conv(v) for k, v in self.data.items()})

    def to_device(self, device) -> "PlutoFeature":
        def mv(v):
            return {k: mv(x) for k, x in v.items()} if isinstance(v, dict) else v.to(device, non_blocking=True)
        return PlutoFeature(data={k: mv(v) for k, v in self.data.items()})

    def serialize(self) -> Dict[str, Any]:
        return {"data": self.data}

    @classmethod
    def deserialize(cls, data: Dict[str, Any]) -> "PlutoFeature":
        return PlutoFeature(data=data["data"])

    @property
    def is_valid(self) -> bool:
        if "reference_line" in self.data:
            return bool(self.data["reference_line"]["valid_mask"].any())
        return self.data["map"]["point_position"].shape[0] > 0
