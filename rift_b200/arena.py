"""Flat fp32 parameter / gradient / optimizer-state arenas.

Layout, decided here and handed to the engine as (name, offset, numel, trainable) entries:

    [ trainable + weight-decay | trainable, no decay | frozen parameters | buffers ]

so that (a) the optimizer apply is ONE launch over a contiguous range with a single index compare
for the decay switch, (b) the data-parallel gradient all-reduce is ONE collective over a
contiguous buffer, (c) checkpoints keep the reference's state-dict names (views into the arena).

The decay / no-decay rule is the outcome of the reference's ``configure_optimizers``
(rift/cbv/planning/fine_tuner/rlft/rift_pluto/rift_trainer.py:287-324) on the Pluto module tree;
the trainable rule is ``freeze_parameters`` (:78-90).
"""
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

from .config import PlutoConfig, param_spec, is_buffer, numel

ALIGN = 64   # elements (256 B): every tensor starts TMA / float4 aligned


def module_names(cfg: PlutoConfig) -> set:
    """Every name ``dict(model.named_modules())`` would contain for these parameters."""
    out = {""}
    for name, _, _ in param_spec(cfg):
        parts = name.split(".")
        for i in range(1, len(parts)):
            out.add(".".join(parts[:i]))
    return out


def trainable_names(cfg: PlutoConfig, trainable_layers: Iterable[str]) -> List[str]:
    mods = module_names(cfg)
    names = []
    for layer in trainable_layers:
        if layer not in mods:
            raise ValueError(f"Layer {layer} not found in the model.")   # same error as the reference
    for name, _, _ in param_spec(cfg):
        if is_buffer(name):
            continue
        if any(name == l or name.startswith(l + ".") for l in trainable_layers):
            names.append(name)
    return names


# Parameters whose outputs never reach an RL objective: autograd leaves their .grad = None in the
# reference, and torch.optim.AdamW skips such parameters entirely (no update, no weight decay).  They are
# therefore laid out with the frozen parameters even when their module is listed as trainable.
NO_GRAD_PREFIXES = ("static_objects_encoder.", "agent_predictor.", "hidden_proj.", "ref_free_decoder.",
                    "planning_decoder.loc_head.", "planning_decoder.yaw_head.", "planning_decoder.vel_head.")


def receives_gradient(name: str) -> bool:
    return not name.startswith(NO_GRAD_PREFIXES)


def is_decay(name: str, shape: Tuple[int, ...]) -> bool:
    """True for Linear / Conv1d / MultiheadAttention weights; biases, LayerNorm / BatchNorm /
    Embedding weights and bare parameters (m_emb, m_pos, query, pos_embed, rpb) do not decay."""
    leaf = name.rsplit(".", 1)[-1]
    if "bias" in leaf:
        return False
    if "weight" in leaf:
        is_embedding = name.endswith("emb.weight") or name.endswith("freqs.weight")
        return len(shape) >= 2 and not is_embedding
    return False


class ParamArena:
    def __init__(self, cfg: PlutoConfig, trainable_layers: Iterable[str] = (), device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        spec = param_spec(cfg)
        self.spec = {n: (tuple(s), k) for n, s, k in spec}
        train = set(n for n in trainable_names(cfg, trainable_layers) if receives_gradient(n))
        groups = {0: [], 1: [], 2: [], 3: []}
        for n, s, k in spec:
            if k != "f32":
                continue
            if is_buffer(n):
                groups[3].append(n)
            elif n in train:
                groups[0 if is_decay(n, tuple(s)) else 1].append(n)
            else:
                groups[2].append(n)
        self.offsets: Dict[str, int] = {}
        off = 0
        self.bounds = []
        for g in range(4):
            for n in sorted(groups[g]) if g < 2 else groups[g]:
                self.offsets[n] = off
                off += (numel(self.spec[n][0]) + ALIGN - 1) // ALIGN * ALIGN
            self.bounds.append(off)
        self.n_decay, self.n_train, _, self.total = self.bounds
        self.trainable = train
        self.decay_names = sorted(groups[0])
        self.no_decay_names = sorted(groups[1])
        self.params = torch.zeros(max(self.total, ALIGN), dtype=torch.float32, device=self.device)
        # + ALIGN tail: slots [n_train, n_train+4) carry (objective sum, valid count) through the gradient all-reduce
        self.grads = torch.zeros(self.n_train + ALIGN, dtype=torch.float32, device=self.device) if train else None
        self.int_buffers = {n: torch.zeros(s, dtype=torch.int64) for n, (s, k) in self.spec.items() if k == "i64"}

    # ---- views
    def view(self, name: str) -> torch.Tensor:
        shape, _ = self.spec[name]
        o = self.offsets[name]
        return self.params[o:o + numel(shape)].view(shape)

    def grad_view(self, name: str) -> torch.Tensor:
        assert name in self.trainable, name
        shape, _ = self.spec[name]
        o = self.offsets[name]
        return self.grads[o:o + numel(shape)].view(shape)

    def entries(self):
        return [(n, self.offsets[n], numel(self.spec[n][0]), int(n in self.trainable)) for n in self.offsets]

    # ---- state dict (reference key names; 'model.'-prefixed variants are handled by the callers)
    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for n, (s, k) in self.spec.items():
            out[n] = self.int_buffers[n].clone() if k == "i64" else self.view(n).detach().clone()
        return out

    def load_state_dict(self, sd, strict: bool = True):
        missing = [n for n in self.spec if n not in sd]
        unexpected = [n for n in sd if n not in self.spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        host = torch.zeros(self.params.numel(), dtype=torch.float32)
        host.copy_(self.params)                       # keep whatever is not in sd
        for n, v in sd.items():
            if n not in self.spec:
                continue
            shape, k = self.spec[n]
            t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v)
            if tuple(t.shape) != shape:
                raise RuntimeError(f"load_state_dict: shape mismatch for {n}: {tuple(t.shape)} vs {shape}")
            if k == "i64":
                self.int_buffers[n] = t.to(torch.int64).cpu().clone()
            else:
                o = self.offsets[n]
                host[o:o + numel(shape)] = t.detach().to(torch.float32).cpu().reshape(-1)
        self.params.copy_(host)
        return missing, unexpected
