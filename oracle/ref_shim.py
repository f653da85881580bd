"""TEST INFRASTRUCTURE ONLY — never imported by the product path (rift_b200/).

Imports the *unmodified* reference modules from /root/reference so that the
restated oracle (oracle/pluto_oracle.py, oracle/loss_oracle.py) can be pinned
against them and golden vectors generated (oracle/make_golden.py).

/root/reference exists only in the build container, never on the GPU box, so
nothing under tests/ -m gpu, smoke() or bench.py may import this file.

Three third-party packages the reference needs are absent here; each is
replaced by a behavioural stand-in restated from its published semantics:

* natten==0.14.6 `NeighborhoodAttention1D` (reference call site
  rift/cbv/planning/pluto/model/layers/embedding.py:4,169-178) — 1-D
  neighbourhood attention, dilation 1: window of k keys starting at
  clamp(i - k//2, 0, L - k), logits + rpb[h, (j - i) + (k - 1)].
  NATTEN's source is not vendored -> "parity unpinned" at this boundary.
* timm==1.0.11 `DropPath` (embedding.py:5, transformer.py:6) — identity in
  eval mode, which is the only mode parity is defined in.
* lightning==2.2.5 `LightningModule` — only used as a base class.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("RIFT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "rift", "cbv", "planning"))


class _DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * r.div_(keep)


class _NeighborhoodAttention1D(nn.Module):
    def __init__(self, dim, num_heads, kernel_size, dilation=1, bias=True, qkv_bias=True,
                 qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.kernel_size = kernel_size
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.rpb = nn.Parameter(torch.zeros(num_heads, 2 * kernel_size - 1)) if bias else None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, L, C = x.shape
        k = self.kernel_size
        q, kk, v = self.qkv(x).reshape(B, L, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q = q * self.scale
        i = torch.arange(L, device=x.device)
        idx = (i - k // 2).clamp(0, L - k)[:, None] + torch.arange(k, device=x.device)[None]
        a = torch.einsum("bhld,bhlkd->bhlk", q, kk[:, :, idx])
        if self.rpb is not None:
            a = a + self.rpb[:, idx - i[:, None] + (k - 1)]
        a = self.attn_drop(a.softmax(-1))
        o = torch.einsum("bhlk,bhlkd->bhld", a, v[:, :, idx]).permute(0, 2, 1, 3).reshape(B, L, C)
        return self.proj_drop(self.proj(o))


class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


_installed = False


def install():
    """Register namespace packages + stand-ins. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    # bypass rift/cbv/planning/__init__.py (eagerly imports carla/hydra/lightning policies)
    for name, sub in [("rift", "rift"), ("rift.cbv", "rift/cbv"), ("rift.cbv.planning", "rift/cbv/planning")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = m
    for modname, attrs in [
        ("timm", {}),
        ("timm.layers", {"DropPath": _DropPath}),
        ("natten", {"NeighborhoodAttention1D": _NeighborhoodAttention1D}),
        ("lightning", {"LightningModule": _LightningModule}),
    ]:
        m = types.ModuleType(modname)
        m.__dict__.update(attrs)
        sys.modules[modname] = m
    sys.modules["timm"].layers = sys.modules["timm.layers"]
    _installed = True


def planning_model_cls():
    install()
    from rift.cbv.planning.pluto.model.pluto_model import PlanningModel
    return PlanningModel


def ppo_pluto_model_cls():
    """PPOPlutoModel restated (ppo_pluto.py:24-37): that file imports hydra/wandb/CARLA, so the
    14-line class is rebuilt here around the importable PlanningModel and CriticPPO, with the
    PlanningModel kwargs opened up so that a medium model can be built too."""
    PlanningModel = planning_model_cls()
    CriticPPO = critic_ppo_cls()

    class PPOPlutoModel(PlanningModel):
        def __init__(self, radius, state_dim, action_dim, hidden_dim, clip_epsilon=0.2,
                     lambda_entropy=0.01, **kw):
            super().__init__(radius=radius, **kw)
            self.clip_epsilon = clip_epsilon
            self.lambda_entropy = lambda_entropy
            self.value_net = CriticPPO(dims=hidden_dim, state_dim=state_dim, action_dim=action_dim)

    return PPOPlutoModel


def trainer_cls(algo: str):
    """algo in {'rift','grpo','ppo','reinforce'} -> the reference LightningTrainer class."""
    install()
    import importlib
    if algo == "ppo":
        name = "rift.cbv.planning.fine_tuner.rlft.ppo_pluto.ppo_pluto"
        if name not in sys.modules:
            pkg = types.ModuleType("rift.cbv.planning.fine_tuner.rlft.ppo_pluto")
            pkg.__path__ = [os.path.join(REF, "rift/cbv/planning/fine_tuner/rlft/ppo_pluto")]
            sys.modules.setdefault("rift.cbv.planning.fine_tuner.rlft.ppo_pluto", pkg)
            stub = types.ModuleType(name)
            stub.PPOPlutoModel = ppo_pluto_model_cls()
            sys.modules[name] = stub
    mod = importlib.import_module(
        f"rift.cbv.planning.fine_tuner.rlft.{algo}_pluto.{algo}_trainer")
    return mod.LightningTrainer


def sft_trainer_cls(kind: str = "sft"):
    """kind in {'sft', 'rtr', 'rs'} -> the reference LightningTrainer of fine_tuner/sft (rtr needs the PPO model stub)."""
    install()
    import importlib
    if kind == "sft":
        return importlib.import_module("rift.cbv.planning.fine_tuner.sft.sft_trainer").LightningTrainer
    trainer_cls("ppo")          # installs the PPOPlutoModel stub that rtr_trainer.py imports
    name = f"rift.cbv.planning.fine_tuner.sft.{kind}_pluto.{kind}_pluto"
    if name not in sys.modules:
        pkg = types.ModuleType(f"rift.cbv.planning.fine_tuner.sft.{kind}_pluto")
        pkg.__path__ = [os.path.join(REF, f"rift/cbv/planning/fine_tuner/sft/{kind}_pluto")]
        sys.modules.setdefault(f"rift.cbv.planning.fine_tuner.sft.{kind}_pluto", pkg)
        stub = types.ModuleType(name)
        stub.RTRPlutoModel = ppo_pluto_model_cls()
        stub.PPOPlutoModel = stub.RTRPlutoModel
        sys.modules[name] = stub
    return importlib.import_module(f"rift.cbv.planning.fine_tuner.sft.{kind}_pluto.{kind}_trainer").LightningTrainer


def pluto_feature_cls():
    install()
    from rift.cbv.planning.pluto.feature_builder.pluto_feature import PlutoFeature
    return PlutoFeature


def critic_ppo_cls():
    """CriticPPO (rift/gym_carla/utils/net.py:420-433). net.py imports only torch/numpy."""
    install()
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_net", os.path.join(REF, "rift/gym_carla/utils/net.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CriticPPO


def ref_class(relpath: str, name: str, extra_globals=None):
    """Execute ONE top-level class of a reference file without importing the file (its module-level imports may
    need hydra / lightning / carla).  Used for the datamodule collate callables and the rollout buffer."""
    import ast
    from typing import Dict, List, Optional
    import numpy as np
    src = open(os.path.join(REF, relpath)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            g = {"torch": torch, "np": np, "Dict": Dict, "List": List, "Optional": Optional,
                 "pad_sequence": torch.nn.utils.rnn.pad_sequence, "PlutoFeature": pluto_feature_cls()}
            g.update(extra_globals or {})
            exec(compile(ast.Module([node], []), relpath, "exec"), g)
            return g[name]
    raise KeyError(name)


def ref_method(relpath: str, cls: str, name: str, extra_globals=None):
    """One method of a reference class as a plain function (first argument = self), without importing the file."""
    import ast
    import numpy as np
    src = open(os.path.join(REF, relpath)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    g = {"torch": torch, "np": np, "npt": __import__("numpy.typing").typing}
                    g.update(extra_globals or {})
                    sub.returns = None
                    for a in sub.args.args:
                        a.annotation = None
                    exec(compile(ast.Module([sub], []), relpath, "exec"), g)
                    return g[name]
    raise KeyError(f"{cls}.{name}")


def pid_controller_cls():
    """PIDController (rift/cbv/planning/pluto/controller/pid_controller.py) - the file imports only numpy / torch."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_pid", os.path.join(REF, "rift/cbv/planning/pluto/controller/pid_controller.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PIDController


def track_propagate_module():
    """rift/cbv/planning/fine_tuner/rlft/traj_eval/track_propogate.py with its two CARLA-side imports stubbed
    (CarlaAgentState is only a type annotation there, get_device_name -> 'cpu')."""
    install()
    import importlib.util
    for name, attrs in (("rift.cbv.planning.pluto.utils.nuplan_state_utils", {"CarlaAgentState": object}),
                        ("rift.util", {}), ("rift.util.torch_util", {"get_device_name": lambda: "cpu"})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    spec = importlib.util.spec_from_file_location(
        "_ref_track_propagate", os.path.join(REF, "rift/cbv/planning/fine_tuner/rlft/traj_eval/track_propogate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dense_reward_model_cls():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_reward", os.path.join(REF, "rift/gym_carla/reward/reward_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DenseRewardModel
