"""``PlanningModel`` — host-side mirror of the reference policy class, running on librift_b200.so.

Same constructor arguments, ``forward(data) -> dict`` keys, externally-read attributes and
state-dict key names as rift/cbv/planning/pluto/model/pluto_model.py:22-225, so that the
fine-tuner plugins (rift/cbv/planning/fine_tuner/rlft/*) can hold one of these instead of the
torch module.  PyTorch is used for device memory and streams only; all arithmetic runs in the
CUDA library and the call fails loudly if that library is missing.

Numerical mode: the deterministic parity mode of SURVEY 8(c) (dropout / drop-path identity,
BatchNorm running statistics) — what the reference computes under ``model.eval()`` and what its
rollout path (``get_action``) uses.
"""
import ctypes as C
import os
from typing import Dict, Iterable, Optional

import torch

from . import _lib
from .arena import ParamArena
from .config import PlutoConfig


def _contig(t: torch.Tensor, dtype) -> torch.Tensor:
    if t.dtype == torch.bool and dtype == torch.uint8:
        t = t.contiguous().view(torch.uint8)
    elif t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class PackedBatch:
    """Raw-pointer view (rift_b200_batch) of a collated PlutoFeature.data dict on the GPU.
    Keeps the (possibly converted) tensors alive for as long as the struct is used."""

    def __init__(self, data: Dict, device):
        a, m, r = data["agent"], data["map"], data["reference_line"]
        f32, u8, i8 = torch.float32, torch.uint8, torch.int8
        dev = torch.device(device)

        def g(t, dt):
            if not torch.is_tensor(t):
                t = torch.as_tensor(t)
            return _contig(t.to(dev, non_blocking=True), dt)

        k = self.keep = {}
        k["agent_position"] = g(a["position"], f32)
        k["agent_heading"] = g(a["heading"], f32)
        k["agent_velocity"] = g(a["velocity"], f32)
        k["agent_shape"] = g(a["shape"], f32)
        k["agent_category"] = g(a["category"], i8)
        k["agent_valid_mask"] = g(a["valid_mask"], u8)
        k["map_point_position"] = g(m["point_position"], f32)
        k["map_point_vector"] = g(m["point_vector"], f32)
        k["map_point_orientation"] = g(m["point_orientation"], f32)
        k["map_polygon_center"] = g(m["polygon_center"], f32)
        k["map_polygon_type"] = g(m["polygon_type"], i8)
        k["map_polygon_on_route"] = g(m["polygon_on_route"], u8)
        k["map_polygon_tl_status"] = g(m["polygon_tl_status"], i8)
        k["map_polygon_has_speed_limit"] = g(m["polygon_has_speed_limit"], u8)
        k["map_polygon_speed_limit"] = g(m["polygon_speed_limit"], f32)
        k["map_valid_mask"] = g(m["valid_mask"], u8)
        k["ref_position"] = g(r["position"], f32)
        k["ref_vector"] = g(r["vector"], f32)
        k["ref_orientation"] = g(r["orientation"], f32)
        k["ref_valid_mask"] = g(r["valid_mask"], u8)
        k["current_state"] = g(data["current_state"], f32)
        so = data.get("static_objects")
        if so is not None and so["position"].shape[1] != 0:
            raise NotImplementedError("static objects: RIFT's feature builder emits none "
                                      "(pluto_feature_builder.py:247-257); N > 0 is not supported")
        bs, A, T = k["agent_heading"].shape
        Mp, P = k["map_valid_mask"].shape[1:]
        R, Pr = k["ref_valid_mask"].shape[1:]
        b = self.struct = _lib.Batch()
        b.bs, b.A, b.Mp, b.P, b.R, b.Pr, b.agent_T = bs, A, Mp, P, R, Pr, T
        for name, t in k.items():
            setattr(b, name, t.data_ptr())
        b.cs_stride = k["current_state"].shape[1]
        self.shape = (bs, A, Mp, P, R, Pr)

    def slice(self, lo: int, hi: int) -> "PackedBatch":
        """Samples [lo, hi) of this batch as a PackedBatch over VIEWS of the same device tensors (micro-batching).  The
        whole batch's reference-line mask travels along, so that the slice reproduces the reference's r2r key-padding
        indexing (r_pad[j % bs] over the WHOLE batch, planning_decoder.py:56-60) exactly."""
        sub = PackedBatch.__new__(PackedBatch)
        sub.keep = {n: t[lo:hi] for n, t in self.keep.items()}
        bs, A, Mp, P, R, Pr = self.shape
        b = sub.struct = _lib.Batch()
        b.bs, b.A, b.Mp, b.P, b.R, b.Pr, b.agent_T = hi - lo, A, Mp, P, R, Pr, self.struct.agent_T
        for name, t in sub.keep.items():
            setattr(b, name, t.data_ptr())
        b.cs_stride = self.struct.cs_stride
        b.ref_valid_mask_global = self.keep["ref_valid_mask"].data_ptr()
        b.bs_global, b.b_offset = bs, lo
        sub.shape = (hi - lo, A, Mp, P, R, Pr)
        sub.parent = self
        return sub


class PlanningModel:
    def __init__(self, radius, dim=128, state_channel=6, polygon_channel=6, history_channel=9, history_steps=21,
                 future_steps=80, encoder_depth=4, decoder_depth=4, drop_path=0.2, dropout=0.1, num_heads=4,
                 num_modes=12, use_ego_history=False, state_attn_encoder=True, state_dropout=0.75,
                 use_hidden_proj=True, cat_x=True, ref_free_traj=True, value_hidden=None,
                 trainable_layers: Iterable[str] = (), device="cuda"):
        if use_ego_history or not state_attn_encoder or not use_hidden_proj or not cat_x or not ref_free_traj:
            raise NotImplementedError("only the RIFT configuration of PlanningModel is implemented "
                                      "(state_attn_encoder, hidden_proj, cat_x, ref_free_traj; no ego history)")
        self.cfg = PlutoConfig(radius=float(radius), dim=dim, state_channel=state_channel,
                               polygon_channel=polygon_channel, history_channel=history_channel,
                               history_steps=history_steps, future_steps=future_steps, encoder_depth=encoder_depth,
                               decoder_depth=decoder_depth, num_heads=num_heads, num_modes=num_modes,
                               value_hidden=tuple(value_hidden) if value_hidden else None)
        # attributes read by the trainers (rift_trainer.py:61-71)
        self.radius, self.dim = radius, dim
        self.history_steps, self.future_steps, self.num_modes = history_steps, future_steps, num_modes
        self.drop_path, self.dropout, self.state_dropout = drop_path, dropout, state_dropout   # inert: parity mode
        self.device = torch.device(device)
        self.training = False
        # True: every GEMM on the exact-fp32 SIMT kernel (validation); False: tcgen05 split-bf16 path
        self.exact_fp32 = bool(int(os.environ.get("RIFT_B200_EXACT_FP32", "0")))
        self._engine = C.c_void_p()
        self._workspace = None
        self._ws_shape = None
        self._extra = []              # further engines over the same parameters (micro-batches in flight together)
        self.ws_generation = 0        # bumped whenever the workspace is re-allocated (captured CUDA graphs go stale)
        self.arena: Optional[ParamArena] = None
        self.set_trainable_layers(trainable_layers)

    @classmethod
    def from_config(cls, cfg: PlutoConfig, **kw):
        return cls(radius=cfg.radius, dim=cfg.dim, state_channel=cfg.state_channel, history_steps=cfg.history_steps,
                   future_steps=cfg.future_steps, encoder_depth=cfg.encoder_depth, decoder_depth=cfg.decoder_depth,
                   num_heads=cfg.num_heads, num_modes=cfg.num_modes, value_hidden=cfg.value_hidden, **kw)

    # ------------------------------------------------------------------ engine / arena
    def _new_engine(self, grads):
        L = _lib.lib()
        eng = C.c_void_p()
        c = self.cfg
        mc = _lib.ModelConfig(c.dim, c.num_heads, c.encoder_depth, c.decoder_depth, c.num_modes, c.history_steps,
                              c.future_steps, c.state_channel, c.ref_points,
                              c.value_hidden[0] if c.value_hidden else 0, c.value_hidden[1] if c.value_hidden else 0)
        ents = self.arena.entries()
        arr = (_lib.ParamEntry * len(ents))()
        self._names = [n.encode() for n, _, _, _ in ents]
        for i, (n, off, ne, tr) in enumerate(ents):
            arr[i].name, arr[i].offset, arr[i].numel, arr[i].trainable = self._names[i], off, ne, tr
        _lib.check(L.rift_b200_create(C.byref(mc), arr, len(ents), C.byref(eng)), "create")
        _lib.check(L.rift_b200_bind_arena(eng, _lib.ptr(self.arena.params), _lib.ptr(grads), self.arena.params.numel()), "bind_arena")
        return eng

    def set_trainable_layers(self, trainable_layers: Iterable[str]):
        """(Re)build the arena layout for a trainable set; parameter values are preserved."""
        old = self.arena.state_dict() if self.arena is not None else None
        self.trainable_layers = list(trainable_layers)
        self.arena = ParamArena(self.cfg, self.trainable_layers, self.device)
        if old is not None:
            self.arena.load_state_dict(old)
        L = _lib.lib()
        for e in [self._engine] + [x["engine"] for x in self._extra]:
            if e:
                L.rift_b200_destroy(e)
        self._extra = []
        self._engine = self._new_engine(self.arena.grads)
        # split-bf16 weight planes + TMA descriptors for the tcgen05 GEMM path
        self._wcache = torch.empty(L.rift_b200_weight_cache_bytes(self._engine), dtype=torch.uint8, device=self.device)
        _lib.check(L.rift_b200_bind_weight_cache(self._engine, _lib.ptr(self._wcache), self._wcache.numel()),
                   "bind_weight_cache")
        self._ws_shape = None
        self.ws_generation += 1       # engine, arenas and weight planes were rebuilt: anything captured against them is stale

    def engine_slot(self, k: int) -> dict:
        """Engine k > 0: same parameters and weight planes as engine 0, own gradient arena / workspace / tape, so that
        several micro-batches can be in flight on different streams (LightningTrainer micro-batching)."""
        while len(self._extra) < k:
            L = _lib.lib()
            grads = torch.zeros_like(self.arena.grads) if self.arena.grads is not None else None
            eng = self._new_engine(grads)
            torch.cuda.synchronize(self.device)      # binding re-writes the (identical) split-job tables inside the shared cache
            _lib.check(L.rift_b200_bind_weight_cache(eng, _lib.ptr(self._wcache), self._wcache.numel()), "bind_weight_cache")
            _lib.check(L.rift_b200_params_updated(eng, 2), "params_updated")      # engine 0 keeps the shared planes fresh
            self._extra.append({"engine": eng, "grads": grads, "workspace": None, "ws_shape": None, "last_batch": None})
        return self._extra[k - 1]

    def refresh_weights(self):
        """Re-split stale weight planes NOW on the current stream (before micro-batches fork onto their own streams)."""
        _lib.check(_lib.lib().rift_b200_refresh_weights(self._engine, _lib.stream_ptr()), "refresh_weights")

    def params_updated(self, trainable_only: bool = False):
        """Tell the engine the fp32 arena changed (optimizer step / checkpoint load): the bf16 weight
        planes of the tensor-core path are re-split by the next forward."""
        _lib.check(_lib.lib().rift_b200_params_updated(self._engine, int(trainable_only)), "params_updated")

    def __del__(self):
        try:
            for e in [self._engine] + [x["engine"] for x in self._extra]:
                if e:
                    _lib.lib().rift_b200_destroy(e)
        except Exception:
            pass

    # ------------------------------------------------------------------ nn.Module-like surface
    def state_dict(self):
        return self.arena.state_dict()

    def load_state_dict(self, sd, strict=True):
        if any(k.startswith("model.") for k in sd):        # training ckpt keys (pluto.py:135-137)
            sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        r = self.arena.load_state_dict(sd, strict)
        self.params_updated(False)
        # eagerly: a captured CUDA graph replays device work only and would keep reading stale frozen planes
        _lib.check(_lib.lib().rift_b200_refresh_weights(self._engine, _lib.stream_ptr()), "refresh_weights")
        return r

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        """Numerics stay in the deterministic parity mode either way (Dropout / DropPath / state dropout identity,
        BatchNorm running statistics, never updated); see LightningTrainer.train for the one-time warning."""
        self.training = mode
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("rift_b200.PlanningModel lives on a CUDA device")
        return self

    def parameters(self):
        return [self.arena.params]

    def __call__(self, data, **kw):
        return self.forward(data, **kw)

    # ------------------------------------------------------------------ forward / backward
    def _ensure_workspace(self, pb: PackedBatch, k: int = 0):
        """Workspace of engine k for this batch shape; returns (engine, workspace tensor)."""
        if k == 0:
            if self._ws_shape != pb.shape:
                need = _lib.lib().rift_b200_workspace_bytes(self._engine, C.byref(pb.struct))
                if need == 0:
                    raise RuntimeError("rift_b200 workspace sizing failed: " + _lib.lib().rift_b200_last_error().decode())
                if self._workspace is None or self._workspace.numel() < need:
                    self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
                    self.ws_generation += 1
                self._ws_shape = pb.shape
            return self._engine, self._workspace
        slot = self.engine_slot(k)
        if slot["ws_shape"] != pb.shape:
            need = _lib.lib().rift_b200_workspace_bytes(slot["engine"], C.byref(pb.struct))
            if need == 0:
                raise RuntimeError("rift_b200 workspace sizing failed: " + _lib.lib().rift_b200_last_error().decode())
            if slot["workspace"] is None or slot["workspace"].numel() < need:
                slot["workspace"] = torch.empty(need, dtype=torch.uint8, device=self.device)
                self.ws_generation += 1
            slot["ws_shape"] = pb.shape
        return slot["engine"], slot["workspace"]

    def pack(self, data) -> PackedBatch:
        return data if isinstance(data, PackedBatch) else PackedBatch(data, self.device)

    def forward(self, data, outputs=("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory",
                                     "candidate_trajectories"), save_for_backward: bool = False, engine: int = 0):
        """data: PlutoFeature.data dict (or a PackedBatch).  Returns the reference's output dict
        (pluto_model.py:219-225) restricted to `outputs` plus the derived entries."""
        pb = self.pack(data)
        eng, ws = self._ensure_workspace(pb, engine)
        bs, A, Mp, P, R, Pr = pb.shape
        c = self.cfg
        T, Mo, D = c.future_steps, c.num_modes, c.dim
        dev = self.device
        f32 = torch.float32
        res = {"probability": torch.empty((bs, R, Mo), dtype=f32, device=dev),
               "r_padding_mask": torch.empty((bs, R), dtype=torch.uint8, device=dev)}
        if "trajectory" in outputs or "candidate_trajectories" in outputs:
            res["trajectory"] = torch.empty((bs, R, Mo, T, 6), dtype=f32, device=dev)
        if "candidate_trajectories" in outputs:
            res["candidate_trajectories"] = torch.empty((bs, R, Mo, T, 3), dtype=f32, device=dev)
        if "prediction" in outputs:
            res["prediction"] = torch.empty((bs, max(A - 1, 0), T, 6), dtype=f32, device=dev)
        if "hidden" in outputs:
            res["hidden"] = torch.empty((bs, D), dtype=f32, device=dev)
        if "ref_free_trajectory" in outputs:
            res["ref_free_trajectory"] = torch.empty((bs, T, 4), dtype=f32, device=dev)
        o = _lib.Outputs()
        for name in ("probability", "trajectory", "prediction", "hidden", "ref_free_trajectory",
                     "candidate_trajectories", "r_padding_mask"):
            setattr(o, name, res[name].data_ptr() if name in res and res[name].numel() else None)
        flags = (_lib.FWD_SAVE_FOR_BACKWARD if save_for_backward else 0) | (_lib.GEMM_SIMT if self.exact_fp32 else 0)
        _lib.check(_lib.lib().rift_b200_forward(eng, C.byref(pb.struct), C.byref(o), _lib.ptr(ws), ws.numel(), flags,
                                                _lib.stream_ptr()), "forward")
        if engine == 0:
            self._last_batch = pb
        else:
            self._extra[engine - 1]["last_batch"] = pb
        res["r_padding_mask"] = res["r_padding_mask"].view(torch.bool)
        return res

    def backward(self, dlogits: torch.Tensor, engine: int = 0):
        """d(loss)/d(probability) -> gradient arena of that engine (trainable parameters only)."""
        assert dlogits.is_cuda and dlogits.dtype == torch.float32 and dlogits.is_contiguous()
        if engine == 0:
            pb, eng, ws = self._last_batch, self._engine, self._workspace
        else:
            slot = self._extra[engine - 1]
            pb, eng, ws = slot["last_batch"], slot["engine"], slot["workspace"]
        _lib.check(_lib.lib().rift_b200_backward(eng, C.byref(pb.struct), _lib.ptr(dlogits), _lib.ptr(ws), ws.numel(),
                                                 _lib.GEMM_SIMT if getattr(self, "exact_bwd", self.exact_fp32) else 0, _lib.stream_ptr()), "backward")

    def merge_micro_batch_grads(self, n: int):
        """grads(engine 0) += grads(engine k), k = 1 .. n - 1, over the trainable range (+ nothing else: the tail slots of
        the arena carry the all-reduce's statistics and are written by the trainer)."""
        L = _lib.lib()
        for k in range(1, n):
            _lib.check(L.rift_b200_op_add_inplace(_lib.ptr(self.arena.grads), _lib.ptr(self._extra[k - 1]["grads"]),
                                                  self.arena.n_train, _lib.stream_ptr()), "add_inplace")
