"""``LightningTrainer`` mirrors — the operator-level drop-in boundary of the fine-tuner.

Constructor signature, ``training_step`` / ``validation_step`` / ``forward`` /
``configure_optimizers`` / ``state_dict`` (``model.``-prefixed keys) follow
rift/cbv/planning/fine_tuner/rlft/{rift,grpo,ppo,reinforce}_pluto/*_trainer.py.  Lightning itself
is not required: ``training_step`` computes the loss AND leaves d(loss)/d(params) in the model's
gradient arena (there is no autograd graph), ``optimizer_step`` performs Lightning's
clip_grad_norm_(0.5) -> AdamW.step (custom_lightning.yaml:40-41), ``on_train_epoch_end`` advances the
WarmupCos schedule (interval = epoch).  ``fit`` strings these together for callers without Lightning.

Data parallelism (not in the reference): when torch.distributed is initialised, each rank
holds a shard of the batch; ONE all-reduce carries the flat gradient arena plus (objective sum,
valid count), and every rank applies the identical update, reproducing the reference's global
masked mean exactly (rift_trainer.py:173-178).
"""
import math
import os
from typing import Dict, List, Optional

import torch

from . import functional as F
from .planning_model import PlanningModel


def allreduce_grads_and_stats(grads: torch.Tensor, n_train: int, stats: torch.Tensor) -> torch.Tensor:
    """The step's single collective: SUM all-reduce of [flat fp32 gradients | objective sum | valid count].

    The gradient arena has a 4-float tail after its n_train gradient elements; the two fp64 scalars travel
    in it as (hi, lo) float pairs (hi = fp32(x), lo = fp32(x - hi)), so there is no second collective and no
    staging copy.  The collective adds the hi parts of the ranks in fp32: valid counts (integers below 2^24) come
    out exact, the objective sum - and with it the LOGGED loss - carries fp32 rounding (~1e-7 relative); the
    parameter update only depends on the count.  Returns the all-reduced (sum, count) as fp64."""
    import torch.distributed as dist
    n = n_train
    comm = grads[:n + 4]
    s = stats.to(torch.float64)
    hi = s.to(torch.float32)
    lo = (s - hi.to(torch.float64)).to(torch.float32)
    comm[n:n + 2].copy_(hi)
    comm[n + 2:n + 4].copy_(lo)
    dist.all_reduce(comm, op=dist.ReduceOp.SUM)
    return comm[n:n + 2].to(torch.float64) + comm[n + 2:n + 4].to(torch.float64)


class WarmupCosLR:
    """pluto/optim/warmup_cos_lr.py:6-54 (per-epoch warm-up + cosine) without the `verbose`
    positional argument that breaks on torch >= 2.2."""

    def __init__(self, optimizer, lr, min_lr, epochs, warmup_epochs):
        self.optimizer, self.lr, self.min_lr, self.epochs, self.warmup_epochs = optimizer, lr, min_lr, epochs, warmup_epochs
        self.last_epoch = 0
        self._apply()

    def get_lr(self, epoch=None):
        e = self.last_epoch if epoch is None else epoch
        if e < self.warmup_epochs:
            return self.lr * (e + 1) / self.warmup_epochs
        return self.min_lr + 0.5 * (self.lr - self.min_lr) * (
            1 + math.cos(math.pi * (e - self.warmup_epochs) / (self.epochs - self.warmup_epochs)))

    def _apply(self):
        for g in self.optimizer.param_groups:
            g["lr"] = self.get_lr()

    def step(self):
        self.last_epoch += 1
        self._apply()


class LightningTrainer:
    ALGO = "rift"

    def __init__(self, model: PlanningModel, lr, cl_lr_decay, weight_decay, epochs, warmup_epochs, frame_rate: int,
                 trainable_layers: List[str], use_drivable_area_loss=True, use_regulate_yaw=True,
                 objective_aggregate_mode: str = "mean", clip=(0.8, 1.2), kl_weight=0.2, gradient_clip_val=0.5):
        self.model = model
        self.lr, self.cl_lr_decay, self.weight_decay = lr, cl_lr_decay, weight_decay
        self.epochs, self.warmup_epochs = epochs, warmup_epochs
        self.objective_aggregate_mode = objective_aggregate_mode
        self.history_steps, self.future_steps = model.history_steps, model.future_steps
        self.frame_rate = int(frame_rate)
        self.trainable_layers = list(trainable_layers)
        self.use_drivable_area_loss, self.use_regulate_yaw = use_drivable_area_loss, use_regulate_yaw
        self.radius, self.num_modes = model.radius, model.num_modes
        self.mode_interval = self.radius / self.num_modes
        self.clip, self.kl_weight, self.gradient_clip_val = clip, kl_weight, gradient_clip_val
        self.training = True
        self.logged: Dict[str, float] = {}
        self.optimizer: Optional[F.ClipAdamW] = None
        self.scheduler: Optional[WarmupCosLR] = None
        self._graphs: Dict[tuple, dict] = {}
        self._graph_seen: Dict[tuple, int] = {}
        self.freeze_parameters(self.trainable_layers)
        self._count = None            # device fp64 valid count of the last training_step (group objectives)
        # CUDA-graph replay of forward + objective + backward (launch-bound: ~1.5k kernels per full update).
        # A batch signature (shapes / dtypes) is captured the second time it is seen; inputs are copied into the
        # graph's static buffers, so varying shapes simply stay on the eager path.  RIFT_B200_CUDA_GRAPH=0 disables.
        self.use_cuda_graph = bool(int(os.environ.get("RIFT_B200_CUDA_GRAPH", "1")))
        # concurrent micro-batches per training step (see _train_core); RIFT_B200_MICRO_BATCHES=1 runs the batch whole
        self.micro_batches = int(os.environ.get("RIFT_B200_MICRO_BATCHES", "1"))
        self._mb_streams: List[torch.cuda.Stream] = []

    # ------------------------------------------------------------------ reference surface
    def freeze_parameters(self, trainable_layers=("planning_decoder.pi_head",)):
        """rift_trainer.py:78-90 — raises ValueError for an unknown layer name.

        Re-lays out the parameter / gradient arenas, so an optimizer built for the previous trainable set (its
        moments are sized and ordered for the old layout) is dropped: ``optimizer_step`` / ``configure_optimizers``
        build a fresh one, exactly like the reference, where a new trainer always gets a new AdamW."""
        self.model.set_trainable_layers(trainable_layers)
        self.optimizer = None
        self.scheduler = None
        self._invalidate_graphs()

    def _invalidate_graphs(self):
        self._graphs = {}
        self._graph_seen = {}

    def close(self):
        """Drop the captured step graphs.  Call before ``torch.distributed.destroy_process_group()`` when world_size > 1: NCCL
        does not finish destroying a communicator while a CUDA graph that captured its all-reduce is alive."""
        self._invalidate_graphs()
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def forward(self, features):
        return self.model(features)

    def train(self, mode=True):
        """Lightning's ``fit`` puts the module in train mode (dropout 0.1, drop-path 0.2, state-token dropout 0.75,
        BatchNorm batch statistics).  This implementation always computes the deterministic parity mode
        (DESIGN.md section 2): ``train(True)`` only selects training_step semantics and says so once."""
        if mode and not getattr(LightningTrainer, "_warned_train_mode", False):
            import warnings
            warnings.warn("rift_b200 trains in the deterministic parity mode: Dropout / DropPath / state-token dropout are "
                          "identity and BatchNorm uses (and does not update) its running statistics; the reference under "
                          "Lightning's model.train() draws dropout masks and updates the BatchNorm buffers "
                          "(DESIGN.md section 2)", stacklevel=2)
            LightningTrainer._warned_train_mode = True
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def log(self, name, value, **kw):
        self.logged[name] = float(value)

    def configure_optimizers(self):
        a = self.model.arena
        self.optimizer = F.ClipAdamW(a.params, a.grads, a.n_train, a.n_decay, lr=self.lr, weight_decay=self.weight_decay,
                                     max_norm=self.gradient_clip_val)
        self.scheduler = WarmupCosLR(self.optimizer, lr=self.lr, min_lr=self.lr * self.cl_lr_decay, epochs=self.epochs,
                                     warmup_epochs=self.warmup_epochs)
        return [self.optimizer], [self.scheduler]

    def state_dict(self):
        return {"model." + k: v for k, v in self.model.state_dict().items()}

    def load_state_dict(self, sd, strict=True):
        """Captured graphs read the fp32 arena in place but the FROZEN weights through their pre-split bf16 planes,
        which a replay never re-splits: PlanningModel.load_state_dict refreshes every plane eagerly, and the graphs
        are dropped anyway so that nothing captured before the load survives it."""
        r = self.model.load_state_dict(sd, strict)
        self._invalidate_graphs()
        return r

    # ------------------------------------------------------------------ objectives
    @staticmethod
    def _features(batch):
        f = batch["cur_pluto_feature_torch"]
        return f.data if hasattr(f, "data") and isinstance(getattr(f, "data"), dict) else f

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _objective(self, res, batch, need_grad):
        """Returns (loss tensor, dlogits or None).  dlogits is NOT divided by the valid count when
        training (the optimizer kernel divides by the all-reduced count)."""
        dev = self.model.device
        loss, dz, stats = F.group_objective(
            self.ALGO, res["probability"], batch["old_group_logits_torch"].to(dev),
            batch["group_advantage_torch"].to(dev), batch["group_advantage_mask_torch"].to(dev),
            res["r_padding_mask"], batch["ref_group_logits_torch"].to(dev) if self.ALGO == "grpo" else None,
            clip=self.clip, kl_weight=self.kl_weight, need_grad=need_grad, scale_by_count=False)
        self._stats = stats
        return loss, dz

    FWD_OUTPUTS = ()       # model outputs the objective reads besides the logits (SFT / RTR: the candidate trajectories)
    MICRO_OK = True        # the objective is a masked sum / count, so micro-batch (sum, count) pairs simply add up

    def _micro_batches(self, bs: int) -> int:
        k = self.micro_batches
        while k > 1 and (bs % k or bs // k < 8):
            k -= 1
        return k if self.MICRO_OK else 1

    def _train_core(self, batch):
        """forward -> objective -> backward; returns the local loss tensor.

        Micro-batching (`micro_batches` > 1): the batch is cut into K contiguous slices that run forward, objective and
        backward CONCURRENTLY on K streams, each on its own engine (own workspace, tape and gradient arena over the same
        parameters and weight planes).  One policy update is a dependent chain of ~600 small kernels that leaves most
        of the 148 SMs idle; K independent chains of half-sized kernels fill them.  The slices' (objective sum, valid
        count) and gradient arenas are added afterwards, which is exactly the data-parallel reduction of SURVEY 8(e)
        inside one GPU; the reference's r2r key-padding indexing over the whole batch is preserved (PackedBatch.slice)."""
        feats = self._features(batch)
        pb = self.model.pack(feats)
        K = self._micro_batches(pb.shape[0])
        if K <= 1:
            res = self.model.forward(pb, outputs=self.FWD_OUTPUTS, save_for_backward=True)
            loss, dz = self._objective(res, batch, need_grad=True)
            self.model.backward(dz)
            return loss
        dev = self.model.device
        cur = torch.cuda.current_stream(dev)
        if len(self._mb_streams) < K:
            self._mb_streams = [torch.cuda.Stream(device=dev) for _ in range(K)]
        self.model.refresh_weights()                     # before the fork: every engine reads the same weight planes
        bs = pb.shape[0]
        per = bs // K
        stats = []
        for k in range(K):
            lo, hi = k * per, (k + 1) * per
            sub = {n: (t[lo:hi] if torch.is_tensor(t) and t.dim() > 0 and t.shape[0] == bs else t) for n, t in batch.items()}
            st = self._mb_streams[k]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                part = pb.slice(lo, hi)
                res = self.model.forward(part, outputs=(), save_for_backward=True, engine=k)
                _, dz = self._objective(res, sub, need_grad=True)
                self.model.backward(dz, engine=k)
                stats.append(self._stats)
        for k in range(K):
            cur.wait_stream(self._mb_streams[k])
        for s_ in stats:
            s_.record_stream(cur)
        self.model.merge_micro_batch_grads(K)
        tot = stats[0].clone()
        for s_ in stats[1:]:
            tot[1:3] += s_[1:3]
        tot[0] = torch.where(tot[2] > 0, -tot[1] / tot[2], torch.zeros_like(tot[1]))
        self._stats = tot
        return tot[0]

    def _step(self, batch, prefix: str):
        if self.training:
            loss = self._graphed_core(batch) if self.use_cuda_graph else None
            if loss is None:
                loss = self._train_core(batch)
            loss = self._reduce(loss)
        else:
            res = self.model.forward(self._features(batch), outputs=self.FWD_OUTPUTS, save_for_backward=False)
            loss, _ = self._objective(res, batch, need_grad=False)
        self._last_loss = loss
        return loss if self.training else 0.0

    # ------------------------------------------------------------------ CUDA-graph replay of the training core
    _FEATURE_KEYS = (("agent", "agent_", ("position", "heading", "velocity", "shape", "category", "valid_mask")),
                     ("map", "map_", ("point_position", "point_vector", "point_orientation", "polygon_center",
                                      "polygon_type", "polygon_on_route", "polygon_tl_status",
                                      "polygon_has_speed_limit", "polygon_speed_limit", "valid_mask")),
                     ("reference_line", "ref_", ("position", "vector", "orientation", "valid_mask")))

    @classmethod
    def _flatten_batch(cls, batch, feats):
        """(name, tensor) pairs of everything the training core reads from a batch dict; feature names are
        PackedBatch.keep's."""
        from .planning_model import PackedBatch
        if isinstance(feats, PackedBatch):
            items = [("f." + k, v) for k, v in feats.keep.items()]
        else:
            items = [("f." + prefix + k, torch.as_tensor(feats[grp][k])) for grp, prefix, keys in cls._FEATURE_KEYS for k in keys]
            items.append(("f.current_state", torch.as_tensor(feats["current_state"])))
        items += [("b." + k, v) for k, v in batch.items() if torch.is_tensor(v)]
        return items

    def _graphed_core(self, batch, full_step: bool = False):
        """Replay (capturing on second sight of a batch signature) forward + objective + backward - and, with
        `full_step`, the gradient all-reduce, clip and AdamW as well (learning rate and step counter are device
        scalars, optim_kernels.cu) - from one CUDA graph.  Returns None when this call has to run eagerly."""
        from .planning_model import PackedBatch
        feats = self._features(batch)
        if not isinstance(feats, PackedBatch):
            so = feats.get("static_objects")
            if so is not None and so["position"].shape[1] != 0:
                return None                              # the eager path raises the documented NotImplementedError
        items = self._flatten_batch(batch, feats)
        # a PackedBatch is used in place as the graph's static input, so its identity is part of the signature (another
        # PackedBatch object of the same shape gets its own graph instead of being copied over the first one)
        key = (bool(full_step), id(feats) if isinstance(feats, PackedBatch) else 0,) + \
            tuple((n, tuple(t.shape), t.dtype) for n, t in items)
        g = self._graphs.get(key)
        if g is not None and g["ws_gen"] != self.model.ws_generation:
            del self._graphs[key]                        # the workspace moved: the captured pointers are stale
            g = None
        if g is None:
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if seen == 0 or len(self._graphs) >= 8:
                return None                              # first sight of this signature (or cache full): eager
            g = self._capture(batch, feats, items, key, full_step)
        # refresh the static inputs (skipped for tensors that already ARE the static ones: a device-resident PackedBatch)
        for (_, t), st in zip(items, g["static"]):
            if t.data_ptr() != st.data_ptr():
                st.copy_(t, non_blocking=True)
        if full_step:
            self.optimizer.sync_lr()
        g["graph"].replay()
        self._stats = g["stats"]
        if full_step:
            self._count = g["count"]
            self.model.params_updated(trainable_only=True)   # host flag: the next eager forward re-splits too
        return g["loss"].clone()                         # the graph's own output buffer is overwritten by the next replay

    def _capture(self, batch, feats, items, key, full_step=False):
        from .planning_model import PackedBatch
        dev = self.model.device
        if isinstance(feats, PackedBatch):
            spb = feats                                  # explicit device-resident batch: used in place
        else:
            data = {grp: {} for grp, _, _ in self._FEATURE_KEYS}
            for grp, prefix, keys in self._FEATURE_KEYS:
                for k in keys:
                    data[grp][k] = torch.as_tensor(feats[grp][k]).to(dev, copy=True)   # private static copies
            data["current_state"] = torch.as_tensor(feats["current_state"]).to(dev, copy=True)
            spb = PackedBatch(data, dev)
        sbatch = dict(batch)
        sbatch["cur_pluto_feature_torch"] = spb
        static = []
        for n, t in items:
            if n.startswith("f."):
                static.append(spb.keep[n[2:]])
            else:
                st = t.to(dev, copy=True)
                sbatch[n[2:]] = st
                static.append(st)
        if full_step and self.optimizer is None:
            self.configure_optimizers()
        self.model.params_updated(trainable_only=True)   # the re-split of the trainable weight planes is part of the graph
        graph = torch.cuda.CUDAGraph()
        # RIFT_B200_STREAM_PRIO=1 (experiment, measured neutral): the capture stream - the step's main chain - gets a high
        # priority and the engine's side stream (parameter gradients) the lowest, see engine.cu::attach_streams
        prio = -1 if os.environ.get("RIFT_B200_STREAM_PRIO", "0") == "1" else 0
        cap = torch.cuda.Stream(device=dev, priority=prio)
        cap.wait_stream(torch.cuda.current_stream(dev))
        count = None
        # thread_local: NCCL's watchdog thread keeps querying events while the data-parallel all-reduce is captured
        with torch.cuda.stream(cap):
            with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):
                loss = self._train_core(sbatch)
                stats = self._stats
                if full_step:
                    loss = self._reduce(loss)
                    count = self._count
                    self.optimizer.step(count=count, sync_lr=False)
        torch.cuda.current_stream(dev).wait_stream(cap)
        g = {"graph": graph, "static": static, "loss": loss, "stats": stats, "batch": sbatch, "count": count,
             "ws_gen": self.model.ws_generation}
        self._graphs[key] = g
        return g

    def _reduce(self, loss):
        """Single collective of the step: [flat grads | objective sum | valid count]."""
        stats = self._stats
        if self._world() > 1:
            import torch.distributed as dist
            a = self.model.arena
            s = allreduce_grads_and_stats(a.grads, a.n_train, stats[1:3])
            self._count = s[1:2].clone()
            return torch.where(s[1] > 0, -s[0] / s[1], torch.zeros_like(s[0]))
        self._count = stats[2:3].clone()
        return loss

    def training_step(self, batch, batch_idx=0):
        self.training = True
        return self._step(batch, "train")

    def validation_step(self, batch, batch_idx=0):
        self.training = False
        try:
            return self._step(batch, "val")
        finally:
            self.training = True

    def validation_loss(self, batch):
        """The value Lightning logs as 'loss/val_loss' (monitored by ModelCheckpoint, training_builder.py:131-140)."""
        self.validation_step(batch)
        return self._last_loss

    def optimizer_step(self):
        if self.optimizer is None:
            self.configure_optimizers()
        self.optimizer.step(count=self._count)
        self.model.params_updated(trainable_only=True)

    def on_train_epoch_end(self):
        if self.scheduler is not None:
            self.scheduler.step()

    # ------------------------------------------------------------------ loop for callers without Lightning
    def step(self, batch):
        """One whole policy update.  With CUDA graphs on, the second and later sights of a batch signature replay
        forward + objective + backward + all-reduce + clip + AdamW from ONE graph."""
        if self.use_cuda_graph:
            self.training = True
            loss = self._graphed_core(batch, full_step=True)
            if loss is not None:
                self._last_loss = loss
                return loss
        loss = self.training_step(batch)
        self.optimizer_step()
        return loss

    def fit(self, batches, epochs: Optional[int] = None):
        """`batches`: a re-iterable of collated batch dicts (one pass = one epoch)."""
        if self.optimizer is None:
            self.configure_optimizers()
        last = None
        for _ in range(epochs if epochs is not None else self.epochs):
            for b in batches:
                last = self.step(b)
            self.on_train_epoch_end()
        return last


class RIFTTrainer(LightningTrainer):
    """rift_pluto/rift_trainer.py — clip [0.8, 1.2] + dual clip 3.0."""
    ALGO = "rift"


class GRPOTrainer(LightningTrainer):
    """grpo_pluto/grpo_trainer.py — clip [0.8, 1.2] - 0.2 * KL(ref || new)."""
    ALGO = "grpo"


class ReinforceTrainer(LightningTrainer):
    """reinforce_pluto/reinforce_trainer.py — -mean(log pi(argmax) * return)."""
    ALGO = "reinforce"
    MICRO_OK = False       # per-sample objectives carry 1 / (global batch): run the batch whole

    def _objective(self, res, batch, need_grad):
        dev = self.model.device
        bs = res["probability"].shape[0]
        loss, dz, _ = F.action_objective("reinforce", res["probability"], res["r_padding_mask"],
                                         batch["return_torch"].to(dev).float(), global_batch=bs * self._world(),
                                         need_grad=need_grad)
        self._stats = None
        return loss, dz

    def _reduce(self, loss):
        self._count = None
        if self._world() > 1:
            import torch.distributed as dist
            a = self.model.arena
            n = a.n_train
            comm = a.grads[:n + 1]
            comm[n:n + 1].copy_(loss.reshape(1))       # per-rank term already carries 1 / global batch
            dist.all_reduce(comm, op=dist.ReduceOp.SUM)
            return comm[n].clone()
        return loss


class PPOPlutoModel(PlanningModel):
    """ppo_pluto/ppo_pluto.py:24-37 — PlanningModel + CriticPPO value net + PPO hyper-parameters."""

    def __init__(self, radius, state_dim=None, action_dim=1, hidden_dim=(256, 256), clip_epsilon=0.2, lambda_entropy=0.01,
                 **kw):
        super().__init__(radius=radius, value_hidden=tuple(hidden_dim), **kw)
        if state_dim is not None and state_dim != self.dim:
            raise ValueError("state_dim must equal the policy's hidden size (the value net reads `hidden`)")
        self.clip_epsilon, self.lambda_entropy = clip_epsilon, lambda_entropy
        self.init_value_net()

    def init_value_net(self, seed: int = 0):
        """CriticPPO's own initialisation (net.py:355-372,420-433): identity normalisation constants, torch's
        default Linear init for the hidden layers, orthogonal(std=0.5) / bias 1e-6 for the output layer."""
        gen = torch.Generator().manual_seed(seed)
        sd = {}
        for n, (shape, kind) in self.arena.spec.items():
            if not n.startswith("value_net."):
                continue
            leaf = n.split(".", 1)[1]
            if leaf in ("state_std", "value_std"):
                sd[n] = torch.ones(shape)
            elif leaf in ("state_avg", "value_avg"):
                sd[n] = torch.zeros(shape)
            elif leaf.endswith(".weight"):
                bound = 1.0 / (shape[1] ** 0.5)
                sd[n] = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            else:
                fan_in = self.arena.spec[n.replace(".bias", ".weight")][0][1]
                sd[n] = (torch.rand(shape, generator=gen) * 2 - 1) / (fan_in ** 0.5)
        last = max(int(n.split(".")[2]) for n in sd if n.startswith("value_net.net."))
        w = torch.empty(self.arena.spec[f"value_net.net.{last}.weight"][0])
        torch.nn.init.orthogonal_(w, 0.5)
        sd[f"value_net.net.{last}.weight"] = w
        sd[f"value_net.net.{last}.bias"] = torch.full(self.arena.spec[f"value_net.net.{last}.bias"][0], 1e-6)
        self.load_state_dict(sd, strict=False)


class PPOTrainer(ReinforceTrainer):
    """ppo_pluto/ppo_trainer.py:126-183 — SmoothL1 value loss - (clipped surrogate + lambda * entropy)."""
    ALGO = "ppo"

    def __init__(self, model, *a, **kw):
        super().__init__(model, *a, **kw)
        self.clip_epsilon = getattr(model, "clip_epsilon", 0.2)
        self.lambda_entropy = getattr(model, "lambda_entropy", 0.01)

    def freeze_parameters(self, trainable_layers=("planning_decoder.pi_head", "value_net")):
        super().freeze_parameters(trainable_layers)
        from .value_net import ValueNet
        self.value_net = ValueNet(self.model.arena)

    def _objective(self, res, batch, need_grad):
        dev = self.model.device
        bs = res["probability"].shape[0]
        gb = bs * self._world()
        value = self.value_net.forward(batch["state_torch"].to(dev).float().contiguous(), save=need_grad)
        vloss, self._dvalue = F.smooth_l1(value, batch["reward_sum_torch"].to(dev).float(), global_batch=gb, need_grad=need_grad)
        loss, dz, _ = F.action_objective("ppo", res["probability"], res["r_padding_mask"], batch["advantage_torch"].to(dev).float(),
                                         action_mode=batch["action_mode_torch"].to(dev),
                                         old_log_prob=batch["old_log_prob_torch"].to(dev).float(),
                                         clip_epsilon=self.clip_epsilon, lambda_entropy=self.lambda_entropy,
                                         extra_loss=vloss, global_batch=gb, need_grad=need_grad)
        self._stats = None
        return loss, dz

    def _train_core(self, batch):
        res = self.model.forward(self._features(batch), outputs=self.FWD_OUTPUTS, save_for_backward=True)
        loss, dz = self._objective(res, batch, need_grad=True)
        self.model.backward(dz)                          # zeroes the gradient span, then the policy gradients
        self.value_net.backward(self._dvalue)            # += value-net gradients
        return loss


class SFTTrainer(ReinforceTrainer):
    """fine_tuner/sft/sft_trainer.py:123-215 — supervised fine-tuning on the teacher's target speed: cross-entropy of the
    candidate logits against (model's best reference line, mode whose PID target speed is closest to the teacher's)."""
    ALGO = "sft"
    FWD_OUTPUTS = ("trajectory",)

    def _objective(self, res, batch, need_grad):
        dev = self.model.device
        bs = res["probability"].shape[0]
        loss, dz, self._label = F.teacher_objective(res["probability"], res["r_padding_mask"], res["trajectory"],
                                                    batch["teacher_infos"].to(dev).float(), frame_rate=self.frame_rate,
                                                    global_batch=bs * self._world(), need_grad=need_grad)
        self._stats = None
        return loss, dz


class RSTrainer(ReinforceTrainer):
    """fine_tuner/sft/rs_pluto/rs_trainer.py:154-170 — rejection-sampling baseline: the REINFORCE objective
    -mean(log pi(argmax) * return) on the buffer its datamodule filtered."""
    ALGO = "rs"


class RTRTrainer(PPOTrainer):
    """fine_tuner/sft/rtr_pluto/rtr_trainer.py:130-195 — RTR baseline: 5 x (PPO value + actor loss) + the SFT teacher loss."""
    ALGO = "rtr"
    FWD_OUTPUTS = ("trajectory",)
    LAMBDA_RL = 5.0

    def _objective(self, res, batch, need_grad):
        dev = self.model.device
        bs = res["probability"].shape[0]
        gb = bs * self._world()
        gb_rl = gb / self.LAMBDA_RL                      # 1 / gb_rl = lambda_rl / gb: scales the PPO loss AND its gradients by lambda_rl
        value = self.value_net.forward(batch["state_torch"].to(dev).float().contiguous(), save=need_grad)
        vloss, self._dvalue = F.smooth_l1(value, batch["reward_sum_torch"].to(dev).float(), global_batch=gb_rl, need_grad=need_grad)
        loss, dz, _ = F.action_objective("ppo", res["probability"], res["r_padding_mask"], batch["advantage_torch"].to(dev).float(),
                                         action_mode=batch["action_mode_torch"].to(dev),
                                         old_log_prob=batch["old_log_prob_torch"].to(dev).float(),
                                         clip_epsilon=self.clip_epsilon, lambda_entropy=self.lambda_entropy,
                                         extra_loss=vloss, global_batch=gb_rl, need_grad=need_grad)
        total = loss.reshape(1).clone()
        loss, dz, self._label = F.teacher_objective(res["probability"], res["r_padding_mask"], res["trajectory"],
                                                    batch["teacher_infos"].to(dev).float(), frame_rate=self.frame_rate,
                                                    global_batch=gb, loss_out=total, dlogits=dz, need_grad=need_grad)
        self._stats = None
        return loss, dz


TRAINERS = {"rift": RIFTTrainer, "grpo": GRPOTrainer, "reinforce": ReinforceTrainer, "ppo": PPOTrainer,
            "sft": SFTTrainer, "rtr": RTRTrainer, "rs": RSTrainer}
