"""Operator-level Python surface of the RL kernels (PyTorch tensors in, PyTorch tensors out).

Each function names the reference code it replaces; all arithmetic happens in librift_b200.so.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def _req(t: torch.Tensor, dtype, name):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.dtype == torch.bool and dtype == torch.uint8:
        t = t.contiguous().view(torch.uint8)
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t.contiguous()


# ------------------------------------------------------------------ group-relative advantage
def group_advantage(returns: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(ret - mean) / (std + 1e-5) per group, float64, bit-identical to the numpy expression at
    rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:466-469.

    returns: (n_groups, G) float64 CUDA tensor, or a flat tensor with int64 `offsets` (n_groups+1)."""
    returns = _req(returns, torch.float64, "returns")
    out = torch.empty_like(returns)
    if offsets is None:
        if returns.dim() != 2:
            raise ValueError("returns must be (n_groups, G) when offsets is None")
        n, G = returns.shape
        op = None
    else:
        offsets = _req(offsets, torch.int64, "offsets")
        n, G, op = offsets.numel() - 1, 0, _lib.ptr(offsets)
    _lib.check(_lib.lib().rift_b200_group_advantage(_lib.ptr(returns), op, n, G, _lib.ptr(out), _lib.stream_ptr()),
               "group_advantage")
    return out


# ------------------------------------------------------------------ RIFT / GRPO objective
def group_objective(algo: str, probability, old_group_logits, group_advantage_, valid_mask, r_padding_mask=None,
                    ref_group_logits=None, clip=(0.8, 1.2), dual_clip=3.0, kl_weight=0.2,
                    need_grad=True, scale_by_count=True) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """RIFT (rift_trainer.py:140-182) / GRPO (grpo_trainer.py:140-194) objective.
    Returns (loss fp64 scalar tensor, d loss / d probability or None, stats = [loss, sum, count] fp64)."""
    z = _req(probability, torch.float32, "probability")
    bs, R, Mo = z.shape
    o = _req(old_group_logits, torch.float32, "old_group_logits")
    a = _req(group_advantage_, torch.float64, "group_advantage")
    v = _req(valid_mask, torch.uint8, "valid_mask")
    rp = _req(r_padding_mask, torch.uint8, "r_padding_mask") if r_padding_mask is not None else None
    f = _req(ref_group_logits, torch.float32, "ref_group_logits") if algo == "grpo" else None
    L = _lib.lib()
    scratch = torch.empty(L.rift_b200_objective_scratch_bytes(bs), dtype=torch.uint8, device=z.device)
    out3 = torch.empty(3, dtype=torch.float64, device=z.device)
    dz = torch.empty_like(z) if need_grad else None
    _lib.check(L.rift_b200_group_objective(
        {"rift": 0, "grpo": 1}[algo], _lib.ptr(z), _lib.ptr(o), _lib.ptr(f), _lib.ptr(a), _lib.ptr(v), _lib.ptr(rp),
        bs, R, Mo, clip[0], clip[1], dual_clip, kl_weight, _lib.ptr(scratch), _lib.ptr(out3), _lib.ptr(dz),
        int(scale_by_count), _lib.stream_ptr()), "group_objective")
    return out3[0], dz, out3


class _GroupObjectiveFn(torch.autograd.Function):
    """Lets the kernels stand in for get_rift_loss / get_grpo_loss inside the reference's own
    torch LightningTrainer (probability produced by a torch model)."""

    @staticmethod
    def forward(ctx, probability, algo, old, adv, valid, r_pad, ref, clip, dual_clip, kl_weight):
        loss, dz, _ = group_objective(algo, probability.detach(), old, adv, valid, r_pad, ref, clip, dual_clip, kl_weight)
        ctx.save_for_backward(dz)
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return (dz * g.to(dz.dtype),) + (None,) * 9


def rift_loss(probability, old_group_logits, group_advantage_, valid_mask, r_padding_mask=None):
    return _GroupObjectiveFn.apply(probability, "rift", old_group_logits, group_advantage_, valid_mask,
                                   r_padding_mask, None, (0.8, 1.2), 3.0, 0.2)


def grpo_loss(probability, old_group_logits, ref_group_logits, group_advantage_, valid_mask, r_padding_mask=None,
              clip=(0.8, 1.2), kl_weight=0.2):
    return _GroupObjectiveFn.apply(probability, "grpo", old_group_logits, group_advantage_, valid_mask,
                                   r_padding_mask, ref_group_logits, clip, 3.0, kl_weight)


# ------------------------------------------------------------------ PPO / REINFORCE objective
def action_objective(mode: str, probability, r_padding_mask, weight, action_mode=None, old_log_prob=None,
                     clip_epsilon=0.2, lambda_entropy=0.01, extra_loss=None, global_batch=None, need_grad=True):
    """PPO policy term (ppo_trainer.py:161-183; `extra_loss` = value loss) or REINFORCE
    (reinforce_trainer.py:120-170).  Returns (loss fp32 scalar tensor, dlogits or None, chosen index)."""
    z = _req(probability, torch.float32, "probability")
    bs, R, Mo = z.shape
    rp = _req(r_padding_mask, torch.uint8, "r_padding_mask")
    w = _req(weight, torch.float32, "weight")
    am = _req(action_mode, torch.int64, "action_mode") if action_mode is not None else None
    ol = _req(old_log_prob, torch.float32, "old_log_prob") if old_log_prob is not None else None
    ex = _req(extra_loss, torch.float32, "extra_loss") if extra_loss is not None else None
    part = torch.empty(bs, dtype=torch.float32, device=z.device)
    loss = torch.empty(1, dtype=torch.float32, device=z.device)
    dz = torch.empty_like(z) if need_grad else None
    chosen = torch.empty(bs, dtype=torch.int32, device=z.device)
    inv_n = 1.0 / float(global_batch or bs)
    _lib.check(_lib.lib().rift_b200_action_objective(
        {"ppo": 0, "reinforce": 1}[mode], _lib.ptr(z), _lib.ptr(rp), _lib.ptr(am), _lib.ptr(w), _lib.ptr(ol), bs, R, Mo,
        clip_epsilon, lambda_entropy, inv_n, _lib.ptr(ex), _lib.ptr(part), _lib.ptr(loss), _lib.ptr(dz),
        _lib.ptr(chosen), _lib.stream_ptr()), "action_objective")
    return loss[0], dz, chosen


def teacher_objective(probability, r_padding_mask, trajectory, teacher_infos, frame_rate=10, global_batch=None, weight=1.0,
                      loss_out=None, dlogits=None, need_grad=True):
    """SFT / RTR teacher cross-entropy (sft_trainer.py:123-215).  With `loss_out` / `dlogits` given the term is ADDED to them
    (RTR: 5 x PPO + teacher).  Returns (loss fp32 scalar tensor, dlogits or None, flat label index per sample)."""
    z = _req(probability, torch.float32, "probability")
    bs, R, Mo = z.shape
    rp = _req(r_padding_mask, torch.uint8, "r_padding_mask")
    tr = _req(trajectory, torch.float32, "trajectory")
    ti = _req(teacher_infos, torch.float32, "teacher_infos")
    T = tr.shape[3]
    acc = loss_out is not None
    loss = loss_out if acc else torch.empty(1, dtype=torch.float32, device=z.device)
    if dlogits is None and need_grad:
        if acc:
            raise ValueError("accumulating into loss_out needs the dlogits to accumulate into as well")
        dlogits = torch.empty_like(z)
    part = torch.empty(bs, dtype=torch.float32, device=z.device)
    label = torch.empty(bs, dtype=torch.int32, device=z.device)
    _lib.check(_lib.lib().rift_b200_teacher_objective(
        _lib.ptr(z), _lib.ptr(rp), _lib.ptr(tr), _lib.ptr(ti), bs, R, Mo, T, int(frame_rate), 1.0 / float(global_batch or bs),
        float(weight), _lib.ptr(part), _lib.ptr(loss), _lib.ptr(dlogits), int(acc), _lib.ptr(label), _lib.stream_ptr()),
        "teacher_objective")
    return loss.reshape(-1)[0], dlogits, label


def smooth_l1(value, target, global_batch=None, need_grad=True):
    """nn.SmoothL1Loss() (mean, beta 1) and d/d value (ppo_trainer.py value_criterion)."""
    v = _req(value, torch.float32, "value")
    t = _req(target, torch.float32, "target")
    loss = torch.empty(1, dtype=torch.float32, device=v.device)
    dv = torch.empty_like(v) if need_grad else None
    _lib.check(_lib.lib().rift_b200_smooth_l1(_lib.ptr(v), _lib.ptr(t), v.numel(), 1.0 / float(global_batch or v.numel()),
                                              _lib.ptr(loss), _lib.ptr(dv), _lib.stream_ptr()), "smooth_l1")
    return loss, dv


# ------------------------------------------------------------------ buffer passes
def gae(rewards, undones, values, next_values, unterminated, gamma=0.98, lambda_gae_adv=0.98, normalise=True):
    """get_advantages_GAE + reward_sum + normalisation (ppo_datamodule.py:22-37,160-163)."""
    args = [_req(x, torch.float32, n) for x, n in ((rewards, "rewards"), (undones, "undones"), (values, "values"),
                                                   (next_values, "next_values"), (unterminated, "unterminated"))]
    n = args[0].numel()
    adv = torch.empty_like(args[0])
    rsum = torch.empty_like(args[0])
    advn = torch.empty_like(args[0]) if normalise else None
    _lib.check(_lib.lib().rift_b200_gae(*[_lib.ptr(a) for a in args], n, gamma, lambda_gae_adv, _lib.ptr(adv),
                                        _lib.ptr(rsum), _lib.ptr(advn), _lib.stream_ptr()), "gae")
    return adv, rsum, advn


def discounted_return(rewards, dones, gamma=0.98):
    """compute_return (reinforce_datamodule.py:19-38)."""
    r = _req(rewards, torch.float32, "rewards")
    d = _req(dones, torch.float32, "dones")
    out = torch.empty_like(r)
    _lib.check(_lib.lib().rift_b200_discounted_return(_lib.ptr(r), _lib.ptr(d), r.numel(), gamma, _lib.ptr(out),
                                                      _lib.stream_ptr()), "discounted_return")
    return out


# ------------------------------------------------------------------ optimizer apply
class ClipAdamW:
    """clip_grad_norm_(max_norm) + torch.optim.AdamW semantics over one flat trainable range
    (custom_lightning.yaml:40-41, rift_trainer.py:333-351).  The first `n_decay` elements decay.

    The learning rate and the count of applied updates live on the device (`hyper`), so ``step`` is pure device
    work and can be captured into a CUDA graph; the scheduler-facing ``param_groups[0]["lr"]`` is pushed to the device
    whenever it changed.  A step whose valid count is 0 is skipped entirely (no decay, no moment update, no counter
    increment), like the reference, whose loss is a constant for such a batch."""

    def __init__(self, params: torch.Tensor, grads: torch.Tensor, n_train: int, n_decay: int, lr=1e-4,
                 weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8, max_norm=0.5):
        self.p, self.g, self.n, self.n_decay = params, grads, int(n_train), int(n_decay)
        self.lr, self.weight_decay, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.m = torch.zeros(max(self.n, 4), dtype=torch.float32, device=params.device)
        self.v = torch.zeros_like(self.m)
        self.scratch = torch.empty(_lib.lib().rift_b200_optim_scratch_bytes(), dtype=torch.uint8, device=params.device)
        self.scal = torch.zeros(8, dtype=torch.float32, device=params.device)   # [grad norm, applied scale, applied?, ...]
        self.hyper = torch.tensor([lr, 0.0], dtype=torch.float32, device=params.device)   # [lr, updates applied]
        self._lr_on_device = lr
        self.param_groups = [{"lr": lr}]        # scheduler-facing, like torch optimizers

    @property
    def step_count(self) -> int:
        """Number of updates applied so far (device counter; reading it synchronises)."""
        return int(self.hyper[1])

    def sync_lr(self):
        """Push a changed learning rate to the device (outside any graph capture)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.hyper[0:1].fill_(lr)
            self._lr_on_device = lr

    def step(self, count: Optional[torch.Tensor] = None, sync_lr: bool = True):
        """count: optional device fp64 scalar; gradients are divided by it first (global valid count)."""
        if sync_lr:
            self.sync_lr()
        _lib.check(_lib.lib().rift_b200_clip_adamw_dev(
            _lib.ptr(self.p), _lib.ptr(self.g), _lib.ptr(self.m), _lib.ptr(self.v), self.n, self.n_decay,
            _lib.ptr(count), self.max_norm, _lib.ptr(self.hyper), self.betas[0], self.betas[1], self.eps,
            self.weight_decay, _lib.ptr(self.scratch), _lib.ptr(self.scal), _lib.stream_ptr()), "clip_adamw_dev")

    def grad_norm(self) -> float:
        return float(self.scal[0])
