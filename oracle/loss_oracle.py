"""TEST INFRASTRUCTURE ONLY — CPU restatement of the RL objectives, advantage and optimizer step.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  Each function cites the reference lines it follows.  Pinned against the
reference's own functions through tests/golden/*.npz (written by oracle/make_golden.py, which imports the unmodified reference)
and tests/test_oracle_golden.py.
"""
import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------- group-relative advantage
def group_advantage(returns: np.ndarray) -> np.ndarray:
    """traj_eval/traj_evaluator.py:466-469 — one group, float64, population std."""
    mean_return = np.mean(returns)
    std_return = np.std(returns) + 1e-5
    return (returns - mean_return) / std_return


def pairwise_sum(x: np.ndarray) -> float:
    """Explicit restatement of numpy's float64 pairwise summation for a contiguous 1-D array
    (numpy/core/src/umath/loops_utils.h.src `pairwise_sum`; numpy 2.3.5 here): n<8 sequential
    from 0.0 (numpy starts at -0.0: identical unless every term is -0.0); n<=128: eight
    accumulators over strides of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail
    added sequentially; n>128: split at n/2 rounded down to a multiple of 8 and recurse."""
    n = x.shape[0]
    if n < 8:
        s = np.float64(0.0)
        for v in x:
            s = s + v
        return s
    if n <= 128:
        r = [np.float64(x[i]) for i in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + x[i + j]
            i += 8
        s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            s = s + x[i]
            i += 1
        return s
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(x[:n2]) + pairwise_sum(x[n2:])


def group_advantage_explicit(returns: np.ndarray) -> np.ndarray:
    """group_advantage() with numpy's reductions spelled out: mean = pairwise_sum/n,
    var = pairwise_sum((x-mean)^2)/n, std = sqrt(var).  Bit-identical to numpy (tested)."""
    x = np.ascontiguousarray(returns, dtype=np.float64)
    n = x.shape[0]
    mean = pairwise_sum(x) / np.float64(n)
    d = x - mean
    var = pairwise_sum(d * d) / np.float64(n)
    std = np.sqrt(var) + np.float64(1e-5)
    return d / std


# ---------------------------------------------------------------- objectives (torch, autograd-able)
def _masked_lsm(logits, r_pad):
    bs = logits.shape[0]
    z = logits.masked_fill(r_pad.unsqueeze(-1), -1e8)
    return F.log_softmax(z.reshape(bs, -1), dim=1)


def rift_loss(probability, old_logits, advantage, valid_mask, r_pad):
    """rift_pluto/rift_trainer.py:140-182 (clip [0.8,1.2], dual clip 3.0, global masked mean)."""
    bs = probability.shape[0]
    lp = _masked_lsm(probability, r_pad)
    lo = _masked_lsm(old_logits, r_pad)
    adv = advantage.reshape(bs, -1)
    ratio = torch.exp(lp - lo)
    mn = torch.min(adv * ratio, adv * torch.clamp(ratio, 0.8, 1.2))
    mx = torch.max(mn, adv * 3.0)
    obj = torch.where(adv < 0, mx, mn)
    v = obj[valid_mask.reshape(bs, -1)]
    return -(v.mean() if v.numel() else torch.tensor(0.0))


def grpo_loss(probability, old_logits, ref_logits, advantage, valid_mask, r_pad,
              clip_lo: float = 0.8, clip_hi: float = 1.2, kl_weight: float = 0.2):
    """grpo_pluto/grpo_trainer.py:140-194.  clip bounds / KL weight are hard-coded there
    (0.8, 1.2, 0.2); they are parameters here for BASELINE cfg4 (PPO-clip eps + KL-to-ref)."""
    bs = probability.shape[0]
    lp = _masked_lsm(probability, r_pad)
    lo = _masked_lsm(old_logits, r_pad)
    pf = F.softmax(ref_logits.masked_fill(r_pad.unsqueeze(-1), -1e8).reshape(bs, -1), dim=1)
    kl = F.kl_div(input=lp, target=pf, reduction="none", log_target=False)
    adv = advantage.reshape(bs, -1)
    ratio = torch.exp(lp - lo)
    obj = torch.min(adv * ratio, adv * torch.clamp(ratio, clip_lo, clip_hi)) - kl_weight * kl
    v = obj[valid_mask.reshape(bs, -1)]
    return -(v.mean() if v.numel() else torch.tensor(0.0))


def ppo_loss(probability, r_pad, action_mode, value, advantage, reward_sum, old_log_prob,
             clip_epsilon: float = 0.2, lambda_entropy: float = 0.01):
    """ppo_pluto/ppo_trainer.py:126-183; `value` = CriticPPO(state)."""
    bs, R, M = probability.shape
    lp = _masked_lsm(probability, r_pad).view(bs, R, M)
    cur = lp[torch.arange(bs), action_mode[:, 0], action_mode[:, 1]]
    entropy = -torch.sum(torch.exp(lp) * lp, dim=(1, 2))
    value_loss = F.smooth_l1_loss(value, reward_sum)
    ratio = (cur - old_log_prob).exp()
    surr = torch.min(advantage * ratio,
                     advantage * torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon)).mean()
    return value_loss - (surr + entropy.mean() * lambda_entropy)


def reinforce_loss(probability, r_pad, returns):
    """reinforce_pluto/reinforce_trainer.py:120-170 (chosen = argmax of the masked logits)."""
    bs, R, M = probability.shape
    z = probability.masked_fill(r_pad.unsqueeze(-1), -1e8)
    best = torch.argmax(z.reshape(bs, -1), dim=1)
    lp = F.log_softmax(z.reshape(bs, -1), dim=1)
    return -torch.mean(lp[torch.arange(bs), best] * returns)


# ---------------------------------------------------------------- SFT / RTR teacher objective
def teacher_label(trajectory, probability, r_pad, teacher_infos, frame_rate: int = 10):
    """fine_tuner/sft/sft_trainer.py:175-215 (generate_target_label) with sft/utils.py:10-33 (global_to_local) and
    pid_controller.py:121-136 (the target-speed half of batch_control_pid).  Returns (flat label index (bs,), masked logits)."""
    bs, R, M = probability.shape
    z = probability.masked_fill(r_pad.unsqueeze(-1), -1e8)
    best_r = torch.argmax(z.reshape(bs, -1), dim=1) // M
    target_speed, origin, heading = teacher_infos[:, 0], teacher_infos[:, 1:3], teacher_infos[:, 3]
    T = trajectory.shape[3]
    pts = trajectory[:, :, :, -1:, :2] if T < frame_rate else trajectory[:, :, :, frame_rate - 1::frame_rate, :2]
    rot = torch.stack([torch.stack([torch.cos(heading), -torch.sin(heading)], dim=1),
                       torch.stack([torch.sin(heading), torch.cos(heading)], dim=1)], dim=1).view(bs, 1, 1, 1, 2, 2)
    local = torch.einsum("brmtc,brmtcd->brmtd", pts - origin.view(bs, 1, 1, 1, 2), rot)
    if local.shape[3] == 1:
        cand_speed = local[..., 0, :].norm(dim=-1, p=2)
    else:
        cand_speed = (local[..., 1:, :] - local[..., :-1, :]).norm(dim=-1, p=2).mean(dim=-1)
    m_idx = torch.argmin(torch.abs(cand_speed - target_speed[:, None, None]).view(bs, -1), dim=1) % M
    return best_r * M + m_idx, z


def sft_loss(trajectory, probability, r_pad, teacher_infos, frame_rate: int = 10):
    """sft_trainer.py:123-173: cross-entropy against the one-hot teacher label (the label is detached)."""
    bs = probability.shape[0]
    label, z = teacher_label(trajectory.detach(), probability, r_pad, teacher_infos, frame_rate)
    return F.cross_entropy(z.reshape(bs, -1), F.one_hot(label, z[0].numel()).to(z.dtype))


def rtr_loss(trajectory, probability, r_pad, teacher_infos, action_mode, value, advantage, reward_sum, old_log_prob,
             frame_rate: int = 10, clip_epsilon: float = 0.2, lambda_entropy: float = 0.01, lambda_rl: float = 5.0):
    """rtr_pluto/rtr_trainer.py:130-195: lambda_rl x (PPO value + actor loss) + the teacher loss."""
    return lambda_rl * ppo_loss(probability, r_pad, action_mode, value, advantage, reward_sum, old_log_prob, clip_epsilon,
                                lambda_entropy) + sft_loss(trajectory, probability, r_pad, teacher_infos, frame_rate)


# ---------------------------------------------------------------- PPO / REINFORCE buffer passes
def gae(rewards, undones, values, next_values, unterminated, gamma=0.98, lam=0.98):
    """ppo_datamodule.py:22-37 — sequential reverse scan, fp32."""
    adv = torch.empty_like(values)
    a = torch.zeros_like(values[0])
    for t in range(rewards.shape[0] - 1, -1, -1):
        delta = rewards[t] + unterminated[t] * gamma * next_values[t] - values[t]
        adv[t] = a = delta + undones[t] * gamma * lam * a
    return adv


def ppo_normalise(adv):
    """ppo_datamodule.py:163 — unbiased std, fp32."""
    return (adv - adv.mean()) / (adv.std(dim=0) + 1e-5)


def discounted_return(rewards, dones, gamma=0.98):
    """reinforce_datamodule.py:19-38 (its trailing 'normalise' acts on a scalar: no effect)."""
    out = torch.zeros_like(rewards)
    g = 0
    for t in range(len(rewards) - 1, -1, -1):
        g = rewards[t] if dones[t] == 1 else rewards[t] + gamma * g
        out[t] = g
    return out


# ---------------------------------------------------------------- optimizer
def warmup_cos_lr(epoch: int, lr: float, min_lr: float, epochs: int, warmup_epochs: int) -> float:
    """pluto/optim/warmup_cos_lr.py:39-54 with last_epoch = epoch."""
    if epoch < warmup_epochs:
        return lr * (epoch + 1) / warmup_epochs
    return min_lr + 0.5 * (lr - min_lr) * (1 + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))


def decay_partition(names_shapes: List[Tuple[str, Tuple[int, ...]]]) -> Tuple[List[str], List[str]]:
    """Outcome of rift_trainer.py:287-324 for the Pluto module tree, as a rule on names:
    'bias' in leaf -> no decay; 'weight' in leaf: Linear/Conv/MHA (ndim>=2, and not an
    Embedding) -> decay, LayerNorm/BatchNorm/Embedding -> no decay; neither substring
    (m_emb, m_pos, query, pos_embed, rpb) -> no decay."""
    decay, no_decay = [], []
    for name, shape in names_shapes:
        leaf = name.rsplit(".", 1)[-1]
        if "bias" in leaf:
            no_decay.append(name)
        elif "weight" in leaf:
            is_embedding = name.endswith("emb.weight") or name.endswith("freqs.weight")
            (decay if (len(shape) >= 2 and not is_embedding) else no_decay).append(name)
        else:
            no_decay.append(name)
    return sorted(decay), sorted(no_decay)


def clip_adamw_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], state: Dict,
                    lr: float, weight_decay: float = 1e-5, max_norm: float = 0.5,
                    betas=(0.9, 0.999), eps: float = 1e-8):
    """Lightning's clip_grad_norm_(0.5, L2) (custom_lightning.yaml:40-41) followed by
    torch.optim.AdamW semantics with the reference's two groups (rift_trainer.py:333-351)."""
    names = sorted(grads)
    total = torch.sqrt(sum((grads[n].double() ** 2).sum() for n in names)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    decay, _ = decay_partition([(n, tuple(params[n].shape)) for n in names])
    decay = set(decay)
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    for n in names:
        g = grads[n] * coef
        m = state.setdefault("m", {}).setdefault(n, torch.zeros_like(g))
        v = state.setdefault("v", {}).setdefault(n, torch.zeros_like(g))
        wd = weight_decay if n in decay else 0.0
        params[n].mul_(1 - lr * wd)
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        params[n].addcdiv_(m, denom, value=-lr / bc1)
    return float(total)
