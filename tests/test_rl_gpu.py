"""GPU: the RL kernels (advantage, objectives, buffer passes, optimizer) through the C ABI against
the oracle (oracle/loss_oracle.py) and the golden vectors written by the reference."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as lo
from rift_b200 import functional as F
from tests.helpers import CASES, case_inputs, golden, check_golden

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------ advantage: bit-exact
def test_group_advantage_golden_bit_exact():
    g = golden("advantage")
    for k in g.files:
        if k.startswith("ret_"):
            G = int(k[4:])
            adv = F.group_advantage(torch.from_numpy(g[k]).cuda()).cpu().numpy()
            assert np.array_equal(adv, g[f"adv_{G}"]), f"G={G}"


@pytest.mark.parametrize("G", [1, 2, 7, 8, 9, 12, 15, 16, 72, 127, 128, 129, 144, 257, 600])
def test_group_advantage_vs_numpy_bit_exact(G):
    rng = np.random.Generator(np.random.PCG64(G))
    ret = rng.normal(-5.0, 20.0, (257, G))
    ref = np.stack([lo.group_advantage(r) for r in ret])
    adv = F.group_advantage(torch.from_numpy(ret).cuda()).cpu().numpy()
    assert np.array_equal(adv, ref)


def test_group_advantage_ragged_offsets_bit_exact():
    rng = np.random.Generator(np.random.PCG64(5))
    sizes = rng.integers(1, 7, 300) * 12                         # R_valid * 12 modes, as in get_grpo_advantage
    sizes[:3] = (1, 133, 1000)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    ret = rng.normal(-5.0, 20.0, offs[-1])
    ref = np.concatenate([lo.group_advantage(ret[offs[i]:offs[i + 1]]) for i in range(len(sizes))])
    adv = F.group_advantage(torch.from_numpy(ret).cuda(), torch.from_numpy(offs).cuda()).cpu().numpy()
    assert np.array_equal(adv, ref)


def test_group_advantage_large_properties():
    """BASELINE cfg5 scale (2^18 groups x 72): spot-check bit-exactness on a sample of groups, and
    the size-independent properties zero mean / unit population variance."""
    n, G = 1 << 18, 72
    gen = torch.Generator(device="cuda").manual_seed(0)
    ret = torch.randn(n, G, dtype=torch.float64, device="cuda", generator=gen) * 20 - 5
    adv = F.group_advantage(ret)
    assert float(adv.mean(dim=1).abs().max()) < 1e-12
    assert float((adv.var(dim=1, unbiased=False) - 1).abs().max()) < 1e-4      # 1e-5 added to std
    idx = torch.arange(0, n, 4099, device="cuda")
    r, a = ret[idx].cpu().numpy(), adv[idx].cpu().numpy()
    for i in range(len(idx)):
        assert np.array_equal(a[i], lo.group_advantage(r[i]))


@pytest.mark.parametrize("G", [1, 5, 12, 24, 31, 32, 33, 36, 72, 128])
def test_group_advantage_many_groups_bit_exact(G):
    """Every kernel form (one lane per group for G <= 32, eight lanes per group up to 128) at a group count that selects
    it, against numpy's vectorised expression (checked against the per-group oracle on a sample)."""
    rng = np.random.Generator(np.random.PCG64(100 + G))
    n = 40009
    ret = rng.normal(-5.0, 20.0, (n, G))
    ref = (ret - ret.mean(1, keepdims=True)) / (ret.std(1, keepdims=True) + 1e-5)
    for i in (0, 1, 4099, n - 1):
        assert np.array_equal(ref[i], lo.group_advantage(ret[i]))
    adv = F.group_advantage(torch.from_numpy(ret).cuda()).cpu().numpy()
    assert np.array_equal(adv, ref)


def test_group_advantage_division_bit_exact_over_dynamic_range():
    """The shared-divisor quotient (reciprocal + two FMA corrections, Markstein) must equal numpy's division bit for
    bit: 2^17 groups whose scales span 2^-350 .. 2^350 (beyond the fast path's range on both sides), 1.5 M quotients."""
    rng = np.random.Generator(np.random.PCG64(11))
    n, G = 1 << 17, 12
    scale = np.exp2(rng.uniform(-350.0, 350.0, (n, 1)))
    ret = rng.normal(-0.25, 1.0, (n, G)) * scale
    ret[5] = 0.0                                                   # zero spread: x - mean = 0, divisor 1e-5
    ret[6, :6] = 0.0
    with np.errstate(all="ignore"):
        ref = (ret - ret.mean(1, keepdims=True)) / (ret.std(1, keepdims=True) + 1e-5)
    for i in (0, 5, 6, 77, n - 1):                                 # the vectorised expression equals the per-group one
        assert np.array_equal(ref[i], lo.group_advantage(ret[i]), equal_nan=True)
    adv = F.group_advantage(torch.from_numpy(ret).cuda()).cpu().numpy()
    assert np.array_equal(adv, ref, equal_nan=True)


# ------------------------------------------------------------------ objectives
def _rand_objective_inputs(bs, R, Mo, seed, ragged=True):
    rng = np.random.Generator(np.random.PCG64(seed))
    nr = rng.integers(1, R + 1, bs) if ragged else np.full(bs, R)
    nr[0] = R
    r_valid = np.arange(R)[None, :] < nr[:, None]
    vm = np.broadcast_to(r_valid[..., None], (bs, R, Mo)).copy()
    z = rng.normal(0, 2.0, (bs, R, Mo)).astype(np.float32)
    old = (z + rng.normal(0, 0.3, z.shape)).astype(np.float32)      # ratios around the clip range
    ref = rng.normal(0, 1.0, z.shape).astype(np.float32)
    adv = rng.normal(0, 1.0, z.shape) * vm
    return [torch.from_numpy(x) for x in (z, old, ref, adv, vm, ~r_valid)]


@pytest.mark.parametrize("algo", ["rift", "grpo"])
@pytest.mark.parametrize("bs,R,Mo", [(1, 1, 12), (64, 6, 12), (257, 3, 12), (8, 40, 12)])
def test_group_objective_vs_oracle(algo, bs, R, Mo):
    z, old, ref, adv, vm, r_pad = _rand_objective_inputs(bs, R, Mo, seed=bs + R)
    zc = z.clone().requires_grad_(True)
    if algo == "rift":
        loss_ref = lo.rift_loss(zc, old, adv, vm, r_pad)
    else:
        loss_ref = lo.grpo_loss(zc, old, ref, adv, vm, r_pad)
    loss_ref.backward()
    loss, dz, stats = F.group_objective(algo, z.cuda(), old.cuda(), adv.cuda(), vm.cuda(), r_pad.cuda(),
                                        ref.cuda() if algo == "grpo" else None)
    assert loss.dtype == torch.float64                       # the reference's loss is fp64 (fp64 advantages)
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(abs(float(loss_ref)), 1e-3)   # tolerance: 1e-3 rel required
    assert int(stats[2]) == int(vm.sum())
    err = (dz.cpu() - zc.grad).abs().max().item()
    assert err <= 1e-5 * zc.grad.abs().max().item() + 1e-9
    # r_pad derived from the valid mask (rift_trainer.py:148-150) gives the same answer
    loss2, dz2, _ = F.group_objective(algo, z.cuda(), old.cuda(), adv.cuda(), vm.cuda(), None,
                                      ref.cuda() if algo == "grpo" else None)
    assert float(loss2) == float(loss) and torch.equal(dz2, dz)


def test_group_objective_no_valid_candidates_is_zero():
    z, old, ref, adv, vm, r_pad = _rand_objective_inputs(4, 2, 12, seed=1, ragged=False)
    vm[:] = False
    loss, dz, stats = F.group_objective("rift", z.cuda(), old.cuda(), adv.cuda(), vm.cuda(), r_pad.cuda())
    assert float(loss) == 0.0 and float(dz.abs().max()) == 0.0


def test_autograd_wrappers_match_reference_loss_functions():
    z, old, ref, adv, vm, r_pad = _rand_objective_inputs(16, 4, 12, seed=3)
    zc = z.cuda().requires_grad_(True)
    loss = F.grpo_loss(zc, old.cuda(), ref.cuda(), adv.cuda(), vm.cuda(), r_pad.cuda())
    (2.0 * loss).backward()
    z2 = z.clone().requires_grad_(True)
    (2.0 * lo.grpo_loss(z2, old, ref, adv, vm, r_pad)).backward()
    assert torch.allclose(zc.grad.cpu(), z2.grad, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("mode", ["ppo", "reinforce"])
def test_action_objective_vs_oracle(mode):
    bs, R, Mo = 33, 5, 12
    z, _, _, _, vm, r_pad = _rand_objective_inputs(bs, R, Mo, seed=9)
    rng = np.random.Generator(np.random.PCG64(2))
    nr = (~r_pad).sum(-1).numpy()
    am = torch.from_numpy(np.stack([rng.integers(0, nr), rng.integers(0, Mo, bs)], -1).astype(np.int64))
    w = torch.from_numpy(rng.normal(0, 1, bs).astype(np.float32))
    olp = torch.from_numpy((-rng.uniform(2.0, 5.0, bs)).astype(np.float32))
    zc = z.clone().requires_grad_(True)
    if mode == "ppo":
        value = torch.from_numpy(rng.normal(0, 2, bs).astype(np.float32)).requires_grad_(True)
        rsum = torch.from_numpy(rng.normal(0, 2, bs).astype(np.float32))
        ref = lo.ppo_loss(zc, r_pad, am, value, w, rsum, olp)
        ref.backward()
        vl, dv = F.smooth_l1(value.detach().cuda(), rsum.cuda())
        loss, dz, _ = F.action_objective("ppo", z.cuda(), r_pad.cuda(), w.cuda(), am.cuda(), olp.cuda(), extra_loss=vl)
        assert torch.allclose(dv.cpu(), value.grad, rtol=1e-5, atol=1e-8)
    else:
        ref = lo.reinforce_loss(zc, r_pad, w)
        ref.backward()
        loss, dz, chosen = F.action_objective("reinforce", z.cuda(), r_pad.cuda(), w.cuda())
        zm = z.masked_fill(r_pad.unsqueeze(-1), -1e8).reshape(bs, -1)
        assert torch.equal(chosen.cpu().long(), zm.argmax(1))          # indices: exact
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(abs(float(ref)), 1e-3)
    assert (dz.cpu() - zc.grad).abs().max().item() <= 1e-5 * zc.grad.abs().max().item() + 1e-9


@pytest.mark.parametrize("name", list(CASES))
def test_objectives_against_reference_golden_logits(name):
    """Feed the reference's own logits (golden) through the kernels: loss and d loss/d logits."""
    cfg, sd, feats, ex = case_inputs(name)
    g = golden(name)
    z = torch.from_numpy(g["out_probability"]).cuda()
    r_pad = torch.from_numpy(~feats["reference_line"]["valid_mask"].any(-1)).cuda()
    t = {k: torch.from_numpy(v).cuda() for k, v in ex.items()}
    for algo in ("rift", "grpo"):
        loss, dz, _ = F.group_objective(algo, z, t["old_group_logits"], t["group_advantage"], t["group_advantage_mask"],
                                        r_pad, t["ref_group_logits"])
        ref = float(g[f"loss_{algo}"])
        assert abs(float(loss) - ref) <= 1e-5 * max(abs(ref), 1e-3)
        check_golden(g, f"dlogits_{algo}", dz.cpu().numpy(), rtol=1e-4, atol=1e-9)
    loss, dz, _ = F.action_objective("reinforce", z, r_pad, t["return"])
    assert abs(float(loss) - float(g["loss_reinforce"])) <= 1e-5 * max(abs(float(g["loss_reinforce"])), 1e-3)
    check_golden(g, "dlogits_reinforce", dz.cpu().numpy(), rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------ buffer passes
def test_gae_and_returns_golden():
    g = golden("buffer_pass")
    t = {k: torch.from_numpy(g[k]).cuda() for k in g.files}
    adv, rsum, advn = F.gae(t["rewards"], 1 - t["dones"], t["values"], t["next_values"], 1 - t["terminated"])
    assert torch.equal(adv, t["gae"])                                   # sequential fp32 order: exact
    assert torch.equal(rsum, t["reward_sum"])
    assert torch.allclose(advn, t["gae_normalised"], rtol=1e-6, atol=1e-6)
    assert torch.equal(F.discounted_return(t["rewards"], t["dones"]), t["discounted_return"])


# ------------------------------------------------------------------ optimizer
@pytest.mark.parametrize("n,n_decay", [(16897 + 63, 16512), (1 << 20, 1 << 19), (7, 3)])
def test_clip_adamw_vs_torch(n, n_decay):
    gen = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=gen) * 0.1
    pa = torch.nn.Parameter(p0[:n_decay].clone())
    pb = torch.nn.Parameter(p0[n_decay:].clone())
    opt = torch.optim.AdamW([{"params": [pa], "weight_decay": 1e-2}, {"params": [pb], "weight_decay": 0.0}], lr=1e-3,
                            weight_decay=1e-2)
    n_pad = (n + 3) // 4 * 4
    p = torch.zeros(n_pad).cuda()
    p[:n] = p0.cuda()
    gbuf = torch.zeros(n_pad).cuda()
    ours = F.ClipAdamW(p, gbuf, n, n_decay, lr=1e-3, weight_decay=1e-2, max_norm=0.5)
    for step in range(4):
        grad = torch.randn(n, generator=gen) * (10.0 if step % 2 == 0 else 1e-4)     # clipped and unclipped steps
        pa.grad, pb.grad = grad[:n_decay].clone(), grad[n_decay:].clone()
        tn = torch.nn.utils.clip_grad_norm_([pa, pb], 0.5)
        opt.step()
        gbuf[:n] = grad.cuda()
        ours.step()
        assert abs(ours.grad_norm() - float(tn)) <= 1e-5 * float(tn)
        ref = torch.cat([pa.detach(), pb.detach()])
        assert (p[:n].cpu() - ref).abs().max().item() <= 2e-6, step


def test_clip_adamw_divides_by_device_count():
    n = 1024
    gen = torch.Generator().manual_seed(1)
    p0, g0 = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    a = F.ClipAdamW(p0.clone().cuda(), (g0 * 7).cuda(), n, n, max_norm=0.5)
    b = F.ClipAdamW(p0.clone().cuda(), g0.clone().cuda(), n, n, max_norm=0.5)
    a.step(count=torch.tensor([7.0], dtype=torch.float64, device="cuda"))
    b.step()
    assert torch.allclose(a.p, b.p, rtol=0, atol=1e-7)
    assert abs(a.grad_norm() - b.grad_norm()) < 1e-4 * b.grad_norm()


def test_clip_adamw_skips_the_update_when_nothing_is_valid():
    """count == 0 (no valid objective term): the reference's loss is a constant, Lightning takes no optimizer step -
    no weight decay, no moment decay, no step-counter increment."""
    n = 2048
    gen = torch.Generator().manual_seed(2)
    p0, g0 = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    o = F.ClipAdamW(p0.clone().cuda(), g0.clone().cuda(), n, n, lr=1e-2, weight_decay=0.1, max_norm=0.5)
    o.step(count=torch.tensor([3.0], dtype=torch.float64, device="cuda"))
    assert o.step_count == 1
    p1, m1, v1 = o.p.clone(), o.m.clone(), o.v.clone()
    o.step(count=torch.tensor([0.0], dtype=torch.float64, device="cuda"))
    assert o.step_count == 1
    assert torch.equal(o.p, p1) and torch.equal(o.m, m1) and torch.equal(o.v, v1)
    o.param_groups[0]["lr"] = 5e-3                      # a scheduler step reaches the device-resident learning rate
    o.step(count=torch.tensor([3.0], dtype=torch.float64, device="cuda"))
    assert o.step_count == 2 and abs(float(o.hyper[0]) - 5e-3) < 1e-9 and not torch.equal(o.p, p1)
