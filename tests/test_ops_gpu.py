"""GPU: kernel-level parity of the primitive operators against plain PyTorch fp32 on the same inputs
(these are floating-point kernels; tolerances are stated per test)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as TF

from rift_b200 import _lib

pytestmark = pytest.mark.gpu


def P(t):
    return _lib.ptr(t)


def S():
    return _lib.stream_ptr()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def close(a, b, rtol, what=""):
    err = (a.double() - b.double()).abs().max().item()
    scale = max(b.double().abs().max().item(), 1e-30)
    assert err <= rtol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("rows,K,N", [(1, 7, 1), (37, 129, 128), (300, 256, 160), (1000, 27, 64), (4608, 512, 1)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_epilogues(rows, K, N, act):
    x, w, b, r = rnd(rows, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3), rnd(rows, N, seed=4)
    y = torch.empty(rows, N, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_linear(P(x), rows, K, P(w), P(b), N, act, P(r), P(y), 1, S()))
    ref = TF.linear(x, w, b)
    ref = TF.relu(ref) if act == 1 else (TF.gelu(ref) if act == 2 else ref)
    close(y, ref + r, 1e-5, "linear")


@pytest.mark.parametrize("rows,K,N", [(128, 64, 128), (64, 32, 16), (300, 256, 160), (1000, 132, 96), (4608, 512, 256),
                                      (3328, 256, 1024), (129, 1024, 256), (46080, 256, 256), (777, 192, 64),
                                      # tall shapes: the weight-resident / TMA-store variant (ragged M, N = 3 x 64, N tail)
                                      (40961, 64, 192), (20000, 128, 384), (41000, 192, 64), (37900, 256, 200)])
@pytest.mark.parametrize("act", [0, 2])
def test_linear_tcgen05_split_bf16(rows, K, N, act):
    """tcgen05 GEMM (A split in-kernel, W pre-split planes via TMA, 3 bf16 MMAs per k-step) vs fp32."""
    x, w, b, r = rnd(rows, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3), rnd(rows, N, seed=4)
    y = torch.full((rows, N), float("nan"), device="cuda")
    L = _lib.lib()
    scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device="cuda")
    _lib.check(L.rift_b200_op_linear_tc(P(x), rows, K, P(w), P(b), N, act, P(r), P(y), P(scratch), scratch.numel(), 1, S()))
    ref = TF.linear(x.double(), w.double(), b.double())
    ref = TF.gelu(ref) if act == 2 else ref
    close(y, ref + r.double(), 3e-5, "linear_tc")     # split-bf16: 16 mantissa bits per operand


@pytest.mark.parametrize("rows,K,N", [(4608, 256, 256), (46081, 256, 256), (40960, 64, 192), (38000, 128, 72)])
@pytest.mark.parametrize("act,beta", [(1, 0.0), (2, 0.0), (0, 1.0)])
def test_linear_tcgen05_all_outputs(rows, K, N, act, beta):
    """fp32 result (stored or accumulated), pre-activation copy and split-bf16 planes of both GEMM variants."""
    x, w, b, r = rnd(rows, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3), rnd(rows, N, seed=4)
    y0 = rnd(rows, N, seed=5)
    y = y0.clone()
    Kp = (N + 63) // 64 * 64
    pre = torch.full((rows, N), float("nan"), device="cuda")
    hi = torch.full((rows, Kp), 0x7fc0, dtype=torch.int16, device="cuda")      # bf16 NaN pattern
    lo = torch.full((rows, Kp), 0x7fc0, dtype=torch.int16, device="cuda")
    L = _lib.lib()
    scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device="cuda")
    _lib.check(L.rift_b200_op_linear_tc_full(P(x), rows, K, P(w), P(b), N, act, P(r), P(y), beta, P(pre), P(hi), P(lo), P(scratch),
                                             scratch.numel(), S()))
    lin = TF.linear(x.double(), w.double(), b.double())
    ref = (TF.relu(lin) if act == 1 else (TF.gelu(lin) if act == 2 else lin)) + r.double()
    close(pre, lin, 3e-5, "pre-activation")
    close(y, ref + beta * y0.double(), 3e-5, "result")
    planes = hi.view(torch.bfloat16).double() + lo.view(torch.bfloat16).double()
    close(planes[:, :N], ref, 3e-5, "planes")                 # the planes carry the result before beta * C
    assert (planes[:, N:] == 0).all(), "plane padding must be zero"


@pytest.mark.parametrize("rows,N,K,splits", [(4608, 256, 256, 8), (4608, 768, 256, 4), (3328, 256, 1024, 2), (40960, 192, 64, 24),
                                             (384, 256, 129, 3), (46080, 128, 256, 24), (100, 64, 64, 1)])
def test_wgrad_tcgen05_atomic_accumulate_and_bias(rows, N, K, splits):
    """dW += dY^T X with split-K parts added by red.global.add and db += colsum(dY) from the second tensor-core
    accumulator, on top of non-zero running sums; fp32 tolerance 3e-5 of the result's scale."""
    dY, X = rnd(rows, N, seed=11), rnd(rows, K, seed=12)
    dW0, db0 = rnd(N, K, seed=13), rnd(N, seed=14)
    dW, db = dW0.clone(), db0.clone()
    L = _lib.lib()
    scratch = torch.empty(L.rift_b200_op_wgrad_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device="cuda")
    _lib.check(L.rift_b200_op_wgrad_tc(P(dY), P(X), rows, N, K, P(dW), P(db), splits, P(scratch), scratch.numel(), S()))
    close(dW, dW0.double() + dY.double().t() @ X.double(), 3e-5, "dW")
    close(db, db0.double() + dY.double().sum(0), 3e-5, "db")


@pytest.mark.parametrize("shapes", [[(4608, 256, 256), (4608, 768, 256), (3328, 256, 1024), (4608, 1024, 256), (3328, 512, 256), (4608, 256, 256)],
                                    [(100, 64, 128)], [(20480, 384, 128), (64, 256, 256), (1000, 100, 132)],
                                    [(1024 + 64 * i, 128 + 64 * (i % 3), 128 + 128 * (i % 2)) for i in range(12)]])
def test_wgrad_grouped_launch(shapes):
    """Up to twelve dW += dY^T X (+ db += colsum(dY)) products of different shapes in ONE persistent launch (wgrad_group_kernel):
    every product against fp64, on top of non-zero running sums, 3e-5 of the result's scale."""
    import ctypes as C
    n = len(shapes)
    L = _lib.lib()
    dYs = [rnd(r, N, seed=20 + i) for i, (r, N, K) in enumerate(shapes)]
    Xs = [rnd(r, K, seed=40 + i) for i, (r, N, K) in enumerate(shapes)]
    dW0 = [rnd(N, K, seed=60 + i) for i, (r, N, K) in enumerate(shapes)]
    db0 = [rnd(N, seed=80 + i) for i, (r, N, K) in enumerate(shapes)]
    dW, db = [t.clone() for t in dW0], [t.clone() for t in db0]
    total = sum(L.rift_b200_op_wgrad_tc_scratch_bytes(r, N, K) for r, N, K in shapes)
    scratch = torch.empty(total, dtype=torch.uint8, device="cuda")
    vp = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    ia = lambda xs: (C.c_int * n)(*xs)
    _lib.check(L.rift_b200_op_wgrad_group(n, vp(dYs), vp(Xs), ia([s_[0] for s_ in shapes]), ia([s_[1] for s_ in shapes]),
                                          ia([s_[2] for s_ in shapes]), vp(dW), vp(db), P(scratch), scratch.numel(), S()))
    for i in range(n):
        close(dW[i], dW0[i].double() + dYs[i].double().t() @ Xs[i].double(), 3e-5, f"dW[{i}]")
        close(db[i], db0[i].double() + dYs[i].double().sum(0), 3e-5, f"db[{i}]")


@pytest.mark.parametrize("layout", ["nt", "nn", "tn", "tt"])
@pytest.mark.parametrize("M,N,K,split", [(65, 70, 33, 1), (256, 192, 4000, 8), (1, 256, 4608, 9)])
def test_gemm_strides_and_split_k(layout, M, N, K, split):
    a = rnd(M, K, seed=5, scale=K ** -0.5)
    b = rnd(N, K, seed=6)
    A = a if layout[0] == "n" else a.t().contiguous()        # stored (K, M) when transposed
    B = b if layout[1] == "t" else b.t().contiguous()        # stored (K, N) for 'n'
    sam, sak = (K, 1) if layout[0] == "n" else (1, M)
    sbn, sbk = (K, 1) if layout[1] == "t" else (1, N)
    c0 = rnd(M, N, seed=7)
    c = c0.clone()
    ws = torch.empty(max(split, 1) * M * N, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_gemm(P(A), sam, sak, P(B), sbn, sbk, P(c), N, M, N, K, 1.0, split, P(ws), 1, S()))
    close(c, c0 + a @ b.t(), 2e-5, "gemm")


@pytest.mark.parametrize("rows,Cn", [(5, 32), (1000, 128), (333, 512), (64, 1024)])
@pytest.mark.parametrize("relu", [0, 1])
def test_layernorm_forward_backward(rows, Cn, relu):
    x = rnd(rows, Cn, seed=1, scale=2.0).requires_grad_(True)
    g = (1 + 0.1 * rnd(Cn, seed=2)).requires_grad_(True)
    b = (0.1 * rnd(Cn, seed=3)).requires_grad_(True)
    y = torch.empty(rows, Cn, device="cuda")
    mean = torch.empty(rows, device="cuda")
    rstd = torch.empty(rows, device="cuda")
    L = _lib.lib()
    _lib.check(L.rift_b200_op_layernorm(P(x), rows, Cn, P(g), P(b), relu, P(y), P(mean), P(rstd), S()))
    ref = TF.layer_norm(x, (Cn,), g, b, 1e-5)
    ref = TF.relu(ref) if relu else ref
    close(y, ref, 2e-6, "layernorm")
    dy = rnd(rows, Cn, seed=4)
    ref.backward(dy)
    dx = torch.empty_like(y)
    dg = torch.zeros(Cn, device="cuda")
    db = torch.zeros(Cn, device="cuda")
    scratch = torch.empty(592 * 2 * Cn, device="cuda")      # rift_b200_op_layernorm_bwd: 592 x 2 x C floats
    _lib.check(L.rift_b200_op_layernorm_bwd(P(x), P(dy), rows, Cn, P(g), P(mean), P(rstd), P(y) if relu else None,
                                            P(dx), P(dg), P(db), P(scratch), S()))
    close(dx, x.grad, 2e-5, "ln dx")
    close(dg, g.grad, 2e-5, "ln dgamma")
    close(db, b.grad, 2e-5, "ln dbeta")


@pytest.mark.parametrize("B,Sq,H,hd", [(3, 52, 4, 32), (7, 6, 8, 32), (2, 200, 4, 32), (5, 12, 4, 64)])
def test_attention_with_key_padding(B, Sq, H, hd):
    D = H * hd
    qkv = rnd(B, Sq, 3 * D, seed=1)
    kpm = torch.zeros(B, Sq, dtype=torch.bool)
    gen = torch.Generator().manual_seed(3)
    for b in range(B):
        n = int(torch.randint(1, Sq + 1, (1,), generator=gen))
        kpm[b, n:] = True
    kpm = kpm.cuda()
    out = torch.empty(B, Sq, D, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_attention(P(qkv), B, Sq, H, hd, P(kpm.view(torch.uint8)), P(out), S()))
    q, k, v = [t.view(B, Sq, H, hd).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    att = (q * hd ** -0.5) @ k.transpose(-1, -2)
    att = att.masked_fill(kpm[:, None, None, :], float("-inf")).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(B, Sq, D)
    close(out, ref, 5e-6, "attention")


@pytest.mark.parametrize("L,heads,hd,k", [(20, 2, 16, 3), (10, 4, 16, 3), (5, 8, 16, 5), (20, 2, 32, 3), (5, 8, 32, 5)])
def test_neighborhood_attention(L, heads, hd, k):
    """NATTEN 0.14 NeighborhoodAttention1D semantics, restated on q/k/v directly."""
    n, dim = 37, heads * hd
    qkv = rnd(n, L, 3 * dim, seed=1)
    rpb = rnd(heads, 2 * k - 1, seed=2, scale=0.3)
    out = torch.empty(n, L, dim, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_nat_attention(P(qkv), n, L, heads, hd, k, P(rpb), P(out), S()))
    q, kk, v = qkv.view(n, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q = q * hd ** -0.5
    i = torch.arange(L, device="cuda")
    idx = (i - k // 2).clamp(0, L - k)[:, None] + torch.arange(k, device="cuda")[None]
    a = torch.einsum("bhld,bhlkd->bhlk", q, kk[:, :, idx]) + rpb[:, idx - i[:, None] + (k - 1)]
    ref = torch.einsum("bhlk,bhlkd->bhld", a.softmax(-1), v[:, :, idx]).permute(0, 2, 1, 3).reshape(n, L, dim)
    close(out, ref, 5e-6, "nat attention")


@pytest.mark.parametrize("B,Sq,H,hd", [(3, 52, 4, 32), (7, 6, 8, 32), (5, 12, 8, 32), (2, 72, 8, 32), (2, 200, 4, 32), (5, 12, 4, 64)])
def test_attention_backward_vs_autograd(B, Sq, H, hd):
    """The cooperative (head_dim 32, short sequences) and the per-row attention backward kernels vs torch autograd."""
    D = H * hd
    qkv = rnd(B, Sq, 3 * D, seed=1)
    d_out = rnd(B, Sq, D, seed=2)
    kpm = torch.zeros(B, Sq, dtype=torch.bool)
    gen = torch.Generator().manual_seed(3)
    for b in range(B):
        n = int(torch.randint(1, Sq + 1, (1,), generator=gen))
        kpm[b, n:] = True
    kpm = kpm.cuda()
    out = torch.empty(B, Sq, D, device="cuda")
    lse = torch.empty(B, H, Sq, device="cuda")
    dqkv = torch.empty_like(qkv)
    _lib.check(_lib.lib().rift_b200_op_attention_bwd(P(qkv), P(d_out), B, Sq, H, hd, P(kpm.view(torch.uint8)), P(out), P(lse),
                                                     P(dqkv), S()))
    x = qkv.clone().requires_grad_(True)
    q, k, v = [t.view(B, Sq, H, hd).transpose(1, 2) for t in x.split(D, dim=-1)]
    att = ((q * hd ** -0.5) @ k.transpose(-1, -2)).masked_fill(kpm[:, None, None, :], float("-inf")).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(B, Sq, D)
    (ref * d_out).sum().backward()
    close(out, ref.detach(), 5e-6, "attention forward")
    close(dqkv, x.grad, 2e-5, "attention backward")       # tolerance: fp32 round-off of a different summation order


@pytest.mark.parametrize("L,heads,hd,k", [(20, 2, 32, 3), (10, 4, 32, 3), (5, 8, 32, 5), (20, 2, 16, 3), (12, 4, 32, 3)])
def test_neighborhood_attention_backward_vs_autograd(L, heads, hd, k):
    """Register-resident (L, k) specialisations and the generic kernel vs torch autograd of the restated semantics."""
    n, dim = 37, heads * hd
    qkv = rnd(n, L, 3 * dim, seed=1)
    rpb = rnd(heads, 2 * k - 1, seed=2, scale=0.3)
    d_out = rnd(n, L, dim, seed=3)
    dqkv = torch.empty_like(qkv)
    part = torch.empty(n * heads, 2 * k - 1, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_nat_attention_bwd(P(qkv), P(d_out), n, L, heads, hd, k, P(rpb), P(dqkv), P(part), S()))
    x = qkv.clone().requires_grad_(True)
    rb = rpb.clone().requires_grad_(True)
    q, kk, v = x.view(n, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q = q * hd ** -0.5
    i = torch.arange(L, device="cuda")
    idx = (i - k // 2).clamp(0, L - k)[:, None] + torch.arange(k, device="cuda")[None]
    a = torch.einsum("bhld,bhlkd->bhlk", q, kk[:, :, idx]) + rb[:, idx - i[:, None] + (k - 1)]
    ref = torch.einsum("bhlk,bhlkd->bhld", a.softmax(-1), v[:, :, idx]).permute(0, 2, 1, 3).reshape(n, L, dim)
    (ref * d_out).sum().backward()
    close(dqkv, x.grad, 2e-5, "nat attention backward")
    close(part.view(n, heads, 2 * k - 1).sum(0), rb.grad, 2e-5, "nat attention rpb gradient")


def test_masked_maxpool():
    groups, n, Cn = 50, 20, 96
    x = rnd(groups, n, Cn, seed=1)
    gen = torch.Generator().manual_seed(2)
    mask = torch.rand(groups, n, generator=gen) > 0.3
    mask[0] = False
    mask[1] = True
    mask = mask.cuda()
    out = torch.empty(groups, Cn, device="cuda")
    arg = torch.empty(groups, Cn, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().rift_b200_op_masked_maxpool(P(x), P(mask.view(torch.uint8)), groups, n, Cn, P(out), P(arg), S()))
    feat = torch.zeros_like(x)
    feat[mask] = x[mask]
    ref, idx = feat.max(dim=1)
    assert torch.equal(out, ref)
    hit = arg >= 0
    assert torch.equal(torch.gather(x, 1, arg.clamp(min=0).long().unsqueeze(1)).squeeze(1)[hit], ref[hit])
    assert (ref[~hit] == 0).all()


def _planes(hi, lo):
    return hi.view(torch.bfloat16).double() + lo.view(torch.bfloat16).double()


@pytest.mark.parametrize("rows,D,Hd,act", [(4608, 256, 1024, 1), (3328, 256, 1024, 2), (300, 128, 512, 2), (40960, 64, 192, 2),
                                           (20480, 128, 384, 2), (10240, 256, 768, 2), (100, 256, 1024, 1), (129, 64, 192, 1)])
def test_fused_mlp_block(rows, D, Hd, act):
    """One cluster kernel for y = x + fc2(act(fc1(LN(x)))) against fp64 torch: result, LayerNorm statistics, and the
    tensors saved for the backward (LN(x) planes, fc1 pre-activation, hidden planes); tolerance 3e-5 of each scale."""
    x = rnd(rows, D, seed=1, scale=2.0)
    g, b = 1 + 0.1 * rnd(D, seed=2), 0.1 * rnd(D, seed=3)
    w1, b1 = rnd(Hd, D, seed=4, scale=D ** -0.5), 0.1 * rnd(Hd, seed=5)
    w2, b2 = rnd(D, Hd, seed=6, scale=Hd ** -0.5), 0.1 * rnd(D, seed=7)
    y = torch.full((rows, D), float("nan"), device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    t2h, t2l = (torch.full((rows, D), 0x7fc0, dtype=torch.int16, device="cuda") for _ in range(2))
    hmh, hml = (torch.full((rows, Hd), 0x7fc0, dtype=torch.int16, device="cuda") for _ in range(2))
    hpre = torch.full((rows, Hd), float("nan"), device="cuda")
    L = _lib.lib()
    scratch = torch.empty(L.rift_b200_op_fused_mlp_scratch_bytes(D, Hd), dtype=torch.uint8, device="cuda")
    _lib.check(L.rift_b200_op_fused_mlp(P(x), rows, D, Hd, act, P(g), P(b), P(w1), P(b1), P(w2), P(b2), P(y), P(mean), P(rstd),
                                        P(t2h), P(t2l), P(hpre), P(hmh), P(hml), P(scratch), scratch.numel(), 1, S()))
    xd = x.double()
    t2 = TF.layer_norm(xd, (D,), g.double(), b.double(), 1e-5)
    pre = TF.linear(t2, w1.double(), b1.double())
    hm = TF.relu(pre) if act == 1 else TF.gelu(pre)
    ref = xd + TF.linear(hm, w2.double(), b2.double())
    close(mean, xd.mean(1), 1e-5, "mean")
    close(rstd, (xd.var(1, unbiased=False) + 1e-5).rsqrt(), 1e-5, "rstd")
    close(_planes(t2h, t2l), t2, 3e-5, "LN planes")
    close(hpre, pre, 3e-5, "pre-activation")
    close(_planes(hmh, hml), hm, 3e-5, "hidden planes")
    close(y, ref, 3e-5, "result")
