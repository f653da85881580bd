"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32, functional) of the reference policy.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product path (rift_b200/) never does.

It restates, in deterministic (eval) mode, `PlanningModel.forward`
(rift/cbv/planning/pluto/model/pluto_model.py:122-225) and every module it calls, as plain
functions over a state dict whose keys are the reference's own.  It is pinned against the
unmodified reference modules through tests/golden/*.npz (written in the container by oracle/make_golden.py, which
imports the reference via oracle/ref_shim.py; checked everywhere by tests/test_oracle_golden.py).

Parity status: pinned against the reference's own torch modules run in this container;
"parity unpinned" only at the NATTEN boundary (natten 0.14.6 is not vendored; see ref_shim.py).
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

NAT_KERNELS = (3, 3, 5)
NAT_HEADS = (2, 4, 8)
NAT_DEPTHS = (2, 2, 2)


def _lin(x, sd: SD, p: str):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(x, sd: SD, p: str):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _bn_eval(x, sd: SD, p: str):
    """BatchNorm1d with running statistics (layers/embedding.py:258,265), eps 1e-5."""
    return (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + 1e-5) \
        * sd[p + ".weight"] + sd[p + ".bias"]


def mlp_layer(x, sd: SD, p: str):
    """layers/mlp_layer.py:4-16."""
    return _lin(F.relu(_ln(_lin(x, sd, p + ".mlp.0"), sd, p + ".mlp.1")), sd, p + ".mlp.3")


def fourier_embedding(x, sd: SD, p: str):
    """layers/fourier_embedding.py:45-55.  x: (..., d)."""
    d = x.shape[-1]
    xf = x.unsqueeze(-1) * sd[p + ".freqs.weight"] * 2 * math.pi
    feat = torch.cat([xf.cos(), xf.sin(), x.unsqueeze(-1)], dim=-1)      # (..., d, 129)
    acc = None
    for i in range(d):
        h = _lin(feat[..., i, :], sd, f"{p}.mlps.{i}.0")
        h = F.relu(_ln(h, sd, f"{p}.mlps.{i}.1"))
        h = _lin(h, sd, f"{p}.mlps.{i}.3")
        acc = h if acc is None else acc + h
    return _lin(F.relu(_ln(acc, sd, p + ".to_out.0")), sd, p + ".to_out.2")


def mha(q_in, k_in, v_in, sd: SD, p: str, heads: int, key_padding_mask: Optional[Tensor] = None):
    """nn.MultiheadAttention(batch_first=True), packed in-projection, eval mode."""
    D = q_in.shape[-1]
    W, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, W[:D], b[:D])
    k = F.linear(k_in, W[D:2 * D], b[D:2 * D])
    v = F.linear(v_in, W[2 * D:], b[2 * D:])
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    hd = D // heads
    q = q.view(B, Sq, heads, hd).transpose(1, 2)
    k = k.view(B, Sk, heads, hd).transpose(1, 2)
    v = v.view(B, Sk, heads, hd).transpose(1, 2)
    att = (q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2)
    if key_padding_mask is not None:
        att = att.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    att = att.softmax(-1)
    o = (att @ v).transpose(1, 2).reshape(B, Sq, D)
    return _lin(o, sd, p + ".out_proj")


def points_encoder(x, mask, sd: SD, p: str, out_dim: int):
    """layers/embedding.py:271-296.  Invalid points contribute zeros to both max-pools."""
    B, n, _ = x.shape
    h = _lin(x[mask], sd, p + ".first_mlp.0")
    h = F.relu(_bn_eval(h, sd, p + ".first_mlp.1"))
    h = _lin(h, sd, p + ".first_mlp.3")
    feat = x.new_zeros(B, n, 256)
    feat[mask] = h
    pooled = feat.max(dim=1)[0]
    feat = torch.cat([feat, pooled.unsqueeze(1).expand(B, n, 256)], dim=-1)
    h = _lin(feat[mask], sd, p + ".second_mlp.0")
    h = F.relu(_bn_eval(h, sd, p + ".second_mlp.1"))
    h = _lin(h, sd, p + ".second_mlp.3")
    res = x.new_zeros(B, n, out_dim)
    res[mask] = h
    return res.max(dim=1)[0]


def neighborhood_attention_1d(x, sd: SD, p: str, heads: int, k: int):
    """NATTEN 0.14 NeighborhoodAttention1D, dilation 1 (call site layers/embedding.py:169-178)."""
    B, L, C = x.shape
    hd = C // heads
    qkv = _lin(x, sd, p + ".qkv").view(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, kk, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    i = torch.arange(L)
    idx = (i - k // 2).clamp(0, L - k)[:, None] + torch.arange(k)[None]          # (L, k)
    a = torch.einsum("bhld,bhlkd->bhlk", q, kk[:, :, idx])
    a = a + sd[p + ".rpb"][:, idx - i[:, None] + (k - 1)]
    a = a.softmax(-1)
    o = torch.einsum("bhlk,bhlkd->bhld", a, v[:, :, idx]).permute(0, 2, 1, 3).reshape(B, L, C)
    return _lin(o, sd, p + ".proj")


def nat_sequence_encoder(x, sd: SD, p: str):
    """layers/embedding.py:62-87.  x: (N, C, T) -> (N, D)."""
    x = F.conv1d(x, sd[p + ".embed.proj.weight"], sd[p + ".embed.proj.bias"], padding=1).permute(0, 2, 1)
    outs = []
    for i in range(3):
        for j in range(NAT_DEPTHS[i]):
            b = f"{p}.levels.{i}.blocks.{j}"
            x = x + neighborhood_attention_1d(_ln(x, sd, b + ".norm1"), sd, b + ".attn", NAT_HEADS[i], NAT_KERNELS[i])
            h = _lin(F.gelu(_lin(_ln(x, sd, b + ".norm2"), sd, b + ".mlp.fc1")), sd, b + ".mlp.fc2")
            x = x + h
        outs.append(_ln(x, sd, f"{p}.norm{i}").permute(0, 2, 1))
        if i < 2:
            d = f"{p}.levels.{i}.downsample"
            x = F.conv1d(x.permute(0, 2, 1), sd[d + ".reduction.weight"], None, stride=2, padding=1).permute(0, 2, 1)
            x = _ln(x, sd, d + ".norm")
    lat = [F.conv1d(outs[i], sd[f"{p}.lateral_convs.{i}.weight"], sd[f"{p}.lateral_convs.{i}.bias"], padding=1)
           for i in range(3)]
    for i in (2, 1):
        lat[i - 1] = lat[i - 1] + F.interpolate(
            lat[i], scale_factor=lat[i - 1].shape[-1] / lat[i].shape[-1], mode="linear", align_corners=False)
    out = F.conv1d(lat[0], sd[p + ".fpn_conv.weight"], sd[p + ".fpn_conv.bias"], padding=1)
    return out[:, :, -1]


def state_attention_encoder(x, sd: SD, p: str):
    """modules/agent_encoder.py:97-140, eval mode (no token dropping); 4 heads hard-coded (:104)."""
    toks = torch.stack([_lin(x[:, i, None], sd, f"{p}.linears.{i}") for i in range(x.shape[1])], dim=1)
    toks = toks + sd[p + ".pos_embed"]
    q = sd[p + ".query"].expand(x.shape[0], 1, -1)
    return mha(q, toks, toks, sd, p + ".attn", 4)[:, 0]


def _to_vector(feat, valid):
    vm = valid[..., :-1] & valid[..., 1:]
    while vm.dim() < feat.dim():
        vm = vm.unsqueeze(-1)
    return torch.where(vm, feat[:, :, 1:] - feat[:, :, :-1], torch.zeros_like(feat[:, :, 1:]))


def agent_encoder(data, sd: SD, T: int, state_channel: int, D: int):
    """modules/agent_encoder.py:54-94."""
    a = data["agent"]
    pos, head, vel, shp = a["position"][:, :, :T], a["heading"][:, :, :T], a["velocity"][:, :, :T], a["shape"][:, :, :T]
    valid = a["valid_mask"][:, :, :T]
    hv = _to_vector(head, valid)
    feat = torch.cat([
        _to_vector(pos, valid), _to_vector(vel, valid),
        torch.stack([hv.cos(), hv.sin()], dim=-1), shp[:, :, 1:],
        (valid[..., 1:] & valid[..., :-1]).float().unsqueeze(-1)], dim=-1)
    bs, A, L, C = feat.shape
    feat = feat.view(bs * A, L, C)
    va = valid.any(-1).flatten()
    x = feat.new_zeros(bs * A, D)
    x[va] = nat_sequence_encoder(feat[va].permute(0, 2, 1).contiguous(), sd, "agent_encoder.history_encoder")
    x = x.view(bs, A, D).clone()
    x[:, 0] = state_attention_encoder(data["current_state"][:, :state_channel], sd, "agent_encoder.ego_state_emb")
    return x + sd["agent_encoder.type_emb.weight"][a["category"].long()]


def map_encoder(data, sd: SD, D: int):
    """modules/map_encoder.py:31-93 (use_lane_boundary=True)."""
    m = data["map"]
    pp, pv, po, pc = m["point_position"], m["point_vector"], m["point_orientation"], m["polygon_center"]
    feat = torch.cat([
        pp[:, :, 0] - pc[..., None, :2], pv[:, :, 0],
        torch.stack([po[:, :, 0].cos(), po[:, :, 0].sin()], dim=-1),
        pp[:, :, 1] - pp[:, :, 0], pp[:, :, 2] - pp[:, :, 0]], dim=-1)
    bs, M, P, C = feat.shape
    x = points_encoder(feat.reshape(bs * M, P, C), m["valid_mask"].view(bs * M, P), sd,
                       "map_encoder.polygon_encoder", D).view(bs, M, D)
    has = m["polygon_has_speed_limit"]
    sl = x.new_zeros(bs, M, D)
    sl[has] = fourier_embedding(m["polygon_speed_limit"][has].unsqueeze(-1), sd, "map_encoder.speed_limit_emb")
    sl[~has] = sd["map_encoder.unknown_speed_emb.weight"]
    return x + (sd["map_encoder.type_emb.weight"][m["polygon_type"].long()]
                + sd["map_encoder.on_route_emb.weight"][m["polygon_on_route"].long()]
                + sd["map_encoder.traffic_light_emb.weight"][m["polygon_tl_status"].long()] + sl)


def decoder_layer(t, mem, r_pad, mem_pad, sd: SD, p: str, heads: int, m_pos):
    """modules/planning_decoder.py:42-86 with every Dropout an identity."""
    bs, R, M, D = t.shape
    x = t.transpose(1, 2).reshape(bs * M, R, D)
    x2 = _ln(x, sd, p + ".norm1")
    x = x + mha(x2, x2, x2, sd, p + ".r2r_attn", heads, r_pad.repeat(M, 1))
    xt = x.reshape(bs, M, R, D).transpose(1, 2).reshape(bs * R, M, D)
    keep = ~r_pad.reshape(-1)
    xv = xt[keep]
    x2 = _ln(xv, sd, p + ".norm2")
    xv = xv + mha(x2 + m_pos, x2 + m_pos, x2, sd, p + ".m2m_attn", heads)
    x = torch.zeros_like(xt)
    x[keep] = xv
    x = x.reshape(bs, R * M, D)
    x = x + mha(_ln(x, sd, p + ".norm3"), mem, mem, sd, p + ".cross_attn", heads, mem_pad)
    h = _lin(F.relu(_lin(_ln(x, sd, p + ".norm4"), sd, p + ".ffn.0")), sd, p + ".ffn.3")
    return (x + h).reshape(bs, R, M, D)


def planning_decoder(data, enc, enc_pad, sd: SD, cfg):
    """modules/planning_decoder.py:135-188."""
    p = "planning_decoder"
    r = data["reference_line"]
    rp, rv, ro, rvalid = r["position"], r["vector"], r["orientation"], r["valid_mask"]
    r_pad = ~rvalid.any(-1)
    feat = torch.cat([rp - rp[..., 0:1, :2], rv, torch.stack([ro.cos(), ro.sin()], dim=-1)], dim=-1)
    bs, R, P, C = feat.shape
    D, Mo, T = cfg.dim, cfg.num_modes, cfg.future_steps
    r_emb = points_encoder(feat.reshape(bs * R, P, C), rvalid.view(bs * R, P), sd, p + ".r_encoder", D).view(bs, R, D)
    r_emb = r_emb + fourier_embedding(torch.cat([rp[:, :, 0], ro[:, :, 0, None]], dim=-1), sd, p + ".r_pos_emb")
    q = torch.cat([r_emb.unsqueeze(2).expand(bs, R, Mo, D), sd[p + ".m_emb"].expand(bs, R, Mo, D)], dim=-1)
    q = _lin(q, sd, p + ".q_proj")
    for i in range(cfg.decoder_depth):
        q = decoder_layer(q, enc, r_pad, enc_pad, sd, f"{p}.decoder_blocks.{i}", cfg.num_heads, sd[p + ".m_pos"])
    x_ego = enc[:, 0, None, None, :].expand(bs, R, Mo, D)
    q = _lin(torch.cat([q, x_ego], dim=-1), sd, p + ".cat_x_proj")
    loc = mlp_layer(q, sd, p + ".loc_head").view(bs, R, Mo, T, 2)
    yaw = mlp_layer(q, sd, p + ".yaw_head").view(bs, R, Mo, T, 2)
    vel = mlp_layer(q, sd, p + ".vel_head").view(bs, R, Mo, T, 2)
    pi = mlp_layer(q, sd, p + ".pi_head").squeeze(-1)
    return torch.cat([loc, yaw, vel], dim=-1), pi, q


def planning_model_forward(data, sd: SD, cfg) -> Dict[str, Tensor]:
    """pluto_model.py:122-225.  `cfg` is a rift_b200.config.PlutoConfig (or anything with its fields)."""
    D, Th, T = cfg.dim, cfg.history_steps, cfg.future_steps
    a, m = data["agent"], data["map"]
    agent_pos = a["position"][:, :, Th - 1]
    agent_head = a["heading"][:, :, Th - 1]
    bs, A = agent_pos.shape[:2]
    position = torch.cat([agent_pos, m["polygon_center"][..., :2]], dim=1)
    angle = torch.cat([agent_head, m["polygon_center"][..., 2]], dim=1)
    angle = (angle + math.pi) % (2 * math.pi) - math.pi
    pos = torch.cat([position, angle.unsqueeze(-1)], dim=-1)
    key_pad = torch.cat([~a["valid_mask"][:, :, :Th].any(-1), ~m["valid_mask"].any(-1)], dim=-1)

    x = torch.cat([agent_encoder(data, sd, Th, cfg.state_channel, D), map_encoder(data, sd, D)], dim=1)
    # static objects: N = 0 in RIFT (pluto_feature_builder.py:247-257); nothing to append
    assert data["static_objects"]["position"].shape[1] == 0, "oracle restates the N=0 case only"
    x = x + fourier_embedding(pos, sd, "pos_emb")
    for i in range(cfg.encoder_depth):
        b = f"encoder_blocks.{i}"
        x2 = _ln(x, sd, b + ".norm1")
        x = x + mha(x2, x2, x2, sd, b + ".attn", cfg.num_heads, key_pad)
        x = x + _lin(F.gelu(_lin(_ln(x, sd, b + ".norm2"), sd, b + ".mlp.fc1")), sd, b + ".mlp.fc2")
    x = _ln(x, sd, "norm")

    xa = x[:, 1:A]
    prediction = torch.cat([mlp_layer(xa, sd, f"agent_predictor.{h}").view(bs, A - 1, T, 2)
                            for h in ("loc_predictor", "yaw_predictor", "vel_predictor")], dim=-1)
    trajectory, probability, q = planning_decoder(data, x, key_pad, sd, cfg)
    out = {"trajectory": trajectory, "probability": probability, "prediction": prediction, "decoder_q": q,
           "enc_emb": x}
    out["hidden"] = _lin(F.relu(_lin(x[:, 0], sd, "hidden_proj.0")), sd, "hidden_proj.2")
    rf = mlp_layer(x[:, 0], sd, "ref_free_decoder").reshape(bs, T, 4)
    out["ref_free_trajectory"] = rf
    out["output_ref_free_trajectory"] = torch.cat(
        [rf[..., :2], torch.atan2(rf[..., 3], rf[..., 2]).unsqueeze(-1)], dim=-1)
    out["output_prediction"] = torch.cat([
        prediction[..., :2] + agent_pos[:, 1:A, None],
        torch.atan2(prediction[..., 3], prediction[..., 2]).unsqueeze(-1) + agent_head[:, 1:A, None, None],
        prediction[..., 4:6]], dim=-1)
    r_pad = ~data["reference_line"]["valid_mask"].any(-1)
    probability = probability.masked_fill(r_pad.unsqueeze(-1), -1e6)
    out["probability"] = probability
    cand = torch.cat([trajectory[..., :2], torch.atan2(trajectory[..., 3], trajectory[..., 2]).unsqueeze(-1)], dim=-1)
    R, Mo = probability.shape[1:]
    best = probability.reshape(bs, R * Mo).argmax(-1)
    out["best_index"] = best
    out["output_trajectory"] = cand.reshape(bs, R * Mo, T, 3)[torch.arange(bs), best]
    out["candidate_trajectories"] = cand
    return out


def critic_ppo(state, sd: SD, p: str = "value_net"):
    """rift/gym_carla/utils/net.py:355-372,420-433."""
    x = (state - sd[p + ".state_avg"]) / sd[p + ".state_std"]
    i = 0
    while f"{p}.net.{2 * i}.weight" in sd:
        x = _lin(x, sd, f"{p}.net.{2 * i}")
        if f"{p}.net.{2 * i + 2}.weight" in sd:
            x = F.relu(x)
        i += 1
    return (x * sd[p + ".value_std"] + sd[p + ".value_avg"]).squeeze(1)


def to_torch(tree):
    """numpy feature tree -> torch tensors (float64->float32 like pluto/utils/utils.py:12-30)."""
    import numpy as np
    if isinstance(tree, dict):
        return {k: to_torch(v) for k, v in tree.items()}
    if isinstance(tree, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(tree))
        return t
    return tree
