"""Seeded synthetic parameters and rollout batches (numpy only; no torch, no CUDA).

The reference ships no weights (``model_ckpt/.gitkeep`` only) and its inputs come from a live
CARLA world, so parity and the benchmark run on synthetic tensors with the exact layout of a
collated ``PlutoFeature`` (rift/cbv/planning/pluto/feature_builder/pluto_feature.py:25-96) and
of ``GRPOCollate`` / ``RIFTCollate`` / ``PPOCollate`` / ``ReinforceCollate`` batches
(fine_tuner/rlft/*/*_datamodule.py).  Everything is generated with numpy's PCG64 so that the
container that writes tests/golden/ and the GPU box that checks them see identical bits.
"""
from typing import Dict, Optional

import numpy as np

from .config import PlutoConfig, param_spec


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def synth_state_dict(cfg: PlutoConfig, seed: int = 7) -> Dict[str, np.ndarray]:
    """Random parameters for every state-dict entry.

    The reference initialisation (pluto_model.py:108-120) leaves every bias at 0, every norm at
    (1, 0) and BatchNorm statistics at (0, 1), which would hide missing-bias / wrong-statistics
    bugs; so every entry is randomised: matrices ~ U(+-sqrt(6/(fan_in+fan_out))) (xavier),
    biases ~ N(0, 0.02), norm gains ~ 1 + N(0, 0.1), running_mean ~ N(0, 0.5),
    running_var ~ U(0.5, 1.5), embeddings / rpb / queries ~ N(0, 0.05), freqs ~ N(0, 0.02).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    for name, shape, kind in param_spec(cfg):
        leaf = name.rsplit(".", 1)[-1]
        if kind == "i64":
            out[name] = np.zeros(shape, np.int64)
            continue
        if leaf == "running_mean":
            v = rng.normal(0.0, 0.5, shape)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, shape)
        elif name.startswith("value_net.") and leaf in ("state_avg", "value_avg"):
            v = rng.normal(0.0, 0.1, shape)
        elif name.startswith("value_net.") and leaf in ("state_std", "value_std"):
            v = rng.uniform(0.8, 1.2, shape)
        elif leaf == "bias" or leaf == "in_proj_bias":
            v = rng.normal(0.0, 0.02, shape)
        elif leaf == "weight" and len(shape) == 1:
            v = 1.0 + rng.normal(0.0, 0.1, shape)
        elif name.endswith("freqs.weight"):
            v = rng.normal(0.0, 0.02, shape)
        elif leaf in ("weight", "in_proj_weight") and len(shape) >= 2 and "emb.weight" not in name \
                and "_emb.weight" not in name:
            fan_out = shape[0]
            fan_in = int(np.prod(shape[1:]))
            a = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-a, a, shape)
        else:  # embeddings, rpb, m_emb, m_pos, query, pos_embed
            v = rng.normal(0.0, 0.05, shape)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    return out


# --------------------------------------------------------------------------------------
# feature batches
# --------------------------------------------------------------------------------------
def synth_features(cfg: PlutoConfig, bs: int, A: int, Mp: int, R: int, seed: int = 1,
                   ragged: bool = False, P: int = 20) -> Dict[str, Dict[str, np.ndarray]]:
    """One collated PlutoFeature.data dict as numpy arrays.

    Padding follows ``pad_sequence``: invalid agents / polygons / reference lines sit at the end
    of their axis and are zero-filled.  ``ragged`` draws per-sample valid counts.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    Th = cfg.history_steps
    Pr = cfg.ref_points
    f32 = np.float32

    # ---- agents: constant-velocity tracks with noise, agent 0 is the centre vehicle at the origin
    p0 = rng.normal(0.0, 25.0, (bs, A, 1, 2))
    p0[:, 0] = 0.0
    h0 = rng.normal(0.0, 1.0, (bs, A, 1))
    h0[:, 0] = 0.0
    speed = np.abs(rng.normal(0.0, 5.0, (bs, A, 1)))
    t = (np.arange(Th) - (Th - 1)).reshape(1, 1, Th) * 0.1
    yaw_rate = rng.normal(0.0, 0.2, (bs, A, 1))
    heading = h0 + yaw_rate * t + rng.normal(0.0, 0.01, (bs, A, Th))
    vel = np.stack([speed * np.cos(heading), speed * np.sin(heading)], -1)
    vel = vel + rng.normal(0.0, 0.1, vel.shape)
    pos = p0 + vel * t[..., None] + rng.normal(0.0, 0.02, (bs, A, Th, 2))
    shape = np.broadcast_to(rng.uniform(1.0, 4.0, (bs, A, 1, 2)), (bs, A, Th, 2)).copy()
    category = rng.integers(1, 4, (bs, A)).astype(np.int8)
    category[:, 0] = 0
    a_valid = np.ones((bs, A, Th), bool)
    if ragged:
        n_agents = rng.integers(min(4, A), A + 1, bs)
        first = rng.integers(0, Th - 1, (bs, A))          # observed from step `first` onwards
        first[:, 0] = 0
        a_valid = (np.arange(Th)[None, None, :] >= first[..., None])
        a_valid &= rng.uniform(size=(bs, A, Th)) > 0.05   # a few dropped detections
        a_valid[:, 0] = True
        a_valid &= (np.arange(A)[None, :, None] < n_agents[:, None, None])
    z = ~a_valid
    pos[z] = 0
    vel[z] = 0
    shape[z] = 0
    heading[z] = 0
    category[~a_valid.any(-1)] = 0
    agent = {
        "position": pos.astype(f32), "heading": heading.astype(f32), "velocity": vel.astype(f32),
        "shape": shape.astype(f32), "category": category, "valid_mask": a_valid,
    }

    # ---- map polylines: straight-ish segments, 3 sides (centre, left, right)
    c = rng.normal(0.0, 30.0, (bs, Mp, 1, 2))
    ang = rng.normal(0.0, 1.5, (bs, Mp, 1))
    curv = rng.normal(0.0, 0.01, (bs, Mp, 1))
    s = (np.arange(P + 1) - P / 2).reshape(1, 1, P + 1) * rng.uniform(0.5, 2.0, (bs, Mp, 1))
    th = ang + curv * s
    pts = c + np.stack([s * np.cos(th), s * np.sin(th)], -1)          # (bs,Mp,P+1,2)
    centre = pts[:, :, :-1]
    vec = pts[:, :, 1:] - pts[:, :, :-1]
    ori = np.arctan2(vec[..., 1], vec[..., 0])
    nrm = np.stack([-np.sin(ori), np.cos(ori)], -1)
    half = rng.uniform(1.5, 2.0, (bs, Mp, 1, 1))
    point_position = np.stack([centre, centre + half * nrm, centre - half * nrm], 2)  # (bs,Mp,3,P,2)
    point_vector = np.stack([vec, vec, vec], 2)
    point_orientation = np.stack([ori, ori, ori], 2)
    m_valid = np.ones((bs, Mp, P), bool)
    if ragged:
        n_poly = rng.integers(max(1, Mp // 2), Mp + 1, bs)
        n_pts = rng.integers(5, P + 1, (bs, Mp))
        m_valid = np.arange(P)[None, None, :] < n_pts[..., None]
        m_valid &= (np.arange(Mp)[None, :, None] < n_poly[:, None, None])
    poly_valid = m_valid.any(-1)
    mid = P // 2
    polygon_center = np.concatenate([centre[:, :, mid], ori[:, :, mid, None]], -1)
    has_sl = rng.uniform(size=(bs, Mp)) > 0.3
    mp = {
        "point_position": point_position, "point_vector": point_vector,
        "point_orientation": point_orientation,
        "point_side": np.broadcast_to(np.arange(3, dtype=np.int8), (bs, Mp, 3)).copy(),
        "polygon_center": polygon_center,
        "polygon_position": centre[:, :, 0].copy(), "polygon_orientation": ori[:, :, 0].copy(),
        "polygon_type": rng.integers(0, 3, (bs, Mp)).astype(np.int8),
        "polygon_on_route": rng.uniform(size=(bs, Mp)) > 0.5,
        "polygon_tl_status": rng.integers(0, 4, (bs, Mp)).astype(np.int8),
        "polygon_has_speed_limit": has_sl,
        "polygon_speed_limit": np.where(has_sl, rng.uniform(5.0, 25.0, (bs, Mp)), 0.0),
        "polygon_road_block_id": rng.integers(0, 50, (bs, Mp)).astype(np.int32),
        "valid_mask": m_valid,
    }
    for k, v in mp.items():   # pad_sequence zero-fill of padded polygons
        if k != "valid_mask":
            v[~poly_valid] = 0
    for k in ("point_position", "point_vector", "point_orientation", "polygon_center",
              "polygon_position", "polygon_orientation", "polygon_speed_limit"):
        mp[k] = mp[k].astype(f32)

    # ---- reference lines: Pr points at 1 m spacing starting near the centre vehicle
    r0 = rng.normal(0.0, 1.5, (bs, R, 1, 2))
    rang = rng.normal(0.0, 0.3, (bs, R, 1))
    rcurv = rng.normal(0.0, 0.01, (bs, R, 1))
    rs = np.arange(Pr + 1).reshape(1, 1, Pr + 1).astype(np.float64)
    rth = rang + rcurv * rs
    rstep = np.stack([np.cos(rth), np.sin(rth)], -1)
    rpts = r0 + np.cumsum(rstep, 2) - rstep
    r_pos = rpts[:, :, :-1]
    r_vec = rpts[:, :, 1:] - rpts[:, :, :-1]
    r_ori = np.arctan2(r_vec[..., 1], r_vec[..., 0])
    r_valid = np.ones((bs, R, Pr), bool)
    if ragged:
        n_ref = rng.integers(1, R + 1, bs)
        n_rp = rng.integers(min(20, Pr), Pr + 1, (bs, R))
        r_valid = np.arange(Pr)[None, None, :] < n_rp[..., None]
        r_valid &= (np.arange(R)[None, :, None] < n_ref[:, None, None])
    rz = ~r_valid.any(-1)
    r_pos[rz] = 0
    r_vec[rz] = 0
    r_ori[rz] = 0
    ref = {
        "position": r_pos.astype(f32), "vector": r_vec.astype(f32), "orientation": r_ori.astype(f32),
        "valid_mask": r_valid,
        "future_projection": np.zeros((bs, R, 8, 2), f32),
    }

    static = {
        "position": np.zeros((bs, 0, 2), f32), "heading": np.zeros((bs, 0), f32),
        "shape": np.zeros((bs, 0, 2), f32), "category": np.zeros((bs, 0), np.int8),
        "valid_mask": np.zeros((bs, 0), bool),
    }
    cur = np.zeros((bs, 7), f32)
    cur[:, 3] = speed[:, 0, 0]
    cur[:, 4] = rng.normal(0.0, 1.0, bs)
    cur[:, 5] = rng.normal(0.0, 0.2, bs)
    cur[:, 6] = yaw_rate[:, 0, 0]
    return {
        "agent": agent, "map": mp, "reference_line": ref, "static_objects": static,
        "current_state": cur,
        "origin": rng.normal(0.0, 100.0, (bs, 2)).astype(f32),
        "angle": rng.normal(0.0, 1.0, bs).astype(f32),
    }


def group_advantage_numpy(returns: np.ndarray) -> np.ndarray:
    """The literal reference expression (traj_eval/traj_evaluator.py:466-468), one group."""
    mean_return = np.mean(returns)
    std_return = np.std(returns) + 1e-5
    return (returns - mean_return) / std_return


def synth_rl_extras(cfg: PlutoConfig, feats: Dict, seed: int = 2) -> Dict[str, np.ndarray]:
    """Per-algorithm tensors that the collate functions add next to the features."""
    rng = np.random.Generator(np.random.PCG64(seed))
    r_valid = feats["reference_line"]["valid_mask"].any(-1)      # (bs, R)
    bs, R = r_valid.shape
    Mo = cfg.num_modes
    adv = np.zeros((bs, R, Mo), np.float64)
    for b in range(bs):
        nr = int(r_valid[b].sum())
        if nr:
            ret = rng.normal(-5.0, 20.0, nr * Mo)                  # dense-reward scale
            adv[b, :nr] = group_advantage_numpy(ret).reshape(nr, Mo)
    vm = np.broadcast_to(r_valid[..., None], (bs, R, Mo)).copy()
    old = rng.normal(0.0, 1.0, (bs, R, Mo)).astype(np.float32) * vm
    ref = rng.normal(0.0, 1.0, (bs, R, Mo)).astype(np.float32) * vm
    nr = np.maximum(r_valid.sum(-1), 1)
    action_mode = np.stack([rng.integers(0, nr), rng.integers(0, Mo, bs)], -1).astype(np.int64)
    return {
        "group_advantage": adv, "group_advantage_mask": vm,
        "old_group_logits": old, "old_group_logits_mask": vm.copy(),
        "ref_group_logits": ref, "ref_group_logits_mask": vm.copy(),
        "state": rng.normal(0.0, 1.0, (bs, cfg.dim)).astype(np.float32),
        "advantage": rng.normal(0.0, 1.0, bs).astype(np.float32),
        "reward_sum": rng.normal(0.0, 5.0, bs).astype(np.float32),
        "old_log_prob": (-rng.uniform(1.0, 5.0, bs)).astype(np.float32),
        "action_mode": action_mode,
        "return": rng.normal(-5.0, 20.0, bs).astype(np.float32),
    }


WORKLOADS = {
    # BASELINE.json configs -> (model, bs, A, Mp, R, future_steps)
    "cfg1": dict(model="small", bs=1, A=8, Mp=20, R=1, future_steps=40),
    "cfg2": dict(model="medium", bs=64, A=32, Mp=20, R=6, future_steps=80),
    "cfg4": dict(model="medium", bs=256, A=48, Mp=20, R=6, future_steps=80),
}
