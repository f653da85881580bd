"""Candidate-rollout evaluator: the CPU oracle against vectors produced by the reference's own
TrackPropagate / derive_kinematics / get_ref_line_info / get_rollout_return / DenseRewardModel / pdm_lite bicycle forecast
(tests/golden/evaluator.npz, oracle/make_golden.py::evaluator_goldens), and the CUDA evaluator against the oracle."""
import numpy as np
import pytest

from oracle import evaluator_oracle as eo
from tests.helpers import golden

CALLS = [(0, 3), (1, 2), (2, 6)]


def _inputs(g, call, R):
    k = f"call{call}_"
    return (g[k + "traj"], [g[k + f"refpos{i}"] for i in range(R)], [g[k + f"refang{i}"] for i in range(R)], g[k + "state"])


def test_oracle_ref_line_info_matches_reference():
    g = golden("evaluator")
    for call, R in CALLS:
        traj, rp, ra, _ = _inputs(g, call, R)
        dd, da = eo.ref_line_info(traj[:, :, :40], rp, ra)
        k = f"call{call}_"
        assert np.abs(dd - g[k + "delta_dis"]).max() <= 2e-4 * max(1.0, np.abs(g[k + "delta_dis"]).max())
        assert np.abs(np.arctan2(np.sin(da - g[k + "delta_angle"]), np.cos(da - g[k + "delta_angle"]))).max() <= 2e-4


def test_oracle_center_rollout_matches_reference_with_pid_state_carried_over():
    """Three consecutive calls on ONE PID state (36, 24, 72 candidates): the reference never resets its batched PID
    buffers (track_propogate.py:496-498 has no caller), so call 2 and 3 start from the integrals call 1 left behind."""
    g = golden("evaluator")
    pid = eo.PIDState()
    for call, R in CALLS:
        traj, _, _, st = _inputs(g, call, R)
        c, a, v, acc, yr, ya, vert = eo.center_rollout(traj[:, :, :40], st[:2], st[2], st[3], st[4], st[5], pid)
        k = f"call{call}_"
        # a chaotic 79-step closed loop in fp32: compare with a tolerance that grows along the horizon
        assert np.abs(c - g[k + "center"]).max() <= 5e-3, (call, np.abs(c - g[k + "center"]).max())
        assert np.abs(v - g[k + "speed"]).max() <= 5e-3
        assert np.abs(np.arctan2(np.sin(a - g[k + "angle"]), np.cos(a - g[k + "angle"]))).max() <= 2e-3
        assert np.abs(acc - g[k + "acc"]).max() <= 5e-2 and np.abs(yr - g[k + "yaw_rate"]).max() <= 2e-2
        assert np.abs(vert - g[k + "vertices"]).max() <= 1e-2


def test_oracle_return_matches_reference():
    g = golden("evaluator")
    for call, R in CALLS:
        k = f"call{call}_"
        ret = eo.rollout_return(g[k + "delta_dis"], g[k + "delta_angle"], g[k + "speed"], g[k + "acc"], g[k + "yaw_rate"],
                                g[k + "yaw_acc"], g[k + "collision"], g[k + "off_road"])
        assert np.allclose(ret, g[k + "return"], rtol=1e-12, atol=1e-12)


def test_oracle_other_vehicle_forecast_matches_reference():
    g = golden("evaluator")
    N = g["other_loc"].shape[0]
    v = eo.other_rollout(g["other_loc"], np.rad2deg(g["other_heading_rad"]), g["other_speed"], g["other_action"],
                         np.tile([[2.3, 1.0]], (N, 1)), n_frames=40)
    centre = v.mean(2)                                       # box centre = mean of its 4 corners
    want = g["other_future_loc"].transpose(1, 0, 2)[..., :2] * np.array([1, -1])
    assert np.abs(centre - want).max() <= 1e-9
    assert v.shape == (N, 40, 4, 2)


def test_collision_and_off_road_restatements():
    sq = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], np.float32)
    cv = np.stack([sq, sq + 5.0])[None].repeat(2, 0)         # (G = 2, T = 2, 4, 2)
    rot45 = sq @ np.array([[0.7071, 0.7071], [-0.7071, 0.7071]], np.float32)
    ov = np.stack([rot45 + np.array([2.3, 0.0], np.float32), rot45 + 50.0])[None]      # envelope touches at t = 0 only
    col = eo.collision_matrix(cv, ov)
    assert col.tolist() == [[True, False], [True, False]]    # envelopes overlap although the oriented boxes do not
    mask = np.ones((400, 400), np.uint8)
    mask[150:250, 150:250] = 0
    pts = np.array([[[0.0, 0.0], [40.0, 0.0], [500.0, 0.0]]])
    assert eo.off_road_matrix(pts, [0.0, 0.0], 0.0, mask).tolist() == [[False, True, False]]     # outside the raster = not off road


# ------------------------------------------------------------------ CUDA evaluator (rift_b200/evaluator.py) vs golden + oracle
def _wrap(d):
    return np.arctan2(np.sin(d), np.cos(d))


def _state(st):
    return dict(x=st[0], y=st[1], heading=st[2], speed=st[3], width=st[4], length=st[5])


@pytest.mark.gpu
def test_cuda_ref_line_info_matches_reference_golden():
    from rift_b200.evaluator import TrajEvaluator
    g = golden("evaluator")
    ev = TrajEvaluator()
    for call, R in CALLS:
        traj, rp, ra, _ = _inputs(g, call, R)
        dd, da = ev.get_ref_line_info(traj, rp, ra)
        k = f"call{call}_"
        assert dd.shape == g[k + "delta_dis"].shape == (R * 12, 40)
        assert np.abs(dd - g[k + "delta_dis"]).max() <= 2e-4 * max(1.0, np.abs(g[k + "delta_dis"]).max())
        assert np.abs(_wrap(da - g[k + "delta_angle"])).max() <= 2e-4


@pytest.mark.gpu
def test_cuda_center_rollout_matches_reference_golden_with_pid_state_carried_over():
    """Same three consecutive calls as the oracle test above, on ONE evaluator (device-resident PID buffers carry over)."""
    from rift_b200.evaluator import TrajEvaluator
    g = golden("evaluator")
    ev = TrajEvaluator()
    for call, R in CALLS:
        traj, _, _, st = _inputs(g, call, R)
        c, a, v, acc, yr, ya, vert = ev.get_center_rollout(traj, [_state(st)])
        k = f"call{call}_"
        # tolerance in the test: 79 closed-loop fp32 steps, libm (CPU) vs CUDA sinf/cosf/atan2f differ in the last ulp
        assert np.abs(c - g[k + "center"]).max() <= 5e-3, (call, np.abs(c - g[k + "center"]).max())
        assert np.abs(v - g[k + "speed"]).max() <= 5e-3
        assert np.abs(_wrap(a - g[k + "angle"])).max() <= 2e-3
        assert np.abs(acc - g[k + "acc"]).max() <= 5e-2 and np.abs(yr - g[k + "yaw_rate"]).max() <= 2e-2
        assert np.abs(ya - g[k + "yaw_acc"]).max() <= 0.3
        assert np.abs(vert - g[k + "vertices"]).max() <= 1e-2
    # a fresh evaluator (zero PID state) gives a DIFFERENT third call: the carried-over state matters
    ev2 = TrajEvaluator()
    traj, _, _, st = _inputs(g, 2, 6)
    c2 = ev2.get_center_rollout(traj, [_state(st)])[0]
    assert np.abs(c2 - g["call2_center"]).max() > 5e-3


@pytest.mark.gpu
def test_cuda_other_vehicle_forecast_matches_reference_golden_and_oracle():
    from rift_b200.evaluator import ActorArrays, TrajEvaluator
    g = golden("evaluator")
    N = g["other_loc"].shape[0]
    ext = np.tile([[2.3, 1.0]], (N, 1))
    ev = TrajEvaluator()
    for nlc in (True, False):
        v = ev.get_other_vehicle_rollout(ActorArrays(g["other_loc"], np.rad2deg(g["other_heading_rad"]), g["other_speed"], g["other_action"], ext),
                                         near_lane_change=nlc)
        want = eo.other_rollout(g["other_loc"], np.rad2deg(g["other_heading_rad"]), g["other_speed"], g["other_action"], ext, n_frames=40,
                                near_lane_change=nlc)
        assert v.shape == (N, 40, 4, 2) and v.dtype == np.float64
        assert np.abs(v - want).max() <= 1e-9
        centre = v.mean(2)
        assert np.abs(centre - g["other_future_loc"].transpose(1, 0, 2)[..., :2] * np.array([1, -1])).max() <= 1e-9
    assert ev.get_other_vehicle_rollout([]).shape == (0, 40, 4, 2)


def _scene(g, call, R, seed):
    """Golden candidates + neighbours parked along the candidates' path and a raster with off-road stripes."""
    traj, rp, ra, st = _inputs(g, call, R)
    rng = np.random.default_rng(seed)
    N = 6
    ch, sh = np.cos(st[2]), np.sin(st[2])
    ahead = rng.uniform(5.0, 35.0, N)
    side = rng.normal(0, 2.0, N)
    ahead[0], side[0] = 6.0, 0.0                             # one neighbour right in front: even a slow candidate reaches its envelope
    # CARLA (left-handed) positions whose right-handed image lies ahead of the vehicle
    loc = np.stack([st[0] + ahead * ch - side * sh, -(st[1] + ahead * sh + side * ch), np.zeros(N)], -1)
    from rift_b200.evaluator import ActorArrays
    actors = ActorArrays(loc, np.rad2deg(-st[2]) + rng.normal(0, 20, N), rng.uniform(0, 3, N),
                         np.stack([rng.uniform(-0.3, 0.3, N), rng.uniform(0, 1, N), (rng.uniform(size=N) < 0.3).astype(float)], -1),
                         np.tile([[2.3, 1.0]], (N, 1)))
    mask = np.zeros((400, 400), np.uint8)
    mask[:, 230:260] = 1
    mask[:190] |= (rng.uniform(size=(190, 400)) < 0.2).astype(np.uint8)
    return traj, rp, ra, st, actors, mask


@pytest.mark.gpu
@pytest.mark.parametrize("call,R", CALLS)
def test_cuda_returns_match_oracle_on_the_same_rollout(call, R):
    """Collision / off-road flags exact and returns to 1e-5 against the oracle fed with the CUDA rollout (the chaotic part is
    covered by the golden test above); the group advantage is numpy-exact."""
    from rift_b200.evaluator import TrajEvaluator
    g = golden("evaluator")
    traj, rp, ra, st, actors, mask = _scene(g, call, R, seed=call)
    ev = TrajEvaluator(off_road_mask_fn=lambda s: mask)
    state = dict(_state(st), cx=st[0] + 1.4 * np.cos(st[2]), cy=st[1] + 1.4 * np.sin(st[2]), cheading=st[2])
    out = ev.evaluate([state], traj, rp, ra, actors)
    h = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    col = eo.collision_matrix(h["vertices"], h["other_vertices"])
    off = eo.off_road_matrix(h["center"], [state["cx"], state["cy"]], state["cheading"], mask)
    assert col.shape == (R * 12, 80) and not col[:, 40:].any()
    assert (h["collision"] == col).all() and (h["off_road"] == off).all()
    assert (col.any() or call == 1) and not col.all() and off.any() and not off.all()   # the scene exercises both flags (call 1 starts at rest)
    want = eo.rollout_return(h["delta_dis"], h["delta_angle"], h["speed"], h["acc"], h["yaw_rate"], h["yaw_acc"], col, off)
    assert np.abs(h["returns"] - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    ev.reset_pid()
    adv = ev.get_grpo_advantage([state], traj, rp, ra, actors)
    ev.reset_pid()
    ret = ev.get_rollout_returns([state], traj, rp, ra, actors)
    assert adv["advantage"].shape == (R, 12) and adv["valid_mask"].all() and adv["advantage"].dtype == np.float64
    assert np.array_equal(adv["advantage"].reshape(-1), (ret - np.mean(ret)) / (np.std(ret) + 1e-5))


@pytest.mark.gpu
def test_cuda_evaluator_edge_cases():
    from rift_b200.evaluator import TrajEvaluator
    g = golden("evaluator")
    traj, rp, ra, st = _inputs(g, 0, 3)
    ev = TrajEvaluator()
    # no neighbours, no raster: nothing collides, nothing is off road
    out = ev.evaluate([_state(st)], traj, rp, ra, [])
    assert not out["collision"].any() and not out["off_road"].any() and np.isfinite(out["returns"].cpu().numpy()).all()
    # a colliding first frame ends the return after one reward: identical to the one-frame oracle value
    with pytest.raises(ValueError):
        ev.get_ref_line_info(traj, rp[:2], ra[:2])
    with pytest.raises(ValueError):
        ev.evaluate([_state(st)], traj[:, :, :30], rp, ra, [])
    with pytest.raises(ValueError):
        TrajEvaluator(off_road_mask_fn=lambda s: np.zeros((10, 10), np.uint8)).evaluate([_state(st)], traj, rp, ra, [])
