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
