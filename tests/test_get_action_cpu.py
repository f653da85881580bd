"""CPU: the host half of get_action (top-k trim, frame changes, PID control) against vectors produced by the reference's
own code (oracle/make_golden.py::get_action_goldens -> tests/golden/get_action.npz) and, where the reference tree is
mounted, against that code directly."""
import numpy as np
import pytest

from rift_b200.controller import PIDController, global_to_local, trim_candidates
from tests.helpers import golden

CASES = [(0, True), (1, True), (2, False), (3, True)]


@pytest.mark.parametrize("case,with_free", CASES)
def test_trim_pid_sequence_matches_reference_golden(case, with_free):
    g = golden("get_action")
    ctrl = PIDController(sample_interval=10)
    for tick in range(4):
        k = f"c{case}t{tick}_"
        x, y, heading, speed = g[k + "state"]
        free = g[k + "free"].astype(np.float64) if with_free else None
        traj, score, orig, n_ref, n_mode = trim_candidates(g[k + "cand"].astype(np.float64), g[k + "prob"], np.array([x, y]),
                                                           float(heading), 10, free)
        assert np.array_equal(orig, g[k + "orig"])                       # index-exact
        assert score.dtype == g[k + "score"].dtype and np.array_equal(score, g[k + "score"])
        assert np.array_equal(traj, g[k + "traj"])
        best = int(score.argmax())
        local = global_to_local(traj[best, 1:], np.array([x, y]), float(heading))
        assert np.array_equal(local, g[k + "local"])
        thr, steer, brake = ctrl.control_pid(local[:, :2], float(speed))
        assert np.array_equal(np.array([float(thr), float(steer), float(bool(brake))]), g[k + "control"]), (case, tick)


@pytest.mark.reference
def test_host_half_matches_reference_code_on_fresh_inputs():
    from scipy.special import softmax
    from oracle import ref_shim
    from oracle.make_golden import _State
    trim = ref_shim.ref_method("rift/cbv/planning/pluto/pluto.py", "PLUTO", "_trim_candidates", {"softmax": softmax})
    g2l = ref_shim.ref_method("rift/cbv/planning/pluto/pluto.py", "PLUTO", "_global_to_local")
    RefPID = ref_shim.pid_controller_cls()

    class _Self:
        _topk = 4
    rng = np.random.default_rng(5)
    ours, ref = PIDController(10), RefPID(sample_interval=10)
    for tick in range(12):
        R = int(rng.integers(1, 7))
        s = np.cumsum(np.full((R, 12, 80), 0.1) * rng.uniform(0, 12, (R, 12, 1)), -1)
        th = rng.normal(0, 0.03, (R, 12, 1)) * s
        cand = np.stack([s * np.cos(th), s * np.sin(th), th], -1)
        prob = rng.normal(0, 2, (R, 12)).astype(np.float32)
        st = _State(rng.normal(0, 30), rng.normal(0, 30), rng.normal(0, 2), rng.uniform(0, 9))
        a = trim(_Self, cand.copy(), prob, st, None)
        b = trim_candidates(cand.copy(), prob, st.rear_axle.array, st.rear_axle.heading, 4, None)
        for x, y in zip(a[:3], b[:3]):
            assert np.array_equal(x, y)
        best = int(a[1].argmax())
        la, lb = g2l(None, a[0][best, 1:], st), global_to_local(b[0][best, 1:], st.rear_axle.array, st.rear_axle.heading)
        assert np.array_equal(la, lb)
        ca = ref.control_pid(la[:, :2], st.dynamic_car_state.center_velocity_2d.magnitude())
        cb = ours.control_pid(lb[:, :2], st.dynamic_car_state.center_velocity_2d.magnitude())
        assert tuple(float(v) for v in ca) == tuple(float(v) for v in cb)
