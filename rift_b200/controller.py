"""Trajectory -> (throttle, steer, brake) for a CBV: the control half of ``get_action``.

Host-side restatement of rift/cbv/planning/pluto/controller/pid_controller.py:12-107 (``PID`` with a 20-sample
window, ``PIDController.control_pid``) and of the trajectory bookkeeping around it in
rift/cbv/planning/pluto/pluto.py:196-278 (``_trim_candidates``, ``_global_to_local``).  These are a few dozen scalar
operations per CBV per simulator tick on numpy float64 - they stay on the host exactly like in the reference; what
moved to the GPU is everything that feeds them (the batched policy forward, candidate selection inputs, advantage
normalisation).  ``tests/golden/get_action.npz`` pins them against the reference's own code.
"""
from collections import deque

import numpy as np


class PID:
    """pid_controller.py:12-35 - P on the error, I on the window mean, D on the last difference."""

    def __init__(self, K_P=1.0, K_I=0.0, K_D=0.0, n=20):
        self.K_P, self.K_I, self.K_D = K_P, K_I, K_D
        self.window = deque([0.0] * n, maxlen=n)

    def step(self, error):
        self.window.append(error)
        if len(self.window) >= 2:
            integral = np.mean(self.window)
            derivative = self.window[-1] - self.window[-2]
        else:
            integral, derivative = 0.0, 0.0
        return self.K_P * error + self.K_I * integral + self.K_D * derivative


class PIDController:
    """pid_controller.py:39-107 (constants identical)."""

    def __init__(self, sample_interval=10, max_throttle=1.0, brake_speed=0.4, brake_ratio=1.1, clip_delta=1.0):
        self.sample_interval = int(sample_interval)
        self.turn_controller = PID(K_P=1.25, K_I=0.75, K_D=0.3, n=20)
        self.speed_controller = PID(K_P=5.0, K_I=0.5, K_D=1.0, n=20)
        self.alpha, self.beta = 0.5, 2.5
        self.min_aim_dis, self.max_aim_dis = 5.0, 8.0
        self.max_throttle, self.brake_speed, self.brake_ratio, self.clip_delta = max_throttle, brake_speed, brake_ratio, clip_delta
        self.desired_speed = None
        self.delta_angle = None

    def control_pid(self, local_pos: np.ndarray, speed: float):
        si = self.sample_interval
        pts = local_pos[si - 1::si] if local_pos.shape[0] >= si else local_pos[-1:]
        seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
        desired_speed = seg.mean()
        aim_dist = np.clip(self.alpha * speed + self.beta, self.min_aim_dis, self.max_aim_dis)
        norms = np.linalg.norm(pts[:-1], axis=1)
        aim = pts[np.abs(norms - aim_dist).argmin()]
        brake = desired_speed < self.brake_speed or (speed / desired_speed) > self.brake_ratio
        delta = np.clip(desired_speed - speed, 0.0, self.clip_delta)
        throttle = np.clip(self.speed_controller.step(delta), 0.0, self.max_throttle)
        throttle = throttle if not brake else 0.0
        angle = np.degrees(-np.arctan2(aim[1], aim[0])) / 90
        if speed < 0.01 or brake:       # no integral wind-up while standing or braking
            angle = 0.0
        steer = np.clip(self.turn_controller.step(angle), -1.0, 1.0)
        self.desired_speed, self.delta_angle = desired_speed, angle
        return throttle, steer, brake


def _softmax(x):
    e = np.exp(x - np.max(x))
    return e / e.sum()


def trim_candidates(candidate_trajectories: np.ndarray, probability: np.ndarray, origin, angle: float, topk: int,
                    ref_free_trajectory=None):
    """pluto.py:196-247.  candidate_trajectories (n_ref, n_mode, T, 3) float64 in the CBV frame, probability
    (n_ref, n_mode) logits.  Returns (trajectories in the world frame with the first point duplicated
    (k [+1], T + 1, 3), softmax scores over the top-k [+ 0.25 for the reference-free trajectory], original flat
    indices [-1 for the reference-free one], n_ref, n_mode)."""
    n_ref, n_mode, T, C = candidate_trajectories.shape
    flat_traj = candidate_trajectories.reshape(-1, T, C)
    flat_prob = probability.reshape(-1)
    order = np.argsort(-flat_prob)
    traj = flat_traj[order][:topk]
    score = _softmax(flat_prob[order][:topk])
    orig = np.arange(n_ref * n_mode)[order][:topk]
    if ref_free_trajectory is not None:
        traj = np.concatenate([traj, ref_free_trajectory[None, ...]], axis=0)
        score = np.concatenate([score, [0.25]], axis=0)
        orig = np.concatenate([orig, [-1]], axis=0)
    rot = np.array([[np.cos(angle), np.sin(angle)], [-np.sin(angle), np.cos(angle)]])
    traj[..., :2] = np.matmul(traj[..., :2], rot) + origin
    traj[..., 2] += angle
    traj = np.concatenate([traj[..., 0:1, :], traj], axis=-2)
    return traj, score, orig, n_ref, n_mode


def global_to_local(global_trajectory: np.ndarray, origin, angle: float) -> np.ndarray:
    """pluto.py:262-278 - back into the CBV frame with the first point pinned to the rear axle."""
    delta = origin - global_trajectory[0, :2]
    pos = global_trajectory[..., :2] + delta
    rot = np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]])
    position = np.matmul(pos - origin, rot)
    heading = global_trajectory[..., 2] - angle
    return np.concatenate([position, heading[..., None]], axis=-1)
