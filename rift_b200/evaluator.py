"""Candidate-rollout evaluator on the GPU.

Host-side mirror of the reference's ``TrajEvaluator`` (rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:83-475)
for the part of ``get_action`` that scores the policy's R x M candidate trajectories: same method names, argument meaning and
return values, with the arithmetic in four CUDA launches (csrc/evaluator_kernels.cu) followed by the bit-exact group-advantage
kernel.  What stays on the host is what touches the simulator: reading the neighbours' CARLA state (``actor.get_control()``,
``get_velocity()``, ...; traj_evaluator.py:175-186) and rasterising the drivable area around the vehicle
(traj_evaluator.py:280-298), which the reference does through its map API, shapely and ``cv2.fillPoly``.  The raster comes
from ``off_road_mask_fn(center_state) -> uint8 (H, W)`` (1 = off road); without one every pixel is drivable.

Reference behaviours kept on purpose (each has a golden in tests/golden/evaluator.npz):
  * the PID circular buffers persist across calls per candidate SLOT (the reference never resets its BatchPIDTorch);
  * "shift the trajectory so that it starts at the origin" only zeroes point 0 (in-place subtraction of an aliased view);
  * collision = overlap of axis-aligned envelopes (shapely's STRtree.query without a predicate);
  * the pixel offset is (map_height / 2, map_width / 2) applied to (x, y).
The dense reward is evaluated in float64 on the float32 rollout quantities, which is what the reference's Python scalars give
under the numpy 1.x it pins; under numpy >= 2 scalar promotion the same expressions round to float32 (differences ~1e-6).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import functional as F

NUM_FRAMES = 40          # TrajEvaluator.num_frames
ROLLOUT_LENGTH = 80      # TrackPropagate.rollout_length


class ActorArrays:
    """Plain-array view of the neighbours: what get_other_vehicle_rollout reads from each CARLA actor."""

    def __init__(self, location, heading_deg, speed, control, extent):
        self.location = np.ascontiguousarray(location, np.float64).reshape(-1, 3)
        self.heading_deg = np.ascontiguousarray(heading_deg, np.float64).reshape(-1)
        self.speed = np.ascontiguousarray(speed, np.float64).reshape(-1)
        self.control = np.ascontiguousarray(control, np.float64).reshape(-1, 3)      # steer, throttle, brake
        self.extent = np.ascontiguousarray(extent, np.float64).reshape(-1, 2)        # bounding_box.extent.x, .y
        n = self.location.shape[0]
        if not (self.heading_deg.shape[0] == self.speed.shape[0] == self.control.shape[0] == self.extent.shape[0] == n):
            raise ValueError("ActorArrays: inconsistent neighbour count")

    def __len__(self):
        return self.location.shape[0]

    @classmethod
    def from_actors(cls, actors):
        """traj_evaluator.py:175-186 - the queries the reference makes on every nearby carla.Actor."""
        if isinstance(actors, cls):
            return actors
        actors = list(actors or [])
        ctrl = [a.get_control() for a in actors]
        loc = [a.get_location() for a in actors]
        return cls(
            [[l.x, l.y, l.z] for l in loc],
            [a.get_transform().rotation.yaw for a in actors],
            [a.get_velocity().length() for a in actors],
            [[c.steer, c.throttle, c.brake] for c in ctrl],
            [[a.bounding_box.extent.x, a.bounding_box.extent.y] for a in actors],
        )


def _state6(center_state):
    """{rear-axle x, y, heading, speed, footprint width, length} of the newest history state (traj_evaluator.py:142-143,
    track_propogate.py:625-628, 677-681)."""
    if isinstance(center_state, (list, tuple)) and len(center_state) and not np.isscalar(center_state[0]):
        center_state = center_state[-1]
    if isinstance(center_state, dict):
        return np.asarray([center_state[k] for k in ("x", "y", "heading", "speed", "width", "length")], np.float32)
    if hasattr(center_state, "rear_axle"):
        ra = center_state.rear_axle
        return np.asarray([ra.array[0], ra.array[1], ra.heading, center_state.dynamic_car_state.speed,
                           center_state.car_footprint.width, center_state.car_footprint.length], np.float32)
    return np.asarray(center_state, np.float32).reshape(6)


def _center_pose(center_state):
    """(x, y, heading) of the vehicle CENTRE, the raster's frame (traj_evaluator.py:281-282)."""
    if isinstance(center_state, (list, tuple)) and len(center_state) and not np.isscalar(center_state[0]):
        center_state = center_state[-1]
    if isinstance(center_state, dict):
        return float(center_state.get("cx", center_state["x"])), float(center_state.get("cy", center_state["y"])), \
            float(center_state.get("cheading", center_state["heading"]))
    if hasattr(center_state, "center"):
        c = center_state.center
        return float(c.array[0]), float(c.array[1]), float(c.heading)
    s = np.asarray(center_state, np.float64).reshape(-1)
    return float(s[0]), float(s[1]), float(s[2])


class TrajEvaluator:
    """Drop-in for the reference's TrajEvaluator on ``policy.set_traj_evaluator``: ``get_grpo_advantage`` has the reference's
    signature and return value; ``get_rollout_returns`` exposes the raw returns, ``evaluate`` every intermediate."""

    def __init__(self, dt=0.1, num_frames=NUM_FRAMES, bbox_inflation_ratio=1.1, map_width=400, map_height=400, resolution=0.5,
                 off_road_mask_fn=None, near_lane_change=True, gamma=0.98, device="cuda"):
        if num_frames != NUM_FRAMES:
            raise ValueError("the evaluator kernels are specialised for the reference's 40 scored frames")
        self.dt, self.num_frames, self.bbox_inflation_ratio = float(dt), num_frames, float(bbox_inflation_ratio)
        self.map_width, self.map_height, self.resolution = int(map_width), int(map_height), float(resolution)
        self.off_road_mask_fn, self.near_lane_change, self.gamma = off_road_mask_fn, bool(near_lane_change), float(gamma)
        self.device = torch.device(device)
        self._pid_slots = 0
        self._pid_buf = self._pid_ptr = self._pid_len = None
        _lib.lib()                                           # fail now, not at the first call, when the library is missing

    # ---- persistent PID state ([turn | speed] x slot x 20), grown with zeros like BatchPIDTorch._ensure_batch_size ----
    def _ensure_pid(self, G):
        if G <= self._pid_slots:
            return
        buf = torch.zeros(2, G, 20, dtype=torch.float32, device=self.device)
        ptr_ = torch.zeros(2, G, dtype=torch.int32, device=self.device)
        len_ = torch.zeros(2, G, dtype=torch.int32, device=self.device)
        if self._pid_slots:
            buf[:, :self._pid_slots] = self._pid_buf
            ptr_[:, :self._pid_slots] = self._pid_ptr
            len_[:, :self._pid_slots] = self._pid_len
        self._pid_buf, self._pid_ptr, self._pid_len, self._pid_slots = buf, ptr_, len_, G

    def reset_pid(self):
        """Not something the reference ever does; for tests that need a fresh controller."""
        self._pid_slots = 0
        self._pid_buf = self._pid_ptr = self._pid_len = None

    # ---- the reference's stages --------------------------------------------------------------------------------------
    def _traj(self, trajectories):
        t = torch.as_tensor(trajectories)
        if t.dim() != 4 or t.shape[2] < self.num_frames or t.shape[3] < 4:
            raise ValueError(f"trajectories must be (R, M, T >= {self.num_frames}, >= 4), got {tuple(t.shape)}")
        t = t.detach().to(self.device, torch.float32)
        if t.shape[3] != 6:
            pad = torch.zeros(*t.shape[:3], 6, dtype=torch.float32, device=self.device)
            pad[..., :min(6, t.shape[3])] = t[..., :6]
            t = pad
        return t.contiguous()

    def get_ref_line_info(self, trajectories, ref_line_pos, ref_line_angle, as_numpy=True):
        """traj_evaluator.py:372-420 -> delta_dis, delta_angle (R * M, 40)."""
        t = self._traj(trajectories)
        R, M, T_in, _ = t.shape
        if len(ref_line_pos) != R or len(ref_line_angle) != R:
            raise ValueError("one reference line per trajectory row is required")
        sizes = [int(torch.as_tensor(p).shape[0]) for p in ref_line_pos]
        if R and min(sizes) == 0:
            raise ValueError("empty reference line")
        pos = torch.cat([torch.as_tensor(p).reshape(-1, 2).to(self.device, torch.float32) for p in ref_line_pos]) if R else \
            torch.zeros(0, 2, device=self.device)
        ang = torch.cat([torch.as_tensor(a).reshape(-1).to(self.device, torch.float32) for a in ref_line_angle]) if R else \
            torch.zeros(0, device=self.device)
        off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=self.device)
        dd = torch.empty(R * M, self.num_frames, dtype=torch.float32, device=self.device)
        da = torch.empty_like(dd)
        _lib.check(_lib.lib().rift_b200_eval_ref_line_info(_lib.ptr(t), R, M, T_in, _lib.ptr(pos), _lib.ptr(ang), _lib.ptr(off),
                                                          _lib.ptr(dd), _lib.ptr(da), _lib.stream_ptr()), "eval_ref_line_info")
        return (dd.cpu().numpy(), da.cpu().numpy()) if as_numpy else (dd, da)

    def get_center_rollout(self, trajectories, center_history_states, as_numpy=True):
        """traj_evaluator.py:115-160 + TrackPropagate.propagate -> center, angle, speed, acc, yaw rate, yaw acc, vertices."""
        t = self._traj(trajectories)
        R, M, T_in, _ = t.shape
        G = R * M
        self._ensure_pid(max(G, 1))
        st = _state6(center_history_states)
        dev, f32 = self.device, torch.float32
        center = torch.empty(G, ROLLOUT_LENGTH, 2, dtype=f32, device=dev)
        angle, speed, acc, yr, ya = (torch.empty(G, ROLLOUT_LENGTH, dtype=f32, device=dev) for _ in range(5))
        vert = torch.empty(G, ROLLOUT_LENGTH, 4, 2, dtype=f32, device=dev)
        st_c = (C.c_float * 6)(*[float(x) for x in st])
        _lib.check(_lib.lib().rift_b200_eval_center_rollout(
            _lib.ptr(t), G, T_in, st_c, self.dt, _lib.ptr(self._pid_buf), _lib.ptr(self._pid_ptr), _lib.ptr(self._pid_len),
            self._pid_slots, _lib.ptr(center), _lib.ptr(angle), _lib.ptr(speed), _lib.ptr(acc), _lib.ptr(yr), _lib.ptr(ya),
            _lib.ptr(vert), _lib.stream_ptr()), "eval_center_rollout")
        out = (center, angle, speed, acc, yr, ya, vert)
        return tuple(o.cpu().numpy() for o in out) if as_numpy else out

    def get_other_vehicle_rollout(self, nearby_actors, num_future_frames=NUM_FRAMES, near_lane_change=None, as_numpy=True):
        """traj_evaluator.py:160-239 -> (N, frames, 4, 2) float64 box corners, right-handed."""
        a = ActorArrays.from_actors(nearby_actors)
        N = len(a)
        vert = torch.empty(N, num_future_frames, 4, 2, dtype=torch.float64, device=self.device)
        if N:
            dev = self.device
            loc, hd, sp, ct, ex = (torch.from_numpy(x).to(dev) for x in (a.location, a.heading_deg, a.speed, a.control, a.extent))
            nlc = self.near_lane_change if near_lane_change is None else near_lane_change
            _lib.check(_lib.lib().rift_b200_eval_other_rollout(_lib.ptr(loc), _lib.ptr(hd), _lib.ptr(sp), _lib.ptr(ct), _lib.ptr(ex), N,
                                                              num_future_frames, int(bool(nlc)), self.bbox_inflation_ratio,
                                                              _lib.ptr(vert), _lib.stream_ptr()), "eval_other_rollout")
        return vert.cpu().numpy() if as_numpy else vert

    def evaluate(self, center_history_states, trajectories, ref_line_pos, ref_line_angle, nearby_actors, off_road_mask=None):
        """All stages on the device; returns a dict of DEVICE tensors (returns float64 (G,), collision / off_road (G, 80) bool,
        and every intermediate)."""
        t = self._traj(trajectories)
        R, M = t.shape[:2]
        G = R * M
        dd, da = self.get_ref_line_info(t, ref_line_pos, ref_line_angle, as_numpy=False)
        center, angle, speed, acc, yr, ya, vert = self.get_center_rollout(t, center_history_states, as_numpy=False)
        other = self.get_other_vehicle_rollout(nearby_actors, self.num_frames, as_numpy=False)
        if off_road_mask is None and self.off_road_mask_fn is not None:
            off_road_mask = self.off_road_mask_fn(center_history_states[-1] if isinstance(center_history_states, (list, tuple))
                                                  else center_history_states)
        mask_t, pose = None, None
        H, W = self.map_height, self.map_width
        if off_road_mask is not None:
            mask_t = torch.as_tensor(off_road_mask).to(self.device, torch.uint8).contiguous()
            H, W = mask_t.shape
            if (H, W) != (self.map_height, self.map_width):
                raise ValueError(f"off-road raster must be ({self.map_height}, {self.map_width}), got {(H, W)}")
            cx, cy, ch = _center_pose(center_history_states)
            pose = (C.c_double * 4)(cx, cy, ch, self.resolution)
        ret = torch.empty(G, dtype=torch.float64, device=self.device)
        col = torch.empty(G, ROLLOUT_LENGTH, dtype=torch.uint8, device=self.device)
        off = torch.empty(G, ROLLOUT_LENGTH, dtype=torch.uint8, device=self.device)
        _lib.check(_lib.lib().rift_b200_eval_returns(
            _lib.ptr(dd), _lib.ptr(da), _lib.ptr(speed), _lib.ptr(acc), _lib.ptr(yr), _lib.ptr(ya), _lib.ptr(center), _lib.ptr(vert),
            _lib.ptr(other) if other.shape[0] else None, int(other.shape[0]), self.num_frames, _lib.ptr(mask_t), H, W, pose, G,
            self.gamma, _lib.ptr(ret), _lib.ptr(col), _lib.ptr(off), _lib.stream_ptr()), "eval_returns")
        return {"returns": ret, "collision": col.bool(), "off_road": off.bool(), "delta_dis": dd, "delta_angle": da,
                "center": center, "angle": angle, "speed": speed, "acc": acc, "yaw_rate": yr, "yaw_acc": ya, "vertices": vert,
                "other_vertices": other, "R": R, "M": M}

    def get_rollout_returns(self, center_history_states, trajectories, ref_line_pos, ref_line_angle, nearby_actors):
        """Raw discounted returns (G,) float64 (traj_evaluator.py:333-370)."""
        return self.evaluate(center_history_states, trajectories, ref_line_pos, ref_line_angle, nearby_actors)["returns"].cpu().numpy()

    def get_grpo_advantage(self, center_history_states, trajectories, ref_line_pos, ref_line_angle, nearby_actors):
        """traj_evaluator.py:422-475: {'advantage': (R, M) float64, 'valid_mask': (R, M) bool}; ONE group of R * M candidates."""
        ev = self.evaluate(center_history_states, trajectories, ref_line_pos, ref_line_angle, nearby_actors)
        R, M = ev["R"], ev["M"]
        adv = F.group_advantage(ev["returns"].reshape(1, -1)).reshape(R, M).cpu().numpy()
        return {"advantage": adv, "valid_mask": np.ones_like(adv, dtype=np.bool_)}
