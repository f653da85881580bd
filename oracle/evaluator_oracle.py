"""CPU restatement of the candidate-rollout evaluator (TEST INFRASTRUCTURE - never imported by rift_b200/).

Follows rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:115-475 and track_propogate.py:160-780 minus the
CARLA queries (actor poses / controls / extents, map polygons arrive as arrays):

    ref_line_info      traj_evaluator.py:372-420   nearest reference-line point, signed lateral offset, wrapped heading error
    center_rollout     traj_evaluator.py:115-158 + track_propogate.py:616-699 (79 PID + kinematic-bicycle steps per candidate,
                       BatchPIDController :403-498, BatchKinematicBicycleModel.forward :213-279, derive_kinematics :500-596
                       with Savitzky-Golay(5, 2) smoothing :105-157, oriented-box vertices :16-74)
    other_rollout      traj_evaluator.py:160-239 (constant-control bicycle forecast, speed-dependent box inflation x 1.1,
                       left -> right-handed flip; rift/ego/pdm_lite/kinematic_bicycle_model.py:33-61, config.py)
    collision_matrix   traj_evaluator.py:241-271: shapely STRtree.query(polygon) without predicate = ENVELOPE (axis-aligned
                       bounding box) intersection of the two oriented boxes - restated as such (shapely is not installed
                       here: this one step is unpinned)
    off_road_matrix    traj_evaluator.py:273-322 minus the raster build: nearest-pixel lookup in a given 400 x 400 mask
    rollout_return     traj_evaluator.py:333-370 + gym_carla/reward/reward_model.py:34-50, gamma = 0.98, stop after the first
                       colliding step (inclusive)
    advantage          traj_evaluator.py:466-470

Pinned by tests/test_evaluator.py against the reference's own TrackPropagate / derive_kinematics / get_ref_line_info /
DenseRewardModel / get_rollout_return executed through oracle/ref_shim.py (tests/golden/evaluator.npz).
Reference quirks reproduced on purpose: (i) the "shift to the first point" of get_center_rollout only zeroes the first point
(in-place subtraction of an aliased view, see center_rollout); (ii) the batched PID keeps its circular buffers across calls (its reset() has no caller, track_propogate.py:496-498):
`PIDState` carries that state explicitly so the leak is reproducible.
"""
import numpy as np

F32 = np.float32
SG_KERNEL = (np.array([-3.0, 12.0, 17.0, 12.0, -3.0]) / 35.0).astype(F32)      # Savitzky-Golay window 5, order 2, 0th derivative

# BatchKinematicBicycleModel constants (track_propogate.py:181-199)
LF, LR, STEER_GAIN = F32(-0.090769015), F32(1.4178275), F32(0.36848336)
BRAKE_POLY = np.array([9.31711370e-03, 8.20967431e-02, -2.83832427e-03, 5.06587474e-05, -4.90357228e-07, 2.44419284e-09,
                       -4.91381935e-12], F32)
THROTTLE_POLY = np.array([9.63873001e-01, 4.37535692e-04, -3.80192912e-01, 1.74950069e+00, 9.16787414e-02, -7.05461530e-02,
                          -1.05996152e-03, 6.71079346e-04], F32)
THROTTLE_THRESHOLD = F32(0.3)


class PIDState:
    """Circular buffers of the two BatchPIDTorch controllers (turn, speed), n = 20, one row per candidate slot."""

    def __init__(self, n_slots=0, n=20):
        self.n = n
        self.buf = np.zeros((2, n_slots, n), F32)
        self.ptr = np.zeros((2, n_slots), np.int64)
        self.len = np.zeros((2, n_slots), np.int64)

    def ensure(self, B):
        if self.buf.shape[1] < B:
            extra = B - self.buf.shape[1]
            self.buf = np.concatenate([self.buf, np.zeros((2, extra, self.n), F32)], 1)
            self.ptr = np.concatenate([self.ptr, np.zeros((2, extra), np.int64)], 1)
            self.len = np.concatenate([self.len, np.zeros((2, extra), np.int64)], 1)

    def step(self, which, error, KP, KI, KD):
        """BatchPIDTorch.step (track_propogate.py:369-400)."""
        B = error.shape[0]
        self.ensure(B)
        ar = np.arange(B)
        idx = self.ptr[which, :B]
        prev = self.buf[which, ar, idx].copy()
        self.buf[which, ar, idx] = error
        self.ptr[which, :B] = (idx + 1) % self.n
        self.len[which, :B] = np.minimum(self.len[which, :B] + 1, self.n)
        integral = self.buf[which, :B].sum(1, dtype=F32) / np.maximum(self.len[which, :B], 1).astype(F32)
        return (F32(KP) * error + F32(KI) * integral + F32(KD) * (error - prev)).astype(F32)


def ref_line_info(trajectories, ref_pos, ref_angle):
    """trajectories (R, M, Ts, >=4) float32; ref_pos / ref_angle: R arrays (n_r, 2) / (n_r,).  -> (delta_dis, delta_angle) (R*M, Ts)."""
    R, M, Ts, _ = trajectories.shape
    dd = np.zeros((R, M, Ts), F32)
    da = np.zeros((R, M, Ts), F32)
    for r in range(R):
        p, a = np.asarray(ref_pos[r], F32), np.asarray(ref_angle[r], F32)
        cand = trajectories[r, ..., :2].astype(F32)
        ang = np.arctan2(trajectories[r, ..., 3], trajectories[r, ..., 2]).astype(F32)
        dist = np.linalg.norm(cand[:, :, None, :] - p[None, None], axis=-1)
        ci = dist.argmin(-1)
        ca = a[ci]
        diff = ang - ca
        da[r] = np.arctan2(np.sin(diff), np.cos(diff))
        rel = cand - p[ci]
        dd[r] = -(rel[..., 0] * np.sin(ca) - rel[..., 1] * np.cos(ca))
    return dd.reshape(-1, Ts), da.reshape(-1, Ts)


def _pid_control(local, speed, pid: PIDState):
    """BatchPIDController.control_pid (track_propogate.py:435-491); local (G, 30, 2)."""
    pts = local[:, 9::10]                                  # sample_interval 10 -> 3 way-points
    seg = np.linalg.norm(pts[:, 1:] - pts[:, :-1], axis=2).astype(F32)
    desired = seg.mean(1, dtype=F32)
    aim_dist = np.clip(F32(0.5) * speed + F32(2.5), F32(5.0), F32(8.0))
    norms = np.linalg.norm(pts[:, :-1], axis=2).astype(F32)
    idx = np.abs(norms - aim_dist[:, None]).argmin(1)
    aim = pts[np.arange(pts.shape[0]), idx]
    brake = (desired < F32(0.4)) | ((speed / np.maximum(desired, F32(1e-4))) > F32(1.1))
    delta = np.clip(desired - speed, F32(0.0), F32(0.25))
    throttle = np.clip(pid.step(1, delta.astype(F32), 5.0, 0.5, 1.0), F32(0.0), F32(0.75)) * (~brake)
    angle = (np.rad2deg(np.arctan2(aim[:, 1], aim[:, 0])) / F32(90.0)).astype(F32)
    angle = np.where((speed < F32(0.01)) | brake, F32(0.0), angle).astype(F32)
    steer = np.clip(pid.step(0, angle, 1.25, 0.75, 0.3), F32(-1.0), F32(1.0))
    return throttle.astype(F32), steer.astype(F32), brake


def _bicycle(loc, heading, speed, throttle, steer, brake, dt):
    """BatchKinematicBicycleModel.forward (track_propogate.py:213-279)."""
    wheel = STEER_GAIN * steer
    slip = np.arctan((LR / (LF + LR)) * np.tan(wheel)).astype(F32)
    nloc = loc.copy()
    nloc[:, 0] += speed * np.cos(heading + slip) * dt
    nloc[:, 1] += speed * np.sin(heading + slip) * dt
    nhead = (heading + (speed / LR) * np.sin(slip) * dt).astype(F32)
    kph = (speed * F32(3.6)).astype(F32)
    vp = np.stack([kph ** i for i in range(1, 8)], 1).astype(F32)
    nb = vp @ BRAKE_POLY
    v2, t, t2 = kph * kph, throttle, throttle * throttle
    feats = np.stack([kph, v2, t, t2, kph * t, kph * t2, v2 * t, v2 * t2], 1).astype(F32)
    nt = feats @ THROTTLE_POLY
    nk = np.where(brake, nb, kph)
    nk = np.where((~brake) & (throttle >= THROTTLE_THRESHOLD), nt, nk)
    return nloc.astype(F32), nhead, np.maximum(nk / F32(3.6), F32(0.0)).astype(F32)


def _sg(x):
    pad = np.pad(x, ((0, 0), (2, 2)), mode="reflect")
    return sum(SG_KERNEL[4 - k] * pad[:, k:k + x.shape[1]] for k in range(5)).astype(F32)     # conv1d = correlation with the flipped-back kernel (symmetric)


def _cdiff(x, dt):
    mid = (x[:, 2:] - x[:, :-2]) / F32(2.0 * dt)
    return np.concatenate([(x[:, 1:2] - x[:, :1]) / F32(dt), mid, (x[:, -1:] - x[:, -2:-1]) / F32(dt)], 1).astype(F32)


def _wrap(d):
    return np.arctan2(np.sin(d), np.cos(d)).astype(F32)


def derive_kinematics(heading, speed, dt=0.1):
    """track_propogate.py:500-596 with speed given, smooth window 5 / order 2."""
    sp = _sg(speed.astype(F32))
    acc = _cdiff(sp, dt)
    hp = _sg(heading.astype(F32))
    yr = np.zeros_like(hp)
    yr[:, 1:-1] = _wrap(hp[:, 2:] - hp[:, :-2]) / F32(2 * dt)
    yr[:, 0] = _wrap(hp[:, 1] - hp[:, 0]) / F32(dt)
    yr[:, -1] = _wrap(hp[:, -1] - hp[:, -2]) / F32(dt)
    return sp, acc, yr.astype(F32), _cdiff(yr.astype(F32), dt)


def box_vertices(center, heading, width, length):
    """compute_agents_vertices_torch (track_propogate.py:16-74): FL, RL, RR, FR; center (..., T, 2)."""
    hw, hl = F32(0.5 * width), F32(0.5 * length)
    dx = np.stack([hl, -hl, -hl, hl], -1) if np.ndim(hl) else np.array([hl, -hl, -hl, hl], F32)
    dy = np.stack([hw, hw, -hw, -hw], -1) if np.ndim(hw) else np.array([hw, hw, -hw, -hw], F32)
    c, s = np.cos(heading)[..., None], np.sin(heading)[..., None]
    return (np.stack([dx * c - dy * s, dx * s + dy * c], -1) + center[..., None, :]).astype(F32)


def center_rollout(trajectories, origin, heading0, speed0, width, length, pid: PIDState, dt=0.1, rollout_length=80):
    """get_center_rollout + TrackPropagate.propagate.  trajectories (R, M, Ts, >=4) raw model output (first Ts = 40 frames).
    -> rollout centre (G, 80, 2), angle, speed, acc, yaw rate, yaw acc (G, 80), vertices (G, 80, 4, 2)."""
    R, M, T, _ = trajectories.shape
    G = R * M
    head = np.arctan2(trajectories[..., 3], trajectories[..., 2]).astype(F32).reshape(G, T)
    pos = trajectories[..., :2].astype(F32).reshape(G, T, 2).copy()
    # traj_evaluator.py:137-138 "force the first point to be (0, 0)": `out[:, :, :2] -= out[:, 0, :2].unsqueeze(1)` subtracts a VIEW
    # of the tensor it updates in place; torch processes t = 0 first, after which the subtrahend is zero - the net effect
    # is that ONLY the first point moves to the origin and all later points stay where they were.  Reproduced as is.
    pos[:, 0] = 0
    ch, sh = F32(np.cos(F32(heading0))), F32(np.sin(F32(heading0)))
    rot = np.array([[ch, sh], [-sh, ch]], F32)
    ref_pos = (pos @ rot + np.asarray(origin, F32)).astype(F32)
    ref_head = (head + F32(heading0)).astype(F32)
    c, a, v = [ref_pos[:, 0].copy()], [ref_head[:, 0].copy()], [np.full(G, speed0, F32)]
    closest = np.zeros(G, np.int64)
    ar = np.arange(G)
    for _ in range(rollout_length - 1):
        idx = closest[:, None] + np.arange(30)[None]
        gathered = ref_pos[ar[:, None], np.minimum(idx, T - 1)]
        gathered = np.where((idx >= T)[..., None], ref_pos[:, -1:], gathered)
        loc = gathered - c[-1][:, None]
        cc, ss = np.cos(a[-1]), np.sin(a[-1])
        local = np.stack([loc[..., 0] * cc[:, None] + loc[..., 1] * ss[:, None],
                          loc[..., 0] * (-ss[:, None]) + loc[..., 1] * cc[:, None]], -1).astype(F32)
        th, st, br = _pid_control(local, v[-1], pid)
        nl, nh, ns = _bicycle(c[-1], a[-1], v[-1], th, st, br, F32(dt))
        c.append(nl); a.append(nh); v.append(ns)
        closest = ((ref_pos - nl[:, None]) ** 2).sum(-1).argmin(1)
    center, angle, speed = np.stack(c, 1), np.stack(a, 1), np.stack(v, 1)
    sp, acc, yr, ya = derive_kinematics(angle, speed, dt)
    return center, angle, sp, acc, yr, ya, box_vertices(center, angle, width, length)


def other_rollout(locations, headings_deg, speeds, controls, extents, n_frames=40, near_lane_change=True, inflation=1.1, cfg=None):
    """get_other_vehicle_rollout with the CARLA reads replaced by arrays: locations (N, 3), yaw in degrees, speed (N,),
    controls (N, 3) = steer, throttle, brake, extents (N, 2) = bounding-box extent x, y.  float64 like the reference."""
    cfg = cfg or OTHER_CFG
    N = locations.shape[0]
    if N == 0:
        return np.zeros((0, n_frames, 4, 2), np.float32)
    loc, hd, sp = locations.astype(np.float64).copy(), np.deg2rad(headings_deg.astype(np.float64)), speeds.astype(np.float64).copy()
    fl, fh, fv = np.empty((n_frames, N, 3)), np.empty((n_frames, N)), np.empty((n_frames, N))
    steer, throttle, brake = controls[:, 0], controls[:, 1], controls[:, 2].astype(np.uint8)
    for i in range(n_frames):
        wheel = cfg["steering_gain"] * steer
        slip = np.arctan(cfg["rear_wheel_base"] / (cfg["front_wheel_base"] + cfg["rear_wheel_base"]) * np.tan(wheel))
        nx = loc[:, 0] + sp * np.cos(hd + slip) * cfg["time_step"]
        ny = loc[:, 1] + sp * np.sin(hd + slip) * cfg["time_step"]
        hd = hd + sp / cfg["rear_wheel_base"] * np.sin(slip) * cfg["time_step"]
        sp = np.maximum(0.0, sp + cfg["time_step"] * np.where(brake, cfg["brake_acceleration"], throttle * cfg["throttle_acceleration"]))
        loc = np.column_stack([nx, ny, loc[:, 2]])
        fl[i], fh[i], fv[i] = loc, hd, sp
    shape = np.empty((n_frames, N, 2))
    s = cfg["high_speed_min_extent_x_other_vehicle_lane_change"] if near_lane_change else cfg["high_speed_min_extent_x_other_vehicle"]
    for i in range(n_frames):
        slow = fv[i] < cfg["extent_other_vehicles_bbs_speed_threshold"]
        ex = extents[:, 0] * np.where(slow, cfg["slow_speed_extent_factor_ego"],
                                      max(s, cfg["high_speed_min_extent_x_other_vehicle"] * float(i) / float(n_frames)))
        ey = extents[:, 1] * np.where(slow, cfg["slow_speed_extent_factor_ego"],
                                      max(cfg["high_speed_min_extent_y_other_vehicle"],
                                          cfg["high_speed_extent_y_factor_other_vehicle"] * float(i) / float(n_frames)))
        shape[i] = np.stack([ey * inflation * 2, ex * inflation * 2], -1)          # [width, length]
    center = fl.transpose(1, 0, 2)[..., :2] * np.array([1, -1])
    ang = -fh.transpose(1, 0)
    sh = shape.transpose(1, 0, 2)
    hw, hl = sh[..., 0] / 2, sh[..., 1] / 2
    cos, sin = np.cos(ang)[..., None], np.sin(ang)[..., None]
    ol = np.stack([hl, -hl, -hl, hl], -1)
    ow = np.stack([hw, hw, -hw, -hw], -1)
    # vertices = [offset_length, offset_width] @ [[cos, sin], [-sin, cos]] + center   (traj_evaluator.py:62-76)
    return np.stack([ol * cos - ow * sin, ol * sin + ow * cos], -1) + center[..., None, :]


# rift/ego/pdm_lite/config.py:186-199,336-347
OTHER_CFG = dict(time_step=0.1, steering_gain=0.36848336, front_wheel_base=-0.090769015, rear_wheel_base=1.4178275,
                 brake_acceleration=-4.952399, throttle_acceleration=0.5633837, slow_speed_extent_factor_ego=1.0,
                 extent_other_vehicles_bbs_speed_threshold=1.0, high_speed_min_extent_x_other_vehicle=1.2,
                 high_speed_min_extent_x_other_vehicle_lane_change=2.0, high_speed_min_extent_y_other_vehicle=1.0,
                 high_speed_extent_y_factor_other_vehicle=1.3)


def collision_matrix(center_vertices, other_vertices):
    """Envelope (AABB) overlap of the candidate's box with any other vehicle's box, per time step (see module docstring)."""
    G, Ts = center_vertices.shape[:2]
    N = other_vertices.shape[0]
    out = np.zeros((G, Ts), bool)
    if N == 0:
        return out
    T = min(Ts, other_vertices.shape[1])
    cmin, cmax = center_vertices[:, :T].min(2), center_vertices[:, :T].max(2)           # (G, T, 2)
    omin, omax = other_vertices[:, :T].min(2), other_vertices[:, :T].max(2)             # (N, T, 2)
    hit = (cmin[:, None] <= omax[None]) & (omin[None] <= cmax[:, None])                 # (G, N, T, 2)
    out[:, :T] = hit.all(-1).any(1)
    return out


def off_road_matrix(rollout_center, origin, angle, mask, resolution=0.5):
    """traj_evaluator.py:300-322: nearest-pixel lookup of the rollout centres in the (H, W) off-road mask (1 = off road)."""
    H, W = mask.shape
    rot = np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]], np.float64)
    G, T, _ = rollout_center.shape
    pts = np.matmul(rollout_center.reshape(-1, 2) - np.asarray(origin, np.float64), rot)
    pix = pts / np.array([resolution, -resolution], np.float32) + np.array([H / 2, W / 2], np.float32)
    pi = np.round(pix).astype(int)
    valid = (pi[:, 0] >= 0) & (pi[:, 0] < W) & (pi[:, 1] >= 0) & (pi[:, 1] < H)
    flags = np.zeros(pts.shape[0], bool)
    flags[valid] = mask[pi[valid, 1], pi[valid, 0]] == 1
    return flags.reshape(G, T)


def dense_reward(delta_dis, delta_angle, speed, acc, ang_speed, ang_acc, collision, offroad):
    """DenseRewardModel.get_reward (reward_model.py:34-50), scalar arguments."""
    cosd = np.cos(delta_angle)
    r_col = -(20.0 + abs(speed)) * collision
    r_off = -5.0 * offroad
    r_comf = -0.8 * (int(abs(acc) > 4) + int(abs(ang_acc) > 4))
    r_align = 0.5 * (min(cosd, 0) + 0.05 * min(cosd * speed, 0) + 0.25 * (1 - abs(delta_angle) / (np.pi / 2)))
    r_center = -0.6 * int(cosd > 0.5) * (abs(delta_dis - 0.0) - 0.05 / np.exp(abs(delta_dis - 0.0) - 0.5))
    r_vel = 0.1 * max(cosd, 0) * int(3 < abs(speed) < 20) * abs(speed)
    r_time = -0.1 * int(abs(speed) > 0 or abs(acc) > 0)
    return r_col + r_off + r_comf + r_align + r_center + r_vel + r_time


def rollout_return(delta_dis, delta_angle, speed, acc, ang_vel, ang_acc, collision, off_road, gamma=0.98):
    """traj_evaluator.py:333-370: the loop runs over delta_angle's Ts (= 40) frames and stops after the first collision."""
    G, Ts = delta_angle.shape
    out = np.zeros(G, np.float64)
    for i in range(G):
        for j in range(Ts):
            out[i] += dense_reward(abs(delta_dis[i, j]), abs(delta_angle[i, j]), speed[i, j], acc[i, j], ang_vel[i, j], ang_acc[i, j],
                                   int(collision[i, j]), int(off_road[i, j])) * gamma ** j
            if collision[i, j]:
                break
    return out


def grpo_advantage(returns):
    return (returns - np.mean(returns)) / (np.std(returns) + 1e-5)
