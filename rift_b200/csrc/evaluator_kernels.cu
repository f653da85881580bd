// Candidate-rollout evaluator on the GPU: everything TrajEvaluator.get_grpo_advantage does between the
// policy's raw trajectories and the group-relative advantage, minus the CARLA queries (actor states and the drivable-area
// raster arrive as arrays).  Reference: rift/cbv/planning/fine_tuner/rlft/traj_eval/traj_evaluator.py:115-475,
// track_propogate.py:160-780, rift/gym_carla/reward/reward_model.py:34-50, rift/ego/pdm_lite/kinematic_bicycle_model.py:33-61.
//
//   eval_ref_line_kernel      one thread per (candidate, frame): nearest reference-line point, signed lateral offset, wrapped
//                             heading error                                                   (traj_evaluator.py:372-420)
//   eval_center_rollout_kernel one thread per candidate: 79 sequential steps of {3 way-points ahead of the closest reference
//                             index in the vehicle frame -> PID (two 20-sample circular buffers that PERSIST across calls,
//                             like the reference's never-reset BatchPIDTorch) -> kinematic bicycle -> arg-min over the
//                             reference points}, then Savitzky-Golay(5, 2) smoothing, central differences, box corners
//                                                                                             (track_propogate.py:403-699)
//   eval_other_rollout_kernel one thread per neighbour, fp64 like the reference's numpy: constant-control bicycle forecast,
//                             speed-dependent box inflation, left -> right-handed flip        (traj_evaluator.py:160-239)
//   eval_return_kernel        one thread per candidate: envelope (AABB) overlap with every neighbour per frame (what shapely's
//                             STRtree.query without predicate tests), nearest-pixel off-road lookup, dense reward,
//                             gamma-discounted sum that stops after the first colliding frame  (traj_evaluator.py:241-370)
// The reference walks G x 40 Python iterations for the return and G x 40 shapely queries per call; here a call is four
// small launches.  The group-relative normalisation is rift_b200_group_advantage (bit-exact, rl_kernels.cu).
#include "common.cuh"
#include "../../include/rift_b200.h"

namespace rift {

constexpr int EV_TS = 40;            // reference frames of a candidate (TrajEvaluator.num_frames)
constexpr int EV_TR = 80;            // rollout length (TrackPropagate.rollout_length)
constexpr int EV_PID_N = 20;

__global__ void __launch_bounds__(128)
eval_ref_line_kernel(const float* __restrict__ traj, int R, int M, int T_in, const float* __restrict__ ref_pos,
                     const float* __restrict__ ref_ang, const int* __restrict__ ref_off, float* __restrict__ delta_dis,
                     float* __restrict__ delta_angle) {
    pdl_grid_sync();
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= R * M * EV_TS) return;
    const int g = i / EV_TS, t = i - g * EV_TS, r = g / M;
    const float* p = traj + ((long long)g * T_in + t) * 6;
    const float x = p[0], y = p[1], ang = atan2f(p[3], p[2]);
    const int o0 = ref_off[r], o1 = ref_off[r + 1];
    float best = INFINITY;
    int bi = o0;
    for (int k = o0; k < o1; ++k) {
        const float dx = x - ref_pos[2 * k], dy = y - ref_pos[2 * k + 1];
        const float d = sqrtf(dx * dx + dy * dy);
        if (d < best) { best = d; bi = k; }
    }
    const float ca = ref_ang[bi];
    const float diff = ang - ca;
    delta_angle[i] = atan2f(sinf(diff), cosf(diff));
    const float rx = x - ref_pos[2 * bi], ry = y - ref_pos[2 * bi + 1];
    delta_dis[i] = -(rx * sinf(ca) - ry * cosf(ca));
}

struct EvalState { float ox, oy, heading, speed, width, length; };

// BatchPIDTorch.step on this candidate's circular buffer (track_propogate.py:369-400): the derivative term uses the value
// that the write OVERWRITES (the error 20 steps ago), exactly like the reference
__device__ __forceinline__ float pid_step(float* buf, int& ptr, int& len, float error, float KP, float KI, float KD) {
    const float prev = buf[ptr];
    buf[ptr] = error;
    ptr = (ptr + 1) % EV_PID_N;
    len = min(len + 1, EV_PID_N);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < EV_PID_N; ++k) s += buf[k];
    return KP * error + KI * (s / (float)max(len, 1)) + KD * (error - prev);
}

__global__ void __launch_bounds__(64)
eval_center_rollout_kernel(const float* __restrict__ traj, int G, int T_in, EvalState st, float dt, float* __restrict__ pid_buf,
                           int* __restrict__ pid_ptr, int* __restrict__ pid_len, int pid_slots, float* __restrict__ center,
                           float* __restrict__ angle, float* __restrict__ speed, float* __restrict__ acc, float* __restrict__ yaw_rate,
                           float* __restrict__ yaw_acc, float* __restrict__ vertices) {
    pdl_grid_sync();
    const int g = blockIdx.x * 64 + threadIdx.x;
    if (g >= G) return;
    // reference trajectory in the world frame (traj_evaluator.py:126-152).  The reference's "shift so that point 0 is the
    // origin" subtracts an aliased view in place, which only zeroes point 0 - reproduced.
    float rx[EV_TS], ry[EV_TS];
    const float ch = cosf(st.heading), sh = sinf(st.heading);
    float head0 = 0.f;
    for (int t = 0; t < EV_TS; ++t) {
        const float* p = traj + ((long long)g * T_in + t) * 6;
        const float px = t == 0 ? 0.f : p[0], py = t == 0 ? 0.f : p[1];
        rx[t] = px * ch - py * sh + st.ox;
        ry[t] = px * sh + py * ch + st.oy;
        if (t == 0) head0 = atan2f(p[3], p[2]) + st.heading;
    }
    // PID state of this candidate slot: [turn | speed] x [slots] x 20
    float tb[EV_PID_N], sb[EV_PID_N];
    for (int k = 0; k < EV_PID_N; ++k) { tb[k] = pid_buf[(long long)g * EV_PID_N + k]; sb[k] = pid_buf[((long long)pid_slots + g) * EV_PID_N + k]; }
    int tptr = pid_ptr[g], sptr = pid_ptr[pid_slots + g], tlen = pid_len[g], slen = pid_len[pid_slots + g];
    float cx = rx[0], cy = ry[0], ca = head0, v = st.speed;
    float* cg = center + (long long)g * EV_TR * 2;
    float* ag = angle + (long long)g * EV_TR;
    float* vg = speed + (long long)g * EV_TR;
    cg[0] = cx; cg[1] = cy; ag[0] = ca; vg[0] = v;
    int closest = 0;
    const float LF = -0.090769015f, LR = 1.4178275f, SG = 0.36848336f;
    for (int step = 1; step < EV_TR; ++step) {
        // three way-points (every 10th of the 30 ahead of the closest index; indices past the end repeat the last point)
        float lx[3], ly[3];
        const float cc = cosf(ca), ss = sinf(ca);
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const int idx = min(closest + 9 + 10 * w, EV_TS - 1);
            const float dx = rx[idx] - cx, dy = ry[idx] - cy;
            lx[w] = dx * cc + dy * ss;
            ly[w] = dx * (-ss) + dy * cc;
        }
        const float s0 = sqrtf((lx[1] - lx[0]) * (lx[1] - lx[0]) + (ly[1] - ly[0]) * (ly[1] - ly[0]));
        const float s1 = sqrtf((lx[2] - lx[1]) * (lx[2] - lx[1]) + (ly[2] - ly[1]) * (ly[2] - ly[1]));
        const float desired = (s0 + s1) / 2.f;
        const float aim_dist = fminf(fmaxf(0.5f * v + 2.5f, 5.f), 8.f);
        const float n0 = sqrtf(lx[0] * lx[0] + ly[0] * ly[0]), n1 = sqrtf(lx[1] * lx[1] + ly[1] * ly[1]);
        const int ai = fabsf(n1 - aim_dist) < fabsf(n0 - aim_dist) ? 1 : 0;
        const bool brake = (desired < 0.4f) || ((v / fmaxf(desired, 1e-4f)) > 1.1f);
        const float delta = fminf(fmaxf(desired - v, 0.f), 0.25f);
        float throttle = fminf(fmaxf(pid_step(sb, sptr, slen, delta, 5.0f, 0.5f, 1.0f), 0.f), 0.75f);
        if (brake) throttle = 0.f;
        float ang = (atan2f(ly[ai], lx[ai]) * 57.29577951308232f) / 90.f;
        if (v < 0.01f || brake) ang = 0.f;
        const float steer = fminf(fmaxf(pid_step(tb, tptr, tlen, ang, 1.25f, 0.75f, 0.3f), -1.f), 1.f);
        // kinematic bicycle (track_propogate.py:213-279)
        const float slip = atanf((LR / (LF + LR)) * tanf(SG * steer));
        const float nx = cx + v * cosf(ca + slip) * dt, ny = cy + v * sinf(ca + slip) * dt;
        const float na = ca + (v / LR) * sinf(slip) * dt;
        const float k1 = v * 3.6f;
        float nk = k1;
        if (brake) {
            const float c[7] = {9.31711370e-03f, 8.20967431e-02f, -2.83832427e-03f, 5.06587474e-05f, -4.90357228e-07f,
                                2.44419284e-09f, -4.91381935e-12f};
            float pw = k1, a = 0.f;
#pragma unroll
            for (int q = 0; q < 7; ++q) { a += pw * c[q]; pw *= k1; }
            nk = a;
        } else if (throttle >= 0.3f) {
            const float v2 = k1 * k1, t = throttle, t2 = t * t;
            nk = k1 * 9.63873001e-01f + v2 * 4.37535692e-04f + t * -3.80192912e-01f + t2 * 1.74950069e+00f + (k1 * t) * 9.16787414e-02f +
                 (k1 * t2) * -7.05461530e-02f + (v2 * t) * -1.05996152e-03f + (v2 * t2) * 6.71079346e-04f;
        }
        v = fmaxf(nk / 3.6f, 0.f);
        cx = nx; cy = ny; ca = na;
        cg[2 * step] = cx; cg[2 * step + 1] = cy; ag[step] = ca; vg[step] = v;
        float best = INFINITY;
        for (int t = 0; t < EV_TS; ++t) {
            const float dx = rx[t] - cx, dy = ry[t] - cy, d = dx * dx + dy * dy;
            if (d < best) { best = d; closest = t; }
        }
    }
    for (int k = 0; k < EV_PID_N; ++k) { pid_buf[(long long)g * EV_PID_N + k] = tb[k]; pid_buf[((long long)pid_slots + g) * EV_PID_N + k] = sb[k]; }
    pid_ptr[g] = tptr; pid_ptr[pid_slots + g] = sptr; pid_len[g] = tlen; pid_len[pid_slots + g] = slen;

    // derive_kinematics (track_propogate.py:500-596): Savitzky-Golay(5, 2) with reflect padding, central differences
    auto refl = [](int i) { return i < 0 ? -i : (i >= EV_TR ? 2 * (EV_TR - 1) - i : i); };
    float sv[EV_TR], hv[EV_TR];
    const float K[5] = {-3.f / 35.f, 12.f / 35.f, 17.f / 35.f, 12.f / 35.f, -3.f / 35.f};
    for (int t = 0; t < EV_TR; ++t) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) { a += K[k] * vg[refl(t + k - 2)]; b += K[k] * ag[refl(t + k - 2)]; }
        sv[t] = a; hv[t] = b;
    }
    auto wrap = [](float d) { return atan2f(sinf(d), cosf(d)); };
    float* accg = acc + (long long)g * EV_TR;
    float* yrg = yaw_rate + (long long)g * EV_TR;
    float* yag = yaw_acc + (long long)g * EV_TR;
    for (int t = 0; t < EV_TR; ++t) {
        if (t == 0) { accg[t] = (sv[1] - sv[0]) / dt; yrg[t] = wrap(hv[1] - hv[0]) / dt; }
        else if (t == EV_TR - 1) { accg[t] = (sv[t] - sv[t - 1]) / dt; yrg[t] = wrap(hv[t] - hv[t - 1]) / dt; }
        else { accg[t] = (sv[t + 1] - sv[t - 1]) / (2.f * dt); yrg[t] = wrap(hv[t + 1] - hv[t - 1]) / (2.f * dt); }
    }
    for (int t = 0; t < EV_TR; ++t) {
        if (t == 0) yag[t] = (yrg[1] - yrg[0]) / dt;
        else if (t == EV_TR - 1) yag[t] = (yrg[t] - yrg[t - 1]) / dt;
        else yag[t] = (yrg[t + 1] - yrg[t - 1]) / (2.f * dt);
    }
    // smoothed speed replaces the raw one; box corners FL, RL, RR, FR around the UNsmoothed pose
    const float hw = 0.5f * st.width, hl = 0.5f * st.length;
    float* vx = vertices + (long long)g * EV_TR * 8;
    for (int t = 0; t < EV_TR; ++t) {
        const float c = cosf(ag[t]), s = sinf(ag[t]);
        const float dx[4] = {hl, -hl, -hl, hl}, dy[4] = {hw, hw, -hw, -hw};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            vx[t * 8 + 2 * q] = dx[q] * c - dy[q] * s + cg[2 * t];
            vx[t * 8 + 2 * q + 1] = dx[q] * s + dy[q] * c + cg[2 * t + 1];
        }
    }
    for (int t = 0; t < EV_TR; ++t) vg[t] = sv[t];
}

__global__ void eval_other_rollout_kernel(const double* __restrict__ loc, const double* __restrict__ heading_deg,
                                          const double* __restrict__ speed, const double* __restrict__ control,
                                          const double* __restrict__ extent, int N, int n_frames, int near_lane_change,
                                          double inflation, double* __restrict__ vertices) {
    pdl_grid_sync();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double dt = 0.1, gain = 0.36848336, fwb = -0.090769015, rwb = 1.4178275, brake_acc = -4.952399, thr_acc = 0.5633837;
    double x = loc[3 * n], y = loc[3 * n + 1], hd = heading_deg[n] * 0.017453292519943295, v = speed[n];
    const double steer = control[3 * n], throttle = control[3 * n + 1];
    const bool brake = ((unsigned char)control[3 * n + 2]) != 0;
    const double slip = atan(rwb / (fwb + rwb) * tan(gain * steer));
    const double s_min = near_lane_change ? 2.0 : 1.2;
    for (int i = 0; i < n_frames; ++i) {
        const double nx = x + v * cos(hd + slip) * dt, ny = y + v * sin(hd + slip) * dt;
        hd = hd + v / rwb * sin(slip) * dt;
        v = fmax(0.0, v + dt * (brake ? brake_acc : throttle * thr_acc));
        x = nx; y = ny;
        const bool slow = v < 1.0;
        const double ex = extent[2 * n] * (slow ? 1.0 : fmax(s_min, 1.2 * (double)i / (double)n_frames)) * inflation;
        const double ey = extent[2 * n + 1] * (slow ? 1.0 : fmax(1.0, 1.3 * (double)i / (double)n_frames)) * inflation;
        const double hw = ey, hl = ex;                        // half width / half length of [2 ey, 2 ex]
        const double cxr = x, cyr = -y, ang = -hd;            // left-handed CARLA -> right-handed
        const double c = cos(ang), s = sin(ang);
        const double ol[4] = {hl, -hl, -hl, hl}, ow[4] = {hw, hw, -hw, -hw};
        double* o = vertices + ((long long)n * n_frames + i) * 8;
        for (int q = 0; q < 4; ++q) { o[2 * q] = ol[q] * c - ow[q] * s + cxr; o[2 * q + 1] = ol[q] * s + ow[q] * c + cyr; }
    }
}

struct EvalMap { const uint8_t* mask; int H, W; double ox, oy, angle, resolution; };

__global__ void __launch_bounds__(64)
eval_return_kernel(const float* __restrict__ delta_dis, const float* __restrict__ delta_angle, const float* __restrict__ speed,
                   const float* __restrict__ acc, const float* __restrict__ yaw_rate, const float* __restrict__ yaw_acc,
                   const float* __restrict__ center, const float* __restrict__ vertices, const double* __restrict__ other, int N,
                   int other_frames, EvalMap mp, int G, double gamma, double* __restrict__ returns, uint8_t* __restrict__ collision_out,
                   uint8_t* __restrict__ offroad_out) {
    pdl_grid_sync();
    const int g = blockIdx.x * 64 + threadIdx.x;
    if (g >= G) return;
    const double ca = cos(mp.angle), sa = sin(mp.angle);
    double ret = 0.0, disc = 1.0;
    bool done = false;
    for (int t = 0; t < EV_TR; ++t) {
        // envelope overlap with any neighbour (only the neighbours' forecast frames exist: traj_evaluator.py:247-250)
        bool col = false;
        if (t < other_frames && N > 0) {
            const float* vx = vertices + ((long long)g * EV_TR + t) * 8;
            float x0 = vx[0], x1 = vx[0], y0 = vx[1], y1 = vx[1];
#pragma unroll
            for (int q = 1; q < 4; ++q) { x0 = fminf(x0, vx[2 * q]); x1 = fmaxf(x1, vx[2 * q]); y0 = fminf(y0, vx[2 * q + 1]); y1 = fmaxf(y1, vx[2 * q + 1]); }
            for (int n = 0; n < N && !col; ++n) {
                const double* o = other + ((long long)n * other_frames + t) * 8;
                double a0 = o[0], a1 = o[0], b0 = o[1], b1 = o[1];
                for (int q = 1; q < 4; ++q) { a0 = fmin(a0, o[2 * q]); a1 = fmax(a1, o[2 * q]); b0 = fmin(b0, o[2 * q + 1]); b1 = fmax(b1, o[2 * q + 1]); }
                col = (double)x0 <= a1 && a0 <= (double)x1 && (double)y0 <= b1 && b0 <= (double)y1;
            }
        }
        // nearest-pixel off-road lookup (traj_evaluator.py:300-330): outside the raster = not off road
        bool off = false;
        if (mp.mask) {
            const double dx = (double)center[((long long)g * EV_TR + t) * 2] - mp.ox, dy = (double)center[((long long)g * EV_TR + t) * 2 + 1] - mp.oy;
            const double px = (dx * ca + dy * sa) / (double)(float)mp.resolution + (double)(float)(mp.H / 2.0);
            const double py = (dx * (-sa) + dy * ca) / (double)(float)(-mp.resolution) + (double)(float)(mp.W / 2.0);
            const long long ix = (long long)rint(px), iy = (long long)rint(py);
            if (ix >= 0 && ix < mp.W && iy >= 0 && iy < mp.H) off = mp.mask[iy * mp.W + ix] == 1;
        }
        if (collision_out) collision_out[(long long)g * EV_TR + t] = col;
        if (offroad_out) offroad_out[(long long)g * EV_TR + t] = off;
        if (t < EV_TS && !done) {
            // DenseRewardModel.get_reward (reward_model.py:34-50), float64 arithmetic on the float32 rollout quantities
            const long long i = (long long)g * EV_TS + t, j = (long long)g * EV_TR + t;
            const double dd = fabs((double)delta_dis[i]), da = fabs((double)delta_angle[i]);
            const double v = speed[j], a = acc[j], w = yaw_rate[j], wa = yaw_acc[j];
            (void)w;
            const double cd = cos(da);
            const double r_col = -(20.0 + fabs(v)) * (col ? 1.0 : 0.0);
            const double r_off = -5.0 * (off ? 1.0 : 0.0);
            const double r_comf = -0.8 * ((fabs(a) > 4 ? 1.0 : 0.0) + (fabs(wa) > 4 ? 1.0 : 0.0));
            const double r_align = 0.5 * (fmin(cd, 0.0) + 0.05 * fmin(cd * v, 0.0) + 0.25 * (1.0 - da / (3.141592653589793 / 2)));
            const double r_center = -0.6 * (cd > 0.5 ? 1.0 : 0.0) * (dd - 0.05 / exp(dd - 0.5));
            const double r_vel = 0.1 * fmax(cd, 0.0) * ((3 < fabs(v) && fabs(v) < 20) ? 1.0 : 0.0) * fabs(v);
            const double r_time = -0.1 * ((fabs(v) > 0 || fabs(a) > 0) ? 1.0 : 0.0);
            ret += (r_col + r_off + r_comf + r_align + r_center + r_vel + r_time) * disc;
            disc *= gamma;
            if (col) done = true;                             // the colliding frame counts, later ones do not
        }
    }
    returns[g] = ret;
}

}  // namespace rift

using namespace rift;

extern "C" {

int rift_b200_eval_ref_line_info(const float* trajectory, int R, int M, int T_in, const float* ref_pos, const float* ref_angle,
                                 const int* ref_offsets, float* delta_dis, float* delta_angle, void* stream) {
    RIFT_REQUIRE(trajectory && ref_pos && ref_angle && ref_offsets && delta_dis && delta_angle, "eval_ref_line_info: null argument");
    RIFT_REQUIRE(T_in >= EV_TS, "eval_ref_line_info: trajectories shorter than 40 frames");
    if (R * M <= 0) return 0;
    launch_k(eval_ref_line_kernel, cdiv((long long)R * M * EV_TS, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream), trajectory, R, M, T_in,
             ref_pos, ref_angle, ref_offsets, delta_dis, delta_angle);
    RIFT_LAUNCH_OK();
    return 0;
}

int rift_b200_eval_center_rollout(const float* trajectory, int G, int T_in, const float* state6_host, float dt, float* pid_buf,
                                  int* pid_ptr, int* pid_len, int pid_slots, float* center, float* angle, float* speed, float* acc,
                                  float* yaw_rate, float* yaw_acc, float* vertices, void* stream) {
    RIFT_REQUIRE(trajectory && state6_host && pid_buf && pid_ptr && pid_len && center && angle && speed && acc && yaw_rate && yaw_acc && vertices,
                 "eval_center_rollout: null argument");
    RIFT_REQUIRE(T_in >= EV_TS && pid_slots >= G, "eval_center_rollout: bad sizes");
    if (G <= 0) return 0;
    EvalState st{state6_host[0], state6_host[1], state6_host[2], state6_host[3], state6_host[4], state6_host[5]};
    launch_k(eval_center_rollout_kernel, cdiv(G, 64), 64, 0, reinterpret_cast<cudaStream_t>(stream), trajectory, G, T_in, st, dt, pid_buf,
             pid_ptr, pid_len, pid_slots, center, angle, speed, acc, yaw_rate, yaw_acc, vertices);
    RIFT_LAUNCH_OK();
    return 0;
}

int rift_b200_eval_other_rollout(const double* location, const double* heading_deg, const double* speed, const double* control,
                                 const double* extent, int N, int n_frames, int near_lane_change, double inflation, double* vertices,
                                 void* stream) {
    if (N <= 0) return 0;
    RIFT_REQUIRE(location && heading_deg && speed && control && extent && vertices, "eval_other_rollout: null argument");
    launch_k(eval_other_rollout_kernel, cdiv(N, 64), 64, 0, reinterpret_cast<cudaStream_t>(stream), location, heading_deg, speed, control,
             extent, N, n_frames, near_lane_change, inflation, vertices);
    RIFT_LAUNCH_OK();
    return 0;
}

int rift_b200_eval_returns(const float* delta_dis, const float* delta_angle, const float* speed, const float* acc, const float* yaw_rate,
                           const float* yaw_acc, const float* center, const float* vertices, const double* other_vertices, int N,
                           int other_frames, const uint8_t* off_road_mask, int H, int W, const double* map_pose4_host, int G,
                           double gamma, double* returns, uint8_t* collision_out, uint8_t* offroad_out, void* stream) {
    RIFT_REQUIRE(delta_dis && delta_angle && speed && acc && yaw_rate && yaw_acc && center && vertices && returns, "eval_returns: null argument");
    RIFT_REQUIRE(N == 0 || other_vertices, "eval_returns: neighbour boxes missing");
    RIFT_REQUIRE(!off_road_mask || map_pose4_host, "eval_returns: the raster needs its pose (origin x, y, angle, resolution)");
    if (G <= 0) return 0;
    EvalMap mp{off_road_mask, H, W, 0, 0, 0, 0.5};
    if (off_road_mask) { mp.ox = map_pose4_host[0]; mp.oy = map_pose4_host[1]; mp.angle = map_pose4_host[2]; mp.resolution = map_pose4_host[3]; }
    launch_k(eval_return_kernel, cdiv(G, 64), 64, 0, reinterpret_cast<cudaStream_t>(stream), delta_dis, delta_angle, speed, acc, yaw_rate,
             yaw_acc, center, vertices, other_vertices, N, other_frames, mp, G, gamma, returns, collision_out, offroad_out);
    RIFT_LAUNCH_OK();
    return 0;
}

}  // extern "C"
