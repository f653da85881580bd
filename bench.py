#!/usr/bin/env python
"""Policy-update throughput benchmark (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--trainable pi_head|full]
                    [--workload auto|cfg2|cfg4] [--scaling auto|strong|weak]

One "step" = one policy update on one pre-collated synthetic rollout batch: PlanningModel forward
-> GRPO objective -> backward -> (N>1: NCCL all-reduce of the flat gradient arena) -> clip_grad_norm_(0.5)
-> AdamW.  Which batch (BASELINE.json `configs`):

    N = 1        configs[1]  cfg2: 64 scenes x 32 agents x 20 polylines x 80 steps, R = 6 x 12 modes, Pluto-medium, GRPO
    N = 2, 4     configs[2]  the SAME cfg2 batch split 32 / 16 scenes per rank           (strong scaling)
    N = 8        configs[3]  cfg4: 256 scenes x 48 agents, 32 per rank, clip eps = 0.2 + 0.2 * KL-to-reference
    --scaling weak           every rank holds its own full batch (also reported as `weak_scaling` at N > 1)

`value`  : 64-scene batch equivalents per second over all ranks = (global scenes / 64) / (median step time);
           inputs resident in HBM, CUDA events around each step, L2 flushed between steps, max over ranks.
`e2e`    : the same step through LightningTrainer.step() starting from pinned HOST buffers (H2D of the
           rank's batch inside the timed region) and ending with the loss read back to the host.
`e2e_device_buffer` (N = 1): the same step fed from the device-resident replay buffer - slot indices drawn on the host,
           one gather launch (DeviceRolloutBuffer.collate_device) into the batch the step graph reads in place, loss read back.
`roofline` / `roofline_step` / `kernels` : the dominant kernel family replayed alone against the measured HBM copy bandwidth,
           the whole step against the measured bf16 rate, and the reference-line GEMM + cfg5 advantage kernels.
`--impl reference` : the CPU oracle port of the reference step on the host cores (torch CPU threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rift_b200.config import MODEL_ZOO  # noqa: E402
from rift_b200.synth import synth_state_dict, synth_features, synth_rl_extras, WORKLOADS  # noqa: E402

METRIC = "policy_update_steps_per_sec"
UNIT = "steps/s (64-scene batch equivalents, all ranks)"
TRAINER_KW = dict(lr=1e-4, cl_lr_decay=0.9, weight_decay=1e-5, epochs=16, warmup_epochs=3, frame_rate=10)
# forward matmul+conv FLOPs PER SAMPLE of the reference modules (FlopCounterMode), keyed by (model, A):
# cfg2 shape (A = 32) and cfg4 shape (A = 48); the step is 3 x forward when the whole policy is differentiated
FWD_GFLOP_PER_SAMPLE = {("medium", 32): 231.086 / 64, ("medium", 48): 1139.578 / 256,
                        ("small", 32): 70.711 / 64, ("small", 48): 334.189 / 256}
PI_HEAD_GFLOP_PER_SAMPLE = {"medium": 0.606 / 64, "small": 0.152 / 64}
EXTRA_KEYS = ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits")


def trainable_layers(mode):
    if mode == "pi_head":
        return ["planning_decoder.pi_head"]
    return ["pos_emb", "agent_encoder", "map_encoder", "static_objects_encoder", "encoder_blocks", "norm",
            "agent_predictor", "planning_decoder", "hidden_proj", "ref_free_decoder"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


def step_gflop(wl, trainable, n_scenes):
    f = FWD_GFLOP_PER_SAMPLE.get((wl["model"], wl["A"]))
    if f is None:
        return None
    if trainable == "full":
        return 3.0 * f * n_scenes
    return (f + 3.0 * PI_HEAD_GFLOP_PER_SAMPLE[wl["model"]]) * n_scenes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.startswith("Active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def tree_slice(tree, sl):
    if isinstance(tree, dict):
        return {k: tree_slice(v, sl) for k, v in tree.items()}
    return np.ascontiguousarray(tree[sl])


def host_batch(cfg, wl, seed, sl=None):
    """The workload's synthetic batch (numpy); `sl` keeps one rank's scenes of it."""
    feats = synth_features(cfg, wl["bs"], wl["A"], wl["Mp"], wl["R"], seed=seed)
    ex = synth_rl_extras(cfg, feats, seed=seed + 1)
    if sl is not None:
        feats, ex = tree_slice(feats, sl), tree_slice(ex, sl)
    return feats, ex


def to_torch(tree, pin=False):
    if isinstance(tree, dict):
        return {k: to_torch(v, pin) for k, v in tree.items()}
    t = torch.from_numpy(np.ascontiguousarray(tree))
    return t.pin_memory() if pin and t.numel() else t


def to_device(tree, dev):
    if isinstance(tree, dict):
        return {k: to_device(v, dev) for k, v in tree.items()}
    return tree.to(dev, non_blocking=True)


def tree_bytes(tree):
    if isinstance(tree, dict):
        return sum(tree_bytes(v) for v in tree.values())
    return tree.numel() * tree.element_size()


def make_batch_dict(feats_t, ex_t):
    b = {"cur_pluto_feature_torch": feats_t}
    for k in EXTRA_KEYS:
        b[k + "_torch"] = ex_t[k]
    return b


def device_buffer_leg(tr, feats, ex, wl, args, sync):
    """Fills a DeviceRolloutBuffer with 2 x bs transitions cut from the workload's scenes (one CBV trajectory, stored through the
    reference's store() protocol) and times sample-indices -> collate_device -> step -> loss."""
    from rift_b200.buffer import DeviceRolloutBuffer
    from rift_b200.feature import PlutoFeature
    bs = wl["bs"]
    keys = ["CBVs_actions", "CBVs_actions_old_group_logits", "CBVs_actions_ref_group_logits", "CBVs_group_advantage", "CBVs_obs",
            "CBVs_next_obs", "CBVs_reward", "CBVs_terminated", "CBVs_done"]
    cap = 2 * bs
    buf = DeviceRolloutBuffer(1, "train_cbv", {"buffer_capacity": cap, "data_keys": keys})
    dd = {k: [] for k in keys}
    dd["CBV_ids"] = []
    n = cap + 1
    for t in range(n):
        i = t % bs
        d = {k: ({kk: vv[i] for kk, vv in v.items()} if isinstance(v, dict) else v[i]) for k, v in feats.items()}
        f = PlutoFeature(data=d)
        vm = np.asarray(ex["group_advantage_mask"][i], bool)
        tr_ = {"CBVs_actions": (0.1, 0.0, False), "CBVs_obs": {"raw_pluto_feature": f}, "CBVs_next_obs": {"raw_pluto_feature": f},
               "CBVs_group_advantage": {"advantage": ex["group_advantage"][i], "valid_mask": vm},
               "CBVs_actions_old_group_logits": {"logits": ex["old_group_logits"][i], "valid_mask": vm},
               "CBVs_actions_ref_group_logits": {"logits": ex["ref_group_logits"][i], "valid_mask": vm},
               "CBVs_reward": np.float32(0.0), "CBVs_terminated": np.float32(0.0), "CBVs_done": t == n - 1}
        dd["CBV_ids"].append([0])
        for k in keys:
            dd[k].append({0: tr_[k]})
    buf.store(dd)
    assert buf.buffer_full and len(buf) == cap
    rng = np.random.Generator(np.random.PCG64(3))

    def one():
        idx = rng.permutation(cap)[:bs]
        return float(tr.step(buf.collate_device(idx.tolist(), "grpo")))
    for _ in range(max(min(args.warmup, 5), 3)):
        one()
    sync()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one()
        ts.append((time.perf_counter() - t0) * 1e3)
    sync()
    return float(np.median(ts)), bs * 8


def pick_workload(args):
    """BASELINE.json configs by GPU count (see the module docstring)."""
    name = args.workload
    if name == "auto":
        name = "cfg4" if args.gpus >= 8 else "cfg2"
    scaling = args.scaling
    if scaling == "auto":
        scaling = "strong" if args.gpus > 1 else "weak"
    wl = dict(WORKLOADS[name])
    if scaling == "strong" and wl["bs"] % args.gpus != 0:
        raise SystemExit(f"--gpus {args.gpus} does not divide the {wl['bs']}-scene batch of {name}")
    return name, wl, scaling


# --------------------------------------------------------------------------------------------------
def cpu_reference_step_times(cfg, wl, trainable, steps, warmup, threads, n_scenes=None):
    """The oracle port of the reference step (forward, GRPO loss, autograd backward over the trainable
    set, clip 0.5, AdamW) on the host cores, on the first `n_scenes` scenes of the workload's batch.
    Returns (list of seconds per step, last loss)."""
    from oracle import pluto_oracle as po, loss_oracle as lo
    torch.set_num_threads(threads)
    feats, ex = host_batch(cfg, wl, seed=1, sl=slice(0, n_scenes) if n_scenes else None)
    data, ext = to_torch(feats), to_torch(ex)
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    from rift_b200.arena import trainable_names
    names = trainable_names(cfg, trainable_layers(trainable))
    for n in names:
        sd[n].requires_grad_(True)
    state = {}
    r_pad = ~data["reference_line"]["valid_mask"].any(-1)
    times = []
    loss = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = po.planning_model_forward(data, sd, cfg)
        loss = lo.grpo_loss(out["probability"], ext["old_group_logits"], ext["ref_group_logits"], ext["group_advantage"],
                            ext["group_advantage_mask"], r_pad)
        loss.backward()
        with torch.no_grad():
            grads = {n: sd[n].grad for n in names if sd[n].grad is not None}
            lo.clip_adamw_step(sd, grads, state, lr=1e-4)
            for n in names:
                sd[n].grad = None
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, float(loss.detach())


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the step (oracle port: a Python reference cannot
    travel to the GPU box) on all host threads, honouring --steps / --warmup.  Each step is a bounded sample of the
    workload: the first 64 scenes of the batch (= one 64-scene batch equivalent, the metric's unit)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl_name, wl, scaling = pick_workload(args)
    cfg = MODEL_ZOO[wl["model"]](future_steps=wl["future_steps"])
    threads = os.cpu_count() or 1
    sample = min(64, wl["bs"])
    times, _ = cpu_reference_step_times(cfg, wl, args.trainable, args.steps, args.warmup, threads, n_scenes=sample)
    sec = float(np.median(times))
    v = (sample / 64.0) / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, **wl, "algo": "grpo", "trainable": args.trainable},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"median of {args.steps} whole policy updates on the first {sample} scenes of the {wl_name} "
                                   f"batch after {args.warmup} warm-up (oracle/ restatement of the reference's torch-CPU path; "
                                   f"torch.set_num_threads({threads}))"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def timed_steps(tr, batch, steps, flush, sync):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync()
    loss = None
    for s, e in evs:
        flush.zero_()
        s.record()
        loss = tr.step(batch)
        e.record()
    sync()
    return [s.elapsed_time(e) for s, e in evs], float(loss)


def kernel_rooflines(L, _lib, dev, flush, wl, hbm, tf):
    """Live CUDA-event timings of the step's time-dominant kernel family and of the cfg5 advantage kernel, each
    replayed alone with the L2 flushed (algorithmic bytes per launch as defined in DESIGN.md section 4)."""
    from rift_b200 import functional as F
    out = {}

    def time_it(fn, reps=20):
        ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            flush.zero_()
            ks.record()
            fn()
            ke.record()
            torch.cuda.synchronize()
            ts.append(ks.elapsed_time(ke))
        return float(np.median(ts))

    # gemm_tc_kernel instances: the decoder-row class (bs*R*12 rows, 256 x 256: the most frequent launch of the step)
    # and the largest launch (reference-line points, bs*R*120 rows)
    for tag, rows in (("decoder_rows", wl["bs"] * wl["R"] * 12), ("refline_points", wl["bs"] * wl["R"] * 120)):
        K = N = 256
        x = torch.randn(rows, K, device=dev)
        w = torch.randn(N, K, device=dev) * K ** -0.5
        y = torch.empty(rows, N, device=dev)
        scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device=dev)

        def gemm_only(flag):     # 1: split W + pack A; 2: planes reused, only gemm_tc_kernel runs
            _lib.check(L.rift_b200_op_linear_tc(_lib.ptr(x), rows, K, _lib.ptr(w), None, N, 1, None, _lib.ptr(y),
                                                _lib.ptr(scratch), scratch.numel(), flag, _lib.stream_ptr()), "op_linear_tc")
        gemm_only(1)
        for _ in range(3):
            gemm_only(2)
        torch.cuda.synchronize()
        ms = time_it(lambda: gemm_only(2))
        alg = rows * K * 4 + rows * N * 4            # bf16 hi+lo planes of A in, fp32 C out, weights from L2
        gbs = alg / (ms * 1e-3) / 1e9
        tfl = 2.0 * rows * K * N / (ms * 1e-3) / 1e12
        out[tag] = {"kernel": f"rift::gemm_tc_kernel {rows}x{K}x{N} (tcgen05 split-bf16 x3)", "kernel_ms": ms,
                    "algorithmic_bytes": alg, "achieved_gbs": gbs, "frac_hbm": gbs / hbm,
                    "tensor_tflops_algorithmic": tfl, "frac_tensor": tfl / tf}
    # BASELINE configs[4]: group-relative advantage, 2^20 groups, G = 72 and 12 (16 * G bytes per group)
    for G in (72, 12):
        n = 1 << 20
        ret = torch.randn(n, G, dtype=torch.float64, device=dev) * 20 - 5
        F.group_advantage(ret)
        torch.cuda.synchronize()
        ms = time_it(lambda: F.group_advantage(ret), reps=10)
        gbs = 16.0 * G * n / (ms * 1e-3) / 1e9
        out[f"advantage_G{G}"] = {"kernel": f"rift::group_advantage 2^20 groups x {G} (fp64, bit-exact)", "kernel_ms": ms,
                                  "algorithmic_bytes": 16 * G * n, "achieved_gbs": gbs, "frac_hbm": gbs / hbm}
        del ret
    return out


DOMINANT = "decoder_rows"      # profiles/r1b_summary.md: gemm_tc_kernel<64> launches of this class = largest share of device time


def dominant_roofline(kern, hbm, src):
    k = kern.get(DOMINANT)
    if not k:
        return None
    traffic = None          # dram bytes per launch of this kernel from the committed ncu --set full capture, if any
    tpath = os.path.join(ROOT, "profiles", "r2_gemm_decoder_rows_ncu_full.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    return {"bound": "hbm", "kernel": k["kernel"] + ", replayed alone, L2 flushed", "achieved": k["achieved_gbs"], "peak": hbm,
            "unit": "GB/s", "frac": k["frac_hbm"], "traffic": traffic, "peak_source": src + " copy bandwidth",
            "algorithmic_bytes": k["algorithmic_bytes"], "kernel_ms": k["kernel_ms"],
            "tensor_tflops_algorithmic": k["tensor_tflops_algorithmic"], "tensor_frac_of_bf16_sustained": k["frac_tensor"],
            "mma_work_factor": 3}


def run_ours(args):
    import torch.distributed as dist
    from rift_b200 import _lib
    from rift_b200.planning_model import PlanningModel
    from rift_b200.trainer import TRAINERS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (printed at communicator creation) goes to stderr
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    wl_name, wl, scaling = pick_workload(args)
    cfg = MODEL_ZOO[wl["model"]](future_steps=wl["future_steps"])

    model = PlanningModel.from_config(cfg, device=dev)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()})
    # configs[3] names "PPO-clip eps = 0.2 with KL-to-reference penalty": the GRPO objective with clip (1 - eps, 1 + eps)
    tr = TRAINERS["grpo"](model, trainable_layers=trainable_layers(args.trainable), clip=(0.8, 1.2), kl_weight=0.2, **TRAINER_KW)
    tr.configure_optimizers()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def rank_batch(mode):
        if mode == "strong":                  # one global batch, contiguous scene shards
            per = wl["bs"] // world
            feats, ex = host_batch(cfg, wl, seed=1, sl=slice(rank * per, (rank + 1) * per))
        else:                                 # weak: every rank its own scenes
            feats, ex = host_batch(cfg, wl, seed=1 + rank)
        feats_h, ex_h = to_torch(feats, pin=True), to_torch(ex, pin=True)
        ex_h = {k: ex_h[k] for k in EXTRA_KEYS}
        return feats_h, ex_h

    feats_h, ex_h = rank_batch(scaling)
    scenes_global = wl["bs"] if scaling == "strong" else wl["bs"] * world
    h2d = tree_bytes(feats_h) + tree_bytes(ex_h)
    batch_dev = make_batch_dict(model.pack(to_device(feats_h, dev)), to_device(ex_h, dev))

    # ---- device-resident timing
    L = _lib.lib()
    graph = tr.use_cuda_graph and not args.ncu and not args.no_graph
    tr.use_cuda_graph = False                 # one eager step: counts the kernels of a policy update
    launches0 = L.rift_b200_launch_count()
    tr.step(batch_dev)
    launches = L.rift_b200_launch_count() - launches0
    tr.use_cuda_graph = graph                 # then (default) the whole update replays from ONE CUDA graph
    for _ in range(args.warmup):
        tr.step(batch_dev)
    sync()
    if args.ncu:
        # profiling aid (never a bench number): exactly one policy update between cudaProfilerStart/Stop, for
        # `ncu --profile-from-start off ... python bench.py --ncu`
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        tr.step(batch_dev)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local)
    sampler.start()
    times, loss_val = timed_steps(tr, batch_dev, args.steps, flush, sync)
    clocks = sampler.stop()
    ms_med, ms_mean = float(np.median(times)), float(np.mean(times))

    # ---- end to end: pinned host buffers -> H2D -> step -> loss on the host
    # the trainer takes the pinned host batch as is: H2D copies into the graph's static inputs happen inside step()
    host_batch_dict = make_batch_dict(feats_h, ex_h)
    for _ in range(max(min(args.warmup, 5), 3)):
        float(tr.step(host_batch_dict))
    sync()
    e2e_t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        float(tr.step(host_batch_dict))
        e2e_t.append((time.perf_counter() - t0) * 1e3)
    sync()
    e2e_ms = float(np.median(e2e_t))

    # ---- end to end from the device-resident replay buffer: the scenes live in the buffer's device arenas
    # (as they would after a rollout), a step = draw 64 slot indices on the host -> ONE gather launch builds the batch -> step ->
    # loss on the host.  Host-to-device traffic: the index vector.
    devbuf_ms, devbuf_h2d = 0.0, 0
    if world == 1 and not args.no_devbuf:
        try:
            feats_np, ex_np = host_batch(cfg, wl, seed=1)
            devbuf_ms, devbuf_h2d = device_buffer_leg(tr, feats_np, ex_np, wl, args, sync)
        except Exception as e:                                   # a secondary record must not cost the headline line
            print(f"device-buffer leg failed: {e!r}", file=sys.stderr)

    # ---- secondary leg at N > 1: weak scaling (every rank its own full batch)
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        wf, we = rank_batch("weak")
        wb = make_batch_dict(model.pack(to_device(wf, dev)), to_device(we, dev))
        for _ in range(max(3, min(args.warmup, 5))):
            tr.step(wb)
        wt, _ = timed_steps(tr, wb, min(args.steps, 20), flush, sync)
        weak = float(np.median(wt))

    # ---- max over ranks
    t = torch.tensor([ms_med, ms_mean, e2e_ms, weak or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_med, ms_mean, e2e_ms, weak_ms = (float(x) for x in t)

    if rank == 0:
        hbm, tf, src = peaks()
        kern = kernel_rooflines(L, _lib, dev, flush, dict(WORKLOADS["cfg2"]), hbm, tf) if not args.no_kernels else {}
        gf = step_gflop(wl, args.trainable, scenes_global)
        step_tf = gf / ms_med / world if gf else None            # GFLOP / ms = TFLOP/s, per GPU
        value = (scenes_global / 64.0) * 1e3 / ms_med
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_med, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 tcgen05 operands, fp32 accumulate; fp32 master weights / activations)",
            "data": "synthetic",
            "config": {"workload": wl_name, **wl, "algo": "grpo", "clip": [0.8, 1.2], "kl_weight": 0.2,
                       "trainable": args.trainable, "l2": "256 MB memset between timed steps",
                       "global_batch": scenes_global, "per_rank_batch": scenes_global // world,
                       "parallelism": f"dp{world}", "loss": loss_val, "timing": "median of per-step CUDA-event times",
                       "ms_per_step_mean": ms_mean,
                       "launch": "forward + objective + backward + all-reduce + clip + AdamW replayed from one CUDA graph"
                                 if graph else "eager launches",
                       "step_gflop_algorithmic": gf, "step_tensor_frac_of_peak": (step_tf / tf) if gf else None},
            "clocks": clocks,
            "e2e": {"value": (scenes_global / 64.0) * 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            # `roofline`: the time-dominant kernel family of the step (profiles/: share of device time per kernel) replayed
            # alone under CUDA events; `roofline_step`: the WHOLE update against the tensor roofline (algorithmic
            # matmul + conv FLOPs of the step over the measured sustained bf16 rate); `kernels`: the other measured kernels
            "roofline": dominant_roofline(kern, hbm, src),
            "roofline_step": {"bound": "tensor", "kernel": "whole policy update (all kernels of the step)",
                              "achieved": step_tf, "peak": tf, "unit": "TFLOP/s", "frac": (step_tf / tf) if gf else None,
                              "mma_work_factor": 3, "peak_source": src},
            "kernels": kern,
        }
        if devbuf_ms > 0:
            line["e2e_device_buffer"] = {"value": (wl["bs"] / 64.0) * 1e3 / devbuf_ms, "unit": UNIT, "ms_per_step": devbuf_ms,
                                         "h2d_bytes_per_step": devbuf_h2d, "d2h_bytes_per_step": 8,
                                         "path": "DeviceRolloutBuffer.collate_device (rift_b200_gather_fields) -> trainer.step"}
        if weak_ms > 0:
            line["weak_scaling"] = {"value": world * (wl["bs"] / 64.0) * 1e3 / weak_ms, "unit": UNIT, "ms_per_step": weak_ms,
                                    "global_batch": wl["bs"] * world}
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = min(64, wl["bs"])
            ct, _ = cpu_reference_step_times(cfg, wl, args.trainable, 5, 1, threads, n_scenes=sample)
            sec = float(np.median(ct))
            line["cpu_baseline"] = {
                "value": (sample / 64.0) / sec, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"median of 5 whole policy updates on {sample} scenes of the {wl_name} batch after 1 warm-up "
                          f"(oracle/ restatement of the reference's torch-CPU path, {threads} threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_process_group(tr)


def shutdown_process_group(*trainers):
    """NCCL does not finish destroying a communicator while CUDA graphs that captured its kernels are alive (the step graph
    holds the all-reduce), so the graphs go first; should the teardown still block, the result line is already out and the
    process leaves without it."""
    import gc
    import threading
    import torch.distributed as dist
    for tr in trainers:
        tr._invalidate_graphs()
    gc.collect()
    torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(20.0)
    if t.is_alive():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trainable", default="full", choices=["pi_head", "full"],
                    help="full = differentiate the whole policy (headline, SURVEY 8d); pi_head = the reference's default "
                         "trainable_layers (rift_training.yaml:26-27)")
    ap.add_argument("--ncu", action="store_true", help="run one step inside cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"])
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-weak", dest="no_weak", action="store_true", help="skip the secondary weak-scaling leg at N > 1")
    ap.add_argument("--no-kernels", dest="no_kernels", action="store_true", help="skip the per-kernel roofline legs")
    ap.add_argument("--no-devbuf", dest="no_devbuf", action="store_true", help="skip the device-resident replay-buffer end-to-end leg")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
