#!/usr/bin/env python
"""Policy-update throughput benchmark (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--trainable pi_head|full]

One "step" = one policy update on one pre-collated synthetic rollout batch: PlanningModel forward
-> GRPO objective -> backward -> (N>1: one NCCL all-reduce of the flat gradient arena) ->
clip_grad_norm_(0.5) -> AdamW.  Workload = BASELINE.json configs[1]: 64 scenes x 32 agents x 20
polylines x 80 steps, R = 6 reference lines x 12 modes, Pluto-medium, GRPO.  Weak scaling: every
rank holds its own 64-scene batch; `value` counts 64-scene batch-equivalents per second over all ranks.

`value`  : inputs already resident in HBM, CUDA events around each step, L2 flushed between steps.
`e2e`    : the same step through LightningTrainer.step() starting from pinned HOST buffers (H2D of the
           whole batch inside the timed region) and ending with the loss read back to the host.
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
# forward matmul+conv FLOPs of the reference modules for the cfg2 batch (FlopCounterMode)
FWD_GFLOP = {"medium": 231.086, "small": 70.711}
PI_HEAD_GFLOP = {"medium": 0.606, "small": 0.152}


def trainable_layers(mode):
    if mode == "pi_head":
        return ["planning_decoder.pi_head"]
    return ["pos_emb", "agent_encoder", "map_encoder", "static_objects_encoder", "encoder_blocks", "norm",
            "agent_predictor", "planning_decoder", "hidden_proj", "ref_free_decoder"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def host_batch(cfg, wl, seed):
    feats = synth_features(cfg, wl["bs"], wl["A"], wl["Mp"], wl["R"], seed=seed)
    ex = synth_rl_extras(cfg, feats, seed=seed + 1)
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
    for k in ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits"):
        b[k + "_torch"] = ex_t[k]
    return b


# --------------------------------------------------------------------------------------------------
def cpu_reference_step_time(cfg, wl, trainable, steps, warmup, threads):
    """The oracle port of the reference step (forward, GRPO loss, autograd backward over the trainable
    set, clip 0.5, AdamW) on the host cores.  Returns seconds per step (best of `steps`)."""
    from oracle import pluto_oracle as po, loss_oracle as lo
    torch.set_num_threads(threads)
    feats, ex = host_batch(cfg, wl, seed=1)
    data, ext = to_torch(feats), to_torch(ex)
    sd = {k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()}
    from rift_b200.arena import trainable_names
    names = trainable_names(cfg, trainable_layers(trainable))
    for n in names:
        sd[n].requires_grad_(True)
    state = {}
    r_pad = ~data["reference_line"]["valid_mask"].any(-1)
    times = []
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
    return min(times), float(loss.detach())


def run_reference(args, wl_name, wl, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 4))                 # bounded sample: a few whole steps of the same batch
    warm = 1 if args.warmup > 0 else 0
    sec, _ = cpu_reference_step_time(cfg, wl, args.trainable, steps, warm, threads)
    v = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, **wl, "algo": "grpo", "trainable": args.trainable},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"best of {steps} whole policy updates of the {wl_name} batch after {warm} warm-up "
                                   f"(oracle/ restatement of the reference's torch-CPU path; torch.set_num_threads({threads}))"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def run_ours(args, wl_name, wl, cfg):
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

    model = PlanningModel.from_config(cfg, device=dev)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()})
    tr = TRAINERS["grpo"](model, trainable_layers=trainable_layers(args.trainable), **TRAINER_KW)
    tr.configure_optimizers()

    feats, ex = host_batch(cfg, wl, seed=1 + rank)                 # every rank its own scenes (weak scaling)
    feats_h, ex_h = to_torch(feats, pin=True), to_torch(ex, pin=True)
    keys = ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits")
    ex_h = {k: ex_h[k] for k in keys}
    h2d = tree_bytes(feats_h) + tree_bytes(ex_h)
    batch_dev = make_batch_dict(model.pack(to_device(feats_h, dev)), to_device(ex_h, dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing
    L = _lib.lib()
    graph = tr.use_cuda_graph and not args.ncu and not args.no_graph
    tr.use_cuda_graph = False                 # one eager step: counts the kernels of a policy update
    launches0 = L.rift_b200_launch_count()
    tr.step(batch_dev)
    launches = L.rift_b200_launch_count() - launches0
    tr.use_cuda_graph = graph                 # then (default) forward + objective + backward replay from a CUDA graph
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
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    for s, e in evs:
        flush.zero_()
        s.record()
        loss = tr.step(batch_dev)
        e.record()
    sync()
    clocks = sampler.stop()
    ms = sum(s.elapsed_time(e) for s, e in evs) / args.steps
    loss_val = float(loss)

    # ---- end to end: pinned host buffers -> H2D -> step -> loss on the host
    e2e_steps = args.steps
    # the trainer takes the pinned host batch as is: H2D copies (into the graph's static inputs, or by PackedBatch on
    # the eager path) happen inside step(), i.e. inside the timed region
    host_batch_dict = make_batch_dict(feats_h, ex_h)
    for _ in range(max(args.warmup, 3)):
        float(tr.step(host_batch_dict))
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        float(tr.step(host_batch_dict))
    sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps

    # ---- dominant kernel live: the largest GEMM of the step, replayed alone under CUDA events
    D = cfg.dim
    rows, K, N = wl["bs"] * wl["R"] * 120, 256, 256         # PointsEncoder second_mlp.0 over reference-line points
    x = torch.randn(rows, K, device=dev)
    w = torch.randn(N, K, device=dev) * K ** -0.5
    y = torch.empty(rows, N, device=dev)
    scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device=dev)
    reps = 20

    def gemm_only(flag):     # flag 3: split W + pack A (first call); 0 afterwards: planes reused, only gemm_tc_kernel runs
        _lib.check(L.rift_b200_op_linear_tc(_lib.ptr(x), rows, K, _lib.ptr(w), None, N, 1, None, _lib.ptr(y), _lib.ptr(scratch),
                                            scratch.numel(), flag, _lib.stream_ptr()), "op_linear_tc")
    gemm_only(1)
    for _ in range(3):
        gemm_only(2)
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    kt = 0.0
    for _ in range(reps):
        flush.zero_()
        ks.record()
        gemm_only(2)
        ke.record()
        torch.cuda.synchronize()
        kt += ks.elapsed_time(ke)
    kernel_ms = kt / reps
    hbm, tf, src = peaks()
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r1b_gemm_big_ncu_full.json")
    if os.path.exists(tpath) and (rows, K, N) == (46080, 256, 256):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    kernel_tf = 2.0 * rows * K * N / (kernel_ms * 1e-3) / 1e12
    # HBM view of the same launch: bf16 hi+lo planes of A (4 B/elem) in, fp32 C out, weights from L2
    kernel_gbs = (rows * K * 4 + rows * N * 4) / (kernel_ms * 1e-3) / 1e9

    # ---- max over ranks
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    step_gflop = FWD_GFLOP[wl["model"]] * (3.0 if args.trainable == "full" else 1.0) + \
        (3.0 * PI_HEAD_GFLOP[wl["model"]] if args.trainable == "pi_head" else 0.0)

    if rank == 0:
        threads = os.cpu_count() or 1
        cpu_sec, _ = cpu_reference_step_time(cfg, wl, args.trainable, 2, 1, threads) if world == 1 and not args.no_cpu \
            else (None, None)
        line = {
            "metric": METRIC, "value": world * 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 tcgen05 operands, fp32 accumulate; fp32 master weights / activations)",
            "data": "synthetic",
            "config": {"workload": wl_name, **wl, "algo": "grpo", "trainable": args.trainable,
                       "l2": "256 MB memset between timed steps", "global_batch": wl["bs"] * world,
                       "parallelism": f"dp{world}", "loss": loss_val,
                       "launch": "forward + objective + backward replayed from one CUDA graph; all-reduce / clip / AdamW eager"
                                 if graph else "eager launches",
                       "step_gflop_algorithmic": step_gflop,
                       "step_tensor_frac_of_peak": step_gflop / ms / tf},
            "clocks": clocks,
            "e2e": {"value": world * 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms},
            "gpu_launches": launches,
            # K, N <= 1024 everywhere in this model: the GEMMs sit below the ridge point and are HBM-bound
            "roofline": {"bound": "hbm", "kernel": "rift::gemm_tc_kernel<128,false> (tcgen05 split-bf16 GEMM, largest shape "
                                                   f"of the step: {rows}x{K}x{N}, replayed alone, L2 flushed)",
                         "achieved": kernel_gbs, "peak": hbm, "unit": "GB/s", "frac": kernel_gbs / hbm,
                         "traffic": traffic, "peak_source": src + " copy bandwidth (MEASURED_PEAKS.json)",
                         "algorithmic_bytes": rows * K * 4 + rows * N * 4, "kernel_ms": kernel_ms,
                         "tensor_tflops_algorithmic": kernel_tf, "tensor_frac_of_bf16_sustained": kernel_tf / tf,
                         "mma_work_factor": 3},
        }
        if cpu_sec is not None:
            line["cpu_baseline"] = {
                "value": 1.0 / cpu_sec, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"best of 2 whole policy updates of the {wl_name} batch after 1 warm-up "
                          f"(oracle/ restatement of the reference's torch-CPU path, {threads} threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trainable", default="full", choices=["pi_head", "full"],
                    help="full = differentiate the whole policy (headline, SURVEY 8d); pi_head = the reference's default "
                         "trainable_layers (rift_training.yaml:26-27)")
    ap.add_argument("--ncu", action="store_true", help="run one step inside cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", dest="no_graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    cfg = MODEL_ZOO[wl["model"]](future_steps=wl["future_steps"])
    if args.impl == "reference":
        run_reference(args, args.workload, wl, cfg)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, args.workload, wl, cfg)


if __name__ == "__main__":
    main()
