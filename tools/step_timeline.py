"""Warm, in-graph kernel timeline of one policy update through torch.profiler (CUPTI): per-kernel device time as it is
inside the replayed CUDA graph (ncu's launch list is cold-cache and serialised), busy time per stream, wall span.
    python tools/step_timeline.py [--trainable full] [--steps 5] > gpurun_out/timeline.json"""
import argparse, json, os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rift_b200.config import MODEL_ZOO
from rift_b200.planning_model import PlanningModel
from rift_b200.trainer import TRAINERS
from rift_b200.synth import synth_state_dict, WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--trainable", default="full"); ap.add_argument("--steps", type=int, default=5); ap.add_argument("--workload", default="cfg2")
args = ap.parse_args()
wl = dict(WORKLOADS[args.workload]); cfg = MODEL_ZOO[wl["model"]](future_steps=wl["future_steps"])
dev = torch.device("cuda", 0)
model = PlanningModel.from_config(cfg, device=dev)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()})
tr = TRAINERS["grpo"](model, trainable_layers=bench.trainable_layers(args.trainable), **bench.TRAINER_KW)
tr.configure_optimizers()
feats, ex = bench.host_batch(cfg, wl, seed=1)
fh, eh = bench.to_torch(feats), bench.to_torch(ex)
batch = bench.make_batch_dict(model.pack(bench.to_device(fh, dev)), bench.to_device({k: eh[k] for k in bench.EXTRA_KEYS}, dev))
for _ in range(6):
    tr.step(batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps):
        tr.step(batch)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and "memcpy" not in e.name.lower()]
agg = collections.defaultdict(lambda: [0, 0.0])
streams = collections.defaultdict(float)
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
for e in evs:
    name = e.name.split("(")[0].replace("void ", "").replace("rift::", "")
    agg[name][0] += 1; agg[name][1] += e.time_range.end - e.time_range.start
    streams[getattr(e, "stream", -1) if hasattr(e, "stream") else -1] += e.time_range.end - e.time_range.start
# union of busy intervals (any kernel running)
iv = sorted((e.time_range.start, e.time_range.end) for e in evs)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None: busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
tot = sum(v[1] for v in agg.values())
out = {"steps": args.steps, "wall_span_us_per_step": (t1 - t0) / args.steps, "sum_kernel_us_per_step": tot / args.steps,
       "any_kernel_busy_us_per_step": busy / args.steps, "kernels_per_step": len(evs) / args.steps,
       "top": [{"kernel": k, "launches_per_step": v[0] / args.steps, "us_per_step": v[1] / args.steps, "avg_us": v[1] / v[0]}
               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]]}
# per-stream view and the launch-ordered list of the LAST step, from the chrome trace (kernel events carry their stream id)
trace_path = os.path.join("gpurun_out", "timeline_trace.json")
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace(trace_path)
tr_ev = [e for e in json.load(open(trace_path))["traceEvents"] if e.get("cat") == "kernel"]
os.remove(trace_path)
tr_ev.sort(key=lambda e: e["ts"])
n_last = len(tr_ev) // args.steps
last = tr_ev[-n_last:]
base = last[0]["ts"]
per_stream = collections.defaultdict(lambda: [0, 0.0])
for e in last:
    per_stream[e["args"].get("stream", -1)][0] += 1
    per_stream[e["args"].get("stream", -1)][1] += e["dur"]
out["last_step_streams"] = {str(k): {"kernels": v[0], "busy_us": v[1]} for k, v in per_stream.items()}
# concurrency profile of the last step: time spent with n kernels resident
pts = sorted([(e["ts"], 1) for e in last] + [(e["ts"] + e["dur"], -1) for e in last])
conc, level, prev = collections.defaultdict(float), 0, pts[0][0]
for t, d in pts:
    conc[level] += t - prev
    prev, level = t, level + d
out["last_step_concurrency_us"] = {str(k): v for k, v in sorted(conc.items())}
with open(os.path.join("gpurun_out", "timeline_last_step.csv"), "w") as f:
    f.write("t_us,dur_us,stream,grid,block,kernel\n")
    for e in last:
        a = e["args"]
        f.write(f"{e['ts'] - base:.1f},{e['dur']:.1f},{a.get('stream', -1)},\"{a.get('grid')}\",\"{a.get('block')}\",\"{e['name'].split('(')[0].replace('void ', '').replace('rift::', '')}\"\n")
print(json.dumps(out, indent=1))
