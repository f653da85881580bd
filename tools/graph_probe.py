"""Experiment: capture one whole policy update in a CUDA graph and compare replay time with eager launches."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from rift_b200.config import MODEL_ZOO
from rift_b200.synth import synth_state_dict, WORKLOADS
from rift_b200.planning_model import PlanningModel
from rift_b200.trainer import TRAINERS

mode = sys.argv[1] if len(sys.argv) > 1 else "full"
wl = dict(WORKLOADS["cfg2"]); cfg = MODEL_ZOO[wl["model"]](future_steps=wl["future_steps"])
dev = torch.device("cuda", 0)
model = PlanningModel.from_config(cfg, device=dev)
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, seed=7).items()})
tr = TRAINERS["grpo"](model, trainable_layers=B.trainable_layers(mode), **B.TRAINER_KW)
tr.configure_optimizers()
feats, ex = B.host_batch(cfg, wl, seed=1)
feats_h, ex_h = B.to_torch(feats), B.to_torch(ex)
batch = B.make_batch_dict(model.pack(B.to_device(feats_h, dev)), B.to_device({k: ex_h[k] for k in ("group_advantage", "group_advantage_mask", "old_group_logits", "ref_group_logits")}, dev))
for _ in range(5):
    tr.step(batch)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    tr.step(batch)
e.record(); torch.cuda.synchronize()
print("eager ms/step", s.elapsed_time(e) / 20)
t0 = time.perf_counter()
for _ in range(20):
    tr.step(batch)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("eager host enqueue ms/step", (t1 - t0) * 1e3 / 20)
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    tr.step(batch)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=st):
        loss = tr.step(batch)
torch.cuda.synchronize()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
s.record()
for _ in range(20):
    g.replay()
e.record(); torch.cuda.synchronize()
print("graph ms/step", s.elapsed_time(e) / 20, "loss", float(loss))
