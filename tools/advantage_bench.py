"""BASELINE config 5: group-relative advantage kernel micro-bench, 1k..1M groups, HBM GB/s vs measured peak.
Algorithmic bytes = 16 * G per group (fp64 returns in, fp64 advantages out).  CPU rows: the reference's literal
per-group numpy expression (traj_evaluator.py:466-469) and a vectorised axis=1 numpy variant."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200 import functional as F
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for G in [int(x) for x in os.environ.get("ADV_G", "12,72").split(",")]:
    for lg in [int(x) for x in os.environ.get("ADV_LG", "10,12,14,16,18,20").split(",")]:
        n = 1 << lg
        ret = (torch.randn(n, G, dtype=torch.float64, device="cuda") * 20 - 5)
        F.group_advantage(ret)
        ts = []
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(10):
            flush.zero_(); s.record(); out = F.group_advantage(ret); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        gbs = 16.0 * G * n / (ms * 1e-3) / 1e9
        row = {"G": G, "n_groups": n, "us": ms * 1e3, "GBps": gbs, "frac_of_measured_hbm": gbs / peak}
        if lg <= 14:
            r = ret.cpu().numpy()
            t0 = time.perf_counter()
            for i in range(r.shape[0]):
                x = r[i]; _ = (x - np.mean(x)) / (np.std(x) + 1e-5)
            row["cpu_loop_us"] = (time.perf_counter() - t0) * 1e6
            t0 = time.perf_counter()
            _ = (r - r.mean(1, keepdims=True)) / (r.std(1, keepdims=True) + 1e-5)
            row["cpu_vectorised_us"] = (time.perf_counter() - t0) * 1e6
        rows.append(row)
        print(json.dumps(row), flush=True)
