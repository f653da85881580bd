"""Micro-timing of the fused sub-block kernels (CUDA events, L2 flushed): python tools/fused_probe.py"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200 import _lib
L = _lib.lib(); P = _lib.ptr; S = _lib.stream_ptr
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=15):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        flush.zero_(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))
for rows, D, Hd, act in [(4608, 256, 1024, 1), (3328, 256, 1024, 2), (40960, 64, 192, 2), (20480, 128, 384, 2), (10240, 256, 768, 2)]:
    x = torch.randn(rows, D, device="cuda"); g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
    w1 = torch.randn(Hd, D, device="cuda") * D ** -0.5; b1 = torch.zeros(Hd, device="cuda")
    w2 = torch.randn(D, Hd, device="cuda") * Hd ** -0.5; b2 = torch.zeros(D, device="cuda")
    y = torch.empty(rows, D, device="cuda"); mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    t2h, t2l = (torch.empty(rows, D, dtype=torch.int16, device="cuda") for _ in range(2))
    hmh, hml = (torch.empty(rows, Hd, dtype=torch.int16, device="cuda") for _ in range(2))
    hpre = torch.empty(rows, Hd, device="cuda")
    scratch = torch.empty(L.rift_b200_op_fused_mlp_scratch_bytes(D, Hd), dtype=torch.uint8, device="cuda")
    def run(tape, resplit=0):
        _lib.check(L.rift_b200_op_fused_mlp(P(x), rows, D, Hd, act, P(g), P(b), P(w1), P(b1), P(w2), P(b2), P(y),
                                            P(mean) if tape else None, P(rstd) if tape else None, P(t2h) if tape else None, P(t2l) if tape else None,
                                            P(hpre) if tape else None, P(hmh) if tape else None, P(hml) if tape else None,
                                            P(scratch), scratch.numel(), resplit, S()), "fused_mlp")
    run(True, 1); torch.cuda.synchronize()
    tr = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.rift_b200_debug_fused_trace(P(tr))
    traces = {}
    for tape in (False, True):
        flush.zero_(); run(tape); torch.cuda.synchronize()
        t = tr.cpu().numpy(); traces["tape" if tape else "notape"] = [round((int(v) - int(t[0])) / 1e3, 2) for v in t[:13]]
    L.rift_b200_debug_fused_trace(None)
    print(json.dumps({"rows": rows, "D": D, "phase_us_since_setup": traces}))
    print(json.dumps({"rows": rows, "D": D, "Hd": Hd, "act": act, "fused_us_notape": timeit(lambda: run(False)), "fused_us_tape": timeit(lambda: run(True))}), flush=True)
