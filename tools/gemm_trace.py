"""GEMM latency anatomy on the GPU box: per-shape kernel time (L2 flushed / warm / back to back), the in-kernel
%globaltimer trace of CTA 0, the bottleneck switches (RIFT_B200_TC_DBG is read once per process, so each switch
runs in a child process) and cuBLAS bf16 on the same shape for the latency floor.  Profiling aid, not a bench."""
import os, sys, json, subprocess, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200 import _lib

SHAPES = [(4608, 256, 256), (128, 256, 256), (4608, 256, 768), (4608, 1024, 256), (3328, 256, 1024), (40960, 64, 192), (40960, 192, 64),
          (46080, 256, 256), (46080, 128, 256)]


def run(shapes):
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    trace = torch.zeros(64, dtype=torch.int64, device="cuda")
    out = []
    for rows, K, N in shapes:
        x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") * K ** -0.5
        b = torch.randn(N, device="cuda"); y = torch.empty(rows, N, device="cuda")
        scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device="cuda")

        def call(flag):
            _lib.check(L.rift_b200_op_linear_tc(_lib.ptr(x), rows, K, _lib.ptr(w), _lib.ptr(b), N, 1, None, _lib.ptr(y),
                                                _lib.ptr(scratch), scratch.numel(), flag, _lib.stream_ptr()))
        call(1)
        for _ in range(3):
            call(2)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        cold = []
        for _ in range(10):
            flush.zero_(); s.record(); call(2); e.record(); torch.cuda.synchronize(); cold.append(s.elapsed_time(e) * 1e3)
        warm = []
        for _ in range(10):
            s.record(); call(2); e.record(); torch.cuda.synchronize(); warm.append(s.elapsed_time(e) * 1e3)
        s.record()
        for _ in range(50):
            call(2)
        e.record(); torch.cuda.synchronize()
        b2b = s.elapsed_time(e) * 1e3 / 50
        # trace of one cold launch
        L.rift_b200_debug_gemm_trace(_lib.ptr(trace))
        flush.zero_(); trace.zero_(); call(2); torch.cuda.synchronize()
        L.rift_b200_debug_gemm_trace(None)
        t = trace.cpu().tolist()
        t0 = t[0]
        rel = {"setup": t[1] - t0, "first_operands": t[2] - t0, "first_tile_mma_issued": t[3] - t0,
               "tiles": [(t[4 + 2 * i] - t0, t[5 + 2 * i] - t0) for i in range(24) if t[4 + 2 * i]],
               "roles_done": t[60] - t0, "end": t[61] - t0}
        # cuBLAS bf16 (same shape, no epilogue) as the latency floor
        xb, wb = x.bfloat16(), w.bfloat16()
        for _ in range(3):
            torch.matmul(xb, wb.t())
        cb = []
        for _ in range(10):
            flush.zero_(); s.record(); torch.matmul(xb, wb.t()); e.record(); torch.cuda.synchronize(); cb.append(s.elapsed_time(e) * 1e3)
        out.append({"shape": [rows, K, N], "cold_us": min(cold), "warm_us": min(warm), "back_to_back_us": b2b,
                    "cublas_bf16_cold_us": min(cb), "trace_ns": rel})
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        print(json.dumps(run(SHAPES if not os.environ.get('TRACE_FEW') else SHAPES[:2] + SHAPES[6:7])))
        sys.exit(0)
    res = {}
    for dbg in [int(x) for x in os.environ.get('TRACE_DBGS', '0,1,2,4,6').split(',')]:
        env = dict(os.environ, RIFT_B200_TC_DBG=str(dbg))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        try:
            res[f"dbg{dbg}"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res[f"dbg{dbg}"] = {"error": r.stderr[-2000:]}
    print(json.dumps(res))
