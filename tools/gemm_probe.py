"""Micro-probe: run the tcgen05 linear op on one shape (for ncu captures and quick timing)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200 import _lib
rows, K, N = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (46080, 256, 256))]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
L = _lib.lib()
x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") * K ** -0.5
b = torch.randn(N, device="cuda"); y = torch.empty(rows, N, device="cuda")
scratch = torch.empty(L.rift_b200_op_linear_tc_scratch_bytes(rows, N, K), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
_lib.check(L.rift_b200_op_linear_tc(_lib.ptr(x), rows, K, _lib.ptr(w), _lib.ptr(b), N, 1, None, _lib.ptr(y), _lib.ptr(scratch), scratch.numel(), 1, _lib.stream_ptr()))
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tt = []
for _ in range(reps):
    flush.zero_()
    s.record()
    L.rift_b200_op_linear_tc(_lib.ptr(x), rows, K, _lib.ptr(w), _lib.ptr(b), N, 1, None, _lib.ptr(y), _lib.ptr(scratch), scratch.numel(), 0, _lib.stream_ptr())
    e.record(); torch.cuda.synchronize(); tt.append(s.elapsed_time(e))
ms = min(tt)
print(f"tc {rows}x{K}x{N}: {ms*1e3:.1f} us  {2*rows*K*N/ms/1e9:.1f} TFLOP/s  A+C traffic {(rows*K+rows*N)*4/ms/1e6:.0f} GB/s")
