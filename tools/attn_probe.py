"""Times (CUDA events, warm, L2-resident like inside the step) the attention forward + backward operator pair at the shapes of
the policy update, and is the command ncu wraps for the attention kernels:
    python tools/attn_probe.py [B S H hd ...]      default: encoder self-attention 64 x 52, decoder m2m 384 x 12, r2r 768 x 6"""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rift_b200 import _lib

P, S = _lib.ptr, _lib.stream_ptr
shapes = [(64, 52, 8, 32), (64, 72, 8, 32), (384, 12, 8, 32), (768, 6, 8, 32)]
if len(sys.argv) > 4:
    shapes = [tuple(int(x) for x in sys.argv[1:5])]
torch.manual_seed(0)
dev = torch.device("cuda", 0)
for B, Sq, H, hd in shapes:
    D = H * hd
    qkv = torch.randn(B * Sq, 3 * D, device=dev)
    d_out = torch.randn(B * Sq, D, device=dev)
    kpm = torch.zeros(B, Sq, dtype=torch.uint8, device=dev)
    kpm[:, Sq - Sq // 5:] = 1
    out, lse, dqkv = torch.empty(B * Sq, D, device=dev), torch.empty(B * H * Sq, device=dev), torch.empty(B * Sq, 3 * D, device=dev)
    def run():
        _lib.check(_lib.lib().rift_b200_op_attention_bwd(P(qkv), P(d_out), B, Sq, H, hd, P(kpm), P(out), P(lse), P(dqkv), S()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    a.record()
    for _ in range(n):
        run()
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"B": B, "S": Sq, "H": H, "hd": hd, "fwd_plus_bwd_us": a.elapsed_time(b) * 1e3 / n}))
