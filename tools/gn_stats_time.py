"""Tuning aid (GPU): rf_cl_gn_stats on the tensors of a 64-chunk refinement step, for the reduction variant RF_GN_MODE
selects (0 scalar loads, 1 float4 loads, 2 bulk copies through shared memory).  L2 is flushed before each timed call."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# (N, S, C1, C2): skip tensor [N,S,S,S,C1] (+ half-resolution [N,S/2,..,C2], virtually upsampled)
SHAPES = [(16384, 16, 8, 0), (16384, 8, 16, 0), (16384, 8, 32, 64), (16384, 8, 56, 0), (16384, 4, 32, 0), (16384, 4, 64, 128),
          (16384, 2, 64, 0), (64, 64, 16, 0), (64, 32, 16, 0), (64, 32, 12, 0)]
tot = 0.0
for N, S, C1, C2 in SHAPES:
    x = torch.randn(N, S, S, S, C1, device=dev)
    x2 = torch.randn(N, S // 2, S // 2, S // 2, C2, device=dev) if C2 else None
    gamma = torch.ones(C1 + C2, device=dev)
    groups = 8 if (C1 + C2) % 8 == 0 else 4
    for _ in range(2):
        ops.cl_gn_stats(x, gamma, groups, 1e-5, x2=x2)
    ms = []
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.cl_gn_stats(x, gamma, groups, 1e-5, x2=x2); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    gb = 4.0 * (x.numel() + (x2.numel() if C2 else 0)) / 1e9
    tot += m
    print(f"N={N} S={S} C={C1}+{C2}: {m:.4f} ms  {gb / m * 1e3:.0f} GB/s")
    del x, x2
print(f"mode {os.environ.get('RF_GN_MODE', 'default')}: total {tot:.3f} ms")
