#!/usr/bin/env python
"""Development aid (GPU): the fused front of the first DoubleConv (rf_unet_front.cu) against the separate launches, 16 384 patches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops  # noqa: E402
from retrieval_fuse_b200.model import unet as U  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
blk = U.DoubleConv(1, 16, encoder=True, order="gcr", num_groups=8).to(dev)
x = torch.randn(N, 16, 16, 16, 1, device=dev)


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


for fused in (True, False):
    U.USE_FUSED_FRONT = fused
    print(f"fused={fused}: DoubleConv(1,16) on {N} patches {timeit(lambda: blk.forward_cl(x)):.3f} ms")
c1, c2 = blk.SingleConv1, blk.SingleConv2
host = (c1.conv.weight.float().cpu().contiguous(), float(c1.groupnorm.weight[0]), float(c1.groupnorm.bias[0]))
bufs = {}
for wp in (False, True):
    t = timeit(lambda: ops.unet_front16(x, host[1], host[2], 1e-5, host[0], c2.groupnorm.weight, c2.groupnorm.bias, 8, 1e-5, 16.0, wp=wp, buffers=bufs))
    print(f"front kernel alone wp={wp}: {t:.3f} ms")
