#!/usr/bin/env python
"""Development aid (GPU): phase timestamps of CTA 0's first items of the shifted-window kernel.
   python tools/wp_phases.py N S C1 Cout pad wp [RF_HALO_GEO] [RF_HALO_FUSED]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops, _lib  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
N, S, C1, Cout, pad, wp = [int(v) for v in sys.argv[1:7]]
if len(sys.argv) > 7 and sys.argv[7] != "-":
    os.environ["RF_HALO_GEO"] = sys.argv[7]
if len(sys.argv) > 8:
    os.environ["RF_HALO_FUSED"] = sys.argv[8]
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(N, S, S, S, C1, device=dev, generator=g)
w = torch.randn(Cout, C1, 3, 3, 3, device=dev, generator=g) / (27 * C1) ** 0.5
img, sw = ops.tc_conv_halo_weight_image(w, C1, 0, wp=bool(wp))
split = ops.cl_norm_split_halo(x, None, None, scale=16.0, pad=pad, wp=bool(wp))
for _ in range(2):
    y = ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw))
e1.record(); e1.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.lib().rf_tc_conv3d_halo_debug_read(ctypes.cast(buf, ctypes.c_void_p))
t0 = buf[0]
print(f"wp={wp} geo={os.environ.get('RF_HALO_GEO')} fused={os.environ.get('RF_HALO_FUSED')} {e0.elapsed_time(e1):.3f} ms")
for it in range(6):
    v = [buf[it * 8 + k] - t0 for k in range(8)]
    print(f"item {it}: issuer start {v[0]} acc-free {v[1]} stage0 {v[2]} issued {v[3]} | epi wait {v[4]} acc-done {v[5]} "
          f"first-ld {v[7]} tile2-ld {buf[48 + it] - t0} stored {v[6]}")
