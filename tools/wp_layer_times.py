#!/usr/bin/env python
"""Development aid (GPU): plain shifted-window kernel vs its W-pair variant (two output voxels per GEMM row) on the
layer shapes of a full-path step (16 384 retrieval patches) and of the conv patch encoders: split + conv times, max
difference of the outputs.

    python tools/wp_layer_times.py [N_scale]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops, _lib  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
sc = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
# (N, S, C1, C2, Cout, pad)
CASES = [(16384, 16, 8, 0, 16, 1), (16384, 8, 16, 0, 16, 1), (16384, 8, 16, 0, 32, 1), (16384, 4, 32, 0, 32, 1), (16384, 8, 56, 0, 16, 1),
         (128, 64, 16, 0, 16, 1), (4096, 30, 8, 0, 16, 0), (1024, 46, 16, 0, 32, 0), (160000, 6, 8, 0, 16, 0)]


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


for N, S, C1, C2, Cout, pad in CASES:
    N = max(1, int(N * sc))
    g = torch.Generator(device=dev).manual_seed(N + S + C1 + Cout)
    x = torch.randn(N, S, S, S, C1, device=dev, generator=g) * 1.5 + 0.3
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, device=dev, generator=g) / (27 * (C1 + C2)) ** 0.5
    o, scores = (ctypes.c_int * 16)(), (ctypes.c_double * 2)()
    _lib.lib().rf_tc_conv3d_halo_wp_geometry(N, S, S, S, Cout, C1, C2, pad, o, scores)
    res = {}
    for wp in (False, True):
        if wp and scores[0] < 0:
            continue
        img, sw = ops.tc_conv_halo_weight_image(w, C1, C2, wp=wp)
        bufs = {}
        t_split = timeit(lambda: ops.cl_norm_split_halo(x, None, None, scale=16.0, pad=pad, buffers=bufs, wp=wp))
        split = ops.cl_norm_split_halo(x, None, None, scale=16.0, pad=pad, buffers=bufs, wp=wp)
        t_conv = timeit(lambda: ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw)))
        y = ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw))
        res[wp] = (t_split, t_conv, y)
        del split, bufs
    line = f"N={N} S={S} C={C1} Cout={Cout} pad={pad}: plain split {res[False][0]:.3f} conv {res[False][1]:.3f} geo {list(o[8:14])}"
    if True in res:
        d = float((res[True][2] - res[False][2]).abs().max())
        line += f" | wp split {res[True][0]:.3f} conv {res[True][1]:.3f} geo {list(o[:6])} model x{scores[0] / scores[1]:.2f} maxdiff {d:.2e}"
    print(line, flush=True)
    del res, x
    torch.cuda.empty_cache()
