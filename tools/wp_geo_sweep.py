#!/usr/bin/env python
"""Development aid (GPU): times the shifted-window kernel (plain and W-pair variant) over forced item shapes
(RF_HALO_GEO / RF_HALO_FUSED) for one layer.   python tools/wp_geo_sweep.py N S C1 Cout pad"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
N, S, C1, Cout, pad = [int(v) for v in sys.argv[1:6]]
C2 = int(sys.argv[6]) if len(sys.argv) > 6 else 0
So = S + 2 * pad - 2
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(N, S, S, S, C1, device=dev, generator=g)
x2 = torch.randn(N, S // 2, S // 2, S // 2, C2, device=dev, generator=g) if C2 else None
w = torch.randn(Cout, C1 + C2, 3, 3, 3, device=dev, generator=g) / (27 * (C1 + C2)) ** 0.5


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


divs = [d for d in (1, 2, 4, 8, 16, 32) if So % d == 0 and d <= So]
for wp in (False, True):
    if wp and 2 * Cout > 128:
        continue
    img, sw = ops.tc_conv_halo_weight_image(w, C1, C2, wp=wp)
    split = ops.cl_norm_split_halo(x, x2, None, scale=16.0, pad=pad, wp=wp)
    res = []
    shapes = [(0, 1, dt, ht, ln, 0) for dt in divs for ht in divs for ln in (0, 1, 2) if ln < 2 or ht == 16]
    if So <= 8:
        shapes += [(1, G, So, So, ln, hl) for G in (1, 2, 3, 4, 6, 8) for ln in (0, 1) for hl in ((0, 1) if pad else (0,))]
    for shp in shapes:
        for fused in (0, 1):
            os.environ["RF_HALO_GEO"] = ",".join(str(v) for v in shp)
            os.environ["RF_HALO_FUSED"] = str(fused)
            try:
                t = timeit(lambda: ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw)))
            except Exception:
                continue
            res.append((t, shp, fused))
    os.environ.pop("RF_HALO_GEO"); os.environ.pop("RF_HALO_FUSED")
    t0 = timeit(lambda: ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw)))
    res.sort()
    print(f"wp={int(wp)} chooser {t0:.3f} ms; best forced:", " ".join(f"{t:.3f}:{s}f{f}" for t, s, f in res[:6]), flush=True)
