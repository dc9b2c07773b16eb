#!/usr/bin/env python
"""Development aid (GPU): convolution with the next layer's GroupNorm in its epilogue (rf_tc_conv3d_halo_gn_fwd) against
convolution + statistics + split as separate launches.   python tools/gn_epi_time.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops, _lib  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")


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


for N, S, C1, C2, Cout in [(16384, 8, 16, 0, 16), (16384, 8, 32, 64, 56)]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(N, S, S, S, C1, device=dev, generator=g)
    x2 = torch.randn(N, S // 2, S // 2, S // 2, C2, device=dev, generator=g) if C2 else None
    w = torch.randn(Cout, C1 + C2, 3, 3, 3, device=dev, generator=g) / (27 * (C1 + C2)) ** 0.5
    gam, bet = torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)
    img, sw = ops.tc_conv_halo_weight_image(w, C1, C2)
    split = ops.cl_norm_split_halo(x, x2, None, scale=16.0)
    bufs = {}
    t_conv = timeit(lambda: ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw)))
    y = ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw))
    t_gn = timeit(lambda: ops.cl_gn_stats(y, gam, 8, 1e-5))
    mu, a = ops.cl_gn_stats(y, gam, 8, 1e-5)
    t_split = timeit(lambda: ops.cl_norm_split_halo(y, None, (mu, a, bet), scale=16.0, buffers=bufs))
    t_fused = timeit(lambda: ops.tc_conv3d_halo_gn(split, img, None, Cout, gam, bet, 8, 1e-5, 16.0, act=ops.ACT_RELU, out_scale=1.0 / (16.0 * sw), buffers=bufs))
    print(f"N={N} S={S} C={C1}+{C2} Cout={Cout}: conv {t_conv:.3f} + stats {t_gn:.3f} + split {t_split:.3f} = {t_conv + t_gn + t_split:.3f} ms | fused {t_fused:.3f} ms", flush=True)
    buf = (ctypes.c_longlong * 64)()
    _lib.lib().rf_tc_conv3d_halo_debug_read(ctypes.cast(buf, ctypes.c_void_p))
    t0 = buf[0]
    for it in range(2, 5):
        v = [buf[it * 8 + k] - t0 for k in range(8)]
        print(f"  item {it}: issuer start {v[0]} acc-free {v[1]} stage0 {v[2]} issued {v[3]} | epi wait {v[4]} acc-done {v[5]}")
