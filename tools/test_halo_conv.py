#!/usr/bin/env python
"""Development aid (GPU): shifted-window conv (rf_tc_conv_halo.cu) vs an fp64 torch evaluation
of GroupNorm -> Conv3d(k3,p1) -> ReLU, and vs the gathering kernel (rf_tc_conv.cu), with timings.

    python tools/test_halo_conv.py [--quick]
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")

# (N, S, C1, C2, Cout, out_ncdhw)
CASES = [
    (5, 8, 16, 0, 32, 0),       # lines mode, one stage
    (3, 8, 32, 64, 56, 0),      # decoder join 96 -> 56 @ 8^3
    (7, 4, 64, 128, 64, 0),     # 192 -> 64 @ 4^3, stacked linear mode, ragged last item
    (4, 4, 32, 0, 32, 0),
    (3, 16, 8, 0, 16, 0),       # pair mode (single channel chunk)
    (2, 8, 56, 0, 16, 1),       # odd chunk count (7 -> 8), NCDHW output
    (9, 2, 64, 0, 128, 0),      # 2^3, Npad 128
    (1, 64, 0, 16, 16, 0),      # decoder: upsampled-only source, slab mode
    (1, 64, 16, 0, 16, 1),
    (2, 32, 16, 0, 16, 0),      # backbone-like
    (2, 16, 12, 24, 24, 0),     # surface-rec style channel counts (nf 12)
]
BIG = [
    (2048, 8, 32, 64, 56, 0),
    (2048, 8, 56, 0, 16, 0),
    (2048, 8, 16, 0, 32, 0),
    (2048, 4, 64, 128, 64, 0),
    (2048, 4, 64, 0, 64, 0),
    (2048, 16, 8, 0, 16, 0),
    (8, 64, 0, 16, 16, 0),
    (8, 64, 16, 0, 16, 0),
]


def reference(x, x2, gamma, beta, groups, w):
    parts = []
    if x is not None:
        parts.append(x.permute(0, 4, 1, 2, 3).double())
    if x2 is not None:
        parts.append(F.interpolate(x2.permute(0, 4, 1, 2, 3).double(), scale_factor=2, mode="nearest"))
    xc = torch.cat(parts, 1)
    y = F.group_norm(xc, groups, gamma.double(), beta.double(), 1e-5)
    return F.relu(F.conv3d(y, w.double(), padding=1))


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def run(case, check=True, old=True):
    N, S, C1, C2, Cout, ncdhw = case
    g = torch.Generator(device=dev).manual_seed(1234 + N + S + C1 + C2 + Cout)
    x = torch.randn(N, S, S, S, C1, device=dev, generator=g) * 1.5 + 0.3 if C1 else None
    x2 = torch.randn(N, S // 2, S // 2, S // 2, C2, device=dev, generator=g) * 0.7 - 0.2 if C2 else None
    C = C1 + C2
    groups = 8 if C >= 8 and C % 8 == 0 else 1
    gamma = torch.rand(C, device=dev, generator=g) + 0.5
    beta = torch.randn(C, device=dev, generator=g) * 0.1
    w = torch.randn(Cout, C, 3, 3, 3, device=dev, generator=g) / (27 * C) ** 0.5
    geo = ops.tc_conv_halo_geometry(N, S, S, S, Cout, C1, C2)
    if geo is None:
        print(f"{case}: UNSUPPORTED by the halo kernel")
        return
    src = x if x is not None else x2
    if x is not None:
        mu, a = ops.cl_gn_stats(x, gamma, groups, 1e-5, x2=x2)
    else:
        mu, a = ops.cl_gn_stats(x2, gamma, groups, 1e-5)
    sa = ops.ACT_SCALE_GN
    img, sw = ops.tc_conv_halo_weight_image(w, C1, C2)

    def halo_split():
        return ops.cl_norm_split_halo(x, x2, (mu, a, beta), scale=sa)

    split = halo_split()

    def halo_conv():
        return ops.tc_conv3d_halo(split, img, None, Cout, act=ops.ACT_RELU, out_ncdhw=bool(ncdhw), out_scale=1.0 / (sa * sw))

    y = halo_conv()
    torch.cuda.synchronize()
    msg = f"{str(case):38s} geo={geo}"
    if check:
        ref = reference(x, x2, gamma, beta, groups, w)
        yy = y if ncdhw else y.permute(0, 4, 1, 2, 3)
        err = float((yy.double() - ref).abs().max())
        scale = float(ref.abs().max())
        msg += f" | max err {err:.2e} (ref max {scale:.2f})"
    t_split, t_conv = timeit(halo_split), timeit(halo_conv)
    flops = 2.0 * N * S ** 3 * 27 * C * Cout
    msg += f" | halo split {t_split:.3f} ms conv {t_conv:.3f} ms ({flops / t_conv / 1e9:.1f} TFLOP/s algorithmic)"
    if old and ops.tc_conv_supported(Cout, C1, C2, 3):
        img_o, sw_o = ops.tc_conv_weight_image(w, C1, C2)

        def old_split():
            xs = ops.cl_norm_split(x, (mu, a, beta), 0, scale=sa) if x is not None else None
            x2s = ops.cl_norm_split(x2, (mu, a, beta), C1, scale=sa) if x2 is not None else None
            return xs, x2s

        xs, x2s = old_split()

        def old_conv():
            return ops.tc_conv3d(xs, x2s, C1, C2, img_o, None, Cout, 3, stride=1, pad=1, act=ops.ACT_RELU,
                                 out_ncdhw=bool(ncdhw), out_scale=1.0 / (sa * sw_o))

        yo = old_conv()
        torch.cuda.synchronize()
        msg += f" | old split {timeit(old_split):.3f} conv {timeit(old_conv):.3f} ms, |halo-old| {float((y - yo).abs().max()):.2e}"
    print(msg, flush=True)


if __name__ == "__main__":
    if "--case" in sys.argv:  # one shape, no checks: the ncu target (profiles/*_halo_conv_*.txt)
        run(tuple(int(v) for v in sys.argv[sys.argv.index("--case") + 1].split(",")), check=False, old=False)
        if "--phases" in sys.argv:  # phase timestamps of CTA 0's first items (cycles relative to the first)
            import ctypes
            from retrieval_fuse_b200 import _lib
            buf = (ctypes.c_longlong * 64)()
            _lib.lib().rf_tc_conv3d_halo_debug_read(ctypes.cast(buf, ctypes.c_void_p))
            t0 = buf[0]
            for it in range(6):
                v = [buf[it * 8 + k] - t0 for k in range(8)]
                print(f"item {it}: issuer start {v[0]} acc-free {v[1]} stage0 {v[2]} issued {v[3]} | epi wait {v[4]} acc-done {v[5]} "
                      f"first-ld {v[7]} tile1-ld {buf[48 + it] - t0} stored {v[6]}")
        sys.exit(0)
    quick = "--quick" in sys.argv
    t0 = time.time()
    for c in CASES:
        run(c)
    if not quick:
        for c in BIG:
            run(c, check=False)
    print(f"done in {time.time() - t0:.1f}s")
