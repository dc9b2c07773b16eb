"""Profiling aid: the re-indexing kernels at the bench sizes, one launch each, for
`ncu --set full -k regex:"fold_unfold|pad_unfold|compose"` (profiles/r01_reindex_*).  Prints CUDA-event times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retrieval_fuse_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.set_grad_enabled(False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


ONCE = "--once" in sys.argv  # one launch per kernel (for ncu)


def timed(name, fn, nbytes, iters=5):
    if ONCE:
        fn()
        return
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        torch.cuda._sleep(300000)  # the host enqueues e0 / kernel / e1 while the GPU spins: no launch latency in the pair
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    print(f"{name:58s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)


x = torch.randn(40, 16, 32, 32, 32, device=dev)
u8 = ops.unfold3d(x, 8)
timed("Unfold3D(8,16) [40,16,32^3]", lambda: ops.unfold3d(x, 8), 2 * x.numel() * 4)
timed("Fold3D(4,8,16)", lambda: ops.fold3d(u8, 4, 8, 16), 2 * x.numel() * 4)
u2 = ops.unfold3d(x, 2)
timed("Unfold3D(2,16) [40,16,32^3]", lambda: ops.unfold3d(x, 2), 2 * x.numel() * 4)
timed("Fold3D(16,2,16)", lambda: ops.fold3d(u2, 16, 2, 16), 2 * x.numel() * 4)
v = torch.randn(40, 1, 64, 64, 64, device=dev)
timed("Unfold3D(16,1) [40,1,64^3]", lambda: ops.unfold3d(v, 16), 2 * v.numel() * 4)
c8 = torch.randn(10000, 1, 8, 8, 8, device=dev)
o = ops.unfold3d_pad_stride(c8, 4, 1, 2, 3.0, norm_sub=0.5, norm_div=1.5)
timed("pad+unfold+normalise 10000 x 8^3 -> 64 x 4^3", lambda: ops.unfold3d_pad_stride(c8, 4, 1, 2, 3.0, norm_sub=0.5, norm_div=1.5),
      (c8.numel() + o.numel()) * 4)
c16 = torch.randn(2500, 1, 16, 16, 16, device=dev)
o = ops.unfold3d_pad_stride(c16, 8, 2, 4, 3.0, norm_sub=0.5, norm_div=1.5)
timed("pad+unfold+normalise 2500 x 16^3 -> 64 x 8^3", lambda: ops.unfold3d_pad_stride(c16, 8, 2, 4, 3.0, norm_sub=0.5, norm_div=1.5),
      (c16.numel() + o.numel()) * 4)
t64 = torch.randn(64, 1, 64, 64, 64, device=dev)
o = ops.unfold3d_pad_stride(t64, 32, 8, 16, 3.0, norm_sub=0.5, norm_div=1.5)
timed("pad+unfold+normalise 64 x 64^3 -> 64 x 32^3", lambda: ops.unfold3d_pad_stride(t64, 32, 8, 16, 3.0, norm_sub=0.5, norm_div=1.5),
      (t64.numel() + o.numel()) * 4)
del o
# compose: 64 chunks x K = 4 x 64 blocks of 16^3 gathered from a 256-scene store
S, B, K = 256, 64, 4
store = torch.randn(S, 64, 64, 64, device=dev)
g = torch.Generator(device="cpu").manual_seed(0)
rows = torch.zeros(B * 64, K, 8)
rows[:, :, 0] = torch.randint(0, S, (B * 64, K), generator=g).float()
st = torch.randint(0, 4, (B * 64, K, 3), generator=g).float() * 16
rows[:, :, 1], rows[:, :, 3], rows[:, :, 5] = st[..., 0], st[..., 1], st[..., 2]
rows[:, :, 2], rows[:, :, 4], rows[:, :, 6] = st[..., 0] + 16, st[..., 1] + 16, st[..., 2] + 16
rows = rows.to(dev)
ext = torch.tensor([[x0, x0 + 16, y0, y0 + 16, z0, z0 + 16] for x0 in range(0, 64, 16) for y0 in range(0, 64, 16) for z0 in range(0, 64, 16)],
                   dtype=torch.int32, device=dev)
timed("compose gather 64 chunks, K=4", lambda: ops.compose_gather(rows, ext, store, B, (64, 64, 64), 3.0, 1.0, norm_sub=0.5, norm_div=1.5),
      2.0 * B * K * 64 ** 3 * 4)
