"""Tuning aid (GPU): the patch attention of a refinement step alone - B chunks, K = 4, nf = 16, retrieval U-Net patches
un-folded (patch_grid 4), channels-last result - against the module-by-module call (Fold3D, NCDHW result).
    python tools/attention_time.py [B] [--once]      (--once: one call of each, for ncu)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops
from retrieval_fuse_b200.model import get_attention_block
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
once = "--once" in sys.argv
nf, K = 16, 4
cfg = dict(nf=nf, attn_patch_extent=4, K=K, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=False,
           attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16)
m = get_attention_block(cfg).to(dev).eval()
g = torch.Generator(device=dev).manual_seed(0)
xb = torch.randn(B, nf, 32, 32, 32, device=dev, generator=g)
feats = torch.randn(B * K * 64, nf, 8, 8, 8, device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(fast):
    if fast:
        return m(xb, feats, patch_grid=4, out_channels_last=True)
    return m(xb, ops.fold3d(feats, 4, 8, nf))


for fast in (False, True):
    reps = 1 if once else 5
    if not once:
        for _ in range(2):
            run(fast)
    ms = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = run(fast); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(f"B={B} {'patches in, channels-last out' if fast else 'Fold3D + attention (NCDHW)      '}: {sorted(ms)[len(ms) // 2]:.3f} ms")
