"""ncu target: Patch32 dictionary encoder on 1024 synthetic 32^3 target patches (16 scenes)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops
from retrieval_fuse_b200.model import get_retrieval_networks
from retrieval_fuse_b200.pipeline import SHAPENET_SR_RETRIEVAL as CFG, init_unit_gain_, synthetic_tsdf_batch, f16_trunc, _encode_normalized
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
d = CFG["dataset"]
n_sc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
targets = synthetic_tsdf_batch(n_sc, 64, d["voxel_size_target"], seed=100, device=dev)
_, fenc = get_retrieval_networks(CFG["retrieval_model"])
init_unit_gain_(fenc, 11)
fenc = fenc.to(dev).eval()
ps, ctx = d["patch_size_target"], d["patch_context_target"]
p = ops.unfold3d_pad_stride(targets.unsqueeze(1), ps + 2 * ctx, ctx, d["patch_stride"], f16_trunc(d["voxel_size_target"]),
                            norm_sub=d["target_mean"], norm_div=d["target_std"])
for _ in range(2):
    e = _encode_normalized(fenc, p, 64)
torch.cuda.synchronize()
print(e.shape, float(e.abs().sum()))
