"""ncu target: Patch08 conv query encoder on 64 x 8^3 patches of 2500 synthetic 16^3 chunks (Matterport SR 16 -> 64)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops
from retrieval_fuse_b200.pipeline import MATTERPORT_SR16 as CFG, RetrievalPipeline, synthetic_tsdf_batch
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
d = CFG["dataset"]
pipe = RetrievalPipeline(CFG, bank=None, device=dev, weight_seed=1234)
chunks = synthetic_tsdf_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 2500, 16, d["voxel_size_input"], seed=9, device=dev, batch=1024).unsqueeze(1).contiguous()
for _ in range(2):
    q = pipe.encode_queries(chunks)
torch.cuda.synchronize()
print(q.shape, float(q.abs().sum()))
