#!/usr/bin/env python
"""Development aid (GPU): how many queries of the full-path workload need the exact re-check, and what the lookup costs."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops
from retrieval_fuse_b200.pipeline import FRONT3D_SR as CFG, RefinementPipeline, build_bank_from_targets, synthetic_tsdf_batch, downsample_tsdf_batch
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
d = CFG["dataset"]
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for prims in (5, 24):
    targets = synthetic_tsdf_batch(S, 64, d["voxel_size_target"], seed=100, device=dev, n_prims=prims)
    bank, _ = build_bank_from_targets(CFG, targets, dev, weight_seed=11, batch_patches=4096)
    pipe = RefinementPipeline(CFG, bank, targets, device=dev, weight_seed=1234)
    qt = synthetic_tsdf_batch(64, 64, d["voxel_size_target"], seed=5000, device=dev, n_prims=prims)
    chunks = downsample_tsdf_batch(qt, 8, d["voxel_size_target"], d["voxel_size_input"])
    q = pipe.encode_queries(chunks)
    uniq = torch.unique(bank.emb, dim=0).shape[0]
    print(f"prims {prims}: bank rows {bank.emb.shape[0]}, distinct rows {uniq}, queries {q.shape[0]}, distinct queries {torch.unique(q, dim=0).shape[0]}")
    for label, kw in (("prepared", dict(image=ops.knn_prepare_bank(bank.emb, 0))), ("per-call", dict(method=0)), ("method3", dict(method=3)), ("exact", dict(method=1))):
        for _ in range(2):
            ops.knn_topk(bank.emb, q, 8, stats=(label != "exact"), **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.knn_topk(bank.emb, q, 8, **kw)
        e1.record(); e1.synchronize()
        print(f"   {label:9s} {e0.elapsed_time(e1) / 5:.3f} ms  stats {dict(ops.last_knn_stats) if label != 'exact' else ''}")
