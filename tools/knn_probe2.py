#!/usr/bin/env python
"""Development aid (GPU): where the time of a prepared lookup goes (host vs device)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from retrieval_fuse_b200 import ops, _lib
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
N, Q, k = 131073, 4096, 8
g = torch.Generator(device=dev).manual_seed(1)
bank = torch.nn.functional.normalize(torch.randn(N, 64, generator=g, device=dev), dim=1)
q = torch.nn.functional.normalize(torch.randn(Q, 64, generator=g, device=dev), dim=1)
img = ops.knn_prepare_bank(bank, 0)
L = _lib.lib()
idx = torch.empty((Q, k), device=dev, dtype=torch.int32)
d = torch.empty((Q, k), device=dev, dtype=torch.float64)
nbytes = L.rf_knn_prepared_workspace_bytes(Q, N, k, img.method)
print("workspace bytes", nbytes)
ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
st = torch.cuda.current_stream().cuda_stream
def call():
    rc = L.rf_knn_l2_topk_prepared(bank.data_ptr(), N, 0, img.image.data_ptr(), img.method, q.data_ptr(), Q, 64, k, idx.data_ptr(), d.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert rc == 0
for _ in range(3):
    call()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    call()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"C call, preallocated ws: host enqueue {(t1 - t0) / 20 * 1e3:.3f} ms per call, with drain {(t2 - t0) / 20 * 1e3:.3f} ms per call")
t0 = time.perf_counter()
for _ in range(20):
    w = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    del w
torch.cuda.synchronize()
print(f"torch.empty(ws): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    ops.knn_topk(bank, q, k, image=img)
torch.cuda.synchronize()
print(f"ops.knn_topk(image): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    ops.knn_topk(bank, q, k, method=0)
torch.cuda.synchronize()
print(f"ops.knn_topk(per call): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
for qq in (65536, 640000):
    qb = torch.nn.functional.normalize(torch.randn(qq, 64, generator=g, device=dev), dim=1)
    for _ in range(2):
        ops.knn_topk(bank, qb, k, method=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ops.knn_topk(bank, qb, k, method=0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"Q={qq}: host enqueue {(t1 - t0) / 3 * 1e3:.3f} ms, total {(time.perf_counter() - t0) / 3 * 1e3:.3f} ms per call")
