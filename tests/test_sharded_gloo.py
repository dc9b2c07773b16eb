"""world_size-2 (and 3) gloo test of the N>1 path: bank sharding, the query
all-gather, the all-to-all of per-shard candidate lists and the merge must
reproduce the single-bank result bit for bit.  The GPU kernels are replaced by
the oracle here (CPU box); tests/test_gpu_parity.py covers the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _topk_cpu(emb, q, k, off):
    i, d64 = _knn_f64(emb.numpy(), q.numpy(), k)
    return torch.from_numpy(i + off).to(torch.int32), torch.from_numpy(d64)


def _knn_f64(db, q, k):
    """oracle ordering with fp64 distances kept (the ABI's contract between shards)."""
    acc = np.zeros((q.shape[0], db.shape[0]))
    for i in range(db.shape[1]):
        diff = q[:, i:i + 1].astype(np.float64) - db[None, :, i].astype(np.float64)
        acc = acc + diff * diff
    order = np.argsort(acc, axis=1, kind="stable")[:, :k]
    return order.astype(np.int32), np.take_along_axis(acc, order, axis=1)


def _merge_cpu(pi, pd):
    S, Q, k = pi.shape
    ci = pi.permute(1, 0, 2).reshape(Q, S * k).numpy()
    cd = pd.permute(1, 0, 2).reshape(Q, S * k).numpy()
    order = np.lexsort((ci, cd), axis=1)[:, :k]
    return torch.from_numpy(np.take_along_axis(ci, order, 1)), torch.from_numpy(np.take_along_axis(cd, order, 1))


def _demote_cpu(idx, d, meta, qs, K):
    qs_np = np.full(idx.shape[0], -1) if qs is None else qs.numpy()
    oi, od = O.demote_same_scene(idx.numpy(), d.numpy().astype(np.float32), meta[:, 0].numpy().astype(np.int64), qs_np, K)
    return torch.from_numpy(O.mapping_rows(meta.numpy(), oi, od)), torch.from_numpy(oi)


def _worker(rank, world, port, N, Ql, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from retrieval_fuse_b200.sharded import ShardedBankQuery
        rng = np.random.default_rng(0)
        emb = rng.normal(size=(N, 64)).astype(np.float32)
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
        emb[N - 1] = emb[1]  # duplicate rows living in different shards
        meta = np.zeros((N, 7), dtype=np.float32)
        meta[:, 0] = rng.integers(0, 5, size=N)
        q_all = rng.normal(size=(world * Ql, 64)).astype(np.float32)
        q_all /= np.linalg.norm(q_all, axis=1, keepdims=True)
        q_all[0] = emb[1]
        qs_all = rng.integers(-1, 5, size=world * Ql).astype(np.int32)

        class Shard:  # the slice of EmbeddingBank that ShardedBankQuery touches
            pass
        per = (N + world - 1) // world
        lo, hi = min(rank * per, N), min((rank + 1) * per, N)
        sh = Shard()
        sh.emb, sh.meta, sh.row_offset, sh.n_total = torch.from_numpy(emb[lo:hi]), torch.from_numpy(meta), lo, N
        sq = ShardedBankQuery(sh, topk_fn=_topk_cpu, merge_fn=_merge_cpu, demote_fn=_demote_cpu)
        rows, idx = sq.query(torch.from_numpy(q_all[rank * Ql:(rank + 1) * Ql]), K,
                             torch.from_numpy(qs_all[rank * Ql:(rank + 1) * Ql]))
        want_rows, want_idx = O.lookup_rows(emb, meta, q_all[rank * Ql:(rank + 1) * Ql], K, qs_all[rank * Ql:(rank + 1) * Ql])
        ok = np.array_equal(idx.numpy(), want_idx) and np.array_equal(rows.numpy(), want_rows)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 1001), (3, 50), (4, 403)])  # 4 = BASELINE configs[2]: bank over 4 ranks
def test_sharded_query_equals_single_bank(world, N):
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, 17, 4, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert all(out.get(r) for r in range(world)), dict(out)
