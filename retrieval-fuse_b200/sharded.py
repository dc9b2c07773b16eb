"""Multi-GPU retrieval: the embedding bank is sharded by rows across the ranks
of one box, chunk batches (and so the queries) are data parallel.

One exchange step per lookup:
  1. all-gather the ranks' query blocks            [W, Ql, 64] fp32
  2. every rank: exact top-2K of ALL queries over ITS shard, with GLOBAL row
     ids and fp64 distances (the canonical order must survive the merge)
  3. all-to-all so that rank r receives, for its own Ql queries, the W per-shard
     candidate lists                               [W, Ql, 2K] x (int32, fp64)
  4. merge the W sorted lists under (d, id), demote same-scene hits, keep K.
Everything else on the hot path (encode, compose, U-Nets, attention) is chunk
data parallel with no communication.  The collectives are torch.distributed
(NCCL over NVLink on the box, gloo in the CPU tests); the compute callbacks
default to the rf_b200 kernels and can be replaced by the tests' oracle."""
import torch
import torch.distributed as dist

from . import ops


class ShardedBankQuery:

    def __init__(self, bank_shard, group=None, topk_fn=None, merge_fn=None, demote_fn=None):
        self.bank = bank_shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._topk = topk_fn or (lambda emb, q, k, off: bank_shard.topk(q, k))  # prepared operand image, reused across calls
        self._merge = merge_fn or ops.knn_merge
        self._demote = demote_fn or ops.knn_demote_rows

    def query(self, q_local, K, query_scene_local=None):
        """q_local [Ql,64] (same Ql on every rank) -> (rows [Ql,K,8], ids [Ql,K])."""
        W, Ql = self.world, q_local.shape[0]
        k2 = min(2 * K, self.bank.n_total)
        if W == 1:
            idx, d = self._topk(self.bank.emb, q_local, k2, self.bank.row_offset)
            return self._demote(idx, d, self.bank.meta, query_scene_local, K)
        q_all = torch.empty((W * Ql, q_local.shape[1]), dtype=q_local.dtype, device=q_local.device)
        dist.all_gather_into_tensor(q_all, q_local.contiguous(), group=self.group)
        n_local = self.bank.emb.shape[0]
        k_local = min(k2, n_local)
        idx, d = self._topk(self.bank.emb, q_all, k_local, self.bank.row_offset)
        if k_local < k2:  # a shard smaller than 2K: pad its lists with +inf sentinels
            pad_i = torch.full((W * Ql, k2 - k_local), 2 ** 31 - 1, dtype=torch.int32, device=idx.device)
            pad_d = torch.full((W * Ql, k2 - k_local), torch.finfo(torch.float64).max, dtype=torch.float64, device=d.device)
            idx, d = torch.cat([idx, pad_i], 1), torch.cat([d, pad_d], 1)
        recv_i = torch.empty((W * Ql, k2), dtype=torch.int32, device=idx.device)
        recv_d = torch.empty((W * Ql, k2), dtype=torch.float64, device=d.device)
        dist.all_to_all_single(recv_i, idx.contiguous(), group=self.group)  # equal splits of Ql rows
        dist.all_to_all_single(recv_d, d.contiguous(), group=self.group)
        midx, md = self._merge(recv_i.reshape(W, Ql, k2), recv_d.reshape(W, Ql, k2))
        return self._demote(midx, md, self.bank.meta, query_scene_local, K)
