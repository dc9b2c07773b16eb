"""The hot path end to end, as one object per GPU:

    raw low-res chunks --pad-unfold+normalise--> 64 query patches / chunk
      --Patch04/Patch08 encoder + L2 normalise--> 64-d unit queries
      --exact kNN over the embedding bank (fetch 2K, demote, keep K)--> [Q,K,8] rows
      --compose gather from the GPU-resident scene store--> K x 64^3 retrieval volumes
      --U-Net backbone, retrieval U-Net on 16^3 blocks, patch attention, decoder--> 64^3 TSDF

It mirrors what `util/retrieval.py --mode map compose` followed by
`RefinementTrainingModule.forward_full` (trainer/train_refinement.py:108-120)
compute for a batch of chunks, without the disk round trip in between
(SURVEY 8f.1).  Used by bench.py, __graft_entry__.smoke() and the tests.
"""
import math

import numpy as np
import torch

from . import ops
from .model import (get_attention_block, get_decoder, get_retrieval_backbone, get_retrieval_networks,
                    get_unet_backbone)
from .util.retrieval import EmbeddingBank, _encode_normalized

# config/super_resolution/ShapeNetV2/retrieval_008_064.yaml (+ base/retrieval_superresolution.yaml)
SHAPENET_SR_RETRIEVAL = dict(
    task="superresolution", K=4,
    dataset=dict(input_chunk_size=8, target_chunk_size=64, patch_size_input=2, patch_context_input=1,
                 patch_size_target=16, patch_context_target=8, patch_stride=16, voxel_size_input=0.166667,
                 voxel_size_target=0.020834, input_mean=0.34774827082940146, input_std=0.16208995673899929,
                 target_mean=0.060043341595512584, target_std=0.009982546908894512),
    retrieval_model=dict(network_input="2+1", network_target="16+8", nf_input=32, nf_target=8, latent_dim=64),
    dictionary=dict(batch_size=512), query=dict(batch_size=512, K=4),
)
# config/super_resolution/3DFront/{retrieval,refinement}_008_064.yaml
FRONT3D_SR = dict(
    task="superresolution", K=4, nf=16, unet_num_level=4, layer_order="gcr", retrieval_fmaps=16, retrieval_num_level=4,
    attn_patch_extent=4, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=False,
    attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16,
    dataset=dict(input_chunk_size=8, target_chunk_size=64, patch_size_input=2, patch_context_input=1,
                 patch_size_target=16, patch_context_target=8, patch_stride=16, voxel_size_input=0.43334,
                 voxel_size_target=0.054167, input_mean=0.8112343966484424, input_std=0.5094238937427482,
                 target_mean=0.15015658121788053, target_std=0.03573221820637578),
    retrieval_model=dict(network_input="2+1", network_target="16+8", nf_input=32, nf_target=8, latent_dim=64),
    dictionary=dict(batch_size=512), query=dict(batch_size=2048, K=4),
)
# config/super_resolution/Matterport3D/{retrieval,refinement}_016_064.yaml
MATTERPORT_SR16 = dict(
    task="superresolution", K=4, nf=16, unet_num_level=4, layer_order="gcr", retrieval_fmaps=16, retrieval_num_level=4,
    attn_patch_extent=4, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=False,
    attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16,
    dataset=dict(input_chunk_size=16, target_chunk_size=64, patch_size_input=4, patch_context_input=2,
                 patch_size_target=16, patch_context_target=8, patch_stride=16, voxel_size_input=15.0,
                 voxel_size_target=3.75, input_mean=35.62394659115317, input_std=14.58642912987053,
                 target_mean=10.502049923464249, target_std=2.3319665041587627),
    retrieval_model=dict(network_input="4+2", network_target="16+8", nf_input=16, nf_target=8, latent_dim=64),
    dictionary=dict(batch_size=512), query=dict(batch_size=1024, K=4),
)
# config/surface_reconstruction/Matterport3D/{retrieval,refinement}_128_064.yaml: 128^3 occupancy grid of a point
# cloud (util/misc.py:73-78) -> 64^3 TSDF; K = 8 as in BASELINE config 4 (the yaml default is 4)
MATTERPORT_SURFACE = dict(
    task="surface_reconstruction", K=8, nf=12, unet_num_level=5, layer_order="gcr", retrieval_fmaps=12, retrieval_num_level=4,
    attn_patch_extent=4, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=False,
    attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16,
    dataset=dict(input_chunk_size=128, target_chunk_size=64, patch_size_input=32, patch_context_input=8,
                 patch_size_target=16, patch_context_target=4, patch_stride=16, voxel_size_input=0, voxel_size_target=3.75,
                 input_mean=0, input_std=1, target_mean=10.502049923464249, target_std=2.3319665041587627, num_points=1000),
    retrieval_model=dict(network_input="pc_32+8", network_target="16+4", nf_input=10, nf_target=12, latent_dim=64),
    dictionary=dict(batch_size=256), query=dict(batch_size=128, K=8),
)


def f16_trunc(voxel_size):
    """dataset/scene.py:32-33: truncation = float16(3 * voxel_size) as float32."""
    return float(np.float16(voxel_size * 3).astype(np.float32))


def init_unit_gain_(module, seed):
    """Deterministic unit-gain initialisation (U(+-sqrt(3/fan_in)) weights, small
    biases) so that a randomly initialised 24-layer network keeps O(1)
    activations: torch's default init makes the synthetic benchmark's outputs
    collapse towards zero.  Not a reference function."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("sig_scale") or name.endswith("sig_shift"):
                continue
            u = torch.rand(p.shape, generator=g, dtype=torch.float32) * 2 - 1
            if "groupnorm.weight" in name:
                v = 1 + 0.2 * u
            elif "groupnorm.bias" in name or name.endswith(".bias"):
                v = 0.1 * u
            else:
                fan_in = p[0].numel() if p.dim() > 1 else p.numel()
                v = u * math.sqrt(3.0 / max(fan_in, 1))
            p.copy_(v.to(p.device))
    return module


class RetrievalPipeline:
    """encode + kNN (+ compose) for batches of low-res chunks on one GPU."""

    def __init__(self, config, bank, scene_store=None, device=None, weight_seed=None, fenc_input=None,
                 sharded_query=None):
        self.cfg = config
        self.ds = config["dataset"]
        self.K = config["K"]
        self.device = torch.device(device if device is not None else "cuda")
        self.latent_dim = config["retrieval_model"]["latent_dim"]
        if fenc_input is None:
            fenc_input, _ = get_retrieval_networks(config["retrieval_model"])
            if weight_seed is not None:
                init_unit_gain_(fenc_input, weight_seed)
        self.fenc_input = fenc_input.to(self.device).eval()
        self.bank = bank
        self.scene_store = scene_store  # [S, 64,64,64] fp32 raw TSDF of the bank's scenes, on the GPU
        self.sharded_query = sharded_query
        d = self.ds
        self.input_trunc = f16_trunc(d["voxel_size_input"])
        self.target_trunc = f16_trunc(d["voxel_size_target"])
        self.in_kernel = d["patch_size_input"] + 2 * d["patch_context_input"]
        self.in_stride = int(d["patch_stride"] * d["patch_size_input"] / d["patch_size_target"])  # dataset/scene.py:40
        n = d["target_chunk_size"] // d["patch_stride"]
        self.patches_per_chunk = n ** 3
        # unpadded destination extents of the chunk's target patches, x-major (dataset/scene.py:153-160)
        ps = d["patch_size_target"]
        ext = [[x * ps, x * ps + ps, y * ps, y * ps + ps, z * ps, z * ps + ps]
               for x in range(n) for y in range(n) for z in range(n)]
        self.dst_extents = torch.tensor(ext, dtype=torch.int32, device=self.device)
        # the target patches tile the chunk in Unfold3D's patch order and have the edge the retrieval U-Net reads (16):
        # compose can hand the refinement its Unfold3D(16, 1) patches directly (rf_compose_gather_patches)
        self.patch_block = (ps, ps, ps) if (ps == 16 and n * ps == d["target_chunk_size"]) else None
        self._pinned = {}

    # ---- a1 + a5/a6 + a9
    def encode_queries(self, chunks):
        """chunks [B,1,s,s,s] raw low-res SDF on the GPU -> unit queries [B*P, 64].
        Padding with the truncation value, patch extraction and (x-mean)/std
        are one kernel (dataset/scene.py:61, patched_scene_dataset.py:127)."""
        d = self.ds
        patches = ops.unfold3d_pad_stride(chunks, self.in_kernel, d["patch_context_input"], self.in_stride,
                                          self.input_trunc, norm_sub=d["input_mean"], norm_div=d["input_std"])
        return _encode_normalized(self.fenc_input, patches, self.latent_dim)

    # ---- a11
    def lookup(self, q, query_scene=None, method=0):
        """q [Q,64] -> (rows [Q,K,8] fp32, ids [Q,K] int32)."""
        if self.sharded_query is not None:
            return self.sharded_query.query(q, self.K, query_scene)
        return self.bank.query(q, self.K, query_scene, method)

    def expand_scene(self, chunk_scene):
        """per-chunk scene id [B] -> per-query [B*P] (int32, device) or None."""
        if chunk_scene is None:
            return None
        cs = torch.as_tensor(chunk_scene, dtype=torch.int32).to(self.device, non_blocking=True)
        return cs.repeat_interleave(self.patches_per_chunk).contiguous()

    def retrieve(self, chunks, chunk_scene=None, method=0):
        q = self.encode_queries(chunks)
        return self.lookup(q, self.expand_scene(chunk_scene), method)

    # ---- a12
    def compose(self, rows, n_chunks, normalize=False, patches=False):
        """rows [n_chunks*P,K,8] -> [n_chunks,K,64,64,64] raw TSDF (or the
        dataloader-normalised volumes the refinement nets consume).
        patches=True (needs self.patch_block): [n_chunks,K,P,16,16,16] - the Unfold3D(16, 1) patches of those volumes
        (trainer/train_refinement.py:110-111), which refine() accepts in place of the volumes."""
        assert self.scene_store is not None, "compose needs the GPU-resident scene store"
        d = self.ds
        c = d["target_chunk_size"]
        if patches and self.patch_block is None:
            raise ValueError("compose(patches=True): the target patches do not tile the chunk in 16^3 blocks")
        return ops.compose_gather(rows, self.dst_extents, self.scene_store, n_chunks, (c, c, c), self.target_trunc, 1.0,
                                  norm_sub=d["target_mean"] if normalize else 0.0,
                                  norm_div=d["target_std"] if normalize else 0.0,
                                  patch_block=self.patch_block if patches else None)

    # ---- reference-facing host call: host buffers in, host buffers out
    def _pin(self, key, shape, dtype):
        buf = self._pinned.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = buf
        return buf

    def retrieve_host(self, chunks_host, chunk_scene_host=None, method=0):
        """What `RetrievalInterface.get_retrieval_mapping` delivers for a batch of
        chunks, with HOST buffers on both sides: chunks_host [B,1,s,s,s] fp32
        (pinned) -> rows [B*P,K,8] fp32 in pinned host memory.  Includes the
        host->device copy of the inputs and the device->host copy of the rows."""
        x = chunks_host.to(self.device, non_blocking=True)
        rows, _ = self.retrieve(x, chunk_scene_host, method)
        out = self._pin("rows", rows.shape, rows.dtype)
        out.copy_(rows, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out


    def retrieve_host_async(self, chunks_host, chunk_scene_host=None, method=0, slot=0):
        """retrieve_host without the final wait, for a two-deep pipeline over batches: the host->device copy, the
        lookup and the device->host copy of batch i are enqueued on stream `slot` (0 / 1, alternate them) so that
        the copies of one batch overlap the kernels of the next.  Returns (rows in pinned host memory, event); the
        rows are valid after event.synchronize(), and slot s may be reused once its previous event was waited for."""
        if not hasattr(self, "_pipe_streams"):
            # first use: fill the lazily built caches (weight images, the bank's operand image) on the current stream
            # and wait, so that the two pipeline streams never race on their first-use initialisation
            self.retrieve(chunks_host.to(self.device, non_blocking=True), chunk_scene_host, method)
            torch.cuda.current_stream(self.device).synchronize()
            self._pipe_streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        st = self._pipe_streams[slot & 1]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            x = chunks_host.to(self.device, non_blocking=True)
            rows, _ = self.retrieve(x, chunk_scene_host, method)
            out = self._pin(("rows", slot & 1), rows.shape, rows.dtype)
            out.copy_(rows, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        return out, ev


class RefinementPipeline(RetrievalPipeline):
    """+ U-Net backbone, retrieval U-Net, patch attention, decoder
    (trainer/train_refinement.py:108-120, inference part)."""

    def __init__(self, config, bank, scene_store=None, device=None, weight_seed=None, modules=None, **kw):
        super().__init__(config, bank, scene_store, device, weight_seed, **kw)
        cfg = dict(config)
        cfg["dataset_train"] = config["dataset"]
        if modules is None:
            modules = dict(unet_backbone=get_unet_backbone(cfg), retrieval_backbone=get_retrieval_backbone(cfg),
                           patched_attention_block=get_attention_block(cfg), decoder=get_decoder(cfg))
            if weight_seed is not None:
                for i, m in enumerate(modules.values()):
                    init_unit_gain_(m, weight_seed + 1 + i)
        self.unet_backbone = modules["unet_backbone"].to(self.device).eval()
        self.retrieval_backbone = modules["retrieval_backbone"].to(self.device).eval()
        self.patched_attention_block = modules["patched_attention_block"].to(self.device).eval()
        self.decoder = modules["decoder"].to(self.device).eval()

    def state_dicts(self):
        sd = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}
        return dict(fenc_input=sd(self.fenc_input), unet_backbone=sd(self.unet_backbone),
                    retrieval_backbone=sd(self.retrieval_backbone), attention=sd(self.patched_attention_block),
                    decoder=sd(self.decoder))

    def normalize_input(self, chunks):
        """(x - input_mean) / input_std of the whole chunk = the refinement dataloader's
        'input' (patch = chunk, context 0; base/refinement_superresolution.yaml)."""
        d = self.ds
        s = d["input_chunk_size"]
        out = ops.unfold3d_pad_stride(chunks, s, 0, s, 0.0, norm_sub=d["input_mean"], norm_div=d["input_std"])
        return out.reshape(chunks.shape)

    def refine_train(self, x_in, retrieval, gumbel_noise=None):
        """refine() with autograd recording (trainer/train_refinement.py:108-120 inside training_step_full): the same
        modules, routed through their differentiable path; returns (pred, x_back, x_retr, x_attn) attached to the graph."""
        from .model.attention import Fold3D, Unfold3D
        nf = self.retrieval_backbone.nf
        B, S = retrieval.shape[0], retrieval.shape[2]
        with torch.enable_grad():
            x_back = self.unet_backbone(x_in)
            retr = retrieval[:, :self.K].reshape(B * self.K, 1, S, S, S)
            patches = Unfold3D(16, 1)(retr)
            x_retr = Fold3D(4, 8, nf)(self.retrieval_backbone(patches))
            x = self.patched_attention_block(x_back, x_retr, gumbel_noise)
            return self.decoder(x), x_back, x_retr, x

    def refine(self, x_in, retrieval, gumbel_noise=None, intermediates=True):
        """x_in [B,1,s,s,s] normalised, retrieval [B,K',64,64,64] normalised (or its Unfold3D(16, 1) patches
        [B,K',64,16,16,16] from compose(patches=True)) -> pred [B,1,64,64,64] in [-1,1].
        Returns (pred, x_back, x_retr, x_attended).  intermediates=False (inference: refine_graphed, infer*): the
        Fold3D before the attention and the Fold3D + layout change after it are folded into the attention call
        (rf_attention_fuse_patched_fwd) and x_retr / x_attended are returned as None; pred is bit-identical."""
        nf = self.retrieval_backbone.nf
        B, S = retrieval.shape[0], retrieval.shape[2]
        # the input U-Net (a few dozen tiny launches on B chunks) and the retrieval U-Net are independent until the
        # attention: fork the former onto a side stream so that it hides under the latter (also inside a CUDA graph)
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_side_stream"):
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = cur if getattr(self, "serial_streams", False) else self._side_stream  # (serial_streams: per-op timing passes)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            x_in.record_stream(side)
            x_back = self.unet_backbone(x_in)
        if retrieval.dim() == 6:   # compose(patches=True): [B,K',P,16,16,16], already Unfold3D(16, 1) of the volumes
            patches = retrieval[:, :self.K].reshape(-1, 1, *retrieval.shape[3:])
        else:
            retr = retrieval[:, :self.K].reshape(B * self.K, 1, S, S, S)   # get_retrievals (:255-257)
            patches = ops.unfold3d(retr, 16)                                # Unfold3D(16, 1) (:34)
        feats = self.retrieval_backbone(patches)                        # [B*K*64, nf, 8,8,8]
        if (not intermediates and self.decoder.tc_path(nf) and feats.is_cuda
                and not ops.grad_needed(x_in, retrieval, feats, x_back, *self.patched_attention_block.parameters(),
                                        *self.decoder.parameters())):
            cur.wait_stream(side)
            x_back.record_stream(cur)
            x_cl = self.patched_attention_block(x_back, feats, gumbel_noise, patch_grid=4, out_channels_last=True)
            return self.decoder(x_cl, channels_last_input=True), x_back, None, None
        x_retr = ops.fold3d(feats, 4, 8, nf)                            # Fold3D(4, 8, nf) (:37)
        cur.wait_stream(side)
        x_back.record_stream(cur)
        x = self.patched_attention_block(x_back, x_retr, gumbel_noise)
        return self.decoder(x), x_back, x_retr, x

    def refine_graphed(self, x_in, retrieval):
        """refine() replayed from a CUDA graph (one capture per input shape): the ~130 kernel
        launches of a forward are submitted as one graph, which removes the host-side launch
        gaps between the many small kernels of the 8^3 backbone.  Returns pred only."""
        if not hasattr(self, "_graphs"):
            self._graphs, self._graphs_gen = {}, ops.persistent_generation()
        if self._graphs_gen != ops.persistent_generation():
            # a persistent buffer the captured kernels point at (operand planes, weight images) was freed since the
            # capture: every graph is stale
            self._graphs.clear()
            self._graphs_gen = ops.persistent_generation()
        key = (tuple(x_in.shape), tuple(retrieval.shape))
        if key not in self._graphs:
            sx, sr = x_in.clone(), retrieval.clone()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up: weight images, function attributes, allocator, operand planes
                for _ in range(2):
                    self.refine(sx, sr, intermediates=False)
            torch.cuda.current_stream(self.device).wait_stream(side)
            for _ in range(3):
                if self._graphs_gen != ops.persistent_generation():
                    # the warm-up of this shape (or a capture attempt) freed buffers of other shapes: their graphs are stale
                    self._graphs.clear()
                    self._graphs_gen = ops.persistent_generation()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self.refine(sx, sr, intermediates=False)[0]
                if self._graphs_gen == ops.persistent_generation():  # nothing was freed while capturing
                    self._graphs[key] = (graph, sx, sr, out)
                    break
                del graph
            else:
                raise RuntimeError("refine_graphed: persistent buffers kept changing during CUDA graph capture")
        graph, sx, sr, out = self._graphs[key]
        sx.copy_(x_in, non_blocking=True)
        sr.copy_(retrieval, non_blocking=True)
        graph.replay()
        return out

    def pred_to_df(self, pred):
        """network_pred_to_df (trainer/train_refinement.py:242-243) - host-side scaling
        of a result tensor, not part of the kernels' work."""
        return (pred + 1) * self.target_trunc / 2

    def infer(self, chunks, chunk_scene=None, method=0, refine_batch=None, graphed=True, out=None, marks=None):
        """The whole hot path for a batch of raw low-res chunks on the GPU: encode -> kNN (fetch 2K, demote, keep K)
        -> compose (the dataloader's normalisation fused) -> U-Nets + patch attention + decoder, i.e. what
        `util/retrieval.py --mode map compose` followed by `forward_full` (trainer/train_refinement.py:108-120)
        computes, without the .npz round trip.  chunks [B,1,s,s,s] -> pred [B,1,64,64,64] in [-1, 1].
        marks: optional list that receives five CUDA events (start, encoded, looked up, composed, refined).
        The refinement runs in sub-batches of `refine_batch` chunks (CUDA-graph replay per sub-batch shape)."""
        B = chunks.shape[0]
        mark = (lambda: None) if marks is None else (lambda: (marks.append(torch.cuda.Event(enable_timing=True)), marks[-1].record()))
        mark()
        q = self.encode_queries(chunks)
        mark()
        rows, _ = self.lookup(q, self.expand_scene(chunk_scene), method)
        mark()
        retr = self.compose(rows, B, normalize=True, patches=self.patch_block is not None)
        x_in = self.normalize_input(chunks)
        mark()
        rb = refine_batch or B
        c = self.ds["target_chunk_size"]
        pred = out if out is not None else torch.empty((B, 1, c, c, c), dtype=torch.float32, device=self.device)
        for lo in range(0, B, rb):
            hi = min(B, lo + rb)
            if graphed:
                pred[lo:hi].copy_(self.refine_graphed(x_in[lo:hi], retr[lo:hi]))
            else:
                pred[lo:hi].copy_(self.refine(x_in[lo:hi], retr[lo:hi], intermediates=False)[0])
        mark()
        return pred

    def infer_host(self, chunks_host, out_host, chunk_scene=None, method=0, refine_batch=None, graphed=True):
        """infer() with HOST buffers on both sides (pinned): chunks_host [B,1,s,s,s] -> out_host [B,1,64,64,64];
        the host->device copy of the inputs and the device->host copy of the prediction are part of the call."""
        x = chunks_host.to(self.device, non_blocking=True)
        pred = self.infer(x, chunk_scene, method, refine_batch, graphed)
        out_host.copy_(pred, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host

    def infer_host_async(self, chunks_host, out_host, chunk_scene=None, method=0, refine_batch=None, graphed=True, slot=0):
        """infer_host without the final wait, for a two-deep pipeline over batches: the kernels of every batch run on the
        current stream, the device->host copy of batch i's prediction runs on copy stream `slot` (0 / 1, alternate them,
        with one out_host buffer per slot) and overlaps the kernels of batch i+1.  Returns (out_host, event): the
        prediction is in host memory after event.synchronize(); a slot may be reused once its previous event was waited
        for."""
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_copy_streams"):
            self._copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
            self._copy_events = [None, None]
            self._pred_slots = {}
        k = slot & 1
        B, c = chunks_host.shape[0], self.ds["target_chunk_size"]
        key = (k, B)
        if key not in self._pred_slots:
            self._pred_slots[key] = torch.empty((B, 1, c, c, c), dtype=torch.float32, device=self.device)
        pred = self._pred_slots[key]
        if self._copy_events[k] is not None:
            cur.wait_event(self._copy_events[k])  # the slot's device buffer is still being copied out
        x = chunks_host.to(self.device, non_blocking=True)
        self.infer(x, chunk_scene, method, refine_batch, graphed, out=pred)
        done = torch.cuda.Event()
        done.record(cur)
        cs = self._copy_streams[k]
        cs.wait_event(done)
        with torch.cuda.stream(cs):
            out_host.copy_(pred, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._copy_events[k] = ev
        return out_host, ev

    def retrieve_and_refine(self, chunks, chunk_scene=None, method=0):
        B = chunks.shape[0]
        rows, idx = self.retrieve(chunks, chunk_scene, method)
        retr_raw = self.compose(rows, B, normalize=False)
        retr = self.compose(rows, B, normalize=True)
        pred, x_back, x_retr, x = self.refine(self.normalize_input(chunks), retr)
        return dict(knn_rows=rows, knn_idx=idx, retrieval=retr_raw, pred=pred, pred_df=self.pred_to_df(pred))


# ---------------------------------------------------------------------------
# synthetic data on the device (setup only; not part of any timed region)
# ---------------------------------------------------------------------------

def synthetic_tsdf_batch(n, size, voxel_size, seed, device, n_prims=5, batch=128):
    """n random TSDF chunks [n, size,size,size]: unsigned distance to a union of
    sphere shells and planes, in voxel units * voxel_size, clamped to the
    float16 truncation.  Generated with torch ops on the device,
    `batch` chunks at a time (setup only)."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    trunc = f16_trunc(voxel_size)
    ax = (torch.arange(size, dtype=torch.float32, device=device) + 0.5)
    X, Y, Z = [t[None] for t in torch.meshgrid(ax, ax, ax, indexing="ij")]
    out = torch.empty((n, size, size, size), dtype=torch.float32, device=device)
    par = torch.rand((n, n_prims, 8), generator=g).to(device)
    for lo in range(0, n, batch):
        p = par[lo: lo + batch]
        d = torch.full((p.shape[0], size, size, size), 1e9, dtype=torch.float32, device=device)
        for j in range(n_prims):
            v = lambda c: p[:, j, c].reshape(-1, 1, 1, 1)
            cx, cy, cz = v(0) * size, v(1) * size, v(2) * size
            r = (0.08 + 0.25 * v(4)) * size
            sphere = (torch.sqrt((X - cx) ** 2 + (Y - cy) ** 2 + (Z - cz) ** 2) - r).abs()
            nrm = p[:, j, 5:8] * 2 - 1
            nrm = nrm / nrm.norm(dim=1, keepdim=True).clamp_min(1e-6)
            w = lambda c: nrm[:, c].reshape(-1, 1, 1, 1)
            plane = ((X - cx) * w(0) + (Y - cy) * w(1) + (Z - cz) * w(2)).abs()
            d = torch.minimum(d, torch.where(v(3) < 0.6, sphere, plane))
        out[lo: lo + batch] = torch.clamp(d * voxel_size, max=trunc)
    return out


def downsample_tsdf_batch(target, factor, voxel_size_target, voxel_size_input):
    """[n,S,S,S] -> [n,1,S/f,S/f,S/f]: min-pool then re-truncate in input units."""
    n, S = target.shape[0], target.shape[1]
    s = S // factor
    v = target.reshape(n, s, factor, s, factor, s, factor).amin(dim=(2, 4, 6))
    v = v / voxel_size_target * voxel_size_input
    return torch.clamp(v, max=f16_trunc(voxel_size_input)).unsqueeze(1).contiguous()


def build_bank_from_targets(config, targets, device, fenc_target=None, weight_seed=None, batch_patches=512,
                            scene_offset=0, n_scenes_total=None):
    """create_dictionary (util/retrieval.py:29-55) for GPU-resident scenes:
    targets [S,64,64,64] raw TSDF -> EmbeddingBank with S*64 rows + the sentinel.
    With scene_offset / n_scenes_total the result is the row shard of a larger
    bank (scenes [scene_offset, scene_offset+S) of n_scenes_total; the sentinel
    row belongs to the last shard; meta is always the full table)."""
    d = config["dataset"]
    lat = config["retrieval_model"]["latent_dim"]
    if fenc_target is None:
        _, fenc_target = get_retrieval_networks(config["retrieval_model"])
        if weight_seed is not None:
            init_unit_gain_(fenc_target, weight_seed)
    fenc_target = fenc_target.to(device).eval()
    ps, ctx = d["patch_size_target"], d["patch_context_target"]
    n = d["target_chunk_size"] // d["patch_stride"]
    P = n ** 3
    S = targets.shape[0]
    S_tot = n_scenes_total if n_scenes_total is not None else S
    last = scene_offset + S == S_tot
    trunc = f16_trunc(d["voxel_size_target"])
    emb = torch.empty((S * P + (1 if last else 0), lat), dtype=torch.float32, device=device)
    per = max(1, batch_patches // P)
    for lo in range(0, S, per):
        hi = min(lo + per, S)
        patches = ops.unfold3d_pad_stride(targets[lo:hi].unsqueeze(1), ps + 2 * ctx, ctx, d["patch_stride"], trunc,
                                          norm_sub=d["target_mean"], norm_div=d["target_std"])
        emb[lo * P: hi * P] = _encode_normalized(fenc_target, patches, lat)
    if last:
        ones = torch.ones((1, 1) + (ps + 2 * ctx,) * 3, dtype=torch.float32, device=device)
        emb[S * P] = _encode_normalized(fenc_target, ones, lat)[0]  # get_zero_patch_entry
    ext = torch.tensor([[x * ps, x * ps + ps, y * ps, y * ps + ps, z * ps, z * ps + ps]
                        for x in range(n) for y in range(n) for z in range(n)], dtype=torch.float32, device=device)
    meta = torch.empty((S_tot * P + 1, 7), dtype=torch.float32, device=device)
    meta[:S_tot * P, 0] = torch.arange(S_tot, device=device, dtype=torch.float32).repeat_interleave(P)
    meta[:S_tot * P, 1:] = ext.repeat(S_tot, 1)
    meta[S_tot * P] = torch.tensor([-1, 0, ps, 0, ps, 0, ps], dtype=torch.float32, device=device)
    bank = EmbeddingBank(emb, meta, [f"scene{i:05d}" for i in range(S_tot)], row_offset=scene_offset * P,
                         n_total=S_tot * P + 1)
    return bank, fenc_target


def build_synthetic_world(n_bank_scenes, n_query_chunks, seed, device, config=None):
    """Bank + scene store + query chunks for smoke / tests (3DFront SR 8->64)."""
    config = config or FRONT3D_SR
    d = config["dataset"]
    targets = synthetic_tsdf_batch(n_bank_scenes, d["target_chunk_size"], d["voxel_size_target"], seed, device)
    bank, fenc_target = build_bank_from_targets(config, targets, device, weight_seed=seed + 7)
    factor = d["target_chunk_size"] // d["input_chunk_size"]
    qt = synthetic_tsdf_batch(n_query_chunks, d["target_chunk_size"], d["voxel_size_target"], seed + 1000, device)
    qt[0] = targets[0]  # one query chunk comes from a bank scene: exercises the demotion
    q_in = downsample_tsdf_batch(qt, factor, d["voxel_size_target"], d["voxel_size_input"])
    scene = [0] + [-1] * (n_query_chunks - 1)
    return dict(config=config, bank=bank, scene_store=targets, query_inputs=q_in, query_targets=qt,
                query_scene=scene, fenc_target=fenc_target)
