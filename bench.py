#!/usr/bin/env python
"""bench.py - chunks/sec of the RetrievalFuse hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU arm (oracle port)

Default workload (`--workload full`, BASELINE.json's metric: encode + kNN + attn-fuse): per step and GPU, `--chunks`
raw 8^3 chunks go through the WHOLE path - pad/unfold/normalise + Patch04 query encoder, exact kNN against a bank of
2 048 encoded synthetic 64^3 targets (131 072 rows + sentinel; fetch 2K = 8, demote, keep K = 4), compose from the
GPU-resident scene store, input U-Net + retrieval U-Net on 16^3 blocks + patch attention + decoder - to 64^3 TSDF
predictions (3DFront super_resolution 008 -> 064 shapes, BASELINE configs[2]).  At N > 1 every rank brings its own
chunks (weak scaling) and the bank is sharded by rows inside groups of `--bank-shards` ranks (all-gather of queries,
all-to-all of per-shard top-2K lists, merge); everything else is chunk data parallel.
Other workloads: `retrieval` (configs[1]: encode + kNN only, 10 000 chunks per step), `surface` (configs[3]: Matterport
surface reconstruction, 128^3 occupancy grid -> 64^3, K = 8), `sweep` (configs[4]: 1 M-row isotropic bank, k sweep +
attention sweep), `refine` (refinement forward alone), `stages` (every SURVEY 8d stage against its own roofline).

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "64^3 TSDF chunks/sec (encode+kNN+attn-fuse)"
METRIC_RETRIEVAL = "64^3 TSDF chunks/sec (encode+kNN only)"
UNIT = "chunks/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None  # set when fd 1 has been redirected to stderr (multi-rank runs)


def emit(line: dict):
    """The one JSON line of the contract, on the real stdout."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "surface", "retrieval", "sweep", "refine", "stages"])
    ap.add_argument("--chunks", type=int, default=0, help="chunks per step per GPU (0 = workload default: full 64, surface 16, "
                                                         "retrieval 10000)")
    ap.add_argument("--bank-scenes", type=int, default=2048, help="64^3 scenes encoded into the bank (x64 rows)")
    ap.add_argument("--knn-method", type=int, default=0, help="0 auto, 1 exact fp64 sweep, 2 tcgen05 fp16, 3 tcgen05 bf16x3")
    ap.add_argument("--bank", default="encoded", choices=["encoded", "random"],
                    help="encoded: Patch32 embeddings of synthetic 64^3 targets (the workload); random: unit Gaussian "
                         "bank AND queries (profiling aid, BASELINE config 5 style)")
    ap.add_argument("--bank-shards", type=int, default=0,
                    help="N > 1: row shards of the bank (0 = auto: 2, the smallest sharding that keeps the NCCL exchange "
                         "on the data path; the N / shards groups of ranks each hold one full copy and exchange "
                         "inside the group).  --bank-shards N = one shard per rank")
    ap.add_argument("--refine-batch", type=int, default=0, help="chunks per refinement sub-batch (0 = workload default: full 16, "
                                                               "surface 4, refine 8)")
    ap.add_argument("--no-cuda-graph", action="store_true", help="refine workload: launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-sample-chunks", type=int, default=0, help="chunks of the workload the CPU arm runs per step "
                                                                    "(0 = workload default: retrieval 2048, full 64 = the whole step, surface 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    full_like = a.workload in ("full", "surface")
    if a.chunks <= 0:
        a.chunks = {"full": 64, "surface": 16}.get(a.workload, 10000)
    if a.refine_batch <= 0:
        a.refine_batch = {"full": 64, "surface": 4}.get(a.workload, 8)
    if a.cpu_sample_chunks <= 0:
        a.cpu_sample_chunks = {"full": 64, "surface": 4}.get(a.workload, 2048)
    if a.workload == "surface" and a.bank_scenes == 2048:
        a.bank_scenes = 512
    a.full_like = full_like
    return a


# ---------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.stop_flag = None, False

    def _nvml_start(self):
        """NVML in-process, one sample every ~3 ms (nvidia-smi -lms delivers its first row after ~100 ms, longer
        than the timed region of the refine workload)."""
        import pynvml as N
        N.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)  # fails here, not in the thread, if unsupported

        def loop():
            while not self.stop_flag:
                try:
                    sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
                    r = int(get_reasons(h))
                    self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active" for k in
                                                           ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                except Exception:
                    pass
                time.sleep(0.003)
        self.nvml = threading.Thread(target=loop, daemon=True)
        self.nvml.start()

    def start(self):
        if os.environ.get("RF_BENCH_NO_CLOCKS"):  # debugging aid: no sampler thread at all
            return
        try:
            self._nvml_start()
            return
        except Exception as e:
            log("[clocks] NVML sampler unavailable, using nvidia-smi:", e)
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing: report nothing rather than fail the bench
            log("[clocks] sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is None and self.proc is None:
            return None
        if self.nvml is not None:
            self.stop_flag = True
            self.nvml.join(timeout=1)
        elif self.proc is None:
            return None
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the same workload on the host cores
# ---------------------------------------------------------------------------

def cpu_retrieval_sample(cfg, sd_fenc, bank_emb, bank_meta, chunks_np, threads):
    """One bounded sample through the reference's CPU arithmetic: patch extraction +
    Patch04 (torch CPU, the reference's own ops) + the kNN baseline BASELINE.md
    names (fp32 torch.cdist + topk brute force; FLANN is not installable) +
    demotion.  Returns seconds."""
    from oracle import rf_oracle as O
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    with torch.no_grad():
        q = torch.from_numpy(O.encode_chunk_queries(cfg, sd_fenc, chunks_np))
        db = torch.from_numpy(bank_emb)
        K2 = 2 * cfg["K"]
        idx = torch.empty((q.shape[0], K2), dtype=torch.int64)
        dist = torch.empty((q.shape[0], K2), dtype=torch.float32)
        for lo in range(0, q.shape[0], 1024):  # the reference queries in batches of 1024 (util/retrieval.py:86)
            d2 = torch.cdist(q[lo:lo + 1024], db) ** 2
            v, i = torch.topk(d2, K2, dim=1, largest=False)
            idx[lo:lo + 1024], dist[lo:lo + 1024] = i, v
        qs = np.full(q.shape[0], -1)
        oi, od = O.demote_same_scene(idx.numpy().astype(np.int32), dist.numpy(), bank_meta[:, 0].astype(np.int64), qs, cfg["K"])
        O.mapping_rows(bank_meta, oi, od)
    return time.perf_counter() - t0


def cpu_refine_sample(cfg, sds, chunks_np, retr_np, threads):
    from oracle import rf_oracle as O
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.refine_chunks(cfg, sds, chunks_np, retr_np)
    return time.perf_counter() - t0


def cpu_full_sample(cfg, sds, bank_emb, bank_meta, scene_store, chunks_np, threads):
    """One bounded sample of the WHOLE path through the reference's CPU arithmetic (the oracle port): patch extraction
    + query encoder + F.normalize, the kNN baseline BASELINE.md names (fp32 torch.cdist + topk brute force in batches of
    1024 queries; FLANN is not installable) + demotion + mapping rows, compose (util/retrieval.py:145-164), the
    dataloader normalisation and forward_full's inference part (U-Nets + attention + decoder).  Returns seconds."""
    from oracle import rf_oracle as O
    torch.set_num_threads(threads)
    n = chunks_np.shape[0]
    t0 = time.perf_counter()
    with torch.no_grad():
        q = torch.from_numpy(O.encode_chunk_queries(cfg, sds["fenc_input"], chunks_np))
        db = torch.from_numpy(bank_emb)
        K2 = 2 * cfg["K"]
        idx = torch.empty((q.shape[0], K2), dtype=torch.int64)
        dist = torch.empty((q.shape[0], K2), dtype=torch.float32)
        for lo in range(0, q.shape[0], 1024):
            d2 = torch.cdist(q[lo:lo + 1024], db) ** 2
            v, i = torch.topk(d2, K2, dim=1, largest=False)
            idx[lo:lo + 1024], dist[lo:lo + 1024] = i, v
        qs = np.full(q.shape[0], -1)
        oi, od = O.demote_same_scene(idx.numpy().astype(np.int32), dist.numpy(), bank_meta[:, 0].astype(np.int64), qs, cfg["K"])
        rows = O.mapping_rows(bank_meta, oi, od)
        retr = O.compose_chunks(cfg, rows, scene_store, n)
        O.refine_chunks(cfg, {k: v for k, v in sds.items() if k != "fenc_input"}, chunks_np, retr)
    return time.perf_counter() - t0


def synth_sds_cpu(cfg, seed=1234):
    """Synthetic state dicts of every network on the path (oracle shape tables), for the reference arm."""
    from oracle import rf_oracle as O
    nf, lv = cfg["nf"], cfg["unet_num_level"]
    kind = {8: "sr08", 16: "sr16", 128: "surface"}[cfg["dataset"]["input_chunk_size"]]
    rm = cfg["retrieval_model"]
    enc = O._INPUT_ENCODER[rm["network_input"]]
    return dict(fenc_input=O.synth_state_dict(O.encoder_param_shapes(enc, rm["nf_input"], rm["latent_dim"]), seed),
                unet_backbone=O.synth_state_dict(O.unet_backbone_shapes(kind, nf, lv), seed),
                retrieval_backbone=O.synth_state_dict(O.retrieval_backbone_shapes(nf, cfg["retrieval_fmaps"], cfg["retrieval_num_level"]), seed),
                attention=O.synth_state_dict(O.attention_shapes(nf, cfg["attn_patch_extent"] // 2), seed),
                decoder=O.synth_state_dict(O.final_decoder_shapes(nf), seed))


def synthetic_chunks_cpu(cfg, n, seed=0):
    """n raw input chunks of the workload on the host: low-res TSDF (super-resolution) or the 128^3 occupancy grid of
    a 1000-point cloud (surface reconstruction; dataset/scene.py:81-90, util/misc.py:73-78)."""
    from oracle import rf_oracle as O
    d = cfg["dataset"]
    rng = np.random.default_rng(seed)
    if cfg["task"] == "surface_reconstruction":
        return np.stack([O.point_cloud_to_grid(rng.random((d.get("num_points", 1000), 3)) * 64, d["input_chunk_size"],
                                               d["input_chunk_size"] / 64.0, 0)[None] for _ in range(n)])
    f = d["target_chunk_size"] // d["input_chunk_size"]
    return np.stack([O.downsample_tsdf(O.synthetic_tsdf(1000 + seed * 131 + i, 64, d["voxel_size_target"]) / d["voxel_size_target"]
                                       * d["voxel_size_input"], f, d["voxel_size_input"])[None] for i in range(n)]).astype(np.float32)


def workload_text(cfg, B, n_rows, rb):
    d = cfg["dataset"]
    enc = cfg["retrieval_model"]["network_input"]
    if cfg["task"] == "surface_reconstruction":
        return (f"Matterport3D surface_reconstruction 128->064 (BASELINE configs[3] shapes): {B} chunks/step/GPU, 128^3 occupancy "
                f"grid of a 1000-point cloud -> 64 x 48^3 patches -> PCPatch48 queries -> exact kNN vs {n_rows} rows (fetch "
                f"{2 * cfg['K']}, demote, keep {cfg['K']}) -> compose -> 5-level U-Net nf 12 + retrieval U-Net + attention K={cfg['K']} "
                f"+ decoder, refine sub-batch {rb}")
    return (f"3DFront super_resolution 008->064 (BASELINE configs[2] shapes): {B} chunks/step/GPU, pad+unfold+normalise + Patch04 "
            f"({enc}) queries -> exact kNN vs {n_rows} rows (fetch {2 * cfg['K']}, demote, keep {cfg['K']}) -> compose -> "
            f"U-Net + retrieval U-Net on 16^3 blocks + patch attention + decoder -> 64^3 TSDF, refine sub-batch {rb}")


def synthetic_bank_cpu(n_rows, seed=1):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n_rows, 64), dtype=np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    meta = np.zeros((n_rows, 7), dtype=np.float32)
    meta[:, 0] = np.arange(n_rows) // 64
    return emb, meta


def run_reference(args, rank, world):
    """--impl reference: rank 0 times the CPU arm on bounded samples of the same workload."""
    if rank != 0:
        return
    from oracle import rf_oracle as O
    from retrieval_fuse_b200.pipeline import FRONT3D_SR, SHAPENET_SR_RETRIEVAL
    threads = os.cpu_count() or 1
    metric = METRIC
    if args.full_like:
        from retrieval_fuse_b200.pipeline import MATTERPORT_SURFACE
        cfg = FRONT3D_SR if args.workload == "full" else MATTERPORT_SURFACE
        n_rows = args.bank_scenes * 64 + 1
        n_store = 32  # distinct synthetic scenes in the host scene store (the bank's scene ids wrap around them)
        emb, meta = synthetic_bank_cpu(n_rows)
        ps = cfg["dataset"]["patch_size_target"]
        ext = np.array([[x * ps, x * ps + ps, y * ps, y * ps + ps, z * ps, z * ps + ps] for x in range(4) for y in range(4)
                        for z in range(4)], dtype=np.float32)
        meta[:-1, 0] = (np.arange(n_rows - 1) // 64) % n_store
        meta[:-1, 1:] = np.tile(ext, (args.bank_scenes, 1))
        meta[-1] = [-1, 0, ps, 0, ps, 0, ps]
        store = np.stack([O.synthetic_tsdf(100 + i, 64, cfg["dataset"]["voxel_size_target"]) for i in range(n_store)])
        sds = synth_sds_cpu(cfg)
        per_step = min(args.cpu_sample_chunks, args.chunks)
        chunks = synthetic_chunks_cpu(cfg, per_step)
        for _ in range(min(args.warmup, 1)):
            cpu_full_sample(cfg, sds, emb, meta, store, chunks[:1], threads)
        times = [cpu_full_sample(cfg, sds, emb, meta, store, chunks, threads) for _ in range(args.steps)]
        config = {"workload": workload_text(cfg, args.chunks, n_rows, args.refine_batch), "chunks_per_step": per_step,
                  "bank_rows": n_rows, "K": cfg["K"],
                  "bank": f"{n_rows} random unit rows (same size as the GPU arm's encoded bank; the content does not change the "
                          f"cost of a brute-force cdist), scene store of {n_store} synthetic 64^3 scenes"}
        sample = (f"{per_step} chunks/step through the whole path (encode, cdist+topk kNN vs {n_rows} rows, compose, refine "
                  f"forward), torch CPU fp32, {threads} threads")
    elif args.workload == "retrieval":
        metric = METRIC_RETRIEVAL
        cfg = SHAPENET_SR_RETRIEVAL
        n_rows = args.bank_scenes * 64 + 1
        emb, meta = synthetic_bank_cpu(n_rows)
        sd = O.synth_state_dict(O.encoder_param_shapes("Patch04", 32, 64), 1234)
        rng = np.random.default_rng(0)
        per_step = min(args.cpu_sample_chunks, args.chunks)
        chunks = (rng.random((per_step, 1, 8, 8, 8), dtype=np.float32) * 0.5).astype(np.float32)
        for _ in range(args.warmup):
            cpu_retrieval_sample(cfg, sd, emb, meta, chunks[:max(1, per_step // 8)], threads)
        times = [cpu_retrieval_sample(cfg, sd, emb, meta, chunks, threads) for _ in range(args.steps)]
        config = {"workload": "ShapeNetV2 SR retrieval_008_064: Patch04 encode + kNN(2K=8, keep 4)", "chunks_per_step": per_step,
                  "bank_rows": n_rows, "queries_per_chunk": 64}
        sample = f"{per_step} chunks/step ({per_step * 64} queries x {n_rows} rows), torch CPU fp32 cdist+topk"
    else:
        metric = "64^3 TSDF chunks/sec (refine forward)"
        cfg = FRONT3D_SR
        sds = dict(unet_backbone=O.synth_state_dict(O.unet_backbone_shapes("sr08", 16, 4), 1234),
                   retrieval_backbone=O.synth_state_dict(O.retrieval_backbone_shapes(16, 16, 4), 1234),
                   attention=O.synth_state_dict(O.attention_shapes(16, 2), 1234),
                   decoder=O.synth_state_dict(O.final_decoder_shapes(16), 1234))
        per_step = 2
        rng = np.random.default_rng(0)
        chunks = rng.random((per_step, 1, 8, 8, 8), dtype=np.float32)
        retr = rng.random((per_step, 4, 64, 64, 64), dtype=np.float32) * 0.16
        for _ in range(min(args.warmup, 1)):
            cpu_refine_sample(cfg, sds, chunks[:1], retr[:1], threads)
        times = [cpu_refine_sample(cfg, sds, chunks, retr, threads) for _ in range(args.steps)]
        config = {"workload": "3DFront SR 008->064 refine forward (K=4)", "chunks_per_step": per_step}
        sample = f"{per_step} chunks/step, torch CPU fp32"
    total = float(sum(times))
    value = per_step * args.steps / total
    line = {"impl": "reference", "metric": metric,
            "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, local, world


def run_stages(args, local):
    """--workload stages: every SURVEY 8(d) stage of the path timed alone (CUDA events, L2 flushed before each
    timed launch) against the roofline that bounds it.  One JSON line: {"stages": {name: {...}}}.  A report next to
    the headline metric, not a replacement for it."""
    from retrieval_fuse_b200 import ops
    from retrieval_fuse_b200.pipeline import (FRONT3D_SR, SHAPENET_SR_RETRIEVAL, RefinementPipeline, RetrievalPipeline,
                                              build_bank_from_targets, synthetic_tsdf_batch)
    from retrieval_fuse_b200.model import get_retrieval_networks
    from retrieval_fuse_b200.pipeline import init_unit_gain_
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_grad_enabled(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs") or 6500.0
    tens = peaks.get("bf16_tflops_sustained") or 1400.0
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters=args.steps):
        for _ in range(max(args.warmup, 3)):
            fn()
        tot = 0.0
        for _ in range(iters):
            flush_buf.fill_(1)
            # keep the GPU busy (~0.15 ms spin) while the host enqueues e0, the stage and e1: otherwise the host-side
            # launch latency of the Python wrapper (20-30 us) is inside the event pair of a 25 us kernel
            torch.cuda._sleep(300000)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    stages = {}

    def add(name, ms, bytes_=None, flops=None, note=None):
        d = {"ms": ms}
        if bytes_ is not None:
            d.update(bound="hbm", algorithmic_bytes=bytes_, achieved_gbs=bytes_ / ms / 1e6, frac=bytes_ / ms / 1e6 / hbm)
        if flops is not None:
            d.update(bound="tensor" if bytes_ is None else "hbm+tensor", algorithmic_flops=flops,
                     achieved_tflops=flops / ms / 1e9, frac_tensor=flops / ms / 1e9 / tens)
        if note:
            d["note"] = note
        stages[name] = d
        log(f"[stages] {name}: {d}")

    cfg = FRONT3D_SR
    d = cfg["dataset"]
    B, K, nf = args.refine_batch, cfg["K"], 16
    # ---- a2 / a3 fold, unfold (pure permutations: 2 x numel x 4 B)
    x = torch.randn(B * (K + 1), nf, 32, 32, 32, device=dev)
    u = ops.unfold3d(x, 8)
    add("a2 Unfold3D(8, nf) [40,16,32^3]", timed(lambda: ops.unfold3d(x, 8)), bytes_=2 * x.numel() * 4)
    add("a3 Fold3D(4, 8, nf)", timed(lambda: ops.fold3d(u, 4, 8, nf)), bytes_=2 * x.numel() * 4)
    del x, u
    # ---- a1/a4 pad + unfold + normalise of the query chunks and of target scenes
    rcfg = SHAPENET_SR_RETRIEVAL
    rd = rcfg["dataset"]
    chunks = synthetic_tsdf_batch(args.chunks, 8, rd["voxel_size_input"], seed=7, device=dev, batch=2048).unsqueeze(1).contiguous()
    pipe_r = RetrievalPipeline(rcfg, bank=None, device=dev, weight_seed=1234)
    pu = lambda: ops.unfold3d_pad_stride(chunks, pipe_r.in_kernel, rd["patch_context_input"], pipe_r.in_stride, pipe_r.input_trunc,
                                         norm_sub=rd["input_mean"], norm_div=rd["input_std"])
    patches = pu()
    add("a1/a4 pad+unfold+normalise, 8^3 chunks -> 64 x 4^3", timed(pu), bytes_=(chunks.numel() + patches.numel()) * 4)
    # ---- a5 + a9 query encoder (Patch04) + row normalisation, one fused launch
    from retrieval_fuse_b200.pipeline import _encode_normalized
    enc = lambda: _encode_normalized(pipe_r.fenc_input, patches, 64)
    n_par = sum(p.numel() for n_, p in pipe_r.fenc_input.named_parameters() if n_.endswith("weight"))
    add("a5+a9 Patch04 MLP + normalise (fused chain)", timed(enc), flops=2.0 * n_par * patches.shape[0],
        note="weights (1.28 MB fp16 hi+lo) are re-streamed from L2 for every 128-row tile")
    del patches
    # ---- a6 conv query encoder (Patch08, Matterport SR 16^3 -> 64^3): 64 x 8^3 patches per chunk
    from retrieval_fuse_b200.pipeline import MATTERPORT_SR16
    mcfg = MATTERPORT_SR16
    md = mcfg["dataset"]
    pipe_m = RetrievalPipeline(mcfg, bank=None, device=dev, weight_seed=1234)
    mchunks = synthetic_tsdf_batch(args.chunks // 4, 16, md["voxel_size_input"], seed=9, device=dev, batch=1024).unsqueeze(1).contiguous()
    mpu = lambda: ops.unfold3d_pad_stride(mchunks, pipe_m.in_kernel, md["patch_context_input"], pipe_m.in_stride, pipe_m.input_trunc,
                                          norm_sub=md["input_mean"], norm_div=md["input_std"])
    mpatches = mpu()
    add("a1/a4 pad+unfold+normalise, 16^3 chunks -> 64 x 8^3", timed(mpu), bytes_=(mchunks.numel() + mpatches.numel()) * 4)
    add(f"a6 Patch08 conv query encoder, {mpatches.shape[0]} x 8^3", timed(lambda: _encode_normalized(pipe_m.fenc_input, mpatches, 64)),
        flops=361e6 / 64 * mpatches.shape[0])
    del mpatches, mchunks
    # ---- a7 dictionary encoder (Patch32) incl. pad-unfold of 64^3 targets
    n_sc = 64
    targets = synthetic_tsdf_batch(n_sc, 64, rd["voxel_size_target"], seed=100, device=dev)
    _, fenc_t = get_retrieval_networks(rcfg["retrieval_model"])
    init_unit_gain_(fenc_t, 11)
    fenc_t = fenc_t.to(dev).eval()
    ps, ctx = rd["patch_size_target"], rd["patch_context_target"]
    tp = lambda: ops.unfold3d_pad_stride(targets.unsqueeze(1), ps + 2 * ctx, ctx, rd["patch_stride"], pipe_r.target_trunc,
                                         norm_sub=rd["target_mean"], norm_div=rd["target_std"])
    tpatches = tp()
    add("a4 pad+unfold of 64^3 targets -> 64 x 32^3", timed(tp), bytes_=(targets.numel() + tpatches.numel()) * 4)
    add("a7 Patch32 dictionary encoder, 4096 x 32^3", timed(lambda: _encode_normalized(fenc_t, tpatches, 64)),
        flops=338.5e6 * tpatches.shape[0])
    del tpatches
    # ---- a12 compose
    bank, _ = build_bank_from_targets(cfg, targets, dev, weight_seed=11)
    pipe = RefinementPipeline(cfg, bank, targets, device=dev, weight_seed=1234)
    rchunks = synthetic_tsdf_batch(B, 8, d["voxel_size_input"], seed=3, device=dev).unsqueeze(1).contiguous()
    rows, _ = pipe.retrieve(rchunks)
    add("a12 compose gather (K=4)", timed(lambda: pipe.compose(rows, B, normalize=True)), bytes_=2.0 * K * 64 ** 3 * 4 * B)
    retr = pipe.compose(rows, B, normalize=True)
    x_in = pipe.normalize_input(rchunks)
    # ---- a13 retrieval U-Net, a15 input U-Net, a14 attention, a16 decoder
    rp = ops.unfold3d(retr.reshape(B * K, 1, 64, 64, 64), 16)
    add("a13 RetrievalUNetBackbone, 2048 x 16^3 patches", timed(lambda: pipe.retrieval_backbone(rp)), flops=19.03e9 * B * K)
    feats = pipe.retrieval_backbone(rp)
    x_retr = ops.fold3d(feats, 4, 8, nf)
    add("a15 Superresolution08UNetBackbone", timed(lambda: pipe.unet_backbone(x_in)), flops=1.91e9 * B)
    x_back = pipe.unet_backbone(x_in)
    add("a14 PatchedAttentionBlock (K=4)", timed(lambda: pipe.patched_attention_block(x_back, x_retr)),
        bytes_=(K + 2.0) * nf * 32 ** 3 * 4 * B, flops=2.18e9 * B)
    xa = pipe.patched_attention_block(x_back, x_retr)
    add("a16 Superresolution08FinalDecoder", timed(lambda: pipe.decoder(xa)), flops=7.26e9 * B)
    line = {"metric": "per-stage rooflines (SURVEY 8d)", "unit": "ms", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "config": {"workload": f"stages: 3DFront SR 008->064 batch {B} K={K}; ShapeNetV2 retrieval {args.chunks} chunks",
                       "l2": "flushed before every timed launch (256 MiB write)"},
            "peaks": {"hbm_gbs": hbm, "bf16_tflops_sustained": tens, "source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "tensor_note": "tensor fractions are ALGORITHMIC flops / bf16 dense peak; the kernels spend 3 fp16 MMAs per product "
                           "(hi/lo split for fp32-level accuracy), so 1/3 is the ceiling of this column",
            "stages": stages}
    print(json.dumps(line), flush=True)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_traffic(key):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS workload
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when there is no capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(key)
    except Exception:
        return None


def build_world(args, cfg, dev, rank, world, need_store):
    """Bank (this rank's row shard), scene store, process group of the shard exchange."""
    import torch.distributed as dist
    from retrieval_fuse_b200.pipeline import build_bank_from_targets, synthetic_tsdf_batch
    from retrieval_fuse_b200.util.retrieval import EmbeddingBank
    d = cfg["dataset"]
    S_tot = args.bank_scenes
    S_b = args.bank_shards if args.bank_shards > 0 else (min(world, 4) if args.full_like else min(world, 2))
    if S_b > world or world % S_b:
        S_b = world
    shard, gidx, my_group = rank % S_b, rank // S_b, None
    if world > 1:
        for gi in range(world // S_b):  # every rank creates every group (collective call)
            grp = dist.new_group(list(range(gi * S_b, (gi + 1) * S_b)))
            if gi == gidx:
                my_group = grp
    per = (S_tot + S_b - 1) // S_b
    lo, hi = min(shard * per, S_tot), min((shard + 1) * per, S_tot)
    store = None
    if args.bank == "encoded":
        if need_store:  # compose reads any scene of the bank: the store is replicated (1 MB per scene)
            store = torch.cat([synthetic_tsdf_batch(min((s + 1) * per, S_tot) - min(s * per, S_tot), 64, d["voxel_size_target"],
                                                    seed=100 + s, device=dev) for s in range(S_b)])
            targets = store[lo:hi]
        else:
            targets = synthetic_tsdf_batch(hi - lo, 64, d["voxel_size_target"], seed=100 + shard, device=dev)
        bank, _ = build_bank_from_targets(cfg, targets, dev, weight_seed=11, scene_offset=lo, n_scenes_total=S_tot,
                                          batch_patches=4096)
        del targets
    else:
        g = torch.Generator(device=dev).manual_seed(1 + shard)
        n_loc = (hi - lo) * 64 + (1 if hi == S_tot else 0)
        emb = torch.nn.functional.normalize(torch.randn(n_loc, 64, generator=g, device=dev), dim=1)
        meta = torch.zeros((S_tot * 64 + 1, 7), device=dev)
        meta[:, 0] = torch.arange(S_tot * 64 + 1, device=dev) // 64
        bank = EmbeddingBank(emb, meta, [f"scene{i:05d}" for i in range(S_tot)], row_offset=lo * 64, n_total=S_tot * 64 + 1)
    return bank, store, my_group, S_b, per


def run_ours(args, rank, local, world):
    import torch.distributed as dist
    from retrieval_fuse_b200 import ops
    from retrieval_fuse_b200.pipeline import (FRONT3D_SR, MATTERPORT_SURFACE, SHAPENET_SR_RETRIEVAL, RefinementPipeline,
                                              RetrievalPipeline, synthetic_tsdf_batch, downsample_tsdf_batch)
    from retrieval_fuse_b200.sharded import ShardedBankQuery
    from retrieval_fuse_b200.util.retrieval import EmbeddingBank

    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_grad_enabled(False)
    if world > 1:
        # stdout carries exactly ONE JSON line, but NCCL writes its version banner there from C: point fd 1 at stderr
        # for the whole run and keep the real stdout for the result line (emit())
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s (of fallback)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def rank_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    t0 = time.time()
    extra = {}
    e2e_pipelined = None
    profile_step = None
    B = args.chunks

    if args.full_like:
        # ------------------------------------------------------------------ the whole path (headline)
        cfg = FRONT3D_SR if args.workload == "full" else MATTERPORT_SURFACE
        d = cfg["dataset"]
        bank, store, my_group, S_b, per = build_world(args, cfg, dev, rank, world, need_store=True)
        sq = ShardedBankQuery(bank, group=my_group) if world > 1 else None
        pipe = RefinementPipeline(cfg, bank, store, device=dev, weight_seed=1234, sharded_query=sq)
        rb = min(args.refine_batch, B)
        if cfg["task"] == "surface_reconstruction":
            g = torch.Generator(device="cpu").manual_seed(7 + rank)
            pts = (torch.rand((B, d["num_points"], 3), generator=g) * 128).long().clamp_(0, 127).to(dev)
            chunks = torch.zeros((B, 1, 128, 128, 128), device=dev)
            chunks[torch.arange(B, device=dev)[:, None], 0, pts[..., 0], pts[..., 1], pts[..., 2]] = 1.0
        else:
            qt = synthetic_tsdf_batch(B, 64, d["voxel_size_target"], seed=5000 + rank, device=dev)
            chunks = downsample_tsdf_batch(qt, 64 // d["input_chunk_size"], d["voxel_size_target"], d["voxel_size_input"])
            del qt
        chunks_host = chunks.cpu().pin_memory()
        out_host = torch.empty((B, 1, 64, 64, 64), dtype=torch.float32).pin_memory()
        n_rows = bank.n_total
        use_graph = not args.no_cuda_graph
        if use_graph:
            try:
                pipe.infer(chunks, None, args.knn_method, rb, True)
            except Exception as e:  # same kernels either way; say so instead of failing the measurement
                log(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); launching eagerly")
                use_graph = False
        pred_buf = torch.empty((B, 1, 64, 64, 64), dtype=torch.float32, device=dev)
        stage_names = ["encode", "knn", "compose", "refine"]

        def step(timed):
            marks = [] if timed else None
            pipe.infer(chunks, None, args.knn_method, rb, use_graph, out=pred_buf, marks=marks)
            return marks

        def e2e_step():
            return pipe.infer_host(chunks_host, out_host, None, args.knn_method, rb, use_graph)

        out_host2 = torch.empty((B, 1, 64, 64, 64), dtype=torch.float32).pin_memory()

        def e2e_pipelined(n):
            """n batches through the two-deep host pipeline: every batch pays its own H2D and D2H inside the timed
            region; the D2H copy of batch i's prediction overlaps the kernels of batch i+1."""
            pending = []
            for i in range(n):
                pending.append(pipe.infer_host_async(chunks_host, (out_host, out_host2)[i & 1], None, args.knn_method, rb, use_graph, slot=i)[1])
                if len(pending) > 1:
                    pending.pop(0).synchronize()  # the prediction of batch i-1 is in host memory
            pending[-1].synchronize()

        def profile_step():
            pipe.infer(chunks, None, args.knn_method, rb, False)

        h2d, d2h = chunks_host.numel() * 4, out_host.numel() * 4
        metric = METRIC
        config = {"workload": workload_text(cfg, B, n_rows, rb), "chunks_per_step_per_gpu": B, "refine_sub_batch": rb,
                  "bank_rows": n_rows, "K": cfg["K"], "bank_shards": S_b if world > 1 else 1,
                  "bank": (f"{S_b} row shards x {world // S_b} replica groups, exchange (all-gather queries, all-to-all "
                           f"top-2K lists) inside a group; scene store replicated") if world > 1 else "single GPU",
                  "l2": "flushed between timed steps (256 MiB write)", "launch": "refinement sub-batches replay a CUDA graph" if use_graph else "eager",
                  "knn_method": args.knn_method, "embeddings": args.bank}
        flops_per_chunk = {"full": 87.4e9 + 0.041e9 + 8192.0 * n_rows, "surface": (50.0 + 85.8 + 3.62 + 4.08 + 91.9) * 1e9 + 8192.0 * n_rows}[args.workload]
        units_per_step = B
        dom_op, dom_name = "rf_tc_conv3d_halo_fwd", "tc_conv3d_halo_kernel (shifted-window 3x3x3 conv, tcgen05, fp16 hi/lo split)"
    elif args.workload == "retrieval":
        # ------------------------------------------------------------------ configs[1]: encode + kNN only
        cfg = SHAPENET_SR_RETRIEVAL
        d = cfg["dataset"]
        bank, _, my_group, S_b, per = build_world(args, cfg, dev, rank, world, need_store=False)
        S_tot = args.bank_scenes
        sq = ShardedBankQuery(bank, group=my_group) if world > 1 else None
        pipe = RetrievalPipeline(cfg, bank, device=dev, weight_seed=1234, sharded_query=sq)
        chunks = synthetic_tsdf_batch(B, 8, d["voxel_size_input"], seed=7 + rank, device=dev, batch=2048).unsqueeze(1).contiguous()
        chunks_host = chunks.cpu().pin_memory()
        n_rows = bank.n_total
        Q = B * pipe.patches_per_chunk
        q_rand = None
        if args.bank == "random":
            g = torch.Generator(device=dev).manual_seed(2 + rank)
            q_rand = torch.nn.functional.normalize(torch.randn(Q, 64, generator=g, device=dev), dim=1)
        stage_names = ["encode", "knn"]

        def step(timed):
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
            if timed:
                marks[0].record()
            q = pipe.encode_queries(chunks)
            if timed:
                marks[1].record()
            pipe.lookup(q if q_rand is None else q_rand, None, args.knn_method)
            if timed:
                marks[2].record()
            return marks

        def e2e_step():
            return pipe.retrieve_host(chunks_host, None, args.knn_method)

        def e2e_pipelined(n):
            """n batches through the two-deep host pipeline: every batch pays its own H2D and D2H inside the timed
            region; the copies of batch i overlap the kernels of batch i+1."""
            pending = []
            for i in range(n):
                pending.append(pipe.retrieve_host_async(chunks_host, None, args.knn_method, slot=i)[1])
                if len(pending) > 1:
                    pending.pop(0).synchronize()  # the rows of batch i-1 are in host memory
            pending[-1].synchronize()

        def profile_step():
            step(False)

        h2d, d2h = chunks_host.numel() * 4, Q * cfg["K"] * 8 * 4
        metric = METRIC_RETRIEVAL
        config = {"workload": f"ShapeNetV2 SR retrieval_008_064 (BASELINE configs[1]): Patch04 encode + exact kNN (fetch 8, demote, keep 4), "
                              f"{B} chunks x 64 queries vs {n_rows} rows", "chunks_per_step_per_gpu": B, "bank_rows": n_rows, "K": cfg["K"],
                  "bank": (f"{S_b} row shards x {world // S_b} replica groups, exchange (all-gather queries, all-to-all "
                           f"top-2K lists) inside a group") if world > 1 else "single GPU", "bank_shards": S_b if world > 1 else 1,
                  "l2": "flushed between timed steps (256 MiB write)", "knn_method": args.knn_method, "embeddings": args.bank}
        units_per_step = B
        dom_op, dom_name = "rf_knn_l2_topk", "rf_knn_l2_topk (operand images + knn_tc_candidates_kernel tcgen05 score GEMM with running top lists + fp64 re-rank)"
    else:
        # ------------------------------------------------------------------ refinement forward alone
        cfg = FRONT3D_SR
        d = cfg["dataset"]
        pipe = RefinementPipeline(cfg, bank=None, device=dev, weight_seed=1234)
        B = args.refine_batch
        x_in = torch.randn(B, 1, 8, 8, 8, device=dev)
        tg = synthetic_tsdf_batch(B * 4, 64, d["voxel_size_target"], seed=3 + rank, device=dev)
        retr = ((tg - d["target_mean"]) / d["target_std"]).reshape(B, 4, 64, 64, 64).contiguous()
        x_in_host, retr_host = x_in.cpu().pin_memory(), retr.cpu().pin_memory()
        out_host = torch.empty((B, 1, 64, 64, 64), dtype=torch.float32).pin_memory()
        use_graph = not args.no_cuda_graph
        if use_graph:
            try:
                pipe.refine_graphed(x_in, retr)
            except Exception as e:
                log(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); launching eagerly")
                use_graph = False
        fwd = (lambda a, b: pipe.refine_graphed(a, b)) if use_graph else (lambda a, b: pipe.refine(a, b)[0])
        stage_names = ["refine"]

        def step(timed):
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if timed else None
            if timed:
                marks[0].record()
            fwd(x_in, retr)
            if timed:
                marks[1].record()
            return marks

        def e2e_step():
            p = fwd(x_in_host.to(dev, non_blocking=True), retr_host.to(dev, non_blocking=True))
            out_host.copy_(p, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return out_host

        def profile_step():
            pipe.refine(x_in, retr)

        h2d, d2h = (x_in.numel() + retr.numel()) * 4, out_host.numel() * 4
        metric = "64^3 TSDF chunks/sec (refine forward)"
        config = {"workload": f"3DFront SR 008->064 refine forward, batch {B}, K=4 (unet + retrieval unet + attention + decoder)",
                  "chunks_per_step_per_gpu": B, "l2": "flushed between timed steps (256 MiB write)",
                  "launch": "CUDA graph replay" if use_graph else "eager"}
        units_per_step = B
        dom_op, dom_name = "rf_tc_conv3d_halo_fwd", "tc_conv3d_halo_kernel (shifted-window 3x3x3 conv, tcgen05, fp16 hi/lo split)"
    torch.cuda.synchronize(dev)
    log(f"[rank {rank}] setup {time.time() - t0:.1f}s")

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()

    # ---- timed region: K steps, per-step CUDA events on the launching stream, L2 flushed in between
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    t_step, t_stage = [], {n: [] for n in stage_names}
    ev_end = torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush_buf.fill_(1)
        marks = step(True)
        ev_end.record()
        ev_end.synchronize()
        t_step.append(marks[0].elapsed_time(ev_end))
        for i, n in enumerate(stage_names):
            t_stage[n].append(marks[i].elapsed_time(marks[i + 1]))
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = rank_max(sum(t_step))
    value = units_per_step * world * args.steps / (total_ms / 1e3)

    # ---- e2e: host buffers in and out through the reference-facing call
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_value = units_per_step * world * args.steps / rank_max(time.perf_counter() - t0)
    e2e_pipe_value = None
    if e2e_pipelined is not None:
        e2e_pipelined(2)
        barrier()
        t0 = time.perf_counter()
        e2e_pipelined(args.steps)
        barrier()
        e2e_pipe_value = units_per_step * world * args.steps / rank_max(time.perf_counter() - t0)

    # ---- launches per step + per-op device times (one eager, untimed pass with a CUDA event pair around every op; the
    # GPU spins first so that the host runs ahead and the ops execute back to back as they do inside the timed steps)
    ops.reset_launches()
    profile_step()
    torch.cuda.synchronize(dev)
    launches_per_step = ops.launches()
    prof = None
    torch.cuda._sleep(int(2e7))
    if rank == 0:
        ops.profile_start()
    # (one stream for this pass: with the input U-Net forked onto its side stream, its launches land between the event pairs
    # of the main stream's ops and inflate them - the convolutions summed to 28 ms of a 33 ms step against 22 ms in the
    # ncu launch list)
    pipe.serial_streams = True
    profile_step()  # every rank takes part (the sharded lookup is a collective); only rank 0 records events
    pipe.serial_streams = False
    if rank == 0:
        prof = ops.profile_stop()
    barrier()

    # ---- the retrieval half alone on the same bank (BASELINE configs[1] style: encode + kNN, no compose / refine)
    if args.full_like and args.workload == "full":
        nR = 10000
        rch = synthetic_tsdf_batch(nR, 8, d["voxel_size_input"], seed=7 + rank, device=dev, batch=2048).unsqueeze(1).contiguous()
        for _ in range(2):
            pipe.retrieve(rch, None, args.knn_method)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(3):
            flush_buf.fill_(1)
            e0.record()
            pipe.retrieve(rch, None, args.knn_method)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        barrier()
        tot = rank_max(tot)
        extra["retrieval_only"] = {"value": nR * world * 3 / (tot / 1e3), "unit": UNIT, "ms_per_step": tot / 3,
                                   "workload": f"encode + kNN only (no compose / refine), {nR} chunks per GPU and step vs the same {n_rows}-row bank"}
        del rch

    if rank == 0:
        step_ms = total_ms / args.steps
        dom = (prof or {}).get(dom_op)
        if dom and dom["ms"] > 0:
            roof = {"bound": "tensor", "kernel": dom_name, "achieved": dom["flops"] / dom["ms"] / 1e9, "peak": peak, "unit": "TFLOP/s",
                    "frac": dom["flops"] / dom["ms"] / 1e9 / peak, "traffic": load_traffic(f"{args.workload}:{dom_op}"),
                    "peak_source": peak_src, "launches_per_step": dom["calls"], "avg_launch_ms": dom["ms"] / dom["calls"],
                    "algorithmic_flops_per_launch": dom["flops"] / dom["calls"], "ms_per_step": dom["ms"],
                    "share_of_step": dom["ms"] / step_ms,
                    "note": "achieved = ALGORITHMIC flops of these launches (2*M*27*Cin*Cout for the conv; 2*Q*N*64 for the kNN) / "
                            "their summed CUDA-event durations in an eager pass of the same step; the convolution spends three "
                            "fp16 MMAs per product (hi/lo split for fp32-level accuracy) on rows that include halo positions, "
                            "so 1/3 is the ceiling of frac"}
        else:
            roof = {"bound": "tensor", "kernel": dom_name, "achieved": None, "peak": peak, "unit": "TFLOP/s", "frac": None,
                    "traffic": None, "peak_source": peak_src}
        roof["whole_step"] = None
        if args.full_like:
            roof["whole_step"] = {"algorithmic_tflops": flops_per_chunk * B / step_ms / 1e9,
                                  "frac_of_peak": flops_per_chunk * B / step_ms / 1e9 / peak,
                                  "flops_per_chunk": flops_per_chunk}
        op_table = None
        if prof:
            tot_ms = sum(v["ms"] for v in prof.values())
            op_table = {k: {"calls": v["calls"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4),
                            **({"tflops": round(v["flops"] / v["ms"] / 1e9, 1)} if v["flops"] else {}),
                            **({"gbs": round(v["bytes"] / v["ms"] / 1e6, 1)} if v["bytes"] else {})}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        e2e = {"value": e2e_pipe_value or e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "call": (("RefinementPipeline.infer_host_async: pinned host chunks in, pinned host predictions out, the D2H copy of a "
                         "batch overlaps the kernels of the next (two slots); sync_value = infer_host, one synchronous call per step"
                         if args.full_like else "RetrievalPipeline.retrieve_host_async: host buffers in and out, two batches in flight")
                        if e2e_pipe_value else
                        ("RefinementPipeline.infer_host: pinned host chunks in, pinned host predictions out, one synchronous "
                         "call per step" if args.full_like else "pipeline call with pinned host tensors in and out, one synchronous call per step"))}
        if e2e_pipe_value:
            e2e["sync_value"] = e2e_value
        if args.workload == "retrieval" and args.bank == "random":
            e2e["note"] = ("the host-buffer call encodes the synthetic chunks, so its queries are the encoded ones against the random "
                           "bank; the isotropic queries of `value` exist on the device only - not comparable with `value`")
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (tcgen05 fp16 hi/lo split products, fp32 accumulate) / f64 distance ranking", "data": "synthetic",
                "config": config, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
                "roofline": roof, "clocks": clocks,
                "breakdown_ms": dict({n: float(np.mean(v)) for n, v in t_stage.items()}, step=step_ms,
                                     wall_per_step_incl_flush=1e3 * wall / args.steps),
                "op_breakdown_eager": op_table}
        line.update(extra)
        if args.workload == "retrieval" and world == 1:
            qq = pipe.encode_queries(chunks) if q_rand is None else q_rand
            ops.knn_topk(bank.emb, qq, 2 * cfg["K"], method=args.knn_method, stats=True)
            line["knn_stats"] = dict(ops.last_knn_stats)
        if not args.no_cpu_baseline and world == 1:
            try:
                threads = os.cpu_count() or 1
                n = min(args.cpu_sample_chunks, B)
                if args.full_like:
                    sds = pipe.state_dicts()
                    secs = cpu_full_sample(cfg, sds, bank.emb.cpu().numpy(), bank.meta.cpu().numpy(), store.cpu().numpy(),
                                           chunks_host[:n].numpy(), threads)
                    sample = (f"{n} of the step's {B} chunks through the whole path on the same bank, scene store and weights "
                              f"(oracle port: torch CPU fp32, cdist+topk kNN), {secs:.1f}s")
                elif args.workload == "retrieval":
                    sd = {k: v.detach().cpu() for k, v in pipe.fenc_input.state_dict().items()}
                    secs = cpu_retrieval_sample(cfg, sd, bank.emb.cpu().numpy(), bank.meta.cpu().numpy(), chunks_host[:n].numpy(), threads)
                    sample = f"{n} of the {B} chunks ({n * 64} queries x {n_rows} rows), torch CPU fp32 cdist+topk, {secs:.1f}s"
                else:
                    n = 1
                    secs = cpu_refine_sample(cfg, {k: v for k, v in pipe.state_dicts().items() if k != "fenc_input"},
                                             x_in_host[:1].numpy(), (retr_host[:1].numpy() * d["target_std"] + d["target_mean"]), threads)
                    sample = f"1 chunk, torch CPU fp32, {secs:.1f}s"
                line["cpu_baseline"] = {"value": n / secs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args, rank, local, world):
    """--workload sweep (BASELINE configs[4]): 1 M-row isotropic bank, brute-force kNN for k in {1, 4, 8, 16} (the lookup
    fetches 2k, util/retrieval.py:92) on a bulk batch (640 000 queries = 10 000 chunks) and on one refinement batch
    (64 chunks), plus the attention fuse for the same k.  At N > 1 the bank is sharded by rows over all ranks and every
    rank brings its own queries (weak scaling).  One JSON line; `value` = chunks/s of the k = 4 bulk lookup."""
    import torch.distributed as dist
    from retrieval_fuse_b200 import ops
    from retrieval_fuse_b200.model import get_attention_block
    from retrieval_fuse_b200.pipeline import FRONT3D_SR, init_unit_gain_
    from retrieval_fuse_b200.sharded import ShardedBankQuery
    from retrieval_fuse_b200.util.retrieval import EmbeddingBank
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.set_grad_enabled(False)
    if world > 1:
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    hbm = peaks.get("hbm_gbs") or 6500.0
    N_rows = 1_000_000
    g = torch.Generator(device=dev).manual_seed(1)
    full = torch.nn.functional.normalize(torch.randn(N_rows, 64, generator=g, device=dev), dim=1)
    per = (N_rows + world - 1) // world
    lo, hi = rank * per, min(N_rows, (rank + 1) * per)
    meta = torch.zeros((N_rows, 7), device=dev)
    meta[:, 0] = torch.arange(N_rows, device=dev) // 64
    bank = EmbeddingBank(full[lo:hi].contiguous(), meta, ["s"], row_offset=lo, n_total=N_rows)
    del full
    sq = ShardedBankQuery(bank) if world > 1 else None
    gq = torch.Generator(device=dev).manual_seed(2 + rank)
    Qb = 640_000
    q_bulk = torch.nn.functional.normalize(torch.randn(Qb, 64, generator=gq, device=dev), dim=1)
    q_small = q_bulk[:4096].contiguous()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        tot = 0.0
        for _ in range(iters):
            flush_buf.fill_(1)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / iters], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rows = []
    for k in (1, 4, 8, 16):
        look = (lambda q: sq.query(q, k)) if sq is not None else (lambda q: bank.query(q, k))
        ms_bulk = timed(lambda: look(q_bulk), max(2, min(args.steps, 3)))
        ms_small = timed(lambda: look(q_small), args.steps)
        st = None
        if world == 1:
            ops.knn_topk(bank.emb, q_bulk, 2 * k, stats=True)
            st = dict(ops.last_knn_stats)
        cfg = dict(FRONT3D_SR, K=k)
        attn = init_unit_gain_(get_attention_block(cfg), 5).to(dev).eval()
        ga = torch.Generator(device=dev).manual_seed(3)
        B = 8
        xb = torch.randn(B, 16, 32, 32, 32, generator=ga, device=dev)
        xr = torch.randn(B * k, 16, 32, 32, 32, generator=ga, device=dev)
        ms_attn = timed(lambda: attn(xb, xr), args.steps)
        flops = 2.0 * Qb * world * (hi - lo) * 64   # per rank: all ranks' queries x its shard
        rows.append({"k": k, "fetch": 2 * k, "knn_bulk_ms": ms_bulk, "knn_bulk_chunks_per_s": Qb / 64 * world / (ms_bulk / 1e3),
                     "knn_bulk_algorithmic_tflops_per_gpu": flops / ms_bulk / 1e9, "knn_bulk_frac_of_bf16_peak": flops / ms_bulk / 1e9 / peak,
                     "knn_64chunks_ms": ms_small, "knn_stats": st, "attention_fuse_ms_batch8": ms_attn,
                     "attention_fuse_gbs": (k + 2.0) * 16 * 32 ** 3 * 4 * B / ms_attn / 1e6,
                     "attention_fuse_frac_of_hbm": (k + 2.0) * 16 * 32 ** 3 * 4 * B / ms_attn / 1e6 / hbm})
        log(f"[sweep] {rows[-1]}")
        del attn, xb, xr
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        r4 = [r for r in rows if r["k"] == 4][0]
        emit({"metric": "1M-row bank brute-force kNN + attention-fuse sweep (BASELINE configs[4])", "value": r4["knn_bulk_chunks_per_s"],
              "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": 2, "ms_per_step": r4["knn_bulk_ms"], "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "fp16 tcgen05 scores (one K = 64 GEMM, proven bound) / f64 distance ranking", "data": "synthetic",
              "config": {"workload": f"1 M-row isotropic unit bank (seed 1), {Qb} isotropic queries per GPU (10 000 chunks), k in 1/4/8/16 "
                                     f"(fetch 2k), bank in {world} row shard(s); attention fuse on randn [8,16,32^3] + [8k,16,32^3]",
                         "l2": "flushed between timed launches (256 MiB write)"},
              "sweep": rows, "clocks": clocks,
              "peaks": {"bf16_tflops_sustained": peak, "hbm_gbs": hbm, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}})
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank, local, world = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "stages":
        if rank == 0:
            run_stages(args, local)
        return
    if args.workload == "sweep":
        run_sweep(args, rank, local, world)
        return
    run_ours(args, rank, local, world)


if __name__ == "__main__":
    main()
