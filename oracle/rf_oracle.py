"""CPU oracle for the RetrievalFuse hot path (TEST INFRASTRUCTURE ONLY).

This file restates, on the CPU, the arithmetic of the reference's hot path
(nihalsid/retrieval-fuse @ fce90fa): patch fold/unfold, the patch encoders, the
exact k-nearest-neighbour lookup that the reference approximates with FLANN,
the compose step, the 3D U-Nets and the patch attention block.  It exists so
that the CUDA path can be checked against it; nothing in the product imports
it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import or execute anything under `oracle/`.

Pinning status
--------------
* fold / unfold / patcher / encoders / U-Nets / attention / decoder: PINNED.
  `tests/golden/make_golden.py` imports the reference's own `model/*.py` and
  `util/patcher.py` from /root/reference (possible in the build container
  only), drives them with the seeded inputs and synthetic weights defined
  below, and commits the outputs under `tests/golden/`.
  `tests/test_oracle_golden.py` checks this restatement against those files.
* patch enumeration (a1), database rows + sentinel (a10), fetch-2K / demotion /
  mapping rows (a11), compose incl. the overlapping-stride rule (a12): PINNED.
  `tests/golden/make_golden_retrieval.py` imports the reference's own
  `util/retrieval.py`, `dataset/scene.py`, `dataset/patched_scene_dataset.py`
  unmodified (pyflann, trimesh, pyrender, marching_cubes, torchmetrics stubbed in
  sys.modules) and executes SceneHandler / PatchedSceneDataset,
  create_dictionary, get_zero_patch_entry, flann_knn_worker (with and without
  ignore_patches_from_source) and create_retrieval_from_mapping on a tiny
  on-disk dataset; `tests/test_oracle_retrieval_golden.py` checks this
  restatement against the committed outputs, bit-exact.
* the nearest-neighbour SEARCH itself: the reference calls `pyflann`
  (un-vendored, unpinned, approximate randomized kd-forest;
  util/retrieval.py:8,50,92), absent here, no golden vectors in the tree.  The
  golden run replaces `FLANN.nn_index` by the thing it approximates - exact
  squared-L2 neighbours under the canonical rule stated at `knn_exact`
  (written out independently in the golden script) - so every line of the
  reference AROUND the search is pinned, and the search is pinned to that rule.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes
import os
import zlib

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------------------
# Deterministic synthetic weights and inputs (shared by golden maker + tests)
# ---------------------------------------------------------------------------


def synth_tensor(name: str, shape, seed: int, kind: str) -> torch.Tensor:
    """Deterministic parameter values from (name, shape, seed).

    Not a reference function: the reference has no fixtures, so the
    golden vectors are made by loading THESE values into the reference's own
    modules.  PCG64 streams are stable across numpy versions.
    kind: 'weight' -> U(-b, b) with b = sqrt(3/fan_in) (unit gain), 'bias' ->
    U(-0.1, 0.1), 'gn_weight' -> 1 + U(-0.2, 0.2), 'gn_bias' -> U(-0.1, 0.1).
    """
    h = zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1 & 0xFFFFFFFF)
    rng = np.random.Generator(np.random.PCG64(h))
    shape = tuple(int(s) for s in shape)
    u = rng.random(size=shape, dtype=np.float64) * 2.0 - 1.0
    if kind == "weight":
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
        # sqrt(3/fan_in): unit-gain uniform, so that activations keep O(1)
        # scale through 20+ layers (torch's default 1/sqrt(fan_in) makes a deep
        # random net's output collapse to ~0, which would make 1e-4 vacuous).
        arr = u * np.sqrt(3.0 / max(fan_in, 1))
    elif kind == "bias":
        arr = u * 0.1
    elif kind == "gn_weight":
        arr = 1.0 + 0.2 * u
    elif kind == "gn_bias":
        arr = 0.1 * u
    elif kind == "const":
        arr = u
    else:
        raise ValueError(kind)
    return torch.from_numpy(arr.astype(np.float32))


def synth_state_dict(shapes: dict, seed: int) -> dict:
    """shapes: {param_name: shape}. Classifies each parameter by its name."""
    sd = {}
    for name, shape in shapes.items():
        if name.endswith("sig_scale"):
            sd[name] = torch.full(tuple(shape), 35.0)  # model/attention.py:60
        elif name.endswith("sig_shift"):
            sd[name] = torch.full(tuple(shape), -27.0)  # model/attention.py:61
        elif "groupnorm.weight" in name:
            sd[name] = synth_tensor(name, shape, seed, "gn_weight")
        elif "groupnorm.bias" in name:
            sd[name] = synth_tensor(name, shape, seed, "gn_bias")
        elif name.endswith("running_mean"):
            sd[name] = synth_tensor(name, shape, seed, "gn_bias")
        elif name.endswith("running_var"):
            sd[name] = synth_tensor(name, shape, seed, "gn_weight")
        elif name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith(".bias"):
            sd[name] = synth_tensor(name, shape, seed, "bias")
        else:
            sd[name] = synth_tensor(name, shape, seed, "weight")
    return sd


def synthetic_tsdf(seed: int, size: int = 64, voxel_size: float = 0.054167,
                   n_prims: int | None = None) -> np.ndarray:
    """Unsigned distance to a union of random spheres / boxes / planes in a
    size^3 grid, voxel units * voxel_size, clamped to trunc = float16(3*voxel)
    (dataset/scene.py:30-33 defines trunc that way).  SURVEY 8(d)."""
    rng = np.random.Generator(np.random.PCG64(0xD15EA5E ^ seed))
    trunc = float(np.float16(voxel_size * 3).astype(np.float32))
    g = np.arange(size, dtype=np.float64) + 0.5
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    d = np.full((size,) * 3, 1e9)
    n = int(rng.integers(3, 9)) if n_prims is None else n_prims
    for _ in range(n):
        kind = int(rng.integers(0, 3))
        c = rng.random(3) * size
        if kind == 0:  # sphere shell
            r = (0.08 + 0.25 * rng.random()) * size
            dd = np.abs(np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r)
        elif kind == 1:  # box surface
            hs = (0.05 + 0.2 * rng.random(3)) * size
            q = np.stack([np.abs(X - c[0]) - hs[0], np.abs(Y - c[1]) - hs[1], np.abs(Z - c[2]) - hs[2]])
            outside = np.sqrt((np.maximum(q, 0) ** 2).sum(0))
            inside = np.minimum(q.max(0), 0)
            dd = np.abs(outside + inside)
        else:  # plane
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            dd = np.abs((X - c[0]) * nrm[0] + (Y - c[1]) * nrm[1] + (Z - c[2]) * nrm[2])
        d = np.minimum(d, dd)
    return np.minimum(d * voxel_size, trunc).astype(np.float32)


def downsample_tsdf(target: np.ndarray, factor: int, voxel_size_input: float) -> np.ndarray:
    """Low-res input = the same field sampled on a coarser grid and re-truncated
    to float16(3*voxel_size_input) (dataset/scene.py:32). SURVEY 8(d)."""
    s = target.shape[0] // factor
    v = target.reshape(s, factor, s, factor, s, factor).min(axis=(1, 3, 5))
    trunc = float(np.float16(voxel_size_input * 3).astype(np.float32))
    return np.minimum(v, trunc).astype(np.float32)


def point_cloud_to_grid(point_cloud: np.ndarray, grid_res: int, scale_factor: float, pad: int) -> np.ndarray:
    """util/misc.py:73-78: occupancy grid from a point cloud."""
    grid = np.zeros([grid_res + 2 * pad] * 3, dtype=np.float32)
    pc = point_cloud * scale_factor
    pg = np.clip(pc, 0, grid_res - 1).astype(np.uint32)
    grid[pad + pg[:, 0], pad + pg[:, 1], pad + pg[:, 2]] = 1
    return grid


# ---------------------------------------------------------------------------
# a2-a4: fold / unfold / patcher (bit-exact index permutations)
# ---------------------------------------------------------------------------


def unfold3d(x: np.ndarray, E: int) -> np.ndarray:
    """model/attention.py:186-188 Unfold3D.forward.
    [B,C,S,S,S] -> [B*(S/E)^3, C, E,E,E]; block index ((b*R+px)*R+py)*R+pz."""
    B, C, S = x.shape[0], x.shape[1], x.shape[2]
    R = S // E
    v = x.reshape(B, C, R, E, R, E, R, E)
    v = np.transpose(v, (0, 2, 4, 6, 1, 3, 5, 7))
    return np.ascontiguousarray(v).reshape(B * R * R * R, C, E, E, E)


def fold3d(x: np.ndarray, R: int, E: int, nf: int) -> np.ndarray:
    """model/attention.py:170-176 Fold3D.forward (inverse of unfold3d; the two
    nn.Fold calls with stride == kernel are pure re-indexing)."""
    v = x.reshape(-1, R, R, R, nf, E, E, E)
    B = v.shape[0]
    v = np.transpose(v, (0, 4, 1, 5, 2, 6, 3, 7))
    return np.ascontiguousarray(v).reshape(B, nf, R * E, R * E, R * E)


def unfold3d_pad_stride(x: np.ndarray, patch_extent: int, pad_size: int, pad_val: float, stride: int) -> np.ndarray:
    """model/attention.py:200-203 Unfold3DPadStride.forward.
    Constant-pad all three axes by pad_size, then overlapping unfold; the
    reference reshapes to (-1, 1, P, P, P), i.e. channels land in the row axis
    in (b, px, py, pz, c) order."""
    xp = np.pad(x, ((0, 0), (0, 0)) + ((pad_size, pad_size),) * 3, mode="constant", constant_values=pad_val)
    B, C = x.shape[:2]
    n = [(xp.shape[2 + a] - patch_extent) // stride + 1 for a in range(3)]
    out = np.empty((B, n[0], n[1], n[2], C, patch_extent, patch_extent, patch_extent), dtype=x.dtype)
    for ix in range(n[0]):
        for iy in range(n[1]):
            for iz in range(n[2]):
                out[:, ix, iy, iz] = xp[:, :, ix * stride: ix * stride + patch_extent,
                                         iy * stride: iy * stride + patch_extent,
                                         iz * stride: iz * stride + patch_extent]
    return out.reshape(-1, 1, patch_extent, patch_extent, patch_extent)


class PatcherOracle:
    """util/patcher.py:4-42 Patcher."""

    def __init__(self, patch_size, side, stride, pad_val, base_size):
        self.pad = list(side)
        self.stride = list(stride)
        self.pad_val = pad_val
        self.base_size = list(base_size)
        self.kernel = [patch_size[i] + 2 * side[i] for i in range(len(patch_size))]

    def __call__(self, x: np.ndarray) -> np.ndarray:  # util/patcher.py:14-19
        p = self.pad
        xp = np.pad(x, ((0, 0), (0, 0), (p[0], p[0]), (p[1], p[1]), (p[2], p[2])), mode="constant",
                    constant_values=self.pad_val)
        B, C = x.shape[:2]
        n = [(xp.shape[2 + a] - self.kernel[a]) // self.stride[a] + 1 for a in range(3)]
        out = np.empty((B, n[0], n[1], n[2], C, self.kernel[0], self.kernel[1], self.kernel[2]), dtype=x.dtype)
        for ix in range(n[0]):
            for iy in range(n[1]):
                for iz in range(n[2]):
                    x0, y0, z0 = ix * self.stride[0], iy * self.stride[1], iz * self.stride[2]
                    out[:, ix, iy, iz] = xp[:, :, x0: x0 + self.kernel[0], y0: y0 + self.kernel[1], z0: z0 + self.kernel[2]]
        return out.reshape(-1, C, self.kernel[0], self.kernel[1], self.kernel[2])

    def recompose_patches(self, original_shape, patches: np.ndarray) -> np.ndarray:  # util/patcher.py:21-30
        p = self.pad
        vol = np.full([original_shape[0], original_shape[1], original_shape[2] + 2 * p[0],
                       original_shape[3] + 2 * p[1], original_shape[4] + 2 * p[2]], self.pad_val, dtype=patches.dtype)
        ctr = 0
        # NB the reference's z loop bound uses patches.shape[2] (sic, :26)
        for x in range(0, vol.shape[2] - patches.shape[2] + 1, self.stride[0]):
            for y in range(0, vol.shape[3] - patches.shape[3] + 1, self.stride[1]):
                for z in range(0, vol.shape[4] - patches.shape[2] + 1, self.stride[2]):
                    vol[:, :, x: x + self.kernel[0], y: y + self.kernel[1], z: z + self.kernel[2]] = patches[:, ctr: ctr + 1]
                    ctr += 1
        return vol[:, :, p[0]: vol.shape[2] - p[0], p[1]: vol.shape[3] - p[1], p[2]: vol.shape[4] - p[2]]

    def get_patch_extents(self):
        return [self.kernel[i] - 2 * self.pad[i] for i in range(3)]

    def get_patch_ratio(self):
        return [self.base_size[i] // (self.kernel[i] - 2 * self.pad[i]) for i in range(3)]

    def get_stride_ratio(self):
        return [self.get_patch_extents()[i] // self.stride[i] for i in range(3)]

    def get_patch_counts(self):
        return [(self.base_size[i] + self.pad[i] * 2 - self.kernel[i]) // self.stride[i] + 1 for i in range(3)]


def get_extents_for_size(size, patch_size, patch_context, patch_stride) -> np.ndarray:
    """dataset/scene.py:153-160 SceneHandler.get_extents_for_size (x-major)."""
    ep = lambda x: x - patch_size
    ls = [np.linspace(0, ep(size[a]), ep(size[a]) // patch_stride + 1).astype(np.int32) for a in range(3)]
    xs, ys, zs = np.meshgrid(ls[0], ls[1], ls[2], indexing="ij")
    e = patch_size + 2 * patch_context
    cols = [xs, xs + e, ys, ys + e, zs, zs + e]
    return np.hstack([c.flatten()[:, None] for c in cols]).astype(np.int32)


def chunk_patches(chunk: np.ndarray, patch_size: int, patch_context: int, patch_stride: int,
                  trunc: float, mean: float, std: float) -> np.ndarray:
    """a1: the patches a PatchedSceneDataset yields for one scene/chunk:
    pad by context with the truncation constant (dataset/scene.py:61,94), slice
    [start, start+patch+2ctx) in get_extents_for_size order (scene.py:153-167),
    then (x-mean)/std (dataset/patched_scene_dataset.py:127-128).
    chunk: [S,S,S] fp32 -> [n,1,P,P,P] fp32."""
    padded = np.pad(chunk, patch_context, mode="constant", constant_values=trunc)
    ext = get_extents_for_size(chunk.shape, patch_size, patch_context, patch_stride)
    out = np.stack([padded[e[0]:e[1], e[2]:e[3], e[4]:e[5]] for e in ext])[:, None]
    return ((out - mean) / std).astype(np.float32)


def patch_occupancy(target_padded: np.ndarray, extent, target_voxel_size) -> int:
    """dataset/scene.py:149-151 calculate_occupancy_for_name: voxels of the PADDED patch extent that lie within
    0.75 * 2 voxels of the surface (the dataset keeps patches with occupancy > occupancy_threshold,
    patched_scene_dataset.py:28).  target_voxel_size is the float16-rounded voxel size (scene.py:31)."""
    e = extent
    return int((target_padded[e[0]:e[1], e[2]:e[3], e[4]:e[5]] <= 0.75 * 2 * target_voxel_size).sum())


def scene_patch_table(inp: np.ndarray, tgt: np.ndarray, d: dict, occupancy_threshold):
    """What PatchedSceneDataset enumerates for one scene (patched_scene_dataset.py:24-29,117-128): inp / tgt are the
    arrays stored on disk (fp16), d the dataset config.  Loading casts through float16 and pads by the context with
    the truncation value (dataset/scene.py:60-61,92-95).  Returns (input extents, target extents (padded coords),
    normalised input patches, normalised target patches, occupancies) of the KEPT patches, in dataset order."""
    itr, ttr = np.float32(f16_trunc(d["voxel_size_input"])), np.float32(f16_trunc(d["voxel_size_target"]))
    vox_t = np.float16(d["voxel_size_target"]).astype(np.float32)
    pin = np.pad(inp.astype(np.float16), d["patch_context_input"], mode="constant", constant_values=itr).astype(np.float32)
    ptg = np.pad(tgt.astype(np.float32), d["patch_context_target"], mode="constant", constant_values=ttr)
    size_t = list(tgt.shape)
    sf = d["patch_size_target"] / d["patch_size_input"]
    size_i = [int(s / sf) for s in size_t]
    stride_i = int(d["patch_stride"] * d["patch_size_input"] / d["patch_size_target"])
    et = get_extents_for_size(size_t, d["patch_size_target"], d["patch_context_target"], d["patch_stride"])
    ei = get_extents_for_size(size_i, d["patch_size_input"], d["patch_context_input"], stride_i)
    occ = np.array([patch_occupancy(ptg, e, vox_t) for e in et])
    keep = np.nonzero(occ > occupancy_threshold)[0]
    cut = lambda a, e: a[e[0]:e[1], e[2]:e[3], e[4]:e[5]]
    p_in = np.stack([(cut(pin, ei[i])[None] - d["input_mean"]) / d["input_std"] for i in keep]).astype(np.float32)
    p_tg = np.stack([(cut(ptg, et[i])[None] - d["target_mean"]) / d["target_std"] for i in keep]).astype(np.float32)
    return ei[keep], et[keep], p_in, p_tg, occ


# ---------------------------------------------------------------------------
# a5-a9: patch encoders (torch fp32 CPU functional restatement)
# ---------------------------------------------------------------------------

# (kernel, stride) per conv for each conv encoder, channels as multiples of nf.
# model/retrieval.py: Patch32 :4, Patch08 :136, PCPatch32 :187, PCPatch48 :217,
# PCPatch64 :247, Patch16 :277, Patch24 :306, Patch24V2 :335, Patch12 :364.
ENCODER_SPECS = {
    "Patch32": dict(patch=32, convs=[(1, 5, 1), (2, 3, 1), (4, 3, 2), (8, 3, 1), (8, 3, 2), (8, 4, 1)]),
    "Patch08": dict(patch=8, convs=[(1, 3, 1), (4, 3, 1), (4, 3, 1), (8, 2, 1)]),
    "PCPatch32": dict(patch=32, convs=[(1, 3, 1), (2, 3, 1), (4, 3, 2), (4, 3, 1), (8, 3, 2), (8, 3, 1), (8, 3, 1)]),
    "PCPatch48": dict(patch=48, convs=[(1, 5, 1), (2, 3, 1), (4, 3, 2), (4, 3, 2), (8, 3, 2), (8, 3, 1), (8, 2, 1)]),
    "PCPatch64": dict(patch=64, convs=[(1, 5, 1), (2, 3, 1), (4, 3, 2), (4, 3, 2), (8, 3, 2), (8, 3, 1), (8, 4, 1)]),
    "Patch16": dict(patch=16, convs=[(1, 3, 1), (2, 3, 1), (2, 3, 1), (4, 3, 1), (4, 3, 1), (8, 3, 1), (8, 4, 1)]),
    "Patch24": dict(patch=24, convs=[(1, 5, 1), (2, 3, 1), (2, 3, 2), (4, 3, 1), (8, 3, 1), (8, 3, 1), (8, 2, 1)]),
    "Patch24V2": dict(patch=24, convs=[(1, 3, 1), (2, 3, 1), (2, 3, 2), (4, 3, 1), (8, 3, 1), (8, 3, 1), (8, 3, 1)]),
    "Patch12": dict(patch=12, convs=[(1, 3, 1), (2, 3, 1), (4, 3, 1), (4, 3, 1), (8, 3, 1), (8, 2, 1)]),
    # BatchNorm variants (model/retrieval.py:31, :160): same convs, BN after each
    "PatchNorm32": dict(patch=32, convs=[(1, 5, 1), (2, 3, 1), (4, 3, 2), (8, 3, 1), (8, 3, 2), (8, 4, 1)], bn=True),
    "PatchNorm08": dict(patch=8, convs=[(1, 3, 1), (4, 3, 1), (4, 3, 1), (8, 2, 1)], bn=True),
}
# MLP encoders: hidden widths as multiples of nf. Patch04 :64, Patch05 :87, Patch04V2 :110
MLP_SPECS = {
    "Patch04": dict(patch=4, hidden=[4, 8, 16, 8]),
    "Patch05": dict(patch=5, hidden=[4, 8, 16, 8]),
    "Patch04V2": dict(patch=4, hidden=[4, 8, 16, 16, 8]),
}


def encoder_param_shapes(name: str, nf: int, z_dim: int) -> dict:
    """Parameter names/shapes of the reference encoder classes (state_dict keys)."""
    shapes = {}
    if name in MLP_SPECS:
        spec = MLP_SPECS[name]
        widths = [spec["patch"] ** 3] + [h * nf for h in spec["hidden"]] + [z_dim]
        for i in range(len(widths) - 1):
            shapes[f"layers.{2 * i}.weight"] = (widths[i + 1], widths[i])
            shapes[f"layers.{2 * i}.bias"] = (widths[i + 1],)
        return shapes
    spec = ENCODER_SPECS[name]
    cin = 1
    step = 3 if spec.get("bn") else 2
    for i, (m, k, _s) in enumerate(spec["convs"]):
        shapes[f"layers.{step * i}.weight"] = (m * nf, cin, k, k, k)
        shapes[f"layers.{step * i}.bias"] = (m * nf,)
        if spec.get("bn"):
            for p in ("weight", "bias", "running_mean", "running_var"):
                shapes[f"layers.{step * i + 1}.{p}"] = (m * nf,)
            shapes[f"layers.{step * i + 1}.num_batches_tracked"] = ()
        cin = m * nf
    shapes["final_layer.weight"] = (z_dim, cin)
    shapes["final_layer.bias"] = (z_dim,)
    return shapes


def encoder_forward(name: str, sd: dict, x: torch.Tensor) -> torch.Tensor:
    """forward of model/retrieval.py encoders: [N,1,P,P,P] -> [N,z,1,1,1]."""
    if name in MLP_SPECS:
        n_lin = len(MLP_SPECS[name]["hidden"]) + 1
        h = x.reshape(x.shape[0], -1)
        for i in range(n_lin):
            h = F.linear(h, sd[f"layers.{2 * i}.weight"], sd[f"layers.{2 * i}.bias"])
            if i < n_lin - 1:
                h = F.relu(h)
        return h.reshape(h.shape[0], h.shape[1], 1, 1, 1)
    spec = ENCODER_SPECS[name]
    step = 3 if spec.get("bn") else 2
    h = x
    for i, (_m, _k, s) in enumerate(spec["convs"]):
        h = F.conv3d(h, sd[f"layers.{step * i}.weight"], sd[f"layers.{step * i}.bias"], stride=s)
        if spec.get("bn"):  # eval-mode BatchNorm3d
            p = f"layers.{step * i + 1}."
            h = F.batch_norm(h, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                             training=False, eps=1e-5)
        h = F.leaky_relu(h, 0.2)
    h = F.linear(h.reshape(h.shape[0], -1), sd["final_layer.weight"], sd["final_layer.bias"])
    return h.reshape(h.shape[0], h.shape[1], 1, 1, 1)


def normalize_features(feat: torch.Tensor, latent_dim: int) -> torch.Tensor:
    """util/retrieval.py:38,66: permute(0,2,3,4,1).reshape(-1,D) then F.normalize(dim=1)."""
    return F.normalize(feat.permute(0, 2, 3, 4, 1).reshape(-1, latent_dim), dim=1)


# ---------------------------------------------------------------------------
# a10-a11: database rows, exact kNN, source-scene demotion
# ---------------------------------------------------------------------------


def zero_patch_row(encode_fn, patch_size: int, patch_context: int, latent_dim: int) -> np.ndarray:
    """util/retrieval.py:21-26 get_zero_patch_entry: the sentinel row is the
    embedding of an all-ONES patch, scene index -1, extent [0, patch_size]x3."""
    ones = torch.ones([1, 1] + [patch_size + 2 * patch_context] * 3, dtype=torch.float32)
    emb = normalize_features(encode_fn(ones), latent_dim).numpy()
    meta = np.array([[-1, 0, patch_size, 0, patch_size, 0, patch_size]], dtype=np.float32)
    return np.hstack([meta, emb]).astype(np.float32)


def database_rows(scene_idx: np.ndarray, extents_padded: np.ndarray, patch_context: int, emb: np.ndarray) -> np.ndarray:
    """util/retrieval.py:39-44: row = [scene_idx, x0,x1,y0,y1,z0,z1 (unpadded:
    end - 2*ctx, patched_scene_dataset.py:101-105), emb(64)] as fp32."""
    e = extents_padded.astype(np.float32).copy()
    e[:, 1::2] -= 2 * patch_context
    return np.hstack([scene_idx.astype(np.float32)[:, None], e, emb.astype(np.float32)]).astype(np.float32)


_KNN_LIB = None


def _knn_lib():
    """Loads oracle/_build/libknn_oracle.so (plain C, built by __graft_entry__.build())."""
    global _KNN_LIB
    if _KNN_LIB is None:
        path = os.path.join(_HERE, "_build", "libknn_oracle.so")
        if not os.path.exists(path):
            return None
        lib = ctypes.CDLL(path)
        lib.rf_oracle_knn.restype = ctypes.c_int
        lib.rf_oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _KNN_LIB = lib
    return _KNN_LIB


def knn_exact(db: np.ndarray, q: np.ndarray, k: int, threads: int = 0, force_numpy: bool = False):
    """Exact kNN under THE CANONICAL RULE (defined by this build; the reference's
    FLANN call util/retrieval.py:92 `nn_index(q, 2K)` is approximate):

        d(q, x) = sum_{i=0..D-1, in that order} (double(q_i) - double(x_i))^2
                  each op correctly rounded in IEEE binary64, no FMA contraction
        result  = the k rows smallest under ascending (d, row_index)
        dist    = float32(d)          (FLANN returns squared L2 as fp32)

    Returns (idx int32 [Q,k], dist fp32 [Q,k]).  Uses the plain-C restatement
    (oracle/knn_oracle.c) when built, else the numpy loop below (same
    arithmetic, vectorised over pairs, sequential over i)."""
    db = np.ascontiguousarray(db, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    N, D = db.shape
    Q = q.shape[0]
    assert k <= N
    lib = None if force_numpy else _knn_lib()
    if lib is not None:
        idx = np.empty((Q, k), dtype=np.int32)
        dist = np.empty((Q, k), dtype=np.float32)
        rc = lib.rf_oracle_knn(db.ctypes.data, N, q.ctypes.data, Q, D, k, idx.ctypes.data, dist.ctypes.data,
                               threads or (os.cpu_count() or 1))
        assert rc == 0
        return idx, dist
    idx = np.empty((Q, k), dtype=np.int32)
    dist = np.empty((Q, k), dtype=np.float32)
    dbd = db.astype(np.float64)
    qb = max(1, min(Q, (1 << 24) // max(N, 1)))
    for s in range(0, Q, qb):
        qd = q[s:s + qb].astype(np.float64)
        acc = np.zeros((qd.shape[0], N), dtype=np.float64)
        for i in range(D):
            diff = qd[:, i:i + 1] - dbd[None, :, i]
            acc = acc + diff * diff
        # ascending (d, index): stable argsort on d keeps index order among ties
        order = np.argsort(acc, axis=1, kind="stable")[:, :k]
        idx[s:s + qb] = order.astype(np.int32)
        dist[s:s + qb] = np.take_along_axis(acc, order, axis=1).astype(np.float32)
    return idx, dist


def demote_same_scene(idx2k: np.ndarray, dist2k: np.ndarray, row_scene: np.ndarray, query_scene: np.ndarray, K: int):
    """util/retrieval.py:94-99: when the query's own scene is in the index,
    hits from that scene are moved BEHIND all other hits (stable on both
    sides: np.concatenate((rows[~M], rows[M]))), then the first K are kept.
    query_scene < 0 means 'not in the index / do not demote' (val split or
    ignore_patches_from_source False)."""
    Q, K2 = idx2k.shape
    out_idx = np.empty((Q, K), dtype=np.int32)
    out_dist = np.empty((Q, K), dtype=np.float32)
    for i in range(Q):
        if query_scene[i] >= 0:
            M = row_scene[idx2k[i]] == query_scene[i]
            order = np.concatenate((np.nonzero(~M)[0], np.nonzero(M)[0]))
        else:
            order = np.arange(K2)
        out_idx[i] = idx2k[i, order[:K]]
        out_dist[i] = dist2k[i, order[:K]]
    return out_idx, out_dist


def mapping_rows(meta: np.ndarray, idx: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """util/retrieval.py:93,100: per query a [K, 8] fp32 array
    [scene_idx, x0,x1,y0,y1,z0,z1, dist]; meta = database[:, 0:7]."""
    return np.concatenate([meta[idx], dist[..., None]], axis=-1).astype(np.float32)


# ---------------------------------------------------------------------------
# a12: compose
# ---------------------------------------------------------------------------


def compose_from_mapping(mapping: np.ndarray, dst_extents: np.ndarray, scene_store: np.ndarray, scene_size,
                         trunc: float, trunc_train: float, no_overlap: bool = True) -> np.ndarray:
    """util/retrieval.py:145-164 create_retrieval_from_mapping for one scene.
    mapping: [P, K, 8] rows for the scene's P patches (in patch order);
    dst_extents: [P, 6] UNPADDED destination extents of those patches;
    scene_store: [S, X, Y, Z] fp32 UNPADDED train targets - `dataset_train` at
    :158 is a PatchedSceneDataset, whose get_scene_target strips the context
    padding again (patched_scene_dataset.py:87-99), and the database extents are
    unpadded too (create_dictionary stores dataset.unpad(extent), :40-44), so
    [X0:X1] addresses the 16^3 core of the retrieved patch.
    Overlapping strides keep the candidate with the lower mean distance (:156)."""
    P, K = mapping.shape[:2]
    out = np.full((K,) + tuple(scene_size), trunc, dtype=np.float32)
    distances = np.full_like(out, 100.0)
    ratio = np.float32(trunc) / np.float32(trunc_train)
    for k in range(K):
        for p in range(P):
            X0, X1, Y0, Y1, Z0, Z1 = mapping[p, k, 1:7].astype(np.int32).tolist()
            cur = mapping[p, k, 7]
            xx0, xx1, yy0, yy1, zz0, zz1 = [int(v) for v in dst_extents[p]]
            if no_overlap or distances[k, xx0:xx1, yy0:yy1, zz0:zz1].mean() > cur:
                ptr = int(mapping[p, k, 0])
                if ptr >= 0:
                    block = scene_store[ptr][X0:X1, Y0:Y1, Z0:Z1]
                else:
                    # :161 np.ones(...) * trunc is float64; it is cast on assignment below
                    block = np.ones((X1 - X0, Y1 - Y0, Z1 - Z0)) * np.float32(trunc)
                out[k, xx0:xx1, yy0:yy1, zz0:zz1] = block * ratio
                distances[k, xx0:xx1, yy0:yy1, zz0:zz1] = float(cur)
    return out


# ---------------------------------------------------------------------------
# a13, a15, a16: 3D U-Nets (model/unet.py, model/refinement.py)
# ---------------------------------------------------------------------------


def _single_conv_shapes(prefix, cin, cout, order, shapes):
    """model/unet.py:19-76 create_conv parameter shapes."""
    for i, ch in enumerate(order):
        if ch == "c":
            shapes[f"{prefix}.conv.weight"] = (cout, cin, 3, 3, 3)
            if not ("g" in order or "b" in order):
                shapes[f"{prefix}.conv.bias"] = (cout,)
        elif ch == "g":
            n = cin if i < order.index("c") else cout
            shapes[f"{prefix}.groupnorm.weight"] = (n,)
            shapes[f"{prefix}.groupnorm.bias"] = (n,)
        elif ch == "b":
            raise NotImplementedError("batchnorm layer order is not used by any shipped config")


def _double_conv_channels(cin, cout, encoder):
    """model/unet.py:125-137 DoubleConv channel rule."""
    if encoder:
        c1 = cout // 2
        if c1 < cin:
            c1 = cin
        return [(cin, c1), (c1, cout)]
    return [(cin, cout), (cout, cout)]


def _stepdown_channels(cin, cout):
    """model/unet.py:147-159 StepDownDoubleConv."""
    mid = (cin + cout) // 2
    return [(cin, mid), (mid, cout)]


def unet3d_plan(in_channels, out_channels, f_maps, num_levels, remove_n_final_layers=0):
    """model/unet.py:418-472 Abstract3DUNet.__init__ with basic_module=DoubleConv,
    final_conv=False. Returns (encoders, decoders) as lists of [(cin,cout),(cin,cout)]."""
    f = [f_maps * 2 ** k for k in range(num_levels)]  # :11-12
    encoders = []
    for i, o in enumerate(f):
        encoders.append(_double_conv_channels(in_channels if i == 0 else f[i - 1], o, True))
    rev = list(reversed(f))
    if remove_n_final_layers > 0:
        rev = rev[:-remove_n_final_layers]
    mod = list(rev)
    mod[-1] = out_channels  # final_conv False (:447-448)
    decoders = []
    for i in range(len(rev) - 1):
        cin = rev[i] + rev[i + 1]
        cout = mod[i + 1]
        if i == len(rev) - 2 and remove_n_final_layers > 0:  # :458-460
            decoders.append(_stepdown_channels(cin, cout))
        else:
            decoders.append(_double_conv_channels(cin, cout, False))
    return encoders, decoders


def unet3d_param_shapes(prefix, in_channels, out_channels, f_maps, num_levels, remove_n_final_layers, order="gcr"):
    enc, dec = unet3d_plan(in_channels, out_channels, f_maps, num_levels, remove_n_final_layers)
    shapes = {}
    for i, convs in enumerate(enc):
        for j, (ci, co) in enumerate(convs):
            _single_conv_shapes(f"{prefix}encoders.{i}.basic_module.SingleConv{j + 1}", ci, co, order, shapes)
    for i, convs in enumerate(dec):
        for j, (ci, co) in enumerate(convs):
            _single_conv_shapes(f"{prefix}decoders.{i}.basic_module.SingleConv{j + 1}", ci, co, order, shapes)
    return shapes


def _single_conv(x, sd, prefix, order, num_groups):
    """model/unet.py:79-100 SingleConv forward for orders over {g,c,r,l,e}."""
    for i, ch in enumerate(order):
        if ch == "g":
            C = x.shape[1]
            g = num_groups if C >= num_groups else 1  # :60-62
            x = F.group_norm(x, g, sd[f"{prefix}.groupnorm.weight"], sd[f"{prefix}.groupnorm.bias"], eps=1e-5)
        elif ch == "c":
            x = F.conv3d(x, sd[f"{prefix}.conv.weight"], sd.get(f"{prefix}.conv.bias"), padding=1)
        elif ch == "r":
            x = F.relu(x)
        elif ch == "l":
            x = F.leaky_relu(x, 0.1)  # :46
        elif ch == "e":
            x = F.elu(x)
        else:
            raise NotImplementedError(ch)
    return x


def _double_conv(x, sd, prefix, order, num_groups):
    x = _single_conv(x, sd, prefix + ".SingleConv1", order, num_groups)
    return _single_conv(x, sd, prefix + ".SingleConv2", order, num_groups)


def unet3d_forward(x, sd, prefix, num_levels, n_decoders, order, num_groups):
    """model/unet.py:492-520 Abstract3DUNet.forward (final_conv Identity, no activation)."""
    feats = []
    for i in range(num_levels):
        if i > 0:
            x = F.max_pool3d(x, 2)  # Encoder.forward :250-254
        x = _double_conv(x, sd, f"{prefix}encoders.{i}.basic_module", order, num_groups)
        feats.insert(0, x)
    feats = feats[1:]
    for i in range(n_decoders):  # zip(decoders, feats) truncates to the decoders that exist
        ef = feats[i]
        x = F.interpolate(x, size=ef.shape[2:], mode="nearest")  # Upsampling :352-358
        x = torch.cat((ef, x), dim=1)  # Decoder._joining :303-306
        x = _double_conv(x, sd, f"{prefix}decoders.{i}.basic_module", order, num_groups)
    return x


def decoder_nojoin_forward(x, sd, prefix, order, num_groups):
    """model/unet.py:311-322 DecoderNoJoining.forward: nearest x2 then DoubleConv
    (the torch.randn it draws only carries a size)."""
    x = F.interpolate(x, size=[2 * s for s in x.shape[2:]], mode="nearest")
    return _double_conv(x, sd, prefix + ".basic_module", order, num_groups)


def retrieval_backbone_shapes(nf, f_maps, num_levels, order="gcr"):
    """model/refinement.py:64-73 RetrievalUNetBackbone."""
    return unet3d_param_shapes("network.", 1, nf, f_maps, num_levels, 1, order)


def retrieval_backbone_forward(x, sd, nf, f_maps, num_levels, order="gcr"):
    _enc, dec = unet3d_plan(1, nf, f_maps, num_levels, 1)
    return unet3d_forward(x, sd, "network.", num_levels, len(dec), order, nf // 2)


def _nojoin_shapes(prefix, cin, cout, order, shapes):
    for j, (ci, co) in enumerate(_double_conv_channels(cin, cout, False)):
        _single_conv_shapes(f"{prefix}.basic_module.SingleConv{j + 1}", ci, co, order, shapes)


def unet_backbone_shapes(kind, nf, num_levels, order="gcr"):
    """model/refinement.py:6-45. kind in {'sr08','sr16','surface'}."""
    if kind == "surface":
        return unet3d_param_shapes("network.", 1, nf, nf, num_levels, 2, order)
    shapes = unet3d_param_shapes("network.0.", 1, 2 * nf, nf, num_levels, 0, order)
    if kind == "sr08":
        _nojoin_shapes("network.1", 2 * nf, 2 * nf, order, shapes)
        _nojoin_shapes("network.2", 2 * nf, nf, order, shapes)
    elif kind == "sr16":
        _nojoin_shapes("network.1", 2 * nf, nf, order, shapes)
    else:
        raise ValueError(kind)
    return shapes


def unet_backbone_forward(kind, x, sd, nf, num_levels, order="gcr"):
    g = nf // 2
    if kind == "surface":
        _e, dec = unet3d_plan(1, nf, nf, num_levels, 2)
        return unet3d_forward(x, sd, "network.", num_levels, len(dec), order, g)
    _e, dec = unet3d_plan(1, 2 * nf, nf, num_levels, 0)
    x = unet3d_forward(x, sd, "network.0.", num_levels, len(dec), order, g)
    x = decoder_nojoin_forward(x, sd, "network.1", order, g)
    if kind == "sr08":
        x = decoder_nojoin_forward(x, sd, "network.2", order, g)
    return x


def final_decoder_shapes(nf, order="gcr"):
    """model/refinement.py:48-61 Superresolution08FinalDecoder."""
    shapes = {}
    _nojoin_shapes("network.0", nf, nf, order, shapes)
    shapes["network.1.weight"] = (1, nf, 1, 1, 1)
    shapes["network.1.bias"] = (1,)
    return shapes


def final_decoder_forward(x, sd, nf, order="gcr"):
    x = decoder_nojoin_forward(x, sd, "network.0", order, nf // 2)
    x = F.conv3d(x, sd["network.1.weight"], sd["network.1.bias"])
    return torch.tanh(x)


# ---------------------------------------------------------------------------
# a14: patch attention (model/attention.py)
# ---------------------------------------------------------------------------


def attention_shapes(nf, e, cf_feat=32, output_mapping=False):
    """PatchedAttentionBlock state_dict; output_mapping: attn_no_output_mapping=False adds the g / o 1x1x1
    convolutions (model/attention.py:56-57)."""
    n_in = nf * e ** 3
    shapes = {"attention_blocks_layer.sig_scale": (1,), "attention_blocks_layer.sig_shift": (1,)}
    if output_mapping:
        for m in ("g", "o"):
            shapes[f"attention_blocks_layer.{m}.weight"] = (nf, nf, 1, 1, 1)
            shapes[f"attention_blocks_layer.{m}.bias"] = (nf,)
    for br in ("theta", "phi"):
        w = [n_in, 128, 128, 128, cf_feat]  # model/attention.py:35-41
        for i in range(4):
            shapes[f"attention_blocks_layer.{br}.encoder.{2 * i}.weight"] = (w[i + 1], w[i])
            shapes[f"attention_blocks_layer.{br}.encoder.{2 * i}.bias"] = (w[i + 1],)
    return shapes


def _attn_mlp(x, sd, br):
    """model/attention.py:29-46 AttentionFeatureEncoder (LeakyReLU default slope 0.01)."""
    h = x
    for i in range(4):
        h = F.linear(h, sd[f"attention_blocks_layer.{br}.encoder.{2 * i}.weight"],
                     sd[f"attention_blocks_layer.{br}.encoder.{2 * i}.bias"])
        if i < 3:
            h = F.leaky_relu(h, 0.01)
    return h


def attention_block_forward(x, p, sd, normalize=True, retrieval_mode=False, blend=True, gumbel_noise=None):
    """model/attention.py:84-113 AttentionBlock.forward; g = o = Identity unless the state dict holds their 1x1x1
    convolutions (attn_no_output_mapping=False, :56-57,95,108).
    x: [b, C, e,e,e]; p: [b, k, C, e,e,e].  In retrieval mode the Gumbel noise
    must be supplied ([b,k]); gumbel_softmax(hard=True) forward value is
    y_hard - y_soft + y_soft (torch/nn/functional.py gumbel_softmax)."""
    b, k, c, e = p.shape[0], p.shape[1], p.shape[2], p.shape[3]
    xf = _attn_mlp(x.reshape(b, -1), sd, "theta")
    pf = _attn_mlp(p.reshape(b * k, -1), sd, "phi").reshape(b, k, -1)
    if normalize:
        xf = F.normalize(xf, dim=1)
        pf = F.normalize(pf, dim=2)
    mapped = "attention_blocks_layer.g.weight" in sd
    if mapped:  # g = Conv3d(C, C, 1): per-voxel channel mixing + bias of every candidate (:95)
        wg, bg = sd["attention_blocks_layer.g.weight"].flatten(1), sd["attention_blocks_layer.g.bias"]
        g = (torch.einsum("oc,bkcs->bkos", wg, p.reshape(b, k, c, -1)) + bg[None, None, :, None]).reshape(b, k, -1)
    else:
        g = p.reshape(b, k, -1)
    scores = torch.einsum("ij,ijk->ik", xf, pf.permute(0, 2, 1))
    switch = F.relu(scores.max(dim=1, keepdim=True).values)  # MaxPool1d(K) over all k, then ReLU (:99)
    if retrieval_mode:
        logits = scores * 25
        y = F.softmax((logits + gumbel_noise) / 1.0, dim=-1)
        ind = y.max(dim=-1, keepdim=True)[1]
        y_hard = torch.zeros_like(y).scatter_(-1, ind, 1.0)
        w = y_hard - y + y
    else:
        sharp = (32 * e * e * e) * 4  # cf_feat * e^3 * 4 (:105)
        w = F.softmax(sharp * scores, dim=1)
    ws = torch.einsum("ij,ijk->ik", w, g)
    if mapped:  # o = Conv3d(C, C, 1) on the weighted sum (:108)
        wo, bo = sd["attention_blocks_layer.o.weight"].flatten(1), sd["attention_blocks_layer.o.bias"]
        ws = (torch.einsum("oc,bcs->bos", wo, ws.reshape(b, c, -1)) + bo[None, :, None]).reshape(b, -1)
    xv = x.reshape(b, -1)
    if blend:
        out = xv * (1 - switch) + ws * switch
    else:
        out = xv + ws * switch
    return out.reshape(b, c, e, e, e)


def patched_attention_forward(x_pred, x_retr, sd, nf, num_patch_x, e, K, **kw):
    """model/attention.py:141-157 PatchedAttentionBlock.forward."""
    xu = torch.from_numpy(unfold3d(x_pred.numpy(), e))
    pu = torch.from_numpy(unfold3d(x_retr.reshape(-1, nf, *x_retr.shape[2:]).numpy(), e))
    R = num_patch_x
    pu = pu.reshape(-1, K, R, R, R, nf, e, e, e).permute(0, 2, 3, 4, 1, 5, 6, 7, 8).reshape(-1, K, nf, e, e, e)
    out = attention_block_forward(xu, pu, sd, **kw)
    return torch.from_numpy(fold3d(out.numpy(), R, e, nf))


def attention_get_features(x_pred, x_tgt, occupancy, sd, nf, e, normalize=True):
    """model/attention.py:132-139 + :74-82 get_features."""
    xu = torch.from_numpy(unfold3d(x_pred.numpy(), e))
    tu = torch.from_numpy(unfold3d(x_tgt.numpy(), e))
    ou = unfold3d(occupancy.numpy(), e)
    xf = _attn_mlp(xu.reshape(xu.shape[0], -1), sd, "theta")
    pf = _attn_mlp(tu.reshape(tu.shape[0], -1), sd, "phi")
    if normalize:
        xf, pf = F.normalize(xf, dim=1), F.normalize(pf, dim=1)
    occ = torch.from_numpy(ou.reshape(ou.shape[0], -1).any(axis=1))
    return xf, pf, occ


# ---------------------------------------------------------------------------
# a17: refinement forward glue (trainer/train_refinement.py:108-120, 255-257)
# ---------------------------------------------------------------------------


def refine_forward(inp, retrieval, sds, cfg):
    """Inference part of RefinementTrainingModule.forward_full: returns pred_shape.
    inp [B,1,s,s,s]; retrieval [B,K',64,64,64]; sds: dict of state dicts
    {unet_backbone, retrieval_backbone, attention, decoder}; cfg: dict with
    kind, nf, unet_num_level, retrieval_fmaps, retrieval_num_level, K, E."""
    nf, K, E = cfg["nf"], cfg["K"], cfg["E"]
    x_back = unet_backbone_forward(cfg["kind"], inp, sds["unet_backbone"], nf, cfg["unet_num_level"])
    b, _k, s = retrieval.shape[0:3]
    retr = retrieval[:, :K].reshape(b * K, 1, s, s, s)  # get_retrievals :255-257
    patches = torch.from_numpy(unfold3d(retr.numpy(), 16))  # Unfold3D(16,1) :34
    feats = retrieval_backbone_forward(patches, sds["retrieval_backbone"], nf, cfg["retrieval_fmaps"],
                                       cfg["retrieval_num_level"])
    x_retr = torch.from_numpy(fold3d(feats.numpy(), 4, 8, nf))  # Fold3D(4,8,nf) :37
    x = patched_attention_forward(x_back, x_retr, sds["attention"], nf, 32 // E, E, K,
                                  retrieval_mode=cfg.get("retrieval_mode", False))
    pred = final_decoder_forward(x, sds["decoder"], nf)
    return pred, x_back, x_retr, x


def network_pred_to_df(pred, trunc):
    """trainer/train_refinement.py:242-243."""
    return (pred + 1) * trunc / 2


# ---------------------------------------------------------------------------
# whole hot path on the CPU (used by smoke(), the tests and bench.py's
# cpu_baseline / --impl reference legs)
# ---------------------------------------------------------------------------

_INPUT_ENCODER = {"2+1": "Patch04", "2+1V2": "Patch04V2", "4+2": "Patch08", "4+2N": "PatchNorm08", "16+4": "Patch24",
                  "pc_16+8": "PCPatch32", "pc_32+8": "PCPatch48", "pc_32+16": "PCPatch64"}  # model/__init__.py:8-23
_TARGET_ENCODER = {"pc_32+16": "PCPatch64", "8+2": "Patch12", "8+4": "Patch16", "16+4": "Patch24", "16+4V2": "Patch24V2",
                   "16+8": "Patch32", "16+8N": "PatchNorm32"}  # model/__init__.py:24-37


def f16_trunc(voxel_size):
    return float(np.float16(voxel_size * 3).astype(np.float32))  # dataset/scene.py:32-33


def encode_chunk_queries(config, sd_fenc_input, chunks: np.ndarray) -> np.ndarray:
    """chunks [B,1,s,s,s] raw low-res SDF -> unit queries [B*P, D]: the dataloader's
    patch extraction + normalisation (a1), the input encoder (a5/a6), F.normalize (a9)."""
    d = config["dataset"]
    stride_in = int(d["patch_stride"] * d["patch_size_input"] / d["patch_size_target"])
    pats = np.concatenate([chunk_patches(c[0], d["patch_size_input"], d["patch_context_input"], stride_in,
                                         f16_trunc(d["voxel_size_input"]), d["input_mean"], d["input_std"])
                           for c in chunks])
    name = _INPUT_ENCODER[config["retrieval_model"]["network_input"]]
    feat = encoder_forward(name, sd_fenc_input, torch.from_numpy(pats))
    return normalize_features(feat, config["retrieval_model"]["latent_dim"]).numpy()


def lookup_rows(bank_emb, bank_meta, q, K, query_scene=None, threads=0):
    """util/retrieval.py:92-100 with the exact kNN: -> (rows [Q,K,8], idx [Q,K])."""
    k2 = min(2 * K, bank_emb.shape[0])
    idx2k, d2k = knn_exact(bank_emb, q, k2, threads=threads)
    qs = np.full(q.shape[0], -1, dtype=np.int64) if query_scene is None else np.asarray(query_scene)
    idx, dist = demote_same_scene(idx2k, d2k, bank_meta[:, 0].astype(np.int64), qs, K)
    return mapping_rows(bank_meta, idx, dist), idx


def chunk_dst_extents(config) -> np.ndarray:
    d = config["dataset"]
    ext = get_extents_for_size([d["target_chunk_size"]] * 3, d["patch_size_target"], d["patch_context_target"], d["patch_stride"])
    ext = ext.copy()
    ext[:, 1::2] -= 2 * d["patch_context_target"]  # unpad
    return ext


def compose_chunks(config, rows, scene_store, n_chunks) -> np.ndarray:
    d = config["dataset"]
    P = rows.shape[0] // n_chunks
    trunc = np.float32(f16_trunc(d["voxel_size_target"]))
    dst = chunk_dst_extents(config)
    c = d["target_chunk_size"]
    return np.stack([compose_from_mapping(rows[i * P:(i + 1) * P], dst, scene_store, (c, c, c), trunc, trunc)
                     for i in range(n_chunks)])


def refine_chunks(config, sds, chunks, retrieval_raw):
    """dataloader normalisation (patched_scene_dataset.py:127-133) + forward_full's inference part."""
    d = config["dataset"]
    x_in = torch.from_numpy(((chunks - d["input_mean"]) / d["input_std"]).astype(np.float32))
    x_re = torch.from_numpy(((retrieval_raw - d["target_mean"]) / d["target_std"]).astype(np.float32))
    kind = {8: "sr08", 16: "sr16", 128: "surface"}[d["input_chunk_size"]]
    cfg = dict(kind=kind, nf=config["nf"], unet_num_level=config["unet_num_level"], retrieval_fmaps=config["retrieval_fmaps"],
               retrieval_num_level=config["retrieval_num_level"], K=config["K"], E=config["attn_patch_extent"] // 2,
               retrieval_mode=config["attn_retrieval_mode"])
    return refine_forward(x_in, x_re, sds, cfg)


# ---------------------------------------------------------------------------
# callers either side of the path (training-side loss and
# normals, evaluation metrics)
# ---------------------------------------------------------------------------

# dataset/patched_scene_dataset.py:194-196 (class attributes of PatchedSceneDataset)
SOBEL_3D_X = np.array([[[+1, +2, +1], [+2, +4, +2], [+1, +2, +1]], [[0, 0, 0], [0, 0, 0], [0, 0, 0]],
                       [[-1, -2, -1], [-2, -4, -2], [-1, -2, -1]]], dtype=np.float32)
SOBEL_3D_Y = np.array([[[+1, +2, +1], [0, 0, 0], [-1, -2, -1]], [[+2, +4, +2], [0, 0, 0], [-2, -4, -2]],
                       [[+1, +2, +1], [0, 0, 0], [-1, -2, -1]]], dtype=np.float32)
SOBEL_3D_Z = np.array([[[-1, 0, +1], [-2, 0, +2], [-1, 0, +1]], [[-2, 0, +2], [-4, 0, +4], [-2, 0, +2]],
                       [[-1, 0, +1], [-2, 0, +2], [-1, 0, +1]]], dtype=np.float32)


def compute_normals(target: torch.Tensor, target_trunc: float) -> torch.Tensor:
    """dataset/patched_scene_dataset.py:139-146 compute_normals, line by line."""
    padded = F.pad(target, [1, 1, 1, 1, 1, 1], mode="constant", value=float(target_trunc))
    k = [torch.from_numpy(a)[None, None] for a in (SOBEL_3D_X, SOBEL_3D_Y, SOBEL_3D_Z)]
    normals = torch.cat([F.conv3d(padded, kk) for kk in k], dim=1)
    normalizer = torch.sqrt(torch.square(normals).sum(dim=1, keepdim=True) + 1e-5)
    return torch.div(normals, normalizer)


def ntxent_loss(zis: torch.Tensor, zjs: torch.Tensor, temperature: float, use_cosine: bool = True, iou_matrix=None,
                sig_scale: float = 80, sig_shift: float = -65) -> torch.Tensor:
    """model/loss.py:48-69 NTXentLoss.forward (CPU restatement: `.cuda()` calls dropped, nothing else changed)."""
    n = zis.shape[0]
    rep = torch.cat([zjs, zis], dim=0)
    if use_cosine:
        sim = torch.nn.CosineSimilarity(dim=-1)(rep.unsqueeze(1), rep.unsqueeze(0))          # :44
    else:
        sim = torch.tensordot(rep.unsqueeze(1), rep.T.unsqueeze(0), dims=2)                  # :35
    l_pos = torch.diag(sim, n)
    r_pos = torch.diag(sim, -n)
    positives = torch.cat([l_pos, r_pos]).view(2 * n, 1)
    mask = torch.from_numpy(1 - (np.eye(2 * n) + np.eye(2 * n, 2 * n, k=-n) + np.eye(2 * n, 2 * n, k=n))).type(torch.bool)  # :25-31
    negatives = sim[mask].view(2 * n, -1)
    logits = torch.cat((positives, negatives), dim=1)
    if iou_matrix is None:
        logits = logits / temperature
    else:
        neg_iou = iou_matrix[mask].view(2 * n, -1)
        logits[:, 0] /= temperature
        logits[:, 1:] /= (temperature + (1 - temperature) * torch.sigmoid(neg_iou * sig_scale + sig_shift))
    labels = torch.zeros(2 * n).long()
    return torch.nn.CrossEntropyLoss(reduction="sum")(logits, labels) / (2 * n)


def sliced_attn_nt_xent(temperature, batch_size, x_fpred, x_ftgt, occupancy, budget=1280):
    """trainer/train_refinement.py:208-221 compute_sliced_attn_nt_xent_loss."""
    split = x_fpred.shape[0] // batch_size
    total = 0
    loss = torch.zeros(1)
    for b in range(batch_size):
        occ = occupancy[b * split:(b + 1) * split] > 0
        if occ.sum() > 0 and total + int(occ.sum()) <= budget:
            loss = ntxent_loss(x_fpred[b * split:(b + 1) * split][occ], x_ftgt[b * split:(b + 1) * split][occ], temperature) + loss
            total += int(occ.sum())
    return loss


def occupancy_metrics(pred: np.ndarray, target: np.ndarray):
    """util/metrics.py:15-24 (IoU), :66-67 (Precision), :83-84 (Recall) for ONE update call on bool [B,1,S,S,S]:
    returns (iou_sum, iou_total, precision_sum, recall_sum, n) exactly as the metric states accumulate them
    (torch float32 arithmetic on the integer sums)."""
    p, t = torch.from_numpy(pred), torch.from_numpy(target)
    inter = (p & t).sum(-1).sum(-1).sum(-1).squeeze(1)
    union = (p | t).sum(-1).sum(-1).sum(-1).squeeze(1)
    valid = union > 0
    iou_sum, iou_total = 0.0, 0
    if union[valid].sum() > 0:
        iou_sum = float((inter[valid] / (union[valid] + 1e-5)).sum())
        iou_total = int(valid.sum())
    prec = float((inter / (p.sum(-1).sum(-1).sum(-1).squeeze(1) + 1e-5)).sum())
    rec = float((inter / (t.sum(-1).sum(-1).sum(-1).squeeze(1) + 1e-5)).sum())
    counts = np.stack([inter.numpy(), union.numpy(), p.sum((-1, -2, -3)).squeeze(1).numpy(), t.sum((-1, -2, -3)).squeeze(1).numpy()], 1)
    return iou_sum, iou_total, prec, rec, counts.astype(np.int64)


def chamfer_nn(a: np.ndarray, b: np.ndarray):
    """Nearest neighbour in b of every point of a, the arithmetic of the (un-vendored) ChamferDistancePytorch
    chamfer3D NmDistanceKernel as nvcc contracts it: d = fma(dz,dz, fma(dy,dy, dx*dx)) in fp32, first minimal index.
    fp32 FMAs are emulated in float64 (exact products; exact for the integer voxel coordinates util/metrics.py:43-44
    feeds it).  Parity unpinned for non-integer clouds (the submodule is absent): tests use a tolerance there."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    dist = np.empty(a.shape[0], dtype=np.float32)
    idx = np.empty(a.shape[0], dtype=np.int32)
    step = 128  # small blocks: the float64 temporaries stay cache-sized
    for lo in range(0, a.shape[0], step):
        aa = a[lo:lo + step, None, :]
        df = (b[None, :, :] - aa).astype(np.float32)                      # fp32 differences
        dd = df.astype(np.float64)
        t = (dd[..., 0] * dd[..., 0]).astype(np.float32)                  # __fmul_rn
        t = (dd[..., 1] * dd[..., 1] + t.astype(np.float64)).astype(np.float32)
        t = (dd[..., 2] * dd[..., 2] + t.astype(np.float64)).astype(np.float32)
        j = np.argmin(t, axis=1)                                          # first minimal index
        idx[lo:lo + step] = j
        dist[lo:lo + step] = t[np.arange(t.shape[0]), j]
    return dist, idx


def chamfer_metric(pred: np.ndarray, target: np.ndarray):
    """util/metrics.py:37-51 Chamfer3D.update on bool [B,1,S,S,S]: (cd_sum, valid count)."""
    cd, valid = 0.0, 0
    for ip in range(pred.shape[0]):
        pp = np.argwhere(pred[ip, 0]).astype(np.float32)
        pt = np.argwhere(target[ip, 0]).astype(np.float32)
        if pp.shape[0] == 0 or pt.shape[0] == 0:      # mean of an empty tensor is NaN and skipped (:48)
            continue
        d1, _ = chamfer_nn(pt, pp)
        d2, _ = chamfer_nn(pp, pt)
        cd += float(torch.from_numpy(d1).mean() + torch.from_numpy(d2).mean())
        valid += 1
    return cd, valid
