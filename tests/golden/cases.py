"""Seeded inputs shared by make_golden.py (reference run) and the tests."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import rf_oracle as O  # noqa: E402

SEED = 1234

ENC_CASES = [  # (class, nf, n_patches)
    ("Patch04", 32, 5), ("Patch04V2", 32, 3), ("Patch05", 8, 3), ("Patch08", 16, 5), ("PatchNorm08", 16, 3),
    ("Patch12", 8, 3), ("Patch16", 8, 3), ("Patch24", 12, 3), ("Patch24V2", 12, 2), ("Patch32", 8, 4),
    ("PatchNorm32", 8, 2), ("PCPatch32", 8, 2), ("PCPatch48", 10, 2), ("PCPatch64", 8, 1),
]
ATTN_CASES = [(16, 4, False), (12, 8, False), (16, 4, True), (16, 1, False)]  # (nf, K, gumbel)
ATTN_MAPPING_CASES = [(16, 4, False), (12, 8, False)]  # attn_no_output_mapping=False (g / o 1x1x1 convolutions; softmax mode only: the
# reference itself fails in retrieval mode, model/attention.py:103 hands o() a 2-D tensor)
UNET_CASES = [("sr08", "Superresolution08UNetBackbone", 16, 4, 8), ("sr16", "Superresolution16UNetBackbone", 16, 4, 16),
              ("surface", "SurfaceReconstructionUNetBackbone", 12, 5, 128)]

# config/super_resolution/3DFront/refinement_008_064.yaml
SR_3DFRONT = dict(voxel_size_input=0.43334, voxel_size_target=0.054167, input_mean=0.8112343966484424,
                  input_std=0.5094238937427482, target_mean=0.15015658121788053, target_std=0.03573221820637578)
REFINE_CFG = dict(task="superresolution", nf=16, unet_num_level=4, layer_order="gcr", retrieval_fmaps=16,
                  retrieval_num_level=4, attn_patch_extent=4, K=4, attn_normalize=True, attn_use_switching=True,
                  attn_retrieval_mode=False, attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16,
                  dataset_train=dict(input_chunk_size=8))


def rnd(name, shape, scale=1.0):
    return (O.synth_tensor(name, shape, SEED, "const") * scale).contiguous()


def encoder_input(cls, n):
    P = (O.MLP_SPECS.get(cls) or O.ENCODER_SPECS[cls])["patch"]
    x = rnd(f"enc.{cls}.x", (n, 1, P, P, P))
    x[-1] = 1.0  # the all-ones sentinel patch (util/retrieval.py:23)
    return x


def retrieval_backbone_input(nf):
    x = rnd(f"rb.{nf}.x", (5, 1, 16, 16, 16))
    x[1] = 0.3125  # constant patch: GroupNorm var = 0 path
    return x


def unet_backbone_input(kind, S):
    if kind == "surface":
        pts = (O.synth_tensor("surface.pts", (1000, 3), SEED, "const").numpy() * 0.5 + 0.5) * 64
        return torch.from_numpy(O.point_cloud_to_grid(pts, 128, 2.0, 0))[None, None]
    return rnd(f"ub.{kind}.x", (2, 1, S, S, S))


def attention_tag(nf, K, mode):
    return f"attention.{nf}.{K}.{'gumbel' if mode else 'softmax'}"


def attention_inputs(nf, K, mode):
    tag = attention_tag(nf, K, mode)
    xb = rnd(f"{tag}.xb", (1, nf, 32, 32, 32))
    xr = rnd(f"{tag}.xr", (K, nf, 32, 32, 32))
    xr[0, :, :8] = xb[0, :, :8]  # candidate 0 == backbone feature in the first slab
    occ = rnd(f"{tag}.occ", (1, 1, 32, 32, 32)) > 0.6
    return xb, xr, occ


def refine_full_inputs():
    """BASELINE config 1: one SR 8^3 -> 64^3 chunk, K = 4."""
    c = SR_3DFRONT
    vs_in, vs_tg = c["voxel_size_input"], c["voxel_size_target"]
    tgt = O.synthetic_tsdf(0, 64, vs_tg)
    inp = O.downsample_tsdf(tgt / vs_tg * vs_in, 8, vs_in)
    retr = np.stack([O.synthetic_tsdf(100 + k, 64, vs_tg) for k in range(4)])
    retr[3, :16, :16, :16] = np.float16(vs_tg * 3)  # an all-trunc 16^3 block (sentinel paste)
    retr[0, 16:48] = tgt[16:48]  # a good retrieval in the middle slab
    x_in = torch.from_numpy((inp - c["input_mean"]) / c["input_std"]).float()[None, None]
    x_re = torch.from_numpy((retr - c["target_mean"]) / c["target_std"]).float()[None]
    return x_in, x_re
