"""The tiny on-disk dataset shared by make_golden_retrieval.py (which feeds it to the
REFERENCE's own dataset/scene.py, dataset/patched_scene_dataset.py and util/retrieval.py) and by the tests (which
feed the same files to the oracle and to the CUDA path).  Everything is derived from seeds; nothing here reads
/root/reference."""
import json
import os
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SEED = 4321
K = 4
LATENT = 64
# voxel sizes / normalisation constants of config/super_resolution/3DFront/retrieval_008_064.yaml
VOXEL_IN, VOXEL_TG = 0.43334, 0.054167
NORM = dict(input_mean=0.8112343966484424, input_std=0.5094238937427482, target_mean=0.15015658121788053,
            target_std=0.03573221820637578)

# (name, target size); inputs are 1/8 of the target (patch_size_target 16 / patch_size_input 2)
TRAIN_SCENES = [("tr_a", (32, 32, 32)), ("tr_b", (32, 32, 32)), ("tr_c", (48, 32, 32)), ("tr_d", (32, 32, 32)),
                ("tr_e", (32, 32, 32)), ("tr_dup", (32, 32, 32))]  # tr_dup == tr_a voxel for voxel: exact distance ties
VAL_SCENES = [("va_a", (32, 32, 32)), ("va_b", (32, 48, 32)), ("tr_b", (32, 32, 32))]  # a val scene NAMED like a train scene

# case -> (patch_stride, occupancy_threshold)
CASES = {"tile": (16, 0), "overlap": (8, -1)}


def _tsdf(rng, size, voxel_size, empty_corner):
    """Distance field of a few spheres / planes, truncated at float16(3 * voxel) like dataset/scene.py:32-33; with
    empty_corner the x < 16, y < 16, z < 16 block holds no surface (occupancy 0 -> the dataset drops that patch)."""
    trunc = np.float16(voxel_size * 3).astype(np.float32)
    ax = [np.arange(s, dtype=np.float32) + 0.5 for s in size]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    d = np.full(size, 1e9, dtype=np.float32)
    for _ in range(4):
        c = rng.random(3) * np.array(size)
        if empty_corner:
            c = np.maximum(c, 28.5)
        r = (0.08 + 0.2 * rng.random()) * min(size)
        if empty_corner:
            r = min(r, 2.5)
        d = np.minimum(d, np.abs(np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r))
    return np.minimum(d * voxel_size, trunc).astype(np.float16)


def _downsample(target, voxel_size_target, voxel_size_input):
    f = 8
    s = [v // f for v in target.shape]
    t = target.astype(np.float32).reshape(s[0], f, s[1], f, s[2], f).min(axis=(1, 3, 5))
    trunc = np.float16(voxel_size_input * 3).astype(np.float32)
    return np.minimum(t / voxel_size_target * voxel_size_input, trunc).astype(np.float16)


def scene_arrays():
    """name -> (input fp16 [s/8...], target fp16 [s...]) per split."""
    rng = np.random.default_rng(SEED)
    out = {"train": {}, "val": {}}
    for split, scenes in (("train", TRAIN_SCENES), ("val", VAL_SCENES)):
        for name, size in scenes:
            if name == "tr_dup":
                out[split][name] = out[split]["tr_a"]
                continue
            tg = _tsdf(rng, size, VOXEL_TG, empty_corner=name in ("tr_d", "va_a"))
            out[split][name] = (_downsample(tg, VOXEL_TG, VOXEL_IN), tg)
    return out


def dataset_config(root, split, case):
    stride, occ = CASES[case]
    return dict(dataset_name=f"Tiny_{split}", data_dir=str(root), scene_dir=str(root), splits_dir="official",
                input_dir="sdf_008", target_dir="sdf_064", input_ext=".npz", target_ext=".npz",
                input_chunk_size=8, target_chunk_size=64, num_points=0, voxel_size_input=VOXEL_IN,
                voxel_size_target=VOXEL_TG, patch_size_input=2, patch_context_input=1, patch_size_target=16,
                patch_context_target=8, patch_stride=stride, occupancy_threshold=occ, skip_occupancy=False,
                preload_scenes=False, preload_retrievals=False, train_multiplier=1, retrieval_dir=str(root), **NORM)


def make_config(root, case):
    """The keys SceneHandler / PatchedSceneDataset / util.retrieval read (config/base/retrieval_superresolution.yaml)."""
    return dict(task="superresolution", fast_visualization=True, no_retrievals=True, K=K,
                retrieval_ckpt="runs/tiny/ckpt.ckpt", dataset_train=dataset_config(root, "train", case),
                dataset_val=dataset_config(root, "val", case),
                retrieval_model=dict(network_input="2+1", network_target="16+8", nf_input=32, nf_target=8, latent_dim=LATENT),
                dictionary=dict(batch_size=7, num_workers=0), query=dict(batch_size=5, num_workers=0, K=K, flann_num_workers=0))


def write_dataset(root):
    """Lays the scenes out the way the reference reads them (dataset/scene.py:44-56,60-61,94)."""
    root = Path(root)
    arrays = scene_arrays()
    for split in ("train", "val"):
        name = f"Tiny_{split}"
        (root / "splits" / name / "official").mkdir(parents=True, exist_ok=True)
        (root / "splits" / name / "official" / f"{split}.txt").write_text("\n".join(arrays[split].keys()) + "\n")
        for d in ("sdf_008", "sdf_064"):
            (root / d / name).mkdir(parents=True, exist_ok=True)
        for s, (inp, tg) in arrays[split].items():
            np.savez_compressed(root / "sdf_008" / name / f"{s}.npz", arr=inp)
            np.savez_compressed(root / "sdf_064" / name / f"{s}.npz", arr=tg)
    return arrays


def encoder_state_dicts():
    from oracle import rf_oracle as O
    return (O.synth_state_dict(O.encoder_param_shapes("Patch04", 32, LATENT), SEED),
            O.synth_state_dict(O.encoder_param_shapes("Patch32", 8, LATENT), SEED))


def load_golden():
    here = os.path.dirname(os.path.abspath(__file__))
    z = np.load(os.path.join(here, "retrieval.npz"))
    names = json.loads(open(os.path.join(here, "retrieval_index.json")).read())
    return z, names
