"""Golden vectors for SURVEY 8 rows a1, a10, a11, a12: the REFERENCE's own retrieval pre-pass, executed.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_retrieval.py      ->  tests/golden/retrieval.npz, retrieval_index.json

`util/retrieval.py`, `dataset/scene.py` and `dataset/patched_scene_dataset.py` are imported UNMODIFIED from
/root/reference.  What they import but this image lacks is stubbed in sys.modules:

  * pyflann          - the kNN itself (un-vendored, unpinned, approximate).  The stub's `FLANN.nn_index` is the thing
                       FLANN approximates: exact squared-L2 neighbours in ascending (fp64 distance, row id) order,
                       distances returned as fp32 - the canonical rule of include/rf_b200.h, written out here in
                       numpy independently of oracle/.
  * trimesh, pyrender, marching_cubes, torchmetrics, the compiled Chamfer extension - visualisation / metrics
    modules that the executed functions never call.
  * Tensor.cuda is the identity (CPU-only box).

Executed reference code: SceneHandler.__init__ (sizes + occupancy caches), get_extents_for_size / get_scene_patches,
PatchedSceneDataset.__init__ / __getitem__ / unpad, create_dictionary + get_zero_patch_entry (util/retrieval.py:21-55),
extract_input_features (:58-72), flann_knn_worker (:79-105; with and without ignore_patches_from_source),
create_retrieval_from_mapping (:145-164; tiling and overlapping strides), Patch04 / Patch32 (model/retrieval.py).
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import retrieval_cases as RC  # noqa: E402


# ----------------------------------------------------------------------------- stubs
class ExactFLANN:
    """Stands in for pyflann.FLANN (call sites util/retrieval.py:49-55,81-83,92)."""

    def __init__(self, **kw):
        self.pts = None

    def build_index(self, pts, **kw):
        self.pts = np.ascontiguousarray(pts, dtype=np.float32)
        return {"algorithm": "exact_stub", "checks": 32, "trees": kw.get("trees", 0)}

    def save_index(self, filename):
        Path(filename.decode("utf-8") if isinstance(filename, bytes) else filename).write_bytes(b"exact_stub")

    def load_index(self, filename, pts):
        self.pts = np.ascontiguousarray(pts, dtype=np.float32)

    def nn_index(self, qpts, num_neighbors=1, **kw):
        q = np.ascontiguousarray(qpts, dtype=np.float32).astype(np.float64)
        x = self.pts.astype(np.float64)
        acc = np.zeros((q.shape[0], x.shape[0]), dtype=np.float64)
        for i in range(x.shape[1]):  # sequential sum over the dimension, every operation rounded once
            diff = q[:, i:i + 1] - x[None, :, i]
            acc = acc + diff * diff
        order = np.argsort(acc, axis=1, kind="stable")[:, :num_neighbors]  # ascending (d, row id)
        return order.astype(np.int32), np.take_along_axis(acc, order, axis=1).astype(np.float32)


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("pyflann", FLANN=ExactFLANN, __all__=["FLANN"])
    tm = mod("trimesh", __path__=[])
    tm.sample = mod("trimesh.sample")
    tm.voxel = mod("trimesh.voxel", __path__=[])
    tm.voxel.ops = mod("trimesh.voxel.ops")
    mod("pyrender")
    mod("marching_cubes")

    class Metric(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default)

    mod("torchmetrics", __path__=[]).metric = mod("torchmetrics.metric", Metric=Metric)
    mod("external.ChamferDistancePytorch.chamfer3D.dist_chamfer_3D", chamfer_3DDist=lambda: None)
    torch.Tensor.cuda = lambda self, *a, **k: self


def main():
    install_stubs()
    from oracle import rf_oracle as O
    from dataset.patched_scene_dataset import PatchedSceneDataset  # reference
    from dataset.scene import SceneHandler  # reference
    from model import get_retrieval_networks  # reference
    from util import retrieval as R  # reference

    torch.set_grad_enabled(False)
    sd_in, sd_tg = RC.encoder_state_dicts()
    out, index = {}, {}
    for case in RC.CASES:
        with tempfile.TemporaryDirectory() as tmp:
            RC.write_dataset(tmp)
            cfg = RC.make_config(tmp, case)
            fenc_input, fenc_target = get_retrieval_networks(cfg["retrieval_model"])
            fenc_input.load_state_dict(sd_in)
            fenc_target.load_state_dict(sd_tg)
            fenc_input.eval(), fenc_target.eval()
            sh_train, sh_val = SceneHandler("train", cfg), SceneHandler("val", cfg)
            ds_train = PatchedSceneDataset("train", cfg["dataset_train"], sh_train)
            ds_val = PatchedSceneDataset("val", cfg["dataset_val"], sh_val)
            tree = Path(tmp) / "tree"
            R.create_dictionary(fenc_target, cfg["dictionary"], RC.LATENT, ds_train, tree)
            database = np.load(tree / "database.npy")
            out[f"{case}.database"] = database
            index[f"{case}.index"] = json.loads((tree / "index.json").read_text())
            handler = R.RetrievalInterface(cfg["query"], RC.LATENT)
            for split, ds, ignore in (("train", ds_train, True), ("val", ds_val, False), ("val_ignore", ds_val, True),
                                      ("train_keep", ds_train, False)):
                names, feats = R.extract_input_features(fenc_input, cfg["query"], RC.LATENT, ds)
                mapping = R.query_dictionary_using_features(cfg["query"], names, feats, ds, tree, ignore)
                assert list(mapping.keys()) == names and all(v is not None for v in mapping.values())
                index[f"{case}.{split}.patch_names"] = names
                out[f"{case}.{split}.features"] = feats
                out[f"{case}.{split}.mapping"] = np.stack([mapping[n] for n in names]).astype(np.float32)
                assert out[f"{case}.{split}.mapping"].shape == (len(names), RC.K, 8)
                # the dataloader's patches (a1): item['input'] / item['target'] of the first scene
                if split in ("train", "val"):
                    items = [ds[i] for i in range(len(ds))]
                    out[f"{case}.{split}.patch_input"] = np.stack([it["input"] for it in items]).astype(np.float32)
                    index[f"{case}.{split}.extent"] = [[int(v) for v in it["extent"]] for it in items]
                    out[f"{case}.{split}.patch_target_sum"] = np.array([np.float64(it["target"].astype(np.float64).sum()) for it in items])
                    index[f"{case}.{split}.occupancy"] = {n: int(v) for n, v in ds.scene_handler.scene_occupancy.items()}
                    index[f"{case}.{split}.scene_size"] = {s: list(ds.get_scene_size(s)) for s in ds.scenes}
                for scene in ds.scenes:
                    vol = R.create_retrieval_from_mapping(scene, mapping, RC.K, ds_train, ds, tree)
                    out[f"{case}.{split}.compose.{scene}"] = vol.numpy().astype(np.float32)
            changed = int((out[f"{case}.train.mapping"] != out[f"{case}.train_keep.mapping"]).any(axis=(1, 2)).sum())
            changed_val = int((out[f"{case}.val.mapping"] != out[f"{case}.val_ignore.mapping"]).any(axis=(1, 2)).sum())
            print(case, "rows", database.shape, "train patches", len(ds_train), "val patches", len(ds_val),
                  "queries changed by the demotion: train", changed, "val", changed_val)
    # sanity: the oracle's canonical kNN agrees with the stub on these queries (both claim the same rule)
    db = out["tile.database"]
    i2, d2 = O.knn_exact(db[:, 7:], out["tile.train.features"], 2 * RC.K)
    f = ExactFLANN(); f.build_index(db[:, 7:])
    si, sd_ = f.nn_index(out["tile.train.features"], 2 * RC.K)
    assert np.array_equal(i2, si) and np.array_equal(d2, sd_), "oracle knn_exact != stub"
    np.savez_compressed(os.path.join(HERE, "retrieval.npz"), **out)
    with open(os.path.join(HERE, "retrieval_index.json"), "w") as fjs:
        json.dump(index, fjs, indent=0, sort_keys=True)
    for fn in ("retrieval.npz", "retrieval_index.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
