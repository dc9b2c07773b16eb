"""Pins the oracle's retrieval half to outputs of the REFERENCE's own code:
tests/golden/make_golden_retrieval.py ran dataset/scene.py, dataset/patched_scene_dataset.py and util/retrieval.py
(create_dictionary, get_zero_patch_entry, flann_knn_worker with / without ignore_patches_from_source,
create_retrieval_from_mapping with tiling and overlapping strides) from /root/reference on the tiny seeded dataset of
tests/golden/retrieval_cases.py, with pyflann replaced by an exact brute force, and committed what they returned."""
import numpy as np
import pytest
import torch

import retrieval_cases as RC
from oracle import rf_oracle as O

torch.set_grad_enabled(False)
Z, IDX = RC.load_golden()
ARR = RC.scene_arrays()


def _dcfg(case, split):
    return RC.dataset_config("/nonexistent", split, case)


def _tables(case, split):
    d = _dcfg(case, split)
    out = {}
    for s, (inp, tg) in ARR[split].items():
        out[s] = O.scene_patch_table(inp, tg, d, d["occupancy_threshold"])
    return out


@pytest.mark.parametrize("case", list(RC.CASES))
@pytest.mark.parametrize("split", ["train", "val"])
def test_patch_enumeration_matches_reference_dataset(case, split):
    """a1: which patches exist, in which order, with which values (PatchedSceneDataset over SceneHandler)."""
    tabs = _tables(case, split)
    ext = np.concatenate([tabs[s][1] for s in ARR[split]])
    assert ext.tolist() == IDX[f"{case}.{split}.extent"]
    names = [f"{s}--" + "_".join(f"{v:04d}" for v in e) for s in ARR[split] for e in tabs[s][1]]
    assert names == IDX[f"{case}.{split}.patch_names"]
    p_in = np.concatenate([tabs[s][2] for s in ARR[split]])
    assert np.array_equal(p_in, Z[f"{case}.{split}.patch_input"])
    tsum = np.array([p.astype(np.float64).sum() for s in ARR[split] for p in tabs[s][3]])
    assert np.array_equal(tsum, Z[f"{case}.{split}.patch_target_sum"])
    # occupancy of every candidate patch (kept or not), as the reference's cache file holds it
    d = _dcfg(case, split)
    occ_ref = IDX[f"{case}.{split}.occupancy"]
    for s, (inp, tg) in ARR[split].items():
        et = O.get_extents_for_size(list(tg.shape), d["patch_size_target"], d["patch_context_target"], d["patch_stride"])
        for e, o in zip(et, tabs[s][4]):
            assert occ_ref[f"{s}--" + "_".join(f"{v:04d}" for v in e)] == o
        assert IDX[f"{case}.{split}.scene_size"][s] == list(tg.shape)


@pytest.mark.parametrize("case", list(RC.CASES))
def test_database_rows_match_create_dictionary(case):
    """a10: [scene_idx, unpadded extents, embedding] rows + the all-ones sentinel (util/retrieval.py:21-55)."""
    db = Z[f"{case}.database"]
    _, sd_tg = RC.encoder_state_dicts()
    d = _dcfg(case, "train")
    tabs = _tables(case, "train")
    scenes = list(ARR["train"])
    assert IDX[f"{case}.index"] == scenes
    enc = lambda x: O.encoder_forward("Patch32", sd_tg, x)
    rows = []
    for si, s in enumerate(scenes):
        _, et, _, p_tg, _ = tabs[s]
        emb = O.normalize_features(enc(torch.from_numpy(p_tg)), RC.LATENT).numpy()
        rows.append(O.database_rows(np.full(len(et), si), et, d["patch_context_target"], emb))
    rows.append(O.zero_patch_row(enc, d["patch_size_target"], d["patch_context_target"], RC.LATENT))
    mine = np.concatenate(rows)
    assert mine.shape == db.shape and mine.dtype == db.dtype
    assert np.array_equal(mine[:, :7], db[:, :7])              # ids and extents: bit-exact
    assert np.abs(mine[:, 7:] - db[:, 7:]).max() <= 2e-6       # embeddings: same torch CPU ops, other batch split


@pytest.mark.parametrize("case", list(RC.CASES))
@pytest.mark.parametrize("split,ignore", [("train", True), ("train_keep", False), ("val", False), ("val_ignore", True)])
def test_knn_demotion_rows_match_flann_knn_worker(case, split, ignore):
    """a11: fetch 2K, stable-partition hits of the query's own scene to the back, keep K; rows [K, 8]
    (util/retrieval.py:79-105), on the reference's own database and query features."""
    db = Z[f"{case}.database"]
    feats = Z[f"{case}.{split}.features"]
    names = IDX[f"{case}.{split}.patch_names"]
    index = IDX[f"{case}.index"]
    qs = np.array([(index.index(n.split("--")[0]) if (ignore and n.split("--")[0] in index) else -1) for n in names])
    rows, idx = O.lookup_rows(db[:, 7:], db[:, :7], feats, RC.K, qs)
    gold = Z[f"{case}.{split}.mapping"]
    assert rows.dtype == gold.dtype and np.array_equal(rows, gold)
    if ignore and split == "train":  # the demotion did something here, and tr_dup's duplicate rows produced ties
        keep = Z[f"{case}.train_keep.mapping"]
        assert (keep != gold).any()
        d2 = O.knn_exact(db[:, 7:], feats, 2 * RC.K)[1]
        assert (d2[:, 1:] == d2[:, :-1]).any()


@pytest.mark.parametrize("case", list(RC.CASES))
@pytest.mark.parametrize("split", ["train", "val", "val_ignore"])
def test_compose_matches_create_retrieval_from_mapping(case, split):
    """a12: paste the retrieved 16^3 cores (util/retrieval.py:145-164); 'tile' = unconditional paste with filtered
    patches left at the truncation value, 'overlap' = keep the candidate with the lower mean distance."""
    ds_split = "val" if split.startswith("val") else "train"
    d = _dcfg(case, ds_split)
    mapping = Z[f"{case}.{split}.mapping"]
    names = IDX[f"{case}.{split}.patch_names"]
    store = [ARR["train"][s][1].astype(np.float32) for s in IDX[f"{case}.index"]]
    trunc = np.float32(O.f16_trunc(d["voxel_size_target"]))
    for s, (_, tg) in ARR[ds_split].items():
        sel = [i for i, n in enumerate(names) if n.split("--")[0] == s]
        ext = np.array([[int(v) for v in names[i].split("--")[1].split("_")] for i in sel])
        ext[:, 1::2] -= 2 * d["patch_context_target"]
        out = O.compose_from_mapping(mapping[sel], ext, store, tg.shape, trunc, trunc, no_overlap=d["patch_stride"] == d["patch_size_target"])
        gold = Z[f"{case}.{split}.compose.{s}"]
        assert out.dtype == gold.dtype and np.array_equal(out, gold), (case, split, s)
