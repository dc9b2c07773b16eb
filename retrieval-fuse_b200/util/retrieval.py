"""Drop-in for the reference's util/retrieval.py: dictionary build, feature
extraction, kNN lookup, compose - same function names, arguments and on-disk
formats, with the arithmetic on the GPU:

  create_dictionary            util/retrieval.py:29-55   (database.npy, index.json)
  extract_features             util/retrieval.py:58-68
  query_dictionary_using_features  :134-142  + flann_knn_worker :79-105
  create_retrieval_from_mapping    :145-164
  RetrievalInterface               :178-207

FLANN's approximate kd-forest is replaced by the exact brute-force kNN of
rf_knn_l2_topk (canonical rule in include/rf_b200.h); the source-scene demotion
and the [K, 8] mapping rows follow the reference exactly.  `EmbeddingBank` is
the GPU-resident form of database.npy and the unit that shards across ranks
(sharded.py).
"""
import json
from pathlib import Path

import numpy as np
import torch
import torch.utils.data

from .. import ops

_QUERY_BATCH = 1 << 16  # queries per kNN launch (bounds the [Q, 2K] temporaries)


def get_retrievals_dir(config):
    """util/misc.py:62-70."""
    ckpt_experiment = Path(config["retrieval_ckpt"]).parents[0].name
    ckpt_epoch = Path(config["retrieval_ckpt"]).name.split(".")[0]
    task_dir = f"{config['task']}_{config['dataset_train']['num_points']:04d}"
    return Path(config["dataset_train"]["retrieval_dir"], "retrieval", task_dir, config["dataset_train"]["dataset_name"],
                config["dataset_train"]["splits_dir"], ckpt_experiment, ckpt_epoch, str(config["K"]))


def _encode_normalized(feature_extractor, x, latent_dim):
    """feature_extractor(x) -> permute/reshape -> F.normalize(dim=1) (util/retrieval.py:38,66).
    The encoder output is [N, D, 1, 1, 1], so the permute is a no-op reshape."""
    if hasattr(feature_extractor, "encode"):  # MLP encoders fuse the normalisation
        return feature_extractor.encode(x, l2_normalize=True).reshape(-1, latent_dim)
    return ops.l2_normalize_rows(feature_extractor(x).reshape(-1, latent_dim))


class EmbeddingBank:
    """GPU-resident database: emb [N,64] fp32 unit rows + meta [N,7] fp32
    ([scene_idx, x0,x1,y0,y1,z0,z1], unpadded extents), plus the scene list
    (`index.json`).  Row N-1 is the all-ones sentinel (scene -1)."""

    def __init__(self, emb, meta, scenes, row_offset=0, n_total=None):
        assert emb.is_cuda and meta.is_cuda, "EmbeddingBank lives on the GPU"
        self.emb = emb.contiguous()
        self.meta = meta.contiguous()   # ALWAYS the full table (global row ids index it)
        self.scenes = list(scenes)
        self.row_offset = int(row_offset)  # first global row id of this shard's emb
        self.n_total = int(n_total if n_total is not None else emb.shape[0])

    @property
    def device(self):
        return self.emb.device

    def __len__(self):
        return self.n_total

    @classmethod
    def from_database(cls, database, scenes, device):
        """database: [N, 7 + D] fp32 array in the reference's database.npy layout."""
        db = torch.as_tensor(np.ascontiguousarray(database), dtype=torch.float32)
        db = db.to(device, non_blocking=True)
        return cls(db[:, 7:].contiguous(), db[:, :7].contiguous(), scenes)

    @classmethod
    def load(cls, tree_path, device):
        tree_path = Path(tree_path)
        database = np.load(tree_path / "database.npy")
        scenes = json.loads((tree_path / "index.json").read_text())
        return cls.from_database(database, scenes, device)

    def to_database(self):
        return torch.cat([self.meta[self.row_offset: self.row_offset + self.emb.shape[0]], self.emb], dim=1).cpu().numpy()

    def save(self, tree_path):
        tree_path = Path(tree_path)
        tree_path.mkdir(exist_ok=True, parents=True)
        np.save(tree_path / "database", self.to_database())
        (tree_path / "index.json").write_text(json.dumps(self.scenes))
        # the reference stores FLANN's build parameters here; consumers only read 'checks'
        (tree_path / "params.json").write_text(json.dumps({"algorithm": "rf_b200_exact", "checks": -1}))

    def shard(self, rank, world_size):
        """Rows [rank*N/W, (rank+1)*N/W) of the embeddings; meta stays whole."""
        n = self.emb.shape[0]
        per = (n + world_size - 1) // world_size
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        return EmbeddingBank(self.emb[lo:hi].contiguous(), self.meta, self.scenes, row_offset=lo, n_total=n)

    def scene_ids(self, scene_names, ignore_patches_from_source):
        """util/retrieval.py:94-95: the index of the query's scene in index.json, -1 when absent / disabled."""
        if not ignore_patches_from_source:
            return None
        lookup = {s: i for i, s in reversed(list(enumerate(self.scenes)))}  # list.index == first occurrence
        return torch.tensor([lookup.get(s, -1) for s in scene_names], dtype=torch.int32)

    _BULK_QUERIES = 32768  # from here on a lookup re-stages the bank in the scan order of ITS queries (~0.15 ms)

    def topk(self, q, k, method=0):
        """Exact top-k of this shard: (global ids int32 [Q,k], fp64 d [Q,k]).  The tensor-core operand image of the
        (static) bank is built once, in bank order, and reused by every lookup - the reference likewise loads its
        FLANN index once per worker (util/retrieval.py:81-83).  Bulk lookups re-stage the bank per call in the scan
        order of their own queries, which pays off from a few ten thousand queries on."""
        if method in (0, 2, 3) and q.shape[0] < self._BULK_QUERIES:
            key = (method, self.emb.data_ptr(), self.emb.shape[0])
            if getattr(self, "_image_key", None) != key:
                self._image = ops.knn_prepare_bank(self.emb, method)
                self._image_key = key
            if self._image is not None:
                return ops.knn_topk(self.emb, q, k, row_offset=self.row_offset, image=self._image)
            return ops.knn_topk(self.emb, q, k, row_offset=self.row_offset, method=1 if method == 0 else method)
        return ops.knn_topk(self.emb, q, k, row_offset=self.row_offset, method=method)

    def query(self, q, K, query_scene=None, method=0):
        """q [Q,64] unit rows on the GPU -> (rows fp32 [Q,K,8], ids int32 [Q,K]):
        fetch 2K, demote same-scene hits, keep K (util/retrieval.py:92-100)."""
        assert self.row_offset == 0 and self.emb.shape[0] == self.n_total, "query() needs the whole bank; see sharded.py"
        k2 = min(2 * K, self.n_total)
        idx2k, d2k = self.topk(q, k2, method)
        qs = None if query_scene is None else query_scene.to(q.device, non_blocking=True)
        return ops.knn_demote_rows(idx2k, d2k, self.meta, qs, K)


def get_zero_patch_entry(feature_extractor, patch_size, patch_context, latent_dim):
    """util/retrieval.py:21-26: row of the all-ONES patch, scene -1, extent [0, patch_size]^3."""
    dev = next(feature_extractor.parameters()).device
    with torch.no_grad():
        ones = torch.ones([1, 1] + [patch_size + 2 * patch_context] * 3, dtype=torch.float32, device=dev)
        emb = _encode_normalized(feature_extractor, ones, latent_dim)
    meta = torch.tensor([[-1, 0, patch_size, 0, patch_size, 0, patch_size]], dtype=torch.float32, device=dev)
    return torch.cat([meta, emb], dim=1)


def _loader(dataset, cfg):
    return torch.utils.data.DataLoader(dataset, batch_size=cfg["batch_size"], shuffle=False,
                                       num_workers=cfg.get("num_workers", 0), drop_last=False, pin_memory=True)


def create_dictionary(feature_extractor, dictionary_config, latent_dim, dataset, tree_path):
    """util/retrieval.py:29-55.  Returns the EmbeddingBank (the reference returns None)
    and writes database.npy / index.json / params.json under tree_path."""
    dev = next(feature_extractor.parameters()).device
    n = len(dataset)
    database = torch.zeros((n + 1, 7 + latent_dim), dtype=torch.float32, device=dev)
    bs = dictionary_config["batch_size"]
    with torch.no_grad():
        for i, item in enumerate(_loader(dataset, dictionary_config)):
            target = item["target"].to(dev, dtype=torch.float32, non_blocking=True)
            emb = _encode_normalized(feature_extractor, target, latent_dim)
            lo, hi = i * bs, i * bs + target.shape[0]
            ext = torch.as_tensor(np.asarray(item["extent"]), dtype=torch.float32).reshape(target.shape[0], 6).clone()
            ext[:, 1::2] -= 2 * dataset.target_patch_context  # dataset.unpad (:40-43)
            scene_index = torch.as_tensor(dataset.get_scene_indices(item["scene"]), dtype=torch.float32)[:, None]
            database[lo:hi, :7] = torch.cat([scene_index, ext], dim=1).to(dev, non_blocking=True)
            database[lo:hi, 7:] = emb
    database[n] = get_zero_patch_entry(feature_extractor, dataset.target_patch_size, dataset.target_patch_context, latent_dim)[0]
    bank = EmbeddingBank(database[:, 7:].contiguous(), database[:, :7].contiguous(), dataset.scenes)
    if tree_path is not None:
        bank.save(tree_path)
    return bank


def extract_features(feature_extractor, query_config, latent_dim, dataset, key, return_device_tensor=False):
    """util/retrieval.py:58-68 -> (patch_names, features [len(dataset), D])."""
    dev = next(feature_extractor.parameters()).device
    feats = torch.zeros((len(dataset), latent_dim), dtype=torch.float32, device=dev)
    names = []
    bs = query_config["batch_size"]
    with torch.no_grad():
        for i, item in enumerate(_loader(dataset, query_config)):
            names.extend(item["name"])
            x = item[key].to(dev, dtype=torch.float32, non_blocking=True)
            feats[i * bs: i * bs + x.shape[0]] = _encode_normalized(feature_extractor, x, latent_dim)
    return names, (feats if return_device_tensor else feats.cpu().numpy())


def extract_input_features(feature_extractor, query_config, latent_dim, dataset):
    return extract_features(feature_extractor, query_config, latent_dim, dataset, "input")


def extract_target_features(feature_extractor, query_config, latent_dim, dataset):
    return extract_features(feature_extractor, query_config, latent_dim, dataset, "target")


_BANK_CACHE = {}


def _bank_for(tree_path, device):
    tree_path = Path(tree_path)
    f = tree_path / "database.npy"
    key = (str(tree_path), str(device), f.stat().st_mtime_ns)
    if key not in _BANK_CACHE:
        _BANK_CACHE.clear()
        _BANK_CACHE[key] = EmbeddingBank.load(tree_path, device)
    return _BANK_CACHE[key]


def query_bank(bank, features, scene_names, K, ignore_patches_from_source, method=0):
    """kNN + demotion for all queries -> fp32 [Q, K, 8] on the GPU."""
    q = torch.as_tensor(features, dtype=torch.float32).to(bank.device, non_blocking=True)
    qs = bank.scene_ids(scene_names, ignore_patches_from_source)
    out = torch.empty((q.shape[0], K, 8), dtype=torch.float32, device=bank.device)
    for lo in range(0, q.shape[0], _QUERY_BATCH):
        hi = min(lo + _QUERY_BATCH, q.shape[0])
        rows, _ = bank.query(q[lo:hi], K, None if qs is None else qs[lo:hi], method)
        out[lo:hi] = rows
    return out


def query_dictionary_using_features(query_config, patch_names, input_features, dataset, tree_path,
                                    ignore_patches_from_source, bank=None):
    """util/retrieval.py:134-142 (+ the worker :79-105): dict patch_name -> fp32 [K, 8].
    `flann_num_workers` is ignored: one GPU sweep replaces the worker processes."""
    dev = torch.device("cuda", torch.cuda.current_device())
    bank = bank if bank is not None else _bank_for(tree_path, dev)
    scene_names = dataset.get_scene_names_from_patches(patch_names)
    rows = query_bank(bank, input_features, scene_names, query_config["K"], ignore_patches_from_source).cpu().numpy()
    return {name: rows[i] for i, name in enumerate(patch_names)}


def _overlap_cells(rows, dst, size):
    """Overlapping target patches (patch_stride < patch_size_target): util/retrieval.py:149-164 pastes the patches of a
    scene in order and a paste happens only while `distances[k, region].mean() > current_distance`, where `distances`
    records the distance of whatever was pasted last.  That accept / reject sequence depends on the mapping alone, not
    on voxel data, so it is replayed here on the host with the reference's own torch fp32 arithmetic; the volume is
    then cut along every patch boundary into cells that have ONE owner (the last accepted patch covering them), and
    the cells go through rf_compose_gather like non-overlapping patches.  Cells nobody owns get scene id -2 (= keep
    the pre-filled truncation value).  rows [P,K,8], dst int [P,6] -> (cell rows [Pc,K,8], cell extents [Pc,6])."""
    P, K = rows.shape[:2]
    cuts = [sorted({0, int(size[a])} | {int(v) for v in dst[:, 2 * a]} | {int(v) for v in dst[:, 2 * a + 1]}) for a in range(3)]
    cuts = [[c for c in cs if 0 <= c <= size[a]] for a, cs in enumerate(cuts)]
    ncell = [len(c) - 1 for c in cuts]
    pos = [{c: i for i, c in enumerate(cs)} for cs in cuts]
    owner = np.full([K] + ncell, -1, dtype=np.int64)
    distances = torch.ones((K,) + tuple(int(v) for v in size), dtype=torch.float32) * 100
    for k in range(K):
        for p in range(P):
            cur = rows[p, k, 7]
            x0, x1, y0, y1, z0, z1 = [int(v) for v in dst[p]]
            if distances[k, x0:x1, y0:y1, z0:z1].mean() > cur:
                distances[k, x0:x1, y0:y1, z0:z1] = float(cur)
                owner[k, pos[0][x0]:pos[0][x1], pos[1][y0]:pos[1][y1], pos[2][z0]:pos[2][z1]] = p
    cells = [(i, j, l) for i in range(ncell[0]) for j in range(ncell[1]) for l in range(ncell[2])]
    cell_ext = np.array([[cuts[0][i], cuts[0][i + 1], cuts[1][j], cuts[1][j + 1], cuts[2][l], cuts[2][l + 1]] for i, j, l in cells],
                        dtype=np.int32)
    cell_rows = np.zeros((len(cells), K, 8), dtype=np.float32)
    cell_rows[:, :, 0] = -2
    for ci, (i, j, l) in enumerate(cells):
        for k in range(K):
            p = owner[k, i, j, l]
            if p < 0:
                continue
            r = rows[p, k]
            off = cell_ext[ci, 0::2] - dst[p, 0::2]            # where the cell starts inside the owner's block
            ext = cell_ext[ci, 1::2] - cell_ext[ci, 0::2]
            src0 = r[1:7:2].astype(np.int32) + off
            src1 = np.minimum(src0 + ext, r[2:7:2].astype(np.int32))
            cell_rows[ci, k] = [r[0], src0[0], src1[0], src0[1], src1[1], src0[2], src1[2], r[7]]
    return cell_rows, cell_ext


def create_retrieval_from_mapping(scene_name, retrieval_mappings, K, dataset_train, dataset, tree_path, dataset_index=None):
    """util/retrieval.py:145-164 for one scene -> CPU tensor [K, X, Y, Z].
    The retrieved scenes are uploaded once (padded to a common size when the bank's scenes differ) and gathered by
    rf_compose_gather; with overlapping strides the reference's accept / reject sequence is replayed first
    (_overlap_cells)."""
    if dataset_index is None:
        dataset_index = json.loads((Path(tree_path) / "index.json").read_text())
    dev = torch.device("cuda", torch.cuda.current_device())
    size = tuple(int(v) for v in dataset.get_scene_size(scene_name))
    patches = dataset.patch_from_scene_lookup[scene_name]
    trunc = float(dataset.target_trunc)
    if len(patches) == 0:
        return torch.from_numpy(np.ones((K,) + size, dtype=np.float32)) * dataset.target_trunc
    rows = np.stack([np.asarray(retrieval_mappings[p], dtype=np.float32)[:K] for p in patches])  # [P,K,8]
    dst = np.array([dataset_train.unpad(*dataset.scene_handler.get_extent_from_name(p)[1]) for p in patches], dtype=np.int32)
    if not dataset.no_overlap:
        rows, dst = _overlap_cells(rows, dst, size)
    used = sorted({int(v) for v in rows[:, :, 0].reshape(-1) if v >= 0})
    if used:
        scenes = [dataset_train.get_scene_target(dataset_index[s]).astype(np.float32) for s in used]
        ssz = tuple(max(sc.shape[a] for sc in scenes) for a in range(3))
        store = np.full((len(scenes),) + ssz, np.float32(dataset_train.target_trunc), dtype=np.float32)
        for i, sc in enumerate(scenes):  # the source extents address the scene's own coordinates: padding is never read
            store[i, :sc.shape[0], :sc.shape[1], :sc.shape[2]] = sc
    else:
        store = np.zeros((1,) + size, dtype=np.float32)
    remap = {s: i for i, s in enumerate(used)}
    rows_local = rows.copy()
    for s, i in remap.items():
        rows_local[:, :, 0][rows[:, :, 0] == s] = i
    ratio = np.float32(dataset.target_trunc) / np.float32(dataset_train.target_trunc)
    out = ops.compose_gather(torch.from_numpy(rows_local).to(dev), torch.from_numpy(np.ascontiguousarray(dst)).to(dev),
                             torch.from_numpy(store).to(dev), 1, size, trunc, float(ratio),
                             prefill=True)  # patches the dataset filtered out keep the truncation value (:148)
    return out[0].cpu()


def get_metrics_for_retrieval(retrievals, dataset):
    """util/retrieval.py:167-175: [IoU, Chamfer, Precision, Recall] of the first retrieved volume of every scene against
    its target, both thresholded at 0.75 voxels.  retrievals[i]: [K, X, Y, Z] array or tensor of scene i."""
    from .metrics import Chamfer3D, IoU, Precision, Recall
    dev = torch.device("cuda", torch.cuda.current_device())
    metrics = [IoU(compute_on_step=False).cuda(), Chamfer3D(compute_on_step=False).cuda(), Precision(compute_on_step=False).cuda(),
               Recall(compute_on_step=False).cuda()]
    thr = 0.75 * dataset.target_voxel_size
    for idx, scene in enumerate(dataset.scenes):
        r = torch.as_tensor(np.asarray(retrievals[idx]))[0]
        nn1 = (r <= thr)[None, None].to(dev)
        target = (torch.as_tensor(dataset.get_scene_target(scene)) <= thr)[None, None].to(dev)
        for metric in metrics:
            metric(nn1, target)
    return [float(m.compute()) for m in metrics]


class RetrievalInterface:
    """util/retrieval.py:178-207."""

    def __init__(self, config_query, latent_dim):
        self.config = config_query
        self.latent_dim = latent_dim

    def get_retrieval_mapping(self, fenc, extraction_func, tree_path, dataset, ignore_patches_from_source):
        patch_names, feats_input = extraction_func(fenc, self.config, self.latent_dim, dataset)
        return query_dictionary_using_features(self.config, patch_names, feats_input, dataset, tree_path, ignore_patches_from_source)

    def get_features(self, fenc_input, fenc_target, dataset):
        names0, feats_input = extract_input_features(fenc_input, self.config, self.latent_dim, dataset)
        names1, feats_target = extract_target_features(fenc_target, self.config, self.latent_dim, dataset)
        assert len(names0) == len(names1) and sorted(names0) == sorted(names1)
        return names0, feats_input, feats_target

    @staticmethod
    def retrieve_nearest_scenes(retrieval_mapping, scene, K, tree_path, dataset_train, dataset):
        return create_retrieval_from_mapping(scene, retrieval_mapping, K, dataset_train, dataset, tree_path)

    @staticmethod
    def retrieve_nearest_scenes_for_all(retrieval_mapping, scenes, K, tree_path, dataset_train, dataset):
        index = json.loads((Path(tree_path) / "index.json").read_text())
        return torch.cat([create_retrieval_from_mapping(s, retrieval_mapping, K, dataset_train, dataset, tree_path, index).unsqueeze(0)
                          for s in scenes], dim=0)

    def create_mapping_and_retrieve_nearest_scenes_for_all(self, fenc_input, tree_path, dataset_train, dataset, K,
                                                           ignore_patches_from_source):
        mapping = self.get_retrieval_mapping(fenc_input, extract_input_features, tree_path, dataset, ignore_patches_from_source)
        return RetrievalInterface.retrieve_nearest_scenes_for_all(mapping, dataset.scenes, K, tree_path, dataset_train, dataset)


def retrievals_to_disk(mode, config, use_target_for_feats, fenc_input=None, fenc_target=None, num_proc=1, proc=0,
                       scene_handlers=None):
    """util/retrieval.py:210-248 'map' and 'compose' modes, same files on disk
    (map_{train,val}.npy pickled dicts, compose/<scene>.npz arr_0 [K,X,Y,Z]).
    Encoders are passed in (checkpoint loading is the caller's business)."""
    from ..dataset.patched_scene_dataset import PatchedSceneDataset
    from ..dataset.scene import SceneHandler
    ckpt_experiment = Path(config["retrieval_ckpt"]).parents[0].name
    ckpt_epoch = Path(config["retrieval_ckpt"]).name.split(".")[0]
    task_dir = f"{config['task']}_{config['dataset_train']['num_points']:04d}"
    retrievals_dir = get_retrievals_dir(config)
    tree_path = Path("runs", "retrieval_scratch", task_dir, config["dataset_train"]["dataset_name"],
                     config["dataset_train"]["splits_dir"], ckpt_experiment, ckpt_epoch, str(config["K"]))
    if scene_handlers is None:
        scene_handlers = {"train": SceneHandler("train", config), "val": SceneHandler("val", config)}
    dataset_train = PatchedSceneDataset("train", config["dataset_train"], scene_handlers["train"])
    dataset_val = PatchedSceneDataset("val", config["dataset_val"], scene_handlers["val"])
    if mode == "map":
        retrievals_dir.mkdir(exist_ok=True, parents=True)
        create_dictionary(fenc_target, config["dictionary"], config["retrieval_model"]["latent_dim"], dataset_train, tree_path)
        handler = RetrievalInterface(config["query"], config["retrieval_model"]["latent_dim"])
        fenc = fenc_target if use_target_for_feats else fenc_input
        extract = extract_target_features if use_target_for_feats else extract_input_features
        np.save(retrievals_dir / "map_train.npy", handler.get_retrieval_mapping(fenc, extract, tree_path, dataset_train, True))
        np.save(retrievals_dir / "map_val.npy", handler.get_retrieval_mapping(fenc, extract, tree_path, dataset_val, False))
    elif mode == "compose":
        (retrievals_dir / "compose").mkdir(exist_ok=True, parents=True)
        for name, dataset in (("map_train.npy", dataset_train), ("map_val.npy", dataset_val)):
            mapping = np.load(retrievals_dir / name, allow_pickle=True)[()]
            for scene in [x for i, x in enumerate(dataset.scenes) if i % num_proc == proc]:
                vol = RetrievalInterface.retrieve_nearest_scenes(mapping, scene, config["K"], tree_path, dataset_train, dataset)
                np.savez_compressed(retrievals_dir / "compose" / f"{scene}.npz", vol.numpy())
    elif mode == "evaluate":  # util/retrieval.py:250-255: metrics of the first retrieval of every val scene
        retrievals = [np.load(retrievals_dir / "compose" / f"{scene}.npz")["arr_0"][:1] for scene in dataset_val.scenes]
        metrics = get_metrics_for_retrieval(retrievals, dataset_val)
        print(metrics)
        return metrics
    else:
        raise ValueError(f"unknown mode '{mode}' (map | compose | evaluate)")
    return tree_path, retrievals_dir
