"""Chunk access with the reference's SceneHandler interface (dataset/scene.py).

`SceneHandler(split, config)` reads the reference's on-disk layout (`.npz`
scenes under data/, splits, size json); `InMemorySceneHandler` serves the same
API from arrays (synthetic data, tests, bench).  Both are host-side numpy: the
reference does this work in dataloader workers, not on the GPU.  Visualisation
helpers and the occupancy cache files are out of scope."""
import json
import random
from pathlib import Path

import numpy as np


def point_cloud_to_grid(point_cloud, grid_res, scale_factor, pad):
    """util/misc.py:73-78."""
    grid = np.zeros([grid_res + 2 * pad] * 3, dtype=np.float32)
    pts = np.clip(point_cloud * scale_factor, 0, grid_res - 1).astype(np.uint32)
    grid[pad + pts[:, 0], pad + pts[:, 1], pad + pts[:, 2]] = 1
    return grid


class _SceneAccess:
    """Everything of SceneHandler that does not touch the file system."""

    def _init_geometry(self, task, dataset_config):
        self.task = task
        c = dataset_config
        self.input_chunk_size = c["input_chunk_size"]
        self.target_chunk_size = c["target_chunk_size"]
        self.number_point_samples = c.get("num_points", 0)
        # dataset/scene.py:30-33: voxel sizes and truncations are rounded through float16
        self.input_voxel_size = np.float16(c["voxel_size_input"]).astype(np.float32)
        self.target_voxel_size = np.float16(c["voxel_size_target"]).astype(np.float32)
        self.input_trunc = np.float16(c["voxel_size_input"] * 3).astype(np.float32)
        self.target_trunc = np.float16(c["voxel_size_target"] * 3).astype(np.float32)
        self.patch_size_target = c["patch_size_target"]
        self.patch_context_target = c["patch_context_target"]
        self.patch_stride_target = c["patch_stride"]
        self.patch_size_input = c["patch_size_input"]
        self.patch_context_input = c["patch_context_input"]
        self.patch_stride_input = int(c["patch_stride"] * c["patch_size_input"] / c["patch_size_target"])  # :40
        self.scale_factor = c["patch_size_target"] / c["patch_size_input"]
        self.scene_size = {}
        self.scene_occupancy = {}
        self.scenes = []

    @staticmethod
    def get_extents_for_size(size, patch_size, patch_context, patch_stride):
        """dataset/scene.py:153-160: x-major patch grid, extents in PADDED coordinates."""
        starts = [np.linspace(0, size[a] - patch_size, (size[a] - patch_size) // patch_stride + 1).astype(np.int32)
                  for a in range(3)]
        xs, ys, zs = np.meshgrid(*starts, indexing="ij")
        e = patch_size + 2 * patch_context
        cols = [xs, xs + e, ys, ys + e, zs, zs + e]
        return np.hstack([c.flatten()[:, np.newaxis] for c in cols])

    def get_scene_patches(self, scene):
        """dataset/scene.py:162-167."""
        size_target = self.scene_size[scene]
        size_input = [int(s / self.scale_factor) for s in size_target]
        ext_t = self.get_extents_for_size(size_target, self.patch_size_target, self.patch_context_target, self.patch_stride_target)
        ext_i = self.get_extents_for_size(size_input, self.patch_size_input, self.patch_context_input, self.patch_stride_input)
        return ext_i, ext_t

    @staticmethod
    def get_name_from_extent(scene, e):
        return f"{scene}--{e[0]:04d}_{e[1]:04d}_{e[2]:04d}_{e[3]:04d}_{e[4]:04d}_{e[5]:04d}"

    @staticmethod
    def get_extent_from_name(identifier):
        scene, rest = identifier.split("--")
        return scene, [int(r) for r in rest.split("_")]

    def get_patch_occupancy(self, scene, target_extent):
        return self.scene_occupancy.get(self.get_name_from_extent(scene, target_extent), 1)  # :208-213

    def calculate_occupancy_for_name(self, patch_identifier):
        scene, e = self.get_extent_from_name(patch_identifier)  # :149-151
        return int((self.get_scene_target(scene)[e[0]:e[1], e[2]:e[3], e[4]:e[5]] <= 0.75 * 2 * self.target_voxel_size).sum())


class InMemorySceneHandler(_SceneAccess):
    """SceneHandler API over arrays. inputs: name -> [s,s,s] SDF (superresolution)
    or [n,3] point cloud (surface_reconstruction); targets: name -> [S,S,S] SDF;
    retrievals (optional): name -> [K,S,S,S]."""

    def __init__(self, task, dataset_config, inputs, targets, retrievals=None, compute_occupancy=False):
        self._init_geometry(task, dataset_config)
        self.scenes = list(targets.keys())
        self.split_shapes = list(self.scenes)
        self._inputs, self._targets, self._retrievals = inputs, targets, retrievals or {}
        self.use_retrievals = retrievals is not None
        for s in self.scenes:
            self.scene_size[s] = list(targets[s].shape)
        if compute_occupancy:
            for s in self.scenes:
                for e in self.get_scene_patches(s)[1]:
                    n = self.get_name_from_extent(s, e)
                    self.scene_occupancy[n] = self.calculate_occupancy_for_name(n)

    def get_scene_input(self, scene):
        if self.task == "surface_reconstruction":  # dataset/scene.py:81-90 (no random subsampling here)
            return point_cloud_to_grid(self._inputs[scene], self.input_chunk_size, 1 / self.scale_factor, self.patch_context_input)
        return np.pad(self._inputs[scene].astype(np.float32), self.patch_context_input, mode="constant",
                      constant_values=self.input_trunc)  # :60-61

    def get_scene_target(self, scene):
        return np.pad(self._targets[scene].astype(np.float32), self.patch_context_target, mode="constant",
                      constant_values=self.target_trunc)  # :92-95

    def get_scene_retrieval(self, scene):
        r = self._retrievals[scene].astype(np.float32)
        c = self.patch_context_target
        return np.pad(r, ((0, 0), (c, c), (c, c), (c, c)), mode="constant", constant_values=self.target_trunc)  # :97-100


class SceneHandler(_SceneAccess):
    """dataset/scene.py:13-229 file-backed handler: same constructor, same
    access methods (get_scene_input / get_scene_target / get_scene_retrieval /
    get_scene_patches / extents <-> names).  Visualisation is not provided."""

    def __init__(self, split, config):
        dc = config[f"dataset_{split}"]
        self._init_geometry(config["task"], dc)
        self.preloaded_scenes_input, self.preloaded_scenes_target, self.preloaded_retrievals = {}, {}, {}
        self.input_ext, self.target_ext = dc["input_ext"], dc["target_ext"]
        self.input_path = Path(dc["scene_dir"], dc["input_dir"], dc["dataset_name"])
        self.target_path = Path(dc["scene_dir"], dc["target_dir"], dc["dataset_name"])
        split_file = Path(f"{dc['data_dir']}/splits/{dc['dataset_name']}/{dc['splits_dir']}/{split}.txt")
        self.split_shapes = [x.strip() for x in split_file.read_text().split("\n") if x.strip() != ""]
        self.scenes = list(self.split_shapes)
        self.use_retrievals = not config["no_retrievals"]
        self.retrievals_dir = None
        if self.use_retrievals:
            from ..util.retrieval import get_retrievals_dir
            self.retrievals_dir = get_retrievals_dir(config)
        self.random_indices_list = None
        if dc["preload_scenes"]:
            for s in self.scenes:
                self.preloaded_scenes_input[s] = self._load_input(s)
                self.preloaded_scenes_target[s] = self._load_target(s, np.float16)
        if self.use_retrievals and dc["preload_retrievals"]:
            for s in self.scenes:
                self.preloaded_retrievals[s] = self._load_retrieval(s, np.float16)
        self._init_random_indices(Path(dc["data_dir"], "random_indices", f"{self.number_point_samples}.npz"))
        self._init_scene_sizes(Path(dc["data_dir"], "size", dc["dataset_name"] + ".json"))
        if not dc["skip_occupancy"]:
            self._init_occupancy(Path(dc["data_dir"], "occupancy",
                                      f"{dc['dataset_name']}_{self.target_chunk_size:03d}_{self.patch_size_target:02d}_{self.patch_context_target:02d}.json"))

    # -- loaders (dataset/scene.py:60-100)
    def _load_input(self, scene):
        if self.task == "surface_reconstruction":
            return np.load(self.input_path / (scene + self.input_ext))["arr_0"]
        return np.pad(np.load(self.input_path / (scene + self.input_ext))["arr"].astype(np.float16), self.patch_context_input,
                      mode="constant", constant_values=self.input_trunc)

    def _load_target(self, scene, dtype):
        return np.pad(np.load(self.target_path / (scene + self.target_ext))["arr"].astype(dtype), self.patch_context_target,
                      mode="constant", constant_values=self.target_trunc)

    def _load_retrieval(self, scene, dtype):
        c = self.patch_context_target
        arr = np.load(self.retrievals_dir / "compose" / (scene + ".npz"))["arr_0"].astype(dtype)
        # the reference pads ALL axes (incl. K) because np.pad gets a scalar width (:99); ctx is 0 in refinement configs
        return np.pad(arr, c, mode="constant", constant_values=self.target_trunc)

    def get_scene_input(self, scene):
        src = self.preloaded_scenes_input[scene] if scene in self.preloaded_scenes_input else self._load_input(scene)
        if self.task != "surface_reconstruction":
            return src.astype(np.float32)
        pc = src
        if pc.shape[0] < 20000:
            pc = np.vstack([pc, pc])
        idx = self.random_indices_list[random.randint(0, self.random_indices_list.shape[0] - 1)]
        return point_cloud_to_grid(pc[idx, :], self.input_chunk_size, 1 / self.scale_factor, self.patch_context_input)

    def get_scene_target(self, scene):
        if scene in self.preloaded_scenes_target:
            return self.preloaded_scenes_target[scene].astype(np.float32)
        return self._load_target(scene, np.float32)

    def get_scene_retrieval(self, scene):
        if scene in self.preloaded_retrievals:
            return self.preloaded_retrievals[scene].astype(np.float32)
        return self._load_retrieval(scene, np.float32)

    # -- caches (dataset/scene.py:102-147)
    def _init_random_indices(self, filepath):
        if filepath.exists():
            self.random_indices_list = np.load(filepath)["arr"]
        elif self.task == "surface_reconstruction":
            lst = [random.sample(range(20000), self.number_point_samples) for _ in range(20000 * 10)]
            self.random_indices_list = np.array(lst)
            filepath.parents[0].mkdir(exist_ok=True)
            np.savez_compressed(filepath, arr=self.random_indices_list)

    def _init_scene_sizes(self, filepath):
        if filepath.exists():
            self.scene_size = json.loads(filepath.read_text())
        if any(s not in self.scene_size for s in self.scenes):
            for s in self.scenes:
                self.scene_size[s] = [v - 2 * self.patch_context_target for v in self.get_scene_target(s).shape]
            filepath.parents[0].mkdir(exist_ok=True)
            filepath.write_text(json.dumps(self.scene_size))

    def _init_occupancy(self, filepath):
        if filepath.exists():
            self.scene_occupancy = json.loads(filepath.read_text())
        names = [self.get_name_from_extent(s, e) for s in self.scenes for e in self.get_scene_patches(s)[1]]
        if any(n not in self.scene_occupancy for n in names):
            for n in names:
                self.scene_occupancy[n] = self.calculate_occupancy_for_name(n)
            filepath.parents[0].mkdir(exist_ok=True)
            filepath.write_text(json.dumps(self.scene_occupancy))
