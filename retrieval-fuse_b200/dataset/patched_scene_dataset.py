"""dataset/patched_scene_dataset.py:13-137: the patch dataset the retrieval
pre-pass iterates (host side).  Same constructor and item dict; scenes come
from the handler (file existence filtering of the reference is the handler's
job here)."""
from collections import defaultdict

import numpy as np
from torch.utils.data.dataset import Dataset

from .scene import _SceneAccess


class PatchedSceneDataset(Dataset):

    def __init__(self, split, dataset_config, scene_handler):
        self.scene_handler = scene_handler
        self.dataset_name = dataset_config.get("dataset_name")
        self.input_mean, self.input_std = dataset_config["input_mean"], dataset_config["input_std"]
        self.target_mean, self.target_std = dataset_config["target_mean"], dataset_config["target_std"]
        self.use_retrievals = scene_handler.use_retrievals
        self.scenes = list(scene_handler.scenes)
        self.data = []
        for s in self.scenes:
            ext_i, ext_t = scene_handler.get_scene_patches(s)
            for ii in range(len(ext_i)):
                if scene_handler.get_patch_occupancy(s, ext_t[ii]) > dataset_config["occupancy_threshold"]:
                    self.data.append([s, ext_i[ii], ext_t[ii]])
        self.patch_from_scene_lookup = defaultdict(list)
        for d in self.data:
            self.patch_from_scene_lookup[d[0]].append(_SceneAccess.get_name_from_extent(d[0], d[2]))
        if split == "train":
            self.data = self.data * dataset_config.get("train_multiplier", 1)

    target_trunc = property(lambda self: self.scene_handler.target_trunc)
    target_voxel_size = property(lambda self: self.scene_handler.target_voxel_size)
    input_trunc = property(lambda self: self.scene_handler.input_trunc)
    input_voxel_size = property(lambda self: self.scene_handler.input_voxel_size)
    target_patch_size = property(lambda self: self.scene_handler.patch_size_target)
    target_patch_context = property(lambda self: self.scene_handler.patch_context_target)
    input_chunk_size = property(lambda self: self.scene_handler.input_chunk_size)
    target_chunk_size = property(lambda self: self.scene_handler.target_chunk_size)
    no_overlap = property(lambda self: self.scene_handler.patch_stride_target == self.scene_handler.patch_size_target)

    def get_scene_size(self, scene):
        return self.scene_handler.scene_size[scene]

    def get_scene_indices(self, scenes):
        return np.array([self.scenes.index(s) for s in scenes])

    def get_scene_names_from_patches(self, patch_names):
        return [self.scene_handler.get_extent_from_name(x)[0] for x in patch_names]

    def __len__(self):
        return len(self.data)

    def _unpadded(self, padded, ctx):
        return padded[ctx: padded.shape[0] - ctx, ctx: padded.shape[1] - ctx, ctx: padded.shape[2] - ctx]

    def get_scene_input(self, scene):
        return self._unpadded(self.scene_handler.get_scene_input(scene), self.scene_handler.patch_context_input)

    def get_scene_target(self, scene):
        return self._unpadded(self.scene_handler.get_scene_target(scene), self.scene_handler.patch_context_target)

    def unpad(self, *extents):  # :101-105
        c = self.scene_handler.patch_context_target
        if len(extents) == 2:
            return [extents[0], extents[1] - 2 * c]
        return self.unpad(*extents[0:2]) + self.unpad(*extents[2:4]) + self.unpad(*extents[4:6])

    def pad(self, *extents):  # :107-111
        c = self.scene_handler.patch_context_target
        if len(extents) == 2:
            return [extents[0], extents[1] + 2 * c]
        return self.pad(*extents[0:2]) + self.pad(*extents[2:4]) + self.pad(*extents[4:6])

    def denormalize_target(self, target):
        return target * self.target_std + self.target_mean

    def compute_normals(self, target):
        """dataset/patched_scene_dataset.py:139-146: Sobel normals of a (denormalised) target batch [B,1,D,H,W] on the
        GPU, padded with the target truncation value."""
        from .. import ops
        return ops.sobel_normals(target, self.scene_handler.target_trunc)

    def __getitem__(self, index):  # :117-137
        scene, ei, et = self.data[index]
        sin = self.scene_handler.get_scene_input(scene)
        stg = self.scene_handler.get_scene_target(scene)
        p_in = sin[ei[0]:ei[1], ei[2]:ei[3], ei[4]:ei[5]]
        p_tg = stg[et[0]:et[1], et[2]:et[3], et[4]:et[5]]
        item = {
            "name": _SceneAccess.get_name_from_extent(scene, et),
            "scene": scene,
            "extent": et,
            "input": (p_in[np.newaxis, ...] - self.input_mean) / self.input_std,
            "target": (p_tg[np.newaxis, ...] - self.target_mean) / self.target_std,
        }
        if self.use_retrievals:
            sre = self.scene_handler.get_scene_retrieval(scene)
            item["retrieval"] = (sre[:, et[0]:et[1], et[2]:et[3], et[4]:et[5]] - self.target_mean) / self.target_std
        else:
            item["retrieval"] = np.ones((4, et[1] - et[0], et[3] - et[2], et[5] - et[4]), dtype=np.float32) * self.target_trunc
        return item
