"""util/patcher.py:4-42 Patcher with the reference's interface, on the GPU."""
from .. import ops


class Patcher(object):

    def __init__(self, patch_size, side, stride, pad_val, base_size):
        super().__init__()
        self.pad = list(side)
        self.stride = list(stride)
        self.pad_val = pad_val
        self.base_size = list(base_size)
        self.kernel = [patch_size[i] + 2 * side[i] for i in range(len(patch_size))]

    def __call__(self, x):
        """[B,C,X,Y,Z] -> [B*n0*n1*n2, C, k0,k1,k2] (util/patcher.py:14-19)."""
        return ops.unfold3d_pad_stride(x, self.kernel, self.pad, self.stride, self.pad_val, keep_channels=True)

    def recompose_patches(self, original_shape, patches):
        """[B, n_patches, k0,k1,k2] -> [B,C,X,Y,Z] (util/patcher.py:21-30): patches
        are written in scan order, later ones overwrite overlaps, padding cropped."""
        padded = [original_shape[2 + a] + 2 * self.pad[a] for a in range(3)]
        # the reference bounds its z loop with patches.shape[2] (sic, :26)
        bound = [patches.shape[2], patches.shape[3], patches.shape[2]]
        count = [len(range(0, padded[a] - bound[a] + 1, self.stride[a])) for a in range(3)]
        if patches.shape[1] < count[0] * count[1] * count[2]:
            raise IndexError("recompose_patches: not enough patches for the volume")
        return ops.recompose_patches(patches, original_shape, self.kernel, self.pad, self.stride, count, self.pad_val)

    def get_patch_extents(self):
        return [self.kernel[i] - 2 * self.pad[i] for i in range(3)]

    def get_patch_ratio(self):
        return [self.base_size[i] // (self.kernel[i] - 2 * self.pad[i]) for i in range(3)]

    def get_stride_ratio(self):
        return [self.get_patch_extents()[i] // self.stride[i] for i in range(3)]

    def get_patch_counts(self):
        return [(self.base_size[i] + self.pad[i] * 2 - self.kernel[i]) // self.stride[i] + 1 for i in range(3)]
