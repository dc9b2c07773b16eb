"""Golden vectors for the SURVEY 8f rows (NT-Xent loss, Sobel normals), produced in the build container from the
reference's own sources under /root/reference (which does not travel to the GPU box).

* model/loss.py imports cleanly (torch + numpy); NTXentLoss.forward calls `mask.cuda(device)`, which has no meaning
  on a CPU-only box, so Tensor.cuda is patched to the identity for the duration of the call - nothing in the
  reference is modified.
* dataset/patched_scene_dataset.py cannot be imported (util.misc -> trimesh); its Sobel kernels are class-level
  literals (:194-196), which this script parses out of the source text, checks against the oracle's constants and
  runs through the same pad + conv3d + normalise sequence as compute_normals (:139-146).

    python tests/golden/make_golden_adjuncts.py     ->  tests/golden/adjuncts.npz
"""
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import rf_oracle as O  # noqa: E402


def main():
    out = {}
    # ---- NT-Xent through the reference class
    from model.loss import NTXentLoss
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        g = torch.Generator().manual_seed(77)
        for tag, n, c, cosine, with_iou, temp in [("cos", 37, 32, True, False, 0.2), ("dot", 20, 64, False, False, 0.5),
                                                  ("cos_iou", 16, 32, True, True, 0.3), ("cos_big", 300, 32, True, False, 0.15)]:
            zis, zjs = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
            zjs = 0.45 * zis + 0.55 * zjs                    # correlated pairs, as the training produces
            if not cosine:
                zis, zjs = 0.15 * zis, 0.15 * zjs            # keeps the dot-product logits O(1)
            iou = torch.rand(2 * n, 2 * n, generator=g) if with_iou else None
            ref = NTXentLoss(temp, cosine)(zis, zjs, iou)
            mine = O.ntxent_loss(zis, zjs, temp, cosine, iou)
            assert abs(float(ref) - float(mine)) <= 1e-6 * max(1.0, abs(float(ref))), (tag, float(ref), float(mine))
            out[f"ntxent.{tag}.zis"], out[f"ntxent.{tag}.zjs"] = zis.numpy(), zjs.numpy()
            if with_iou:
                out[f"ntxent.{tag}.iou"] = iou.numpy()
            out[f"ntxent.{tag}.loss"] = np.float32(float(ref))
            out[f"ntxent.{tag}.cfg"] = np.array([temp, float(cosine)], dtype=np.float32)
            print(tag, float(ref))
    finally:
        torch.Tensor.cuda = orig_cuda
    # ---- Sobel kernels: literals of the reference source vs the oracle's constants
    src = open(os.path.join(REF, "dataset", "patched_scene_dataset.py")).read()
    for name, mine in (("sobel_3d_x", O.SOBEL_3D_X), ("sobel_3d_y", O.SOBEL_3D_Y), ("sobel_3d_z", O.SOBEL_3D_Z)):
        m = re.search(name + r" = torch\.from_numpy\(np\.array\((\[.*?\]), dtype=np\.float32\)\)", src)
        assert m, name
        ref = np.array(eval(m.group(1)), dtype=np.float32)  # a nested list literal of integers
        assert ref.shape == (3, 3, 3) and np.array_equal(ref, mine), name
    from oracle.rf_oracle import synthetic_tsdf
    tgt = torch.from_numpy(np.stack([synthetic_tsdf(s, 32, 0.05) for s in (3, 4)])[:, None])
    trunc = float(np.float16(0.15))
    nrm = O.compute_normals(tgt, trunc)
    out["normals.target"], out["normals.trunc"], out["normals.out"] = tgt.numpy(), np.float32(trunc), nrm.numpy()
    np.savez_compressed(os.path.join(HERE, "adjuncts.npz"), **out)
    print("wrote adjuncts.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
