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
    # ---- evaluation metrics through the reference's own util/metrics.py classes.  torchmetrics is absent:
    # its Metric base is stubbed (add_state = setattr); the compiled Chamfer extension is replaced by the pure-torch
    # implementation the submodule itself ships to validate that extension (external/ChamferDistancePytorch/
    # chamfer_python.py distChamfer, cf. its unit_test.py) - exact on the integer voxel coordinates the metric feeds it.
    import types

    class Metric(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def add_state(self, name, default, dist_reduce_fx=None):
            setattr(self, name, default)

        def __call__(self, *a):
            self.update(*a)
    tmm = types.ModuleType("torchmetrics.metric"); tmm.Metric = Metric
    tm = types.ModuleType("torchmetrics"); tm.__path__ = []; tm.metric = tmm
    sys.modules["torchmetrics"], sys.modules["torchmetrics.metric"] = tm, tmm
    sys.path.insert(0, os.path.join(REF, "external", "ChamferDistancePytorch"))
    import chamfer_python
    dc = types.ModuleType("external.ChamferDistancePytorch.chamfer3D.dist_chamfer_3D")
    def dist_stub(a, b):
        if a.shape[1] == 0 or b.shape[1] == 0:  # the metric relies on the NaN mean of such a pair and skips it (:48)
            nan = float("nan")
            return (torch.full((1, a.shape[1]), nan), torch.full((1, b.shape[1]), nan), torch.zeros((1, a.shape[1]), dtype=torch.int32),
                    torch.zeros((1, b.shape[1]), dtype=torch.int32))
        return chamfer_python.distChamfer(a, b)
    dc.chamfer_3DDist = lambda: dist_stub
    sys.modules["external.ChamferDistancePytorch.chamfer3D.dist_chamfer_3D"] = dc
    from util import metrics as ref_metrics
    rng = np.random.default_rng(8)
    tgt_m = np.stack([synthetic_tsdf(40 + i, 32, 0.05) for i in range(5)])[:, None]
    pred_m = tgt_m + rng.normal(size=tgt_m.shape).astype(np.float32) * 0.02
    p, t = pred_m <= 0.05 * 0.75, tgt_m <= 0.05 * 0.75
    p[3] = False
    p[4] = False
    t[4] = False
    pt_, tt_ = torch.from_numpy(p), torch.from_numpy(t)
    vals = []
    for cls in (ref_metrics.IoU, ref_metrics.Chamfer3D, ref_metrics.Precision, ref_metrics.Recall):
        m = cls(compute_on_step=False)
        m(pt_[:3], tt_[:3])
        m(pt_[3:], tt_[3:])
        vals.append(float(m.compute()))
    out["metrics.pred"], out["metrics.target"] = p, t
    out["metrics.values"] = np.array(vals, dtype=np.float64)  # IoU, Chamfer, Precision, Recall
    a = torch.nonzero(tt_[0, 0], as_tuple=False).float()[None]
    b = torch.nonzero(pt_[0, 0], as_tuple=False).float()[None]
    d1, d2, i1, i2 = chamfer_python.distChamfer(a, b)
    out["chamfer.d1"], out["chamfer.d2"] = d1[0].numpy(), d2[0].numpy()
    out["chamfer.i1"], out["chamfer.i2"] = i1[0].numpy(), i2[0].numpy()
    iou_sum, iou_n, prec, rec, _ = O.occupancy_metrics(p, t)
    cd, valid = O.chamfer_metric(p, t)
    mine = [iou_sum / iou_n, cd / valid, prec / p.shape[0], rec / p.shape[0]]
    assert np.allclose(mine, vals, rtol=1e-5, atol=1e-7), (mine, vals)
    print("metrics", vals)
    np.savez_compressed(os.path.join(HERE, "adjuncts.npz"), **out)
    print("wrote adjuncts.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
