"""util/metrics.py of the reference without torchmetrics: the same update / compute semantics, the
integer sums and the nearest-neighbour searches on the GPU (rf_occupancy_counts, rf_chamfer_nn)."""
import torch

from .. import ops


class _Metric:
    """The slice of torchmetrics.Metric the reference's callers use: `IoU(compute_on_step=False).cuda()` then
    `metric(preds, target)` per batch and `metric.compute()` (util/retrieval.py:168-175, trainer/*.py)."""

    def __init__(self, *args, **kwargs):  # torchmetrics options (compute_on_step, dist_sync_on_step, ...) are accepted
        self._reset_state()

    def _reset_state(self):
        raise NotImplementedError

    def __call__(self, preds, target):
        self.update(preds, target)

    def reset(self):
        self._reset_state()

    def cuda(self, *args, **kwargs):  # the state lives on the host; the kernels run where the inputs are
        return self

    def to(self, *args, **kwargs):
        return self

    def cpu(self):
        return self


class IoU(_Metric):
    """util/metrics.py:6-25."""

    def _reset_state(self):
        self.iou_sum = 0.0
        self.total = 0.0

    def update(self, preds, target):
        c = ops.occupancy_counts(preds, target).cpu()
        inter, union = c[:, 0].float(), c[:, 1].float()
        valid = union > 0
        inter, union = inter[valid], union[valid]
        if union.sum() > 0:
            self.iou_sum += float((inter / (union + 1e-5)).sum())
            self.total += inter.shape[0]

    def compute(self):
        return torch.tensor(self.iou_sum, dtype=torch.float32) / self.total


class Precision(_Metric):
    """util/metrics.py:56-70."""

    def _reset_state(self):
        self.precision_sum = 0.0
        self.total = 0.0

    def update(self, preds, target):
        c = ops.occupancy_counts(preds, target).cpu()
        self.precision_sum += float((c[:, 0].float() / (c[:, 2].float() + 1e-5)).sum())
        self.total += c.shape[0]

    def compute(self):
        return torch.tensor(self.precision_sum, dtype=torch.float32) / self.total


class Recall(_Metric):
    """util/metrics.py:73-87."""

    def _reset_state(self):
        self.recall_sum = 0.0
        self.total = 0.0

    def update(self, preds, target):
        c = ops.occupancy_counts(preds, target).cpu()
        self.recall_sum += float((c[:, 0].float() / (c[:, 3].float() + 1e-5)).sum())
        self.total += c.shape[0]

    def compute(self):
        return torch.tensor(self.recall_sum, dtype=torch.float32) / self.total


def chamfer_3d_dist(xyz1, xyz2):
    """external/ChamferDistancePytorch chamfer_3DDist.forward for one cloud pair: xyz1 [n,3], xyz2 [m,3] ->
    (dist1 [n], dist2 [m], idx1 [n], idx2 [m])."""
    d1, i1 = ops.chamfer_nn(xyz1, xyz2)
    d2, i2 = ops.chamfer_nn(xyz2, xyz1)
    return d1, d2, i1, i2


class Chamfer3D(_Metric):
    """util/metrics.py:28-53.  An empty cloud yields a NaN mean in the reference, which it skips (:48); here the
    pair is skipped before the search."""

    def _reset_state(self):
        self.cd_sum = 0.0
        self.total = 0.0

    def update(self, preds, target):
        preds, target = preds.squeeze(1), target.squeeze(1)
        for ip in range(preds.shape[0]):
            points_pred = torch.nonzero(preds[ip], as_tuple=False).float()
            points_target = torch.nonzero(target[ip], as_tuple=False).float()
            if points_pred.shape[0] == 0 or points_target.shape[0] == 0:
                continue
            dist1, dist2, _, _ = chamfer_3d_dist(points_target, points_pred)
            self.cd_sum += float(dist1.mean() + dist2.mean())
            self.total += 1

    def compute(self):
        return torch.tensor(self.cd_sum, dtype=torch.float32) / self.total
