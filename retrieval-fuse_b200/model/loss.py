"""model/loss.py of the reference, forward only, on the GPU."""
import torch

from .. import ops


class NTXentLoss(torch.nn.Module):
    """model/loss.py:5-69.  Same constructor and call signature; the 2N x 2N similarity matrix is never
    materialised (rf_ntxent_fwd).  Forward only: the reference trains through this loss, this build evaluates it
    (validation logging, train_refinement.py:125)."""

    def __init__(self, temperature, use_cosine_similarity, sig_scale=80, sig_shift=-65):
        super().__init__()
        self.temperature = temperature
        self.use_cosine_similarity = bool(use_cosine_similarity)
        self.sig_scale = sig_scale
        self.sig_shift = sig_shift

    def forward(self, zis, zjs, iou_matrix=None):
        return ops.ntxent(zis, zjs, self.temperature, cosine=self.use_cosine_similarity, iou_matrix=iou_matrix,
                          sig_scale=self.sig_scale, sig_shift=self.sig_shift)


def compute_sliced_attn_nt_xent_loss(loss_ntxent, batch_size, x_attn_fpred, x_attn_ftgt, occupancy_attn, budget=1280):
    """trainer/train_refinement.py:208-221: the loss over the occupied sub-patches of each slice, slices taken in
    order while the running count of occupied rows stays within `budget`."""
    split_size = x_attn_fpred.shape[0] // batch_size
    total = 0
    loss = torch.zeros(1, dtype=torch.float32, device=x_attn_fpred.device)
    occ_counts = (occupancy_attn.reshape(batch_size, split_size) > 0).sum(1).tolist()  # one D2H copy for the whole batch
    for b in range(batch_size):
        n = int(occ_counts[b])
        if n > 0 and total + n <= budget:
            sl = slice(b * split_size, (b + 1) * split_size)
            b_occ = occupancy_attn[sl] > 0
            loss = loss_ntxent(x_attn_fpred[sl][b_occ], x_attn_ftgt[sl][b_occ]) + loss
            total += n
    return loss
