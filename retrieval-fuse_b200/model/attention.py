"""model/attention.py: 3-D fold / unfold modules and the patch attention block
with the reference's names, signatures and state_dict keys."""
import torch
from torch import nn

from .. import ops
from ._base import RfModule


class Conv3dAttentionOutput(nn.Conv3d):
    """model/attention.py:5-15: the g / o 1x1x1 convolutions of attn_no_output_mapping=False (near-identity init)."""

    def __init__(self, nf_in, nf_out):
        super().__init__(nf_in, nf_out, kernel_size=1, stride=1, padding=0)

    def reset_parameters(self) -> None:
        nn.init.dirac_(self.weight)
        with torch.no_grad():
            self.weight[:] = self.weight[:] + torch.randn_like(self.weight[:]) * 0.01
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class Conv3dAttentionFeature(nn.Conv3d):
    """model/attention.py:18-26 (defined by the reference, used by none of its modules)."""

    def __init__(self, nf_in, nf_out):
        super().__init__(nf_in, nf_out, kernel_size=1, stride=1, padding=0)

    def reset_parameters(self) -> None:
        nn.init.normal_(self.weight, 0, 0.01)
        if self.bias is not None:
            nn.init.zeros_(self.bias)


class AttentionFeatureEncoder(nn.Module):
    """model/attention.py:29-46 (parameter container; the MLP runs inside
    rf_attention_fuse_fwd / rf_attention_features)."""

    def __init__(self, n_in, n_out, e):
        super().__init__()
        self.n_in = n_in * (e ** 3)
        self.n_out = n_out
        self.encoder = nn.Sequential(nn.Linear(self.n_in, 128), nn.LeakyReLU(), nn.Linear(128, 128), nn.LeakyReLU(),
                                     nn.Linear(128, 128), nn.LeakyReLU(), nn.Linear(128, self.n_out))

    def linears(self):
        return [m for m in self.encoder if isinstance(m, nn.Linear)]


class AttentionBlock(RfModule):
    """model/attention.py:49-116."""

    def __init__(self, num_output_channels, patch_extent, K, normalize, use_switching, retrieval_mode,
                 no_output_mapping, blend):
        super().__init__()
        self.cf_op = num_output_channels
        self.cf_feat = 32
        self.patch_extent = patch_extent
        self.K = K
        self.theta = AttentionFeatureEncoder(num_output_channels, self.cf_feat, patch_extent)
        self.phi = AttentionFeatureEncoder(num_output_channels, self.cf_feat, patch_extent)
        self.g = Conv3dAttentionOutput(num_output_channels, self.cf_op) if not no_output_mapping else nn.Identity()
        self.o = Conv3dAttentionOutput(self.cf_op, num_output_channels) if not no_output_mapping else nn.Identity()
        self.init_scale = 35
        self.init_shift = -27
        # unused in the reference's forward (:97-99) but part of its state_dict
        self.sig_scale = nn.Parameter(torch.ones(1) * self.init_scale)
        self.sig_shift = nn.Parameter(torch.ones(1) * self.init_shift)
        self.retrieval_mode = retrieval_mode
        self.blend_mode = blend
        self.use_switching = use_switching
        self.normalize = normalize

    use_tensor_cores = True  # theta / phi MLPs as tcgen05 fp16-split GEMMs; False -> fp32 FMA kernels

    def _branch(self, enc):
        lin = enc.linears()
        imgs = None
        widths = [lin[0].in_features] + [m.out_features for m in lin]
        if self.use_tensor_cores and len(lin) == 4 and ops.tc_mlp_supported(widths):
            tag = "theta" if enc is self.theta else "phi"
            imgs = [self._wcache.derived((tag, j), [m.weight], ops.tc_mlp_weight_image) for j, m in enumerate(lin)]
        return [self._wt(m.weight) for m in lin], [m.bias for m in lin], imgs

    def output_mapping(self):
        """None for g = o = Identity; else (Wo Wg [nf,nf], Wo bg [nf], bo [nf]): both 1x1x1 convolutions are linear and act
        per voxel, so o(sum_k w_k g(p_k)) = (Wo Wg) sum_k w_k p_k + (Wo bg) sum_k w_k + bo (model/attention.py:95,108)."""
        if isinstance(self.g, nn.Identity):
            return None
        if self.retrieval_mode:
            # the reference cannot run this combination either (its forward raises): in retrieval mode the weighted sum
            # stays 2-D (model/attention.py:103, no reshape) and o = Conv3d rejects it
            raise ValueError("attn_no_output_mapping=False has no retrieval (Gumbel) mode in the reference (model/attention.py:103)")

        def compose(wg, bg, wo, bo):
            wg2, wo2 = wg.detach().float().flatten(1), wo.detach().float().flatten(1)
            return ((wo2 @ wg2).contiguous(), (wo2 @ bg.detach().float()).contiguous(), bo.detach().float().contiguous())
        return self._wcache.derived(("output_mapping",), [self.g.weight, self.g.bias, self.o.weight, self.o.bias], compose)

    def forward(self, x, p, gumbel_noise=None):
        """model/attention.py:84-113.  x [b, C, E,E,E] predicted sub-patches, p [b, K, C, E,E,E] their K retrieved
        candidates -> [b, C, E,E,E].  Every sub-patch is handed to rf_attention_fuse_fwd as a volume of edge E (one
        sub-patch per volume), which is exactly the computation PatchedAttentionBlock runs on unfolded volumes."""
        ops._forward_only(x, p, *self.parameters())
        b, k, c, e = p.shape[0], p.shape[1], p.shape[2], p.shape[3]
        if k != self.K:
            raise ValueError(f"AttentionBlock was built for K = {self.K} retrievals, got {k} (MaxPool1d(kernel_size=K), :59)")
        if b == 0:
            return x.new_empty(x.shape)
        mode = 1 if self.retrieval_mode else 0
        if mode == 1 and gumbel_noise is None:
            gumbel_noise = -torch.empty(b, k, device=x.device, dtype=torch.float32).exponential_().log()
        return ops.attention_fuse(x.reshape(b, c, e, e, e), p.reshape(b * k, c, e, e, e), self._branch(self.theta),
                                  self._branch(self.phi), e, k, normalize=self.normalize, mode=mode, blend=self.blend_mode,
                                  gumbel_noise=gumbel_noise, output_mapping=self.output_mapping())

    def get_regularization_losses(self):
        return ((self.sig_scale - self.init_scale) ** 2 + (self.sig_shift - self.init_shift) ** 2) if self.use_switching else 0


class PatchedAttentionBlock(nn.Module):
    """model/attention.py:119-157."""

    def __init__(self, nf, num_patch_x, patch_extent, num_nearest_neighbors, attention_block):
        super().__init__()
        self.num_patch_x = num_patch_x
        self.patch_extent = patch_extent
        self.num_nearest_neighbors = num_nearest_neighbors
        self.nf = nf
        self.attention_blocks_layer = attention_block
        self.fold_3d = Fold3D(num_patch_x, patch_extent, self.nf)
        self.unfold_3d = Unfold3D(patch_extent, self.nf)
        self.unfold_3d_occ = Unfold3D(patch_extent, 1)

    def get_features(self, x_predicted, x_target, occupancy):
        ab = self.attention_blocks_layer
        ops._forward_only(x_predicted, x_target, *ab.parameters())
        return ops.attention_features(x_predicted, x_target, occupancy, ab._branch(ab.theta), ab._branch(ab.phi),
                                      self.patch_extent, normalize=ab.normalize)

    def forward(self, x_predicted, x_retrieved, gumbel_noise=None, patch_grid=1, out_channels_last=False):
        """x_predicted [B,F,S,S,S], x_retrieved [B*K,F,S,S,S] -> [B,F,S,S,S].
        In retrieval (Gumbel) mode the noise [B*R^3, K] may be injected; when it
        is not, it is drawn on the device as torch's gumbel_softmax does.
        Inference shortcuts (not in the reference's signature): patch_grid = P > 1 takes the retrieval U-Net's
        un-folded patches [B*K*P^3,F,S/P,S/P,S/P] (the Fold3D of train_refinement.py:112 becomes index arithmetic),
        out_channels_last returns [B,S,S,S,F] for the decoder's channels-last path."""
        ab = self.attention_blocks_layer
        K = self.num_nearest_neighbors
        if x_predicted.shape[0] == 0:  # empty batch: what torch's ops would return
            S = x_predicted.shape[2]
            return x_predicted.new_empty((0, S, S, S, self.nf) if out_channels_last else x_predicted.shape)
        if patch_grid > 1 or out_channels_last:
            if ops.grad_needed(x_predicted, x_retrieved, *ab.parameters()):
                raise ValueError("patch_grid / out_channels_last are inference shortcuts; fold the patches for training")
            if x_retrieved.shape[0] != x_predicted.shape[0] * K * patch_grid ** 3:
                raise ValueError(f"x_retrieved has {x_retrieved.shape[0]} patches, expected B*K*P^3")
            mode = 1 if ab.retrieval_mode else 0
            if mode == 1 and gumbel_noise is None:
                rows = x_predicted.shape[0] * self.num_patch_x ** 3
                gumbel_noise = -torch.empty(rows, K, device=x_predicted.device, dtype=torch.float32).exponential_().log()
            if out_channels_last and ab.output_mapping() is not None:
                raise ValueError("out_channels_last is not available with attn_no_output_mapping=False")
            return ops.attention_fuse(x_predicted, x_retrieved, ab._branch(ab.theta), ab._branch(ab.phi), self.patch_extent, K,
                                      normalize=ab.normalize, mode=mode, blend=ab.blend_mode, gumbel_noise=gumbel_noise,
                                      patch_grid=patch_grid, out_channels_last=out_channels_last,
                                      output_mapping=ab.output_mapping())
        if x_retrieved.shape[0] != x_predicted.shape[0] * K:
            raise ValueError(f"x_retrieved has {x_retrieved.shape[0]} volumes, expected B*K = {x_predicted.shape[0] * K}")
        if x_predicted.shape[2] != self.num_patch_x * self.patch_extent:
            raise ValueError("feature volume edge must equal attn_num_patch * patch_extent")
        if ops.grad_needed(x_predicted, x_retrieved, *ab.parameters()):  # training: differentiable composition
            from ..autograd import patched_attention
            return patched_attention(self, x_predicted, x_retrieved, gumbel_noise)
        mode = 1 if ab.retrieval_mode else 0
        if mode == 1 and gumbel_noise is None:
            rows = x_predicted.shape[0] * self.num_patch_x ** 3
            gumbel_noise = -torch.empty(rows, K, device=x_predicted.device, dtype=torch.float32).exponential_().log()
        return ops.attention_fuse(x_predicted, x_retrieved.reshape(-1, self.nf, *x_retrieved.shape[2:]),
                                  ab._branch(ab.theta), ab._branch(ab.phi), self.patch_extent, K,
                                  normalize=ab.normalize, mode=mode, blend=ab.blend_mode, gumbel_noise=gumbel_noise,
                                  output_mapping=ab.output_mapping())


class Fold3D(nn.Module):
    """model/attention.py:160-176."""

    def __init__(self, num_patch_x, patch_extent, nf):
        super().__init__()
        self.nf = nf
        self.num_patch_x = num_patch_x
        self.patch_extent = patch_extent

    def forward(self, x):
        if ops.grad_needed(x):
            from ..autograd import Fold3DFn
            return Fold3DFn.apply(x.contiguous(), self.num_patch_x, self.patch_extent, self.nf)
        return ops.fold3d(x, self.num_patch_x, self.patch_extent, self.nf)


class Unfold3D(nn.Module):
    """model/attention.py:179-188."""

    def __init__(self, patch_extent, nf):
        super().__init__()
        self.patch_extent = patch_extent
        self.nf = nf

    def forward(self, x):
        if x.shape[1] != self.nf:
            raise ValueError(f"Unfold3D was built for nf={self.nf}, input has {x.shape[1]} channels")
        if ops.grad_needed(x):
            from ..autograd import Unfold3DFn
            return Unfold3DFn.apply(x.contiguous(), self.patch_extent)
        return ops.unfold3d(x, self.patch_extent)


class Unfold3DPadStride(nn.Module):
    """model/attention.py:191-203."""

    def __init__(self, patch_extent, pad_size, pad_val, stride):
        super().__init__()
        self.patch_extent = patch_extent
        self.pad_size = pad_size
        self.pad_val = pad_val
        self.stride = stride

    def forward(self, x):
        return ops.unfold3d_pad_stride(x, self.patch_extent, self.pad_size, self.stride, self.pad_val)
