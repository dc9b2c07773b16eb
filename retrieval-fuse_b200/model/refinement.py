"""model/refinement.py:6-73 - the task-specific U-Net wirings."""
import torch
from torch import nn

from .. import ops
from ._base import RfModule
from .unet import UNet3D, DecoderNoJoining


class _Chain(nn.Module):
    def forward(self, x):
        from . import unet as U
        nets = list(self.network)
        training_path = ops.grad_needed(x, *self.parameters())  # autograd is recording: differentiable NCDHW path
        if not training_path and U.USE_TENSOR_CORES and x.is_cuda and nets[0].tc_ok(x.shape[1]) and all(
                n.basic_module.tc_ok(0, n.basic_module.SingleConv1.in_channels) for n in nets[1:]):
            h = nets[0].forward_cl(ops.cl_from_ncdhw(x))          # channels-last all the way through
            for i, n in enumerate(nets[1:]):
                h = n.forward_cl(h, out_ncdhw=i == len(nets) - 2)
            return h
        for net in nets:
            x = net(x)
        return x


class Superresolution08UNetBackbone(_Chain):
    """model/refinement.py:6-19: [B,1,8^3] -> [B,nf,32^3]."""

    def __init__(self, nf, num_levels, layer_order):
        super().__init__()
        self.network = nn.ModuleList([
            UNet3D(in_channels=1, out_channels=2 * nf, final_sigmoid=False, final_conv=False, f_maps=nf,
                   num_groups=nf // 2, layer_order=layer_order, num_levels=num_levels, is_segmentation=False),
            DecoderNoJoining(2 * nf, 2 * nf, conv_layer_order=layer_order, num_groups=nf // 2),
            DecoderNoJoining(2 * nf, nf, conv_layer_order=layer_order, num_groups=nf // 2),
        ])


class Superresolution16UNetBackbone(_Chain):
    """model/refinement.py:22-34: [B,1,16^3] -> [B,nf,32^3]."""

    def __init__(self, nf, num_levels, layer_order):
        super().__init__()
        self.network = nn.ModuleList([
            UNet3D(in_channels=1, out_channels=2 * nf, final_sigmoid=False, final_conv=False, f_maps=nf,
                   num_groups=nf // 2, layer_order=layer_order, num_levels=num_levels, is_segmentation=False),
            DecoderNoJoining(2 * nf, nf, conv_layer_order=layer_order, num_groups=nf // 2),
        ])


class SurfaceReconstructionUNetBackbone(nn.Module):
    """model/refinement.py:37-45: [B,1,128^3] -> [B,nf,32^3]."""

    def __init__(self, nf, num_levels, layer_order):
        super().__init__()
        self.network = UNet3D(in_channels=1, out_channels=nf, final_sigmoid=False, final_conv=False,
                              remove_n_final_layers=2, f_maps=nf, layer_order=layer_order, num_groups=nf // 2,
                              num_levels=num_levels, is_segmentation=False)

    def forward(self, x):
        return self.network(x)


class Superresolution08FinalDecoder(RfModule):
    """model/refinement.py:48-61: nearest x2 + DoubleConv, 1x1x1 conv + bias, tanh
    (bias and tanh are fused into the conv epilogue)."""

    def __init__(self, nf, layer_order):
        super().__init__()
        self.network = nn.ModuleList([
            DecoderNoJoining(nf, nf, conv_layer_order=layer_order, num_groups=nf // 2),
            nn.Conv3d(nf, 1, 1, padding=0),
            nn.Tanh(),
        ])

    def tc_path(self, nf):
        """True when forward() takes the channels-last tensor-core path (and therefore accepts channels-last input)."""
        from . import unet as U
        return bool(U.USE_TENSOR_CORES and self.network[0].basic_module.tc_ok(0, nf))

    def forward(self, x, channels_last_input=False):
        """channels_last_input: x is [B,S,S,S,nf] (the attention's channels-last result); inference only."""
        from . import unet as U
        head = self.network[1]
        nf = head.in_channels
        if channels_last_input:
            if not (x.is_cuda and self.tc_path(nf)) or ops.grad_needed(x, *self.parameters()):
                raise ValueError("channels-last decoder input needs the tensor-core inference path")
            h = self.network[0].forward_cl(x)
            return ops.cl_pointwise_head(h, head.weight, head.bias, act=ops.ACT_TANH)
        if ops.grad_needed(x, *self.parameters()):  # training: DoubleConv + (1x1x1 conv, bias, tanh) as differentiable ops
            from ..autograd import ConvGnAct
            h = self.network[0](x)
            return ConvGnAct.apply(h, None, head.weight, head.bias, None, None, 0, 0.0, 0, ops.ACT_TANH, 0.0)
        if U.USE_TENSOR_CORES and x.is_cuda and self.network[0].basic_module.tc_ok(0, nf):
            h = self.network[0].forward_cl(ops.cl_from_ncdhw(x))
            # the 1x1x1 head is memory-bound: one fp32 FMA pass over the channels-last volume, bias + tanh fused
            return ops.cl_pointwise_head(h, head.weight, head.bias, act=ops.ACT_TANH)
        x = self.network[0](x)
        return ops.conv3d(x, self._wt(head.weight), head.bias, cout=1, ks=1, stride=1, pad=0, act=ops.ACT_TANH)


class RetrievalUNetBackbone(nn.Module):
    """model/refinement.py:64-73: per 16^3 patch [N,1,16^3] -> [N,nf,8^3].
    (positional order is (f_maps, nf, ...) as in the reference; the factory
    passes keywords.)"""

    def __init__(self, f_maps, nf, num_levels, layer_order):
        super().__init__()
        self.nf = nf
        self.network = UNet3D(in_channels=1, out_channels=nf, num_groups=nf // 2, final_sigmoid=False,
                              final_conv=False, remove_n_final_layers=1, f_maps=f_maps, layer_order=layer_order,
                              num_levels=num_levels, is_segmentation=False)

    def forward(self, x):
        return self.network(x)
