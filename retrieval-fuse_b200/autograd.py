"""Differentiable path of the drop-in modules.

`trainer/train_refinement.py:74-89` (training_step_full) back-propagates a loss on `pred_shape` through the final
decoder, the patch attention, the retrieval U-Net, the input U-Net and the Fold3D / Unfold3D re-indexing.  When
autograd is recording, the modules in `model/` route their forward through the `torch.autograd.Function`s below:
the forward halves are the fp32 NCDHW kernels of the C ABI (rf_conv3d_fwd, rf_groupnorm_stats, rf_maxpool3d_2,
rf_linear_fwd, rf_attention_epilogue_fwd, rf_unfold3d / rf_fold3d), the backward halves their hand-written adjoints
(csrc/rf_backward.cu; a convolution's input gradient is the forward kernel on the flipped, transposed filter).
torch is used for the graph bookkeeping and for re-laying out the (small) weight tensors only.
Inference (`torch.no_grad()`) keeps using the tcgen05 kernels; nothing here is on that path.
"""
import torch

from . import ops


def _wt(weight):
    """[Cout, Cin, k,k,k] (or [N, K]) -> the [Cin*k^3, Cout] GEMM layout rf_conv3d_fwd / rf_linear_fwd read."""
    return weight.detach().reshape(weight.shape[0], -1).t().contiguous()


class ConvGnAct(torch.autograd.Function):
    """model/unet.py:19-100 SingleConv as ONE differentiable op: [GroupNorm ->] Conv3d(k, stride 1, pad) [+ bias]
    -> activation, on the virtual input concat(x, nearest_up2(x2)) (either may be None)."""

    @staticmethod
    def forward(ctx, x, x2, weight, bias, gn_weight, gn_bias, groups, eps, pad, act, slope):
        ks, cout = weight.shape[2], weight.shape[0]
        mu = rstd = None
        gn = None
        if gn_weight is not None:
            # gamma = 1 makes the statistics kernel return (mean, 1/sqrt(var + eps)); the per-channel scale a = rstd * gamma
            mu, rstd = ops.groupnorm_stats(x, torch.ones_like(gn_weight), groups, eps, x2=x2)
            gn = (mu, (rstd * gn_weight.detach()[None, :]).contiguous(), gn_bias.detach())
        y = ops.conv3d(x, _wt(weight), None if bias is None else bias.detach(), cout=cout, ks=ks, stride=1, pad=pad, act=act,
                       slope=slope, x2=x2, gn=gn)
        ctx.save_for_backward(x, x2, weight, bias, gn_weight, gn_bias, mu, rstd, y)
        ctx.cfg = (groups, pad, act, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x2, weight, bias, gn_weight, gn_bias, mu, rstd, y = ctx.saved_tensors
        groups, pad, act, slope = ctx.cfg
        ks, cin = weight.shape[2], weight.shape[1]
        c1 = x.shape[1] if x is not None else 0
        dz = ops.act_bwd(dy.contiguous(), y, act, slope)
        gn = None
        if gn_weight is not None:
            gn = (mu, (rstd * gn_weight.detach()[None, :]).contiguous(), gn_bias.detach())
        d_w = ops.conv3d_wgrad(x, x2, gn, dz, ks, stride=1, pad=pad) if ctx.needs_input_grad[2] else None
        d_b = ops.channel_sum(dz) if (bias is not None and ctx.needs_input_grad[3]) else None
        dx = dx2 = d_gw = d_gb = None
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if need_in or gn_weight is not None:
            # gradient wrt the (normalised) input: the forward kernel on the flipped filter with in / out swapped
            wflip = weight.detach().flip(2, 3, 4).transpose(0, 1).contiguous()  # [Cin, Cout, k,k,k]
            g = ops.conv3d(dz, _wt(wflip), None, cout=cin, ks=ks, stride=1, pad=ks - 1 - pad)
            if gn_weight is not None:
                dx, dx2, d_gw, d_gb = ops.gn_bwd(x, x2, g, mu, rstd, gn_weight.detach(), groups)
            else:
                dx = g[:, :c1].contiguous() if x is not None else None
                dx2 = ops.upsample2_bwd(g, c1) if x2 is not None else None
        return dx, dx2, d_w, d_b, d_gw, d_gb, None, None, None, None, None


class MaxPool3d2(torch.autograd.Function):
    """nn.MaxPool3d(2) (model/unet.py:230)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool3d_2(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.maxpool3d_2_bwd(x, dy.contiguous())


class LinearAct(torch.autograd.Function):
    """nn.Linear followed by an activation (model/attention.py:35-41)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, slope):
        y = ops.linear(x, _wt(weight), bias.detach(), act, slope)
        ctx.save_for_backward(x, weight, y)
        ctx.cfg = (act, slope)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        act, slope = ctx.cfg
        dz = ops.act_bwd(dy.contiguous(), y, act, slope)
        M, K = x.shape
        dx = ops.linear(dz, weight.detach().contiguous(), None) if ctx.needs_input_grad[0] else None  # dz @ W
        d_w = ops.conv3d_wgrad(x.reshape(M, K, 1, 1, 1), None, None, dz.reshape(M, -1, 1, 1, 1), 1).reshape(weight.shape)
        d_b = ops.channel_sum(dz)
        return dx, d_w, d_b, None, None


class Unfold3DFn(torch.autograd.Function):
    """model/attention.py:186-188; the adjoint of an unfold is the fold."""

    @staticmethod
    def forward(ctx, x, E):
        ctx.cfg = (x.shape[1], x.shape[2] // E, E)
        return ops.unfold3d(x, E)

    @staticmethod
    def backward(ctx, dy):
        C, R, E = ctx.cfg
        return ops.fold3d(dy.contiguous(), R, E, C), None


class Fold3DFn(torch.autograd.Function):
    """model/attention.py:170-176."""

    @staticmethod
    def forward(ctx, x, R, E, nf):
        ctx.cfg = (E, x.shape)
        return ops.fold3d(x, R, E, nf)

    @staticmethod
    def backward(ctx, dy):
        E, shape = ctx.cfg
        return ops.unfold3d(dy.contiguous(), E).reshape(shape), None, None, None


class AttentionEpilogue(torch.autograd.Function):
    """model/attention.py:92-113 on the theta / phi features: normalise, scores, ReLU-max switch, softmax(1024 s) or
    hard Gumbel (straight-through), weighted sum of the candidates, blend."""

    @staticmethod
    def forward(ctx, xf, pf, xu, pu, noise, rp3, K, normalize, mode, blend, sharp):
        ctx.save_for_backward(xf, pf, xu, pu, noise)
        ctx.cfg = (rp3, K, normalize, mode, blend, sharp)
        return ops.attention_epilogue(xf, pf, xu, pu, noise, rp3, K, normalize, mode, blend, sharp)

    @staticmethod
    def backward(ctx, dout):
        xf, pf, xu, pu, noise = ctx.saved_tensors
        dxf, dpf, dxu, dpu = ops.attention_epilogue_bwd(xf, pf, xu, pu, noise, dout.contiguous(), *ctx.cfg)
        return dxf, dpf, dxu, dpu, None, None, None, None, None, None, None


def single_conv(module, x, x2=None):
    """Differentiable forward of model.unet.SingleConv (NCDHW)."""
    g = module.groupnorm if "g" in module.order else None
    return ConvGnAct.apply(x, x2, module.conv.weight, module.conv.bias, None if g is None else g.weight,
                           None if g is None else g.bias, 0 if g is None else g.num_groups, 0.0 if g is None else g.eps,
                           module.padding, module.act, 0.1)


def attention_mlp(encoder, x):
    """AttentionFeatureEncoder (model/attention.py:29-46): four Linear layers, LeakyReLU(0.01) between them."""
    lin = encoder.linears()
    for i, m in enumerate(lin):
        last = i == len(lin) - 1
        x = LinearAct.apply(x, m.weight, m.bias, ops.ACT_NONE if last else ops.ACT_LEAKY, 0.01)
    return x


def patched_attention(block, x_predicted, x_retrieved, gumbel_noise=None):
    """Differentiable forward of PatchedAttentionBlock (model/attention.py:141-157)."""
    ab = block.attention_blocks_layer
    if ab.output_mapping() is not None:
        raise NotImplementedError("training through attn_no_output_mapping=False (the g / o 1x1x1 convolutions) is not "
                                  "implemented; the forward is (rf_attention_fuse_patched_fwd)")
    E, K, nf = block.patch_extent, block.num_nearest_neighbors, block.nf
    B, S = x_predicted.shape[0], x_predicted.shape[2]
    Rp = S // E
    rp3, V = Rp ** 3, nf * E ** 3
    xu = Unfold3DFn.apply(x_predicted.contiguous(), E).reshape(B * rp3, V)
    pu = Unfold3DFn.apply(x_retrieved.reshape(B * K, nf, S, S, S).contiguous(), E).reshape(B * K * rp3, V)
    xf = attention_mlp(ab.theta, xu)
    pf = attention_mlp(ab.phi, pu)
    mode = 1 if ab.retrieval_mode else 0
    if mode == 1 and gumbel_noise is None:
        gumbel_noise = -torch.empty(B * rp3, K, device=xu.device, dtype=torch.float32).exponential_().log()
    sharp = float(ab.cf_feat * E ** 3 * 4)
    rows = AttentionEpilogue.apply(xf, pf, xu, pu, gumbel_noise, rp3, K, ab.normalize, mode, ab.blend_mode, sharp)
    return Fold3DFn.apply(rows.reshape(B * rp3, nf, E, E, E), Rp, E, nf)
