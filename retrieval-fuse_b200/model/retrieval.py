"""Patch encoders with the reference's class names, constructor signatures and
state_dict keys (model/retrieval.py:4-388), executed by rf_conv3d_fwd /
rf_mlp_encode_fwd.  The architectures come from one table instead of thirteen
hand-written classes; `layers` keeps the reference's ModuleList indices so
checkpoints load unchanged."""
import torch
from torch import nn

from .. import ops
from ._base import RfModule

# (channel multiple of nf, kernel, stride) per Conv3d; LeakyReLU(0.2) after each.
_CONV_SPECS = {
    "Patch32": [(1, 5, 1), (2, 3, 1), (4, 3, 2), (8, 3, 1), (8, 3, 2), (8, 4, 1)],        # :4
    "Patch08": [(1, 3, 1), (4, 3, 1), (4, 3, 1), (8, 2, 1)],                              # :136
    "PCPatch32": [(1, 3, 1), (2, 3, 1), (4, 3, 2), (4, 3, 1), (8, 3, 2), (8, 3, 1), (8, 3, 1)],  # :187
    "PCPatch48": [(1, 5, 1), (2, 3, 1), (4, 3, 2), (4, 3, 2), (8, 3, 2), (8, 3, 1), (8, 2, 1)],  # :217
    "PCPatch64": [(1, 5, 1), (2, 3, 1), (4, 3, 2), (4, 3, 2), (8, 3, 2), (8, 3, 1), (8, 4, 1)],  # :247
    "Patch16": [(1, 3, 1), (2, 3, 1), (2, 3, 1), (4, 3, 1), (4, 3, 1), (8, 3, 1), (8, 4, 1)],    # :277
    "Patch24": [(1, 5, 1), (2, 3, 1), (2, 3, 2), (4, 3, 1), (8, 3, 1), (8, 3, 1), (8, 2, 1)],    # :306
    "Patch24V2": [(1, 3, 1), (2, 3, 1), (2, 3, 2), (4, 3, 1), (8, 3, 1), (8, 3, 1), (8, 3, 1)],  # :335
    "Patch12": [(1, 3, 1), (2, 3, 1), (4, 3, 1), (4, 3, 1), (8, 3, 1), (8, 2, 1)],               # :364
}
_CONV_SPECS["PatchNorm32"] = _CONV_SPECS["Patch32"]  # :31 (+BatchNorm3d after every conv)
_CONV_SPECS["PatchNorm08"] = _CONV_SPECS["Patch08"]  # :160
_MLP_SPECS = {  # (input patch edge, hidden widths as multiples of nf)
    "Patch04": (4, [4, 8, 16, 8]),       # :64
    "Patch05": (5, [4, 8, 16, 8]),       # :87
    "Patch04V2": (4, [4, 8, 16, 16, 8]),  # :110
}


class _ConvPatchEncoder(RfModule):
    _spec = None
    _batchnorm = False

    def __init__(self, nf, z_dim):
        super().__init__()
        mods, cin = [], 1
        for mult, k, s in self._spec:
            mods.append(nn.Conv3d(cin, mult * nf, kernel_size=k, stride=s, padding=0))
            if self._batchnorm:
                mods.append(nn.BatchNorm3d(mult * nf))
            mods.append(nn.LeakyReLU(0.2, inplace=True))
            cin = mult * nf
        self.layers = nn.ModuleList(mods)  # parameter containers; never called
        self.final_layer = nn.Linear(cin, z_dim)

    use_tensor_cores = True  # tcgen05 implicit-GEMM convs on channels-last fp16-split activations
    use_halo_conv = True     # stride-1 3x3x3 layers through the shifted-window kernel (rf_tc_conv_halo.cu)
    use_wrun_conv = True     # single-channel first layer (3^3 / 5^3) through the same kernel on W-run operand planes

    def _forward_tc(self, x):
        """Channels-last tensor-core path: split -> tc_conv3d (bias + LeakyReLU fused) per layer."""
        convs = [m for m in self.layers if isinstance(m, nn.Conv3d)]
        h = ops.cl_from_ncdhw(x)
        for li, (conv, (_mult, k, s)) in enumerate(zip(convs, self._spec)):
            cin = conv.in_channels
            if (cin == 1 and s == 1 and self.use_wrun_conv and (k == 5 or conv.out_channels >= 16) and
                    ops.tc_conv_wrun_supported(h.shape[0], h.shape[1], h.shape[2], h.shape[3], conv.out_channels, k, 0)):
                # first layer on tensor cores: W-run operand planes, one K chunk = all kw taps of a (kd,kh) line
                img, sw = self._wcache.derived(("wrun", li), [conv.weight], ops.tc_conv_wrun_weight_image)
                h = ops.tc_conv3d_wrun(h, img, conv.bias, conv.out_channels, k, pad=0, act=ops.ACT_LEAKY, slope=0.2, out_scale=1.0 / sw)
                continue
            if cin == 1 and conv.out_channels <= 32:  # first layer: direct convolution on the raw single-channel patch
                h = ops.conv3d_cin1_cl(h, conv.weight, conv.bias, None, ks=k, stride=s, pad=0, act=ops.ACT_LEAKY, slope=0.2)
                continue
            if (self.use_halo_conv and k == 3 and s == 1 and
                    ops.tc_conv_halo_supported(h.shape[0], h.shape[1], h.shape[2], h.shape[3], conv.out_channels, cin, 0, pad=0)):
                # 'valid' 3x3x3 layers: shifted-window kernel, the patch block staged in shared memory once (small Cout:
                # its W-pair variant, two output voxels per GEMM row, where the kernel's cost model prefers it)
                wp = ops.tc_conv_halo_wp_wanted(h.shape[0], h.shape[1], h.shape[2], h.shape[3], conv.out_channels, cin, 0, pad=0)
                img, sw = self._wcache.derived(("halo", li, wp), [conv.weight],
                                               lambda w, c=cin, q=wp: ops.tc_conv_halo_weight_image(w, c, 0, wp=q))
                h = ops.tc_conv3d_halo(ops.cl_norm_split_halo(h, None, None, scale=1.0, pad=0, wp=wp), img, conv.bias, conv.out_channels,
                                       act=ops.ACT_LEAKY, slope=0.2, out_scale=1.0 / sw)
                continue
            # (measured, 1 024 patches: 64 -> 64 @ 20^3 1.31 ms against 1.49 for the gathering kernel, 32 -> 64 @ 42^3 13.5 against 9.8 -
            # the parity planes of a large input cost more than the gather saves, so only small extents take this path)
            if (self.use_halo_conv and k == 3 and s == 2 and h.shape[1] <= 24 and
                    ops.tc_conv_halo_s2_supported(h.shape[0], h.shape[1], h.shape[2], h.shape[3], conv.out_channels, cin)):
                # stride-2 'valid' 3x3x3 layers: same kernel, the block staged as its 8 parity sub-blocks
                img, sw = self._wcache.derived(("halo", li), [conv.weight], lambda w, c=cin: ops.tc_conv_halo_weight_image(w, c, 0))
                h = ops.tc_conv3d_halo_s2(ops.cl_split_parity_planes(h), img, conv.bias, conv.out_channels,
                                          act=ops.ACT_LEAKY, slope=0.2, out_scale=1.0 / sw)
                continue
            img, sw = self._wcache.derived(("tcconv", li), [conv.weight], lambda w, c=cin: ops.tc_conv_weight_image(w, c, 0))
            h = ops.tc_conv3d(ops.cl_norm_split(h), None, cin, 0, img, conv.bias, conv.out_channels, k, stride=s, pad=0,
                              act=ops.ACT_LEAKY, slope=0.2, out_scale=1.0 / sw)
        h = h.reshape(h.shape[0], -1)  # spatial extent is 1^3 here
        fl = self.final_layer
        if h.shape[0] >= 128 and ops.tc_supported(*fl.weight.shape):
            z = ops.tc_linear(h, self._wcache.derived("tcfinal", [fl.weight], ops.tc_weight_image), fl.bias, fl.out_features)
        else:
            z = ops.linear(h, self._wt(fl.weight), fl.bias)
        return z.reshape(z.shape[0], z.shape[1], 1, 1, 1)

    def forward(self, x):
        ops._forward_only(x, *self.parameters())
        if self._batchnorm and self.training:
            raise NotImplementedError("PatchNorm* encoders run in eval mode only (running statistics)")
        if (self.use_tensor_cores and not self._batchnorm and x.is_cuda and x.shape[0] >= 8 and
                all(ops.tc_conv_supported(m.out_channels, m.in_channels, 0, m.kernel_size[0])
                    for m in self.layers if isinstance(m, nn.Conv3d))):
            return self._forward_tc(x)
        h = x
        i = 0
        for _mult, k, s in self._spec:
            conv = self.layers[i]
            i += 1
            oscale = oshift = None
            if self._batchnorm:
                bn = self.layers[i]
                i += 1
                oscale, oshift = self._wcache.derived(
                    ("bn", i), [bn.weight, bn.bias, bn.running_mean, bn.running_var],
                    lambda w, b, m, v, eps=bn.eps: ((w / torch.sqrt(v + eps)).contiguous(),
                                                    (b - m * w / torch.sqrt(v + eps)).contiguous()))
            i += 1  # LeakyReLU slot
            h = ops.conv3d(h, self._wt(conv.weight), conv.bias, cout=conv.out_channels, ks=k, stride=s, pad=0,
                           act=ops.ACT_LEAKY, slope=0.2, oscale=oscale, oshift=oshift)
        h = h.reshape(h.shape[0], -1)  # spatial is 1^3 here (squeeze x3 in the reference)
        z = ops.linear(h, self._wt(self.final_layer.weight), self.final_layer.bias)
        return z.reshape(z.shape[0], z.shape[1], 1, 1, 1)


class _MlpPatchEncoder(RfModule):
    _spec = None

    def __init__(self, nf, z_dim):
        super().__init__()
        edge, hidden = self._spec
        widths = [edge ** 3] + [h * nf for h in hidden] + [z_dim]
        mods = []
        for j in range(len(widths) - 1):
            mods.append(nn.Linear(widths[j], widths[j + 1]))
            if j < len(widths) - 2:
                mods.append(nn.ReLU())
        self.layers = nn.ModuleList(mods)

    def _linears(self):
        return [m for m in self.layers if isinstance(m, nn.Linear)]

    def forward(self, x):
        return self.encode(x, l2_normalize=False)

    use_tensor_cores = True  # tcgen05 fp16-split GEMMs (~2e-7 relative); False -> fp32 FMA kernels
    use_fused_chain = True   # one launch for the whole MLP (rf_tc_mlp_fwd) instead of one rf_tc_linear_fwd per layer

    def encode(self, x, l2_normalize):
        """forward (+ optionally util/retrieval.py:66 row normalisation fused)."""
        ops._forward_only(x, *self.parameters())
        lin = self._linears()
        h = x.reshape(x.shape[0], -1)
        widths = [lin[0].in_features] + [m.out_features for m in lin]
        if self.use_tensor_cores and self.use_fused_chain and h.shape[0] >= 128 and ops.tc_mlp_supported(widths):
            # all layers in one launch: hidden activations stay in shared memory / TMEM (rf_tc_mlp.cu)
            imgs = [self._wcache.derived(("mlpimg", j), [m.weight], ops.tc_mlp_weight_image) for j, m in enumerate(lin)]
            z = ops.tc_mlp(h, imgs, [m.bias for m in lin], widths, act=ops.ACT_RELU, l2_normalize=l2_normalize)
        elif self.use_tensor_cores and h.shape[0] >= 128 and all(ops.tc_supported(*m.weight.shape) for m in lin):
            for j, m in enumerate(lin):
                img = self._wcache.derived(("tcimg", j), [m.weight], ops.tc_weight_image)
                h = ops.tc_linear(h, img, m.bias, m.out_features, act=ops.ACT_RELU if j < len(lin) - 1 else ops.ACT_NONE)
            z = ops.l2_normalize_rows(h) if l2_normalize else h
        else:
            z = ops.mlp_encode(h, [self._wt(m.weight) for m in lin], [m.bias for m in lin], l2_normalize=l2_normalize)
        return z.reshape(z.shape[0], z.shape[1], 1, 1, 1)


def _make(name):
    if name in _MLP_SPECS:
        return type(name, (_MlpPatchEncoder,), {"_spec": _MLP_SPECS[name], "__doc__": f"model/retrieval.py {name}"})
    return type(name, (_ConvPatchEncoder,), {"_spec": _CONV_SPECS[name], "_batchnorm": name.startswith("PatchNorm"),
                                             "__doc__": f"model/retrieval.py {name}"})


Patch32 = _make("Patch32")
PatchNorm32 = _make("PatchNorm32")
Patch04 = _make("Patch04")
Patch05 = _make("Patch05")
Patch04V2 = _make("Patch04V2")
Patch08 = _make("Patch08")
PatchNorm08 = _make("PatchNorm08")
PCPatch32 = _make("PCPatch32")
PCPatch48 = _make("PCPatch48")
PCPatch64 = _make("PCPatch64")
Patch16 = _make("Patch16")
Patch24 = _make("Patch24")
Patch24V2 = _make("Patch24V2")
Patch12 = _make("Patch12")
