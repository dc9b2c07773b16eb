"""Tensor-level wrappers over the C ABI (include/rf_b200.h).

torch is plumbing here: device memory, the current stream, dtype/shape checks.
Every function launches the hand-written sm_100a kernels through ctypes on
`torch.cuda.current_stream()`; CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes
import functools
import math
import os

import torch

from . import _lib
from ._lib import check, int3, ptr_array

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3

# launches issued through this module since the last reset (bench.py's gpu_launches)
_launches = 0
_LAUNCHES_PER_CALL = {}


def launches() -> int:
    return _launches


def reset_launches() -> None:
    global _launches
    _launches = 0


def _count(n=1):
    global _launches
    _launches += n


# ---- per-call device timing (bench.py's live roofline figures) --------------------------------
# profile_start() makes every op below bracket its launches with a CUDA event pair on the launching
# stream; profile_stop() returns {op name: {"calls", "ms", "flops", "bytes"}}.  Off by default (two event
# records per call); never on inside a timed region or a graph capture.
_prof = None


def profile_start():
    global _prof
    _prof = []


def profile_stop():
    global _prof
    recs, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, flops, nbytes in recs:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
        d["bytes"] += nbytes
    return out


class _timed:
    __slots__ = ("name", "flops", "bytes", "e0")

    def __init__(self, name, flops=0.0, nbytes=0.0):
        self.name, self.flops, self.bytes, self.e0 = name, flops, nbytes, None

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.name, self.e0, e1, self.flops, self.bytes))
        return False


# ---- persistent device buffers referenced by captured CUDA graphs -----------------------------
# The halo operand planes and the derived weight images are allocated once and reused across calls, so a captured
# graph holds raw pointers to them.  Whenever one of them is FREED (a plane set evicted, a weight image rebuilt after
# load_state_dict / an optimizer step) the generation changes and RefinementPipeline drops its graphs.
_generation = 0
MAX_PLANE_SHAPES = 4  # operand-plane sets kept per layer (one per input shape, least recently used evicted)


def persistent_generation() -> int:
    return _generation


def bump_generation() -> None:
    global _generation
    _generation += 1


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _dev(t: torch.Tensor, dtype=torch.float32, name="tensor") -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.RfError(f"{name} must be a CUDA tensor: the rf_b200 ops have no CPU path")
    if t.dtype != dtype:
        raise _lib.RfError(f"{name} must be {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def grad_needed(*tensors) -> bool:
    """True when autograd is recording and one of the tensors takes part in it: the modules then run their
    differentiable path (retrieval_fuse_b200.autograd: fp32 NCDHW kernels + the adjoint kernels of rf_backward.cu)."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def _forward_only(*tensors):
    if grad_needed(*tensors):
        raise NotImplementedError(
            "this rf_b200 op has no backward pass: call it under torch.no_grad().  Differentiable: the U-Nets, the patch "
            "attention, the final decoder, Fold3D / Unfold3D (what training_step_full back-propagates through)")


def _ptr(t):
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------
# a2-a4
# ---------------------------------------------------------------------------

def unfold3d(x: torch.Tensor, E: int) -> torch.Tensor:
    """model/attention.py:186-188."""
    _forward_only(x)
    x = _dev(x, name="x")
    B, C, S = x.shape[0], x.shape[1], x.shape[2]
    assert x.dim() == 5 and x.shape[3] == S and x.shape[4] == S, "Unfold3D expects a cubic [B,C,S,S,S] volume"
    R = S // E
    out = torch.empty((B * R * R * R, C, E, E, E), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device), _timed("rf_unfold3d", nbytes=2.0 * x.numel() * 4):
        check(_lib.lib().rf_unfold3d(x.data_ptr(), out.data_ptr(), B, C, S, E, _stream(x)), "rf_unfold3d")
    _count()
    return out


def fold3d(x: torch.Tensor, R: int, E: int, nf: int) -> torch.Tensor:
    """model/attention.py:170-176 (returns a contiguous tensor with the same values)."""
    _forward_only(x)
    x = _dev(x, name="x")
    per = R * R * R * nf * E * E * E
    assert x.numel() % per == 0, "Fold3D: input size does not match num_patch_x / patch_extent / nf"
    B = x.numel() // per
    out = torch.empty((B, nf, R * E, R * E, R * E), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device), _timed("rf_fold3d", nbytes=2.0 * x.numel() * 4):
        check(_lib.lib().rf_fold3d(x.data_ptr(), out.data_ptr(), B, nf, R, E, _stream(x)), "rf_fold3d")
    _count()
    return out


def unfold3d_pad_stride(x, kernel, pad, stride, pad_val, norm_sub=0.0, norm_div=0.0, keep_channels=False):
    """model/attention.py:200-203 / util/patcher.py:14-19. kernel/pad/stride: int or 3-sequence."""
    _forward_only(x)
    x = _dev(x, name="x")
    B, C = x.shape[0], x.shape[1]
    size = tuple(x.shape[2:])
    k, p, s = int3(kernel), int3(pad), int3(stride)
    cnt = [(size[a] + 2 * p[a] - k[a]) // s[a] + 1 for a in range(3)]
    rows = B * cnt[0] * cnt[1] * cnt[2]
    shape = (rows, C, k[0], k[1], k[2]) if keep_channels else (rows * C, 1, k[0], k[1], k[2])
    out = torch.empty(shape, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device), _timed("rf_unfold3d_pad_stride", nbytes=(x.numel() + out.numel()) * 4.0):
        check(_lib.lib().rf_unfold3d_pad_stride(x.data_ptr(), out.data_ptr(), B, C, int3(size), k, p, s, float(pad_val),
                                                float(norm_sub), float(norm_div), _stream(x)), "rf_unfold3d_pad_stride")
    _count()
    return out


def recompose_patches(patches, out_shape, kernel, pad, stride, count, pad_val):
    """util/patcher.py:21-30."""
    _forward_only(patches)
    patches = _dev(patches, name="patches")
    B, C = int(out_shape[0]), int(out_shape[1])
    size = tuple(int(v) for v in out_shape[2:])
    n = int(count[0]) * int(count[1]) * int(count[2])
    if patches.shape[1] != n:  # extra trailing patches are never read by the reference either
        patches = patches[:, :n].contiguous()
    out = torch.empty((B, C) + size, device=patches.device, dtype=patches.dtype)
    with torch.cuda.device(patches.device), _timed("rf_recompose_patches"):
        check(_lib.lib().rf_recompose_patches(patches.data_ptr(), out.data_ptr(), B, C, int3(size), int3(kernel),
                                              int3(pad), int3(stride), int3(count), float(pad_val), _stream(patches)),
              "rf_recompose_patches")
    _count()
    return out


# ---------------------------------------------------------------------------
# conv / linear / norm building blocks
# ---------------------------------------------------------------------------

def conv3d(x, wt, bias=None, *, cout, ks, stride=1, pad=0, act=ACT_NONE, slope=0.0, x2=None, gn=None, oscale=None,
           oshift=None):
    """rf_conv3d_fwd. x [N,C1,D,H,W] (or None when everything comes from x2),
    x2 optional [N,C2,D/2,H/2,W/2] (virtually upsampled + concatenated),
    wt [Cin*ks^3, cout]; gn = (gn_mu, gn_a, gn_beta) or None."""
    ref = x if x is not None else x2
    if x is not None:
        x = _dev(x, name="x")
        N, C1, D, H, W = x.shape
    else:
        C1 = 0
    C2 = 0
    if x2 is not None:
        x2 = _dev(x2, name="x2")
        C2 = x2.shape[1]
        if x is None:
            N, D, H, W = x2.shape[0], 2 * x2.shape[2], 2 * x2.shape[3], 2 * x2.shape[4]
        else:
            assert x2.shape[0] == N and tuple(x2.shape[2:]) == (D // 2, H // 2, W // 2), "x2 must be half resolution"
    cin = C1 + C2
    wt = _dev(wt, name="wt")
    assert tuple(wt.shape) == (cin * ks ** 3, cout), f"wt shape {tuple(wt.shape)} != {(cin * ks ** 3, cout)}"
    Do, Ho, Wo = [(v + 2 * pad - ks) // stride + 1 for v in (D, H, W)]
    y = torch.empty((N, cout, Do, Ho, Wo), device=ref.device, dtype=torch.float32)
    g = gn if gn is not None else (None, None, None)
    with torch.cuda.device(ref.device), _timed("rf_conv3d_fwd"):
        check(_lib.lib().rf_conv3d_fwd(_ptr(x), _ptr(x2), C2, wt.data_ptr(), _ptr(bias), _ptr(oscale), _ptr(oshift),
                                       _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), y.data_ptr(), N, cin, D, H, W, cout, ks,
                                       stride, pad, act, float(slope), _stream(ref)), "rf_conv3d_fwd")
    _count()
    return y


def linear(x, wt, bias=None, act=ACT_NONE, slope=0.0):
    """y = act(x @ wt + bias); wt [K, N] (= nn.Linear.weight.T)."""
    x = _dev(x, name="x")
    wt = _dev(wt, name="wt")
    M, K = x.shape
    assert wt.shape[0] == K
    N = wt.shape[1]
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_linear_fwd"):
        check(_lib.lib().rf_linear_fwd(x.data_ptr(), wt.data_ptr(), _ptr(bias), y.data_ptr(), M, K, N, act, float(slope),
                                       _stream(x)), "rf_linear_fwd")
    _count()
    return y


def tc_supported(N, K):
    """True when the tensor-core linear kernel handles an [N, K] weight."""
    return _lib.lib().rf_tc_weight_image_bytes(int(N), int(K)) > 0 and K % 4 == 0


def tc_weight_image(weight):
    """nn.Linear.weight [N, K] -> the pre-split, pre-swizzled fp16 operand image (uint8 tensor view, 1024-aligned)."""
    weight = _dev(weight.detach(), name="weight")
    N, K = weight.shape
    L = _lib.lib()
    nbytes = L.rf_tc_weight_image_bytes(N, K)
    if nbytes == 0:
        raise _lib.RfError(f"tensor-core linear does not support a weight of shape {(N, K)}")
    buf = torch.empty(nbytes + 1024, device=weight.device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % 1024
    img = buf[off: off + nbytes]
    with torch.cuda.device(weight.device), _timed("rf_tc_weight_image"):
        check(L.rf_tc_weight_image(weight.data_ptr(), N, K, img.data_ptr(), _stream(weight)), "rf_tc_weight_image")
    _count()
    return img


def tc_linear(x, weight_image, bias, N, act=ACT_NONE, slope=0.0):
    """y = act(x @ W^T + bias) on the tensor cores (fp16 hi/lo split, fp32 accumulate)."""
    x = _dev(x, name="x")
    M, K = x.shape
    y = torch.empty((M, N), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_tc_linear_fwd"):
        check(_lib.lib().rf_tc_linear_fwd(x.data_ptr(), K, weight_image.data_ptr(), _ptr(bias), y.data_ptr(), M, K, N, act,
                                          float(slope), _stream(x)), "rf_tc_linear_fwd")
    _count()
    return y


# ---- channels-last tensor-core convolution path ------------------------------------------------

def _widths_arr(widths):
    return (ctypes.c_int * 9)(*(list(widths) + [0] * (9 - len(widths))))


def tc_mlp_supported(widths):
    """widths = [in, hidden..., out]: can the fused tcgen05 MLP chain run these layers?"""
    return len(widths) <= 9 and bool(_lib.lib().rf_tc_mlp_supported(_widths_arr(widths), len(widths) - 1))


def tc_mlp_weight_image(weight):
    """nn.Linear weight [N, K] -> operand image of rf_tc_mlp_fwd."""
    weight = _dev(weight.detach(), name="weight")
    N, K = weight.shape
    L = _lib.lib()
    nbytes = L.rf_tc_mlp_weight_image_bytes(N, K)
    if nbytes == 0:
        raise _lib.RfError(f"fused MLP does not support a weight of shape {tuple(weight.shape)}")
    img = _aligned_bytes(nbytes, weight.device)
    with torch.cuda.device(weight.device), _timed("rf_tc_mlp_weight_image"):
        check(L.rf_tc_mlp_weight_image(weight.data_ptr(), N, K, img.data_ptr(), _stream(weight)), "rf_tc_mlp_weight_image")
    _count()
    return img


def tc_mlp(x, images, biases, widths, act=ACT_RELU, slope=0.0, l2_normalize=False, eps=1e-12):
    """All layers of an MLP in one launch (hidden activations stay on chip).  x [M, widths[0]] fp32 (row stride = its
    leading dimension), images[l] = tc_mlp_weight_image(layer l weight) -> [M, widths[-1]]."""
    x = _dev(x, name="x")
    M = x.shape[0]
    assert x.shape[1] == widths[0] and len(images) == len(widths) - 1
    y = torch.empty((M, widths[-1]), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_tc_mlp_fwd"):
        check(_lib.lib().rf_tc_mlp_fwd(x.data_ptr(), x.shape[1], ptr_array([i.data_ptr() for i in images]),
                                       ptr_array([_ptr(b) for b in biases]), _widths_arr(widths), len(widths) - 1, act,
                                       float(slope), int(bool(l2_normalize)), float(eps), y.data_ptr(), widths[-1], M,
                                       _stream(x)), "rf_tc_mlp_fwd")
    _count()
    return y


def _aligned_bytes(nbytes, device, align=1024):
    buf = torch.empty(nbytes + align, device=device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % align
    return buf[off: off + nbytes]


def cl_from_ncdhw(x):
    """[N,C,D,H,W] fp32 -> channels-last [N,D,H,W,C] (a view when C == 1)."""
    x = _dev(x, name="x")
    N, C, D, H, W = x.shape
    if C == 1:
        return x.reshape(N, D, H, W, 1)
    y = torch.empty((N, D, H, W, C), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_cl_transpose"):
        check(_lib.lib().rf_cl_transpose(x.data_ptr(), y.data_ptr(), N, D * H * W, C, 1, _stream(x)), "rf_cl_transpose")
    _count()
    return y


def cl_to_ncdhw(x):
    x = _dev(x, name="x")
    N, D, H, W, C = x.shape
    if C == 1:
        return x.reshape(N, 1, D, H, W)
    y = torch.empty((N, C, D, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_cl_transpose"):
        check(_lib.lib().rf_cl_transpose(x.data_ptr(), y.data_ptr(), N, D * H * W, C, 0, _stream(x)), "rf_cl_transpose")
    _count()
    return y


def cl_gn_stats(x, gamma, groups, eps=1e-5, x2=None):
    """GroupNorm statistics of the virtual channels-last input concat(x, up2(x2)) -> (mu, a) [N, C]."""
    x = _dev(x, name="x")
    N, D, H, W, C1 = x.shape
    C2 = 0
    if x2 is not None:
        x2 = _dev(x2, name="x2")
        C2 = x2.shape[-1]
    C = C1 + C2
    mu = torch.empty((N, C), device=x.device, dtype=torch.float32)
    a = torch.empty((N, C), device=x.device, dtype=torch.float32)
    ws = torch.empty((N, C, 2), device=x.device, dtype=torch.float64)
    with torch.cuda.device(x.device), _timed("rf_cl_gn_stats", nbytes=4.0 * x.numel()):
        check(_lib.lib().rf_cl_gn_stats(x.data_ptr(), _ptr(x2), C2, gamma.data_ptr(), mu.data_ptr(), a.data_ptr(), N, C, D, H,
                                        W, groups, float(eps), ws.data_ptr(), _stream(x)), "rf_cl_gn_stats")
    _count(3 if x2 is not None else 2)
    return mu, a


ACT_SCALE_GN = 16.0  # GroupNorm-ed activations are O(1): x16 keeps their fp16 lo parts normal, far from overflow


def cl_norm_split(x, gn=None, c_off=0, scale=1.0):
    """fp32 channels-last -> (hi, lo) fp16 [N,D,H,W,Cp] of scale * value; gn = (mu, a, beta) applies the
    GroupNorm affine with this tensor's channels starting at c_off of the statistics."""
    x = _dev(x, name="x")
    N, D, H, W, C = x.shape
    Cp = 1 if C == 1 else (C + 7) // 8 * 8  # single-channel inputs stay unpadded (tap-major conv mode)
    hi = torch.empty((N, D, H, W, Cp), device=x.device, dtype=torch.int16)
    lo = torch.empty((N, D, H, W, Cp), device=x.device, dtype=torch.int16)
    mu, a, beta = gn if gn is not None else (None, None, None)
    c_tot = mu.shape[1] if mu is not None else C
    with torch.cuda.device(x.device), _timed("rf_cl_norm_split"):
        check(_lib.lib().rf_cl_norm_split(x.data_ptr(), _ptr(mu), _ptr(a), _ptr(beta), c_off, c_tot, hi.data_ptr(),
                                          lo.data_ptr(), N, D * H * W, C, Cp, float(scale), _stream(x)), "rf_cl_norm_split")
    _count()
    return hi, lo


def cl_maxpool3d_2(x):
    x = _dev(x, name="x")
    N, D, H, W, C = x.shape
    y = torch.empty((N, D // 2, H // 2, W // 2, C), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_cl_maxpool3d_2", nbytes=4.5 * x.numel()):
        check(_lib.lib().rf_cl_maxpool3d_2(x.data_ptr(), y.data_ptr(), N, D, H, W, C, _stream(x)), "rf_cl_maxpool3d_2")
    _count()
    return y


def conv3d_cin1_cl(x, weight, bias, gn=None, ks=3, stride=1, pad=0, act=ACT_NONE, slope=0.0):
    """Direct convolution of a single-channel channels-last volume x [N,D,H,W,1] -> fp32 [N,Do,Ho,Wo,Cout]."""
    x = _dev(x, name="x")
    N, D, H, W = x.shape[:4]
    weight = _dev(weight.detach(), name="weight")
    cout = weight.shape[0]
    Do, Ho, Wo = [(v + 2 * pad - ks) // stride + 1 for v in (D, H, W)]
    y = torch.empty((N, Do, Ho, Wo, cout), device=x.device, dtype=torch.float32)
    mu, a, beta = gn if gn is not None else (None, None, None)
    with torch.cuda.device(x.device), _timed("rf_conv3d_cin1_cl_fwd"):
        check(_lib.lib().rf_conv3d_cin1_cl_fwd(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(mu), _ptr(a), _ptr(beta),
                                               y.data_ptr(), N, D, H, W, cout, ks, stride, pad, act, float(slope), _stream(x)),
              "rf_conv3d_cin1_cl_fwd")
    _count()
    return y


def tc_conv_halo_s2_supported(N, D, H, W, cout, c1):
    """Whether the stride-2 'valid' 3x3x3 layer runs on the shifted-window kernel (parity sub-blocks)."""
    return bool(_lib.lib().rf_tc_conv3d_halo_s2_supported(int(N), int(D), int(H), int(W), int(cout), int(c1)))


def cl_split_parity_planes(x, scale=1.0):
    """fp32 channels-last x [N,D,H,W,C] -> (hi, lo) parity planes [chunk][parity][N][D/2][H/2][W/2][8] of scale * x (the 8
    sub-lattices of every sample as dense volumes): the operand layout of tc_conv3d_halo_s2."""
    x = _dev(x, name="x")
    N, D, H, W, C = x.shape
    L = _lib.lib()
    nbytes = L.rf_halo_s2_act_bytes(N, D, H, W, C)
    hi = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    lo = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device), _timed("rf_cl_split_parity_planes", nbytes=8.0 * x.numel()):
        check(L.rf_cl_split_parity_planes(x.data_ptr(), C, hi.data_ptr(), lo.data_ptr(), N, D, H, W, float(scale), _stream(x)),
              "rf_cl_split_parity_planes")
    _count()
    return hi, lo, (N, D, H, W, C, 0, 0)


def tc_conv3d_halo_s2(split, img, bias, cout, act=ACT_NONE, slope=0.0, out_scale=1.0):
    """split = cl_split_parity_planes(x) result.  Conv3d(k3, stride 2, no padding) -> fp32 channels-last [N,Do,Ho,Wo,Cout]."""
    hi, lo, (N, D, H, W, c1, c2, pad) = split
    assert c2 == 0 and pad == 0
    Do, Ho, Wo = (D - 3) // 2 + 1, (H - 3) // 2 + 1, (W - 3) // 2 + 1
    y = torch.empty((N, Do, Ho, Wo, cout), device=hi.device, dtype=torch.float32)
    with torch.cuda.device(hi.device), _timed("rf_tc_conv3d_halo_s2_fwd", flops=2.0 * N * Do * Ho * Wo * 27 * c1 * cout):
        check(_lib.lib().rf_tc_conv3d_halo_s2_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W,
                                                  int(cout), int(c1), act, float(slope), float(out_scale), 0,
                                                  torch.cuda.current_stream(hi.device).cuda_stream), "rf_tc_conv3d_halo_s2_fwd")
    _count()
    return y


def tc_conv_wrun_supported(N, D, H, W, cout, ks, pad):
    """Whether the single-input-channel ks^3 layer runs on the tensor-core kernel (W-run operand planes)."""
    return bool(_lib.lib().rf_tc_conv3d_wrun_supported(int(N), int(D), int(H), int(W), int(cout), int(ks), int(pad)))


def tc_conv_wrun_weight_image(weight):
    """Conv3d weight [Cout, 1, ks,ks,ks] -> (pre-split fp16 operand image for rf_tc_conv3d_wrun_fwd, weight scale)."""
    weight = _dev(weight.detach(), name="weight")
    cout, cin, ks = weight.shape[0], weight.shape[1], weight.shape[2]
    assert cin == 1 and tuple(weight.shape[2:]) == (ks, ks, ks)
    wmax = float(weight.abs().max())
    scale = 2.0 ** (4 - math.floor(math.log2(wmax))) if wmax > 0 and math.isfinite(wmax) else 1.0
    scale = min(max(scale, 2.0 ** -8), 2.0 ** 24)
    L = _lib.lib()
    nbytes = L.rf_tc_conv_wrun_weight_image_bytes(cout, ks)
    if nbytes == 0:
        raise _lib.RfError(f"W-run conv does not support weight {tuple(weight.shape)}")
    img = _aligned_bytes(nbytes, weight.device)
    with torch.cuda.device(weight.device), _timed("rf_tc_conv_wrun_weight_image"):
        check(L.rf_tc_conv_wrun_weight_image(weight.data_ptr(), cout, ks, scale, img.data_ptr(), _stream(weight)),
              "rf_tc_conv_wrun_weight_image")
    _count()
    return img, scale


def tc_conv3d_wrun(x, img, bias, cout, ks, pad=0, gn=None, scale=1.0, act=ACT_NONE, slope=0.0, out_scale=1.0, buffers=None):
    """Single-channel channels-last volume x [N,D,H,W,1] -> fp32 [N,Do,Ho,Wo,Cout] through the shifted-window tensor-core
    kernel: normalise (GroupNorm with one group, gn = (mu, a, beta)) + W-run split, then the convolution."""
    x = _dev(x, name="x")
    N, D, H, W = x.shape[:4]
    L = _lib.lib()
    nbytes = L.rf_wrun_act_bytes(N, D, H, W, int(ks), int(pad))
    if nbytes == 0:
        raise _lib.RfError(f"W-run layout does not support {tuple(x.shape)} ks={ks} pad={pad}")
    key = (N, D, H, W, int(ks), int(pad), x.device)
    if buffers is not None:
        if key not in buffers:
            while len(buffers) >= MAX_PLANE_SHAPES:
                buffers.pop(next(iter(buffers)))
                bump_generation()
            buffers[key] = (torch.empty(nbytes, device=x.device, dtype=torch.uint8), torch.empty(nbytes, device=x.device, dtype=torch.uint8))
        else:
            buffers[key] = buffers.pop(key)
        hi, lo = buffers[key]
    else:
        hi = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        lo = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    mu, a, beta = gn if gn is not None else (None, None, None)
    Do, Ho, Wo = [v + 2 * pad - ks + 1 for v in (D, H, W)]
    y = torch.empty((N, Do, Ho, Wo, cout), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        with _timed("rf_cl_norm_split_wrun", nbytes=N * D * H * (4.0 * W + 32.0 * Wo)):
            check(L.rf_cl_norm_split_wrun(x.data_ptr(), _ptr(mu), _ptr(a), _ptr(beta), hi.data_ptr(), lo.data_ptr(), N, D, H, W, int(ks),
                                          int(pad), float(scale), _stream(x)), "rf_cl_norm_split_wrun")
        _count()
        with _timed("rf_tc_conv3d_wrun_fwd", flops=2.0 * N * Do * Ho * Wo * ks ** 3 * cout):
            check(L.rf_tc_conv3d_wrun_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W, int(ks),
                                          int(pad), int(cout), act, float(slope), float(out_scale), 0, _stream(x)), "rf_tc_conv3d_wrun_fwd")
        _count()
    return y


def cl_pointwise_head(x, weight, bias, act=ACT_NONE, slope=0.0):
    """Conv3d(C, 1, 1) + bias + activation of a channels-last volume x [N,D,H,W,C] -> NCDHW [N,1,D,H,W]."""
    x = _dev(x, name="x")
    N, D, H, W, C = x.shape
    weight = _dev(weight.detach().reshape(-1), name="weight")
    assert weight.numel() == C
    y = torch.empty((N, 1, D, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_cl_pointwise_head"):
        check(_lib.lib().rf_cl_pointwise_head(x.data_ptr(), weight.data_ptr(), _ptr(bias), y.data_ptr(), N * D * H * W, C, act,
                                              float(slope), _stream(x)), "rf_cl_pointwise_head")
    _count()
    return y


def tc_conv_supported(cout, c1, c2, ks):
    return _lib.lib().rf_tc_conv_weight_image_bytes(int(cout), int(c1), int(c2), int(ks)) > 0


def tc_conv_weight_image(weight, c1, c2):
    """Conv3d weight [Cout, C1+C2, k,k,k] -> (pre-split, pre-swizzled fp16 operand image, weight scale).
    The scale is the power of two that brings max|w| into [16, 32)."""
    weight = _dev(weight.detach(), name="weight")
    wmax = float(weight.abs().max())
    scale = 2.0 ** (4 - math.floor(math.log2(wmax))) if wmax > 0 and math.isfinite(wmax) else 1.0
    scale = min(max(scale, 2.0 ** -8), 2.0 ** 24)
    cout, cin, ks = weight.shape[0], weight.shape[1], weight.shape[2]
    assert cin == c1 + c2
    L = _lib.lib()
    nbytes = L.rf_tc_conv_weight_image_bytes(cout, c1, c2, ks)
    if nbytes == 0:
        raise _lib.RfError(f"tensor-core conv does not support weight {tuple(weight.shape)}")
    img = _aligned_bytes(nbytes, weight.device)
    with torch.cuda.device(weight.device), _timed("rf_tc_conv_weight_image"):
        check(L.rf_tc_conv_weight_image(weight.data_ptr(), cout, c1, c2, ks, scale, img.data_ptr(), _stream(weight)),
              "rf_tc_conv_weight_image")
    _count()
    return img, scale


def tc_conv3d(xs, x2s, c1, c2, img, bias, cout, ks, stride=1, pad=0, act=ACT_NONE, slope=0.0, out_ncdhw=False,
              out_scale=1.0):
    """xs / x2s: (hi, lo) split tensors or None; img: operand image; out_scale = 1 / (activation scale x
    weight scale).  Returns fp32 channels-last [N,Do,Ho,Wo,Cout] or NCDHW."""
    if xs is not None:
        N, D, H, W = xs[0].shape[:4]
    else:
        N, D, H, W = x2s[0].shape[0], 2 * x2s[0].shape[1], 2 * x2s[0].shape[2], 2 * x2s[0].shape[3]
    dev = (xs or x2s)[0].device
    Do, Ho, Wo = [(v + 2 * pad - ks) // stride + 1 for v in (D, H, W)]
    shape = (N, cout, Do, Ho, Wo) if out_ncdhw else (N, Do, Ho, Wo, cout)
    y = torch.empty(shape, device=dev, dtype=torch.float32)
    xh, xl = (xs[0].data_ptr(), xs[1].data_ptr()) if xs is not None else (None, None)
    x2h, x2l = (x2s[0].data_ptr(), x2s[1].data_ptr()) if x2s is not None else (None, None)
    with torch.cuda.device(dev), _timed("rf_tc_conv3d_fwd", flops=2.0 * N * Do * Ho * Wo * ks ** 3 * (c1 + c2) * cout):
        check(_lib.lib().rf_tc_conv3d_fwd(xh, xl, c1, x2h, x2l, c2, img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W, cout,
                                          ks, stride, pad, act, float(slope), float(out_scale), int(bool(out_ncdhw)),
                                          torch.cuda.current_stream(dev).cuda_stream), "rf_tc_conv3d_fwd")
    _count()
    return y


# ---- shifted-window ("halo") tensor-core convolution: 3x3x3, stride 1, pad 1 -------------------

def tc_conv_halo_supported(N, D, H, W, cout, c1, c2, pad=1):
    """D, H, W: input extents; pad 1 = 'same' (U-Nets), pad 0 = 'valid' (conv patch encoders)."""
    return bool(_halo_query("rf_tc_conv3d_halo_supported", int(N), int(D), int(H), int(W), int(cout), int(c1), int(c2), int(pad)))


def tc_conv_halo_geometry(N, D, H, W, cout, c1, c2, pad=1):
    """Item shape the kernel picks: dict(stacked, G, Dt, Ht, lines, n_tiles, n_items, smem) or None."""
    out = (ctypes.c_int * 8)()
    if not _lib.lib().rf_tc_conv3d_halo_geometry(int(N), int(D), int(H), int(W), int(cout), int(c1), int(c2), int(pad), out):
        return None
    return dict(zip(["stacked", "G", "Dt", "Ht", "lines", "n_tiles", "n_items", "smem"], list(out)))


class _SplitShape(tuple):
    """(N, D, H, W, c1, c2, pad) of a pair of operand planes; wp: W-pair layout (rf_cl_norm_split_halo_wp)."""
    wp = False


def tc_conv_halo_wp_mode():
    """RF_HALO_WP: '0' never use the W-pair variant, '1' wherever it runs, unset: where the cost model prefers it."""
    return os.environ.get("RF_HALO_WP", "")


# The per-layer questions below are pure functions of the shape (the item chooser runs on the host): memoised, because a
# U-Net forward asks several of them per layer and an eager forward on 8 chunks is host-bound.
@functools.lru_cache(maxsize=4096)
def _halo_query(fn, *args):
    return int(getattr(_lib.lib(), fn)(*args))


def tc_conv_halo_wp_wanted(N, D, H, W, cout, c1, c2, pad=1):
    """Whether the W-pair variant of the shifted-window kernel (one GEMM row = two output voxels, N = 2 Cout) should run
    this 3x3x3 stride-1 layer.  D, H, W: input extents."""
    mode = tc_conv_halo_wp_mode()
    if mode == "0":
        return False
    r = _halo_query("rf_tc_conv3d_halo_wp_supported", int(N), int(D), int(H), int(W), int(cout), int(c1), int(c2), int(pad))
    return r >= (1 if mode == "1" else 2)


def cl_norm_split_halo(x, x2=None, gn=None, scale=1.0, pad=1, buffers=None, wp=False):
    """fp32 channels-last x [N,D,H,W,C1] (or None) and half-resolution x2 [N,D/2,H/2,W/2,C2] (or None) ->
    (hi, lo) compact fp16 slot planes [chunk][N][D][H][W][8] of scale * GroupNorm(concat(x, up2(x2))); the zero halo
    of width pad is produced by the convolution's TMA loads (out-of-bounds zero fill).  wp: W-pair planes
    [chunk][w parity][N][D][H][W/2][8] for tc_conv3d_halo on a weight image made with wp=True."""
    src = x if x is not None else x2
    src = _dev(src, name="x")
    c1 = x.shape[-1] if x is not None else 0
    c2 = x2.shape[-1] if x2 is not None else 0
    if x is not None:
        N, D, H, W = x.shape[:4]
    else:
        N, D, H, W = x2.shape[0], 2 * x2.shape[1], 2 * x2.shape[2], 2 * x2.shape[3]
    L = _lib.lib()
    nbytes = L.rf_halo_act_bytes(N, D, H, W, c1, c2, int(pad))
    if nbytes == 0:
        raise _lib.RfError(f"halo layout does not support C={c1}+{c2}")
    # buffers: a dict owned by the caller (one per layer): the planes for this shape are allocated once and reused
    # (compact planes, every slot is rewritten by every call; the conv's TMA loads make the zero halo on the fly)
    interior_only = 0
    if buffers is not None:
        key = (N, D, H, W, c1, c2, int(pad), src.device, bool(wp))
        if key not in buffers:
            # planes are kept per input shape: a CUDA graph captured for another batch size still points at its own
            # set.  Only when more than MAX_PLANE_SHAPES shapes are live is the least recently used set freed, and
            # that invalidates every captured graph (bump_generation)
            while len(buffers) >= MAX_PLANE_SHAPES:
                buffers.pop(next(iter(buffers)))
                bump_generation()
            buffers[key] = (torch.empty(nbytes, device=src.device, dtype=torch.uint8),
                            torch.empty(nbytes, device=src.device, dtype=torch.uint8))
        else:
            buffers[key] = buffers.pop(key)  # most recently used last
        hi, lo = buffers[key]
    else:
        hi = torch.empty(nbytes, device=src.device, dtype=torch.uint8)
        lo = torch.empty(nbytes, device=src.device, dtype=torch.uint8)
    mu, a, beta = gn if gn is not None else (None, None, None)
    with torch.cuda.device(src.device), _timed("rf_cl_norm_split_halo", nbytes=(c1 + c2) * 8.0 * N * D * H * W):
        if wp:
            check(L.rf_cl_norm_split_halo_wp(_ptr(x), c1, _ptr(x2), c2, _ptr(mu), _ptr(a), _ptr(beta), hi.data_ptr(), lo.data_ptr(),
                                             N, D, H, W, int(pad), float(scale), _stream(src)), "rf_cl_norm_split_halo_wp")
        else:
            check(L.rf_cl_norm_split_halo(_ptr(x), c1, _ptr(x2), c2, _ptr(mu), _ptr(a), _ptr(beta), hi.data_ptr(), lo.data_ptr(),
                                          N, D, H, W, int(pad), float(scale), interior_only, _stream(src)), "rf_cl_norm_split_halo")
    _count()
    shape = _SplitShape((N, D, H, W, c1, c2, int(pad)))
    shape.wp = bool(wp)
    return hi, lo, shape


def unet_front16(x, gn1_w, gn1_b, eps1, conv_w_host, gn2_w, gn2_b, groups2, eps2, scale, wp=False, buffers=None):
    """Fused front of a 'gcr' DoubleConv on 16^3 single-channel samples (rf_unet_front16_fwd): x [N,16,16,16,1] ->
    the operand planes cl_norm_split_halo(relu(conv(GroupNorm(x))), GroupNorm 2) would produce, as a split tuple for
    tc_conv3d_halo.  gn1_w / gn1_b: floats; conv_w_host: contiguous CPU fp32 [8,1,3,3,3]; gn2_w / gn2_b: device [8]."""
    x = _dev(x, name="x")
    N = x.shape[0]
    assert tuple(x.shape[1:]) == (16, 16, 16, 1) and conv_w_host.device.type == "cpu" and conv_w_host.numel() == 216
    L = _lib.lib()
    nbytes = L.rf_halo_act_bytes(N, 16, 16, 16, 8, 0, 1)
    key = (N, 16, 16, 16, 8, 0, 1, x.device, bool(wp))
    if buffers is not None:
        if key not in buffers:
            while len(buffers) >= MAX_PLANE_SHAPES:
                buffers.pop(next(iter(buffers)))
                bump_generation()
            buffers[key] = (torch.empty(nbytes, device=x.device, dtype=torch.uint8), torch.empty(nbytes, device=x.device, dtype=torch.uint8))
        else:
            buffers[key] = buffers.pop(key)
        hi, lo = buffers[key]
    else:
        hi = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        lo = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device), _timed("rf_unet_front16_fwd", flops=2.0 * N * 4096 * 27 * 8):
        check(L.rf_unet_front16_fwd(x.data_ptr(), float(gn1_w), float(gn1_b), float(eps1), conv_w_host.data_ptr(), gn2_w.data_ptr(),
                                         gn2_b.data_ptr(), int(groups2), float(eps2), float(scale), hi.data_ptr(), lo.data_ptr(), N,
                                         int(bool(wp)), _stream(x)), "rf_unet_front16_fwd")
    _count()
    shape = _SplitShape((N, 16, 16, 16, 8, 0, 1))
    shape.wp = bool(wp)
    return hi, lo, shape


def tc_conv_halo_weight_image(weight, c1, c2, wp=False):
    """Conv3d weight [Cout, C1+C2, 3,3,3] -> (pre-split fp16 operand image for rf_tc_conv3d_halo_fwd, weight scale).
    wp: the image of the W-pair variant (rf_tc_conv3d_halo_wp_fwd)."""
    weight = _dev(weight.detach(), name="weight")
    assert tuple(weight.shape[2:]) == (3, 3, 3) and weight.shape[1] == c1 + c2
    wmax = float(weight.abs().max())
    scale = 2.0 ** (4 - math.floor(math.log2(wmax))) if wmax > 0 and math.isfinite(wmax) else 1.0
    scale = min(max(scale, 2.0 ** -8), 2.0 ** 24)
    cout = weight.shape[0]
    L = _lib.lib()
    nbytes = (L.rf_tc_conv_halo_wp_weight_image_bytes if wp else L.rf_tc_conv_halo_weight_image_bytes)(cout, c1, c2)
    if nbytes == 0:
        raise _lib.RfError(f"halo conv does not support weight {tuple(weight.shape)}")
    img = _aligned_bytes(nbytes, weight.device)
    with torch.cuda.device(weight.device), _timed("rf_tc_conv_halo_weight_image"):
        check((L.rf_tc_conv_halo_wp_weight_image if wp else L.rf_tc_conv_halo_weight_image)(
            weight.data_ptr(), cout, c1, c2, scale, img.data_ptr(), _stream(weight)), "rf_tc_conv_halo_weight_image")
    _count()
    return img, scale


def tc_conv_halo_wp_pool_supported(N, D, H, W, cout, c1, c2):
    """Whether the W-pair variant can max-pool 2x2x2 in its epilogue for this 'same' 3x3x3 layer (D, H, W: input extents)."""
    if tc_conv_halo_wp_mode() == "0":
        return False
    return bool(_halo_query("rf_tc_conv3d_halo_wp_pool_supported", int(N), int(D), int(H), int(W), int(cout), int(c1), int(c2)))


def tc_conv3d_halo(split, img, bias, cout, act=ACT_NONE, slope=0.0, out_ncdhw=False, out_scale=1.0, pool=False):
    """split = cl_norm_split_halo(...) result.  Returns fp32 channels-last [N,Do,Ho,Wo,Cout] or NCDHW
    (output extents = input + 2 pad - 2).  pool (W-pair planes only): MaxPool3d(2) of the activated output is taken in
    the epilogue, the result is [N,Do/2,Ho/2,Wo/2,Cout]."""
    hi, lo, (N, D, H, W, c1, c2, pad) = split
    Do, Ho, Wo = D + 2 * pad - 2, H + 2 * pad - 2, W + 2 * pad - 2
    if pool:
        assert getattr(split[2], "wp", False) and pad == 1 and not out_ncdhw, "the pooling epilogue belongs to the W-pair variant"
        y = torch.empty((N, Do // 2, Ho // 2, Wo // 2, cout), device=hi.device, dtype=torch.float32)
        with torch.cuda.device(hi.device), _timed("rf_tc_conv3d_halo_fwd", flops=2.0 * N * Do * Ho * Wo * 27 * (c1 + c2) * cout):
            check(_lib.lib().rf_tc_conv3d_halo_wp_pool_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W,
                                                           cout, c1, c2, act, float(slope), float(out_scale),
                                                           torch.cuda.current_stream(hi.device).cuda_stream), "rf_tc_conv3d_halo_wp_pool_fwd")
        _count()
        return y
    shape = (N, cout, Do, Ho, Wo) if out_ncdhw else (N, Do, Ho, Wo, cout)
    y = torch.empty(shape, device=hi.device, dtype=torch.float32)
    with torch.cuda.device(hi.device), _timed("rf_tc_conv3d_halo_fwd", flops=2.0 * N * Do * Ho * Wo * 27 * (c1 + c2) * cout):
        if getattr(split[2], "wp", False):  # W-pair planes (img must come from tc_conv_halo_weight_image(..., wp=True))
            assert not out_ncdhw, "the W-pair variant writes channels-last"
            check(_lib.lib().rf_tc_conv3d_halo_wp_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W,
                                                      pad, cout, c1, c2, act, float(slope), float(out_scale),
                                                      torch.cuda.current_stream(hi.device).cuda_stream), "rf_tc_conv3d_halo_wp_fwd")
        else:
            check(_lib.lib().rf_tc_conv3d_halo_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), y.data_ptr(), N, D, H, W,
                                                   pad, cout, c1, c2, act, float(slope), float(out_scale), int(bool(out_ncdhw)),
                                                   torch.cuda.current_stream(hi.device).cuda_stream), "rf_tc_conv3d_halo_fwd")
    _count()
    return y


def tc_conv_halo_gn_supported(N, D, H, W, cout, c1, c2, groups2, pad=1):
    """Whether the shifted-window kernel can apply the next layer's GroupNorm in its epilogue (one whole sample per item)."""
    r = _halo_query("rf_tc_conv3d_halo_gn_supported", int(N), int(D), int(H), int(W), int(pad), int(cout), int(c1), int(c2), int(groups2))
    return r >= (1 if os.environ.get("RF_HALO_GN", "") == "1" else 2)  # RF_HALO_GN=1: wherever it runs (tests)


def tc_conv3d_halo_gn(split, img, bias, cout, gn2_w, gn2_b, groups2, eps2, scale2, act=ACT_NONE, slope=0.0, out_scale=1.0, out_wp=False,
                      buffers=None):
    """Convolution (plain operand planes in `split`) whose epilogue applies the NEXT layer's GroupNorm and writes that layer's
    operand planes: returns the split tuple cl_norm_split_halo(conv_output, GroupNorm 2, scale2, wp=out_wp) would give."""
    hi, lo, (N, D, H, W, c1, c2, pad) = split
    assert not getattr(split[2], "wp", False)
    Do, Ho, Wo = D + 2 * pad - 2, H + 2 * pad - 2, W + 2 * pad - 2
    L = _lib.lib()
    nbytes = L.rf_halo_act_bytes(N, Do, Ho, Wo, cout, 0, 1)
    key = (N, Do, Ho, Wo, cout, 0, 1, hi.device, bool(out_wp))
    if buffers is not None:
        if key not in buffers:
            while len(buffers) >= MAX_PLANE_SHAPES:
                buffers.pop(next(iter(buffers)))
                bump_generation()
            buffers[key] = (torch.empty(nbytes, device=hi.device, dtype=torch.uint8), torch.empty(nbytes, device=hi.device, dtype=torch.uint8))
        else:
            buffers[key] = buffers.pop(key)
        ohi, olo = buffers[key]
    else:
        ohi = torch.empty(nbytes, device=hi.device, dtype=torch.uint8)
        olo = torch.empty(nbytes, device=hi.device, dtype=torch.uint8)
    with torch.cuda.device(hi.device), _timed("rf_tc_conv3d_halo_fwd", flops=2.0 * N * Do * Ho * Wo * 27 * (c1 + c2) * cout):
        check(L.rf_tc_conv3d_halo_gn_fwd(hi.data_ptr(), lo.data_ptr(), img.data_ptr(), _ptr(bias), N, D, H, W, pad, int(cout), c1, c2, act,
                                         float(slope), float(out_scale), gn2_w.data_ptr(), gn2_b.data_ptr(), int(groups2), float(eps2),
                                         float(scale2), ohi.data_ptr(), olo.data_ptr(), int(bool(out_wp)),
                                         torch.cuda.current_stream(hi.device).cuda_stream), "rf_tc_conv3d_halo_gn_fwd")
    _count()
    shape = _SplitShape((N, Do, Ho, Wo, int(cout), 0, 1))
    shape.wp = bool(out_wp)
    return ohi, olo, shape


def groupnorm_stats(x, gamma, groups, eps=1e-5, x2=None):
    """Returns (gn_mu [N,C], gn_a [N,C]) for the virtual input concat(x, up2(x2))."""
    ref = x if x is not None else x2
    C1 = 0
    if x is not None:
        x = _dev(x, name="x")
        N, C1, D, H, W = x.shape
    C2 = 0
    if x2 is not None:
        x2 = _dev(x2, name="x2")
        C2 = x2.shape[1]
        if x is None:
            N, D, H, W = x2.shape[0], 2 * x2.shape[2], 2 * x2.shape[3], 2 * x2.shape[4]
    C = C1 + C2
    mu = torch.empty((N, C), device=ref.device, dtype=torch.float32)
    a = torch.empty((N, C), device=ref.device, dtype=torch.float32)
    with torch.cuda.device(ref.device), _timed("rf_groupnorm_stats"):
        check(_lib.lib().rf_groupnorm_stats(_ptr(x), _ptr(x2), C2, gamma.data_ptr(), mu.data_ptr(), a.data_ptr(), N, C,
                                            D, H, W, groups, float(eps), _stream(ref)), "rf_groupnorm_stats")
    _count()
    return mu, a


def maxpool3d_2(x):
    x = _dev(x, name="x")
    N, C, D, H, W = x.shape
    y = torch.empty((N, C, D // 2, H // 2, W // 2), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_maxpool3d_2"):
        check(_lib.lib().rf_maxpool3d_2(x.data_ptr(), y.data_ptr(), N, C, D, H, W, _stream(x)), "rf_maxpool3d_2")
    _count()
    return y


def upsample_nearest_2(x):
    x = _dev(x, name="x")
    N, C, D, H, W = x.shape
    y = torch.empty((N, C, 2 * D, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_upsample_nearest_2"):
        check(_lib.lib().rf_upsample_nearest_2(x.data_ptr(), y.data_ptr(), N, C, D, H, W, _stream(x)),
              "rf_upsample_nearest_2")
    _count()
    return y


def l2_normalize_rows(x, eps=1e-12):
    """F.normalize(x, dim=1) for a [M, D] matrix."""
    x = _dev(x, name="x")
    M, D = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device), _timed("rf_l2_normalize_rows"):
        check(_lib.lib().rf_l2_normalize_rows(x.data_ptr(), y.data_ptr(), M, D, float(eps), _stream(x)),
              "rf_l2_normalize_rows")
    _count()
    return y


def mlp_encode(x, wts, biases, l2_normalize=True):
    """rf_mlp_encode_fwd: ReLU MLP (Patch04 family) + optional row normalisation."""
    x = _dev(x, name="x")
    M = x.shape[0]
    n = len(wts)
    widths = [x.shape[1]] + [int(w.shape[1]) for w in wts]
    import ctypes
    warr = (ctypes.c_int * 9)(*(widths + [0] * (9 - len(widths))))
    L = _lib.lib()
    ws_bytes = L.rf_mlp_encode_workspace_bytes(M, warr, n)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    out = torch.empty((M, widths[-1]), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_mlp_encode_fwd"):
        check(L.rf_mlp_encode_fwd(x.data_ptr(), ptr_array([w.data_ptr() for w in wts]),
                                  ptr_array([b.data_ptr() for b in biases]), warr, n, int(bool(l2_normalize)),
                                  out.data_ptr(), M, ws.data_ptr(), ws_bytes, _stream(x)), "rf_mlp_encode_fwd")
    _count(n + (1 if l2_normalize else 0))
    return out


# ---------------------------------------------------------------------------
# kNN
# ---------------------------------------------------------------------------

last_knn_stats = {}


_knn_ws = {}


def _knn_workspace(nbytes, device):
    """One lookup workspace per (device, stream), grown on demand and reused: a bulk lookup needs a few hundred MB of
    scratch, and taking that from the caching allocator on every call made the lookup's latency depend on the
    allocator's state.  Lookups on one stream are ordered, so they can share it; other streams get their own."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, device=device, dtype=torch.uint8)  # graph-private pool
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _knn_ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _knn_ws.pop(key, None)
        buf = torch.empty(int(nbytes * 1.25) + 4096, device=device, dtype=torch.uint8)
        _knn_ws[key] = buf
    return buf


class KnnBankImage:
    """The prepared tensor-core operand image of one bank (shard): rf_knn_bank_prepare's output, built once and
    reused by every lookup (the reference loads its FLANN index once per worker, util/retrieval.py:81-83)."""

    def __init__(self, bank, method, image):
        self.bank_ptr, self.n_rows, self.method, self.image = bank.data_ptr(), bank.shape[0], method, image


def knn_prepare_bank(bank, method=0, q_sample=None):
    """bank [n,64] fp32 -> KnnBankImage, or None when the bank is too small for the tensor-core path."""
    bank = _dev(bank, name="bank")
    n = bank.shape[0]
    L = _lib.lib()
    m = L.rf_knn_bank_method(n, int(method))
    if m < 2:
        return None
    image = torch.empty(L.rf_knn_bank_image_bytes(n, m), device=bank.device, dtype=torch.uint8)
    scratch = torch.empty(L.rf_knn_bank_scratch_bytes(n, m), device=bank.device, dtype=torch.uint8)
    if q_sample is not None:
        q_sample = _dev(q_sample, name="q_sample")
    with torch.cuda.device(bank.device), _timed("rf_knn_bank_prepare"):
        check(L.rf_knn_bank_prepare(bank.data_ptr(), n, m, _ptr(q_sample), 0 if q_sample is None else q_sample.shape[0],
                                    image.data_ptr(), image.numel(), scratch.data_ptr(), scratch.numel(), _stream(bank)),
              "rf_knn_bank_prepare")
    _count(5)
    return KnnBankImage(bank, m, image)


def knn_topk(bank, q, k, row_offset=0, method=0, stats=False, image=None):
    """Exact top-k under the canonical (fp64 d, row id) rule.
    bank [n,64] fp32, q [Q,64] fp32 -> (idx int32 [Q,k] global ids, d fp64 [Q,k]).
    image: a KnnBankImage of this bank (skips the per-call operand staging).
    stats=True also fills ops.last_knn_stats for the tensor-core methods."""
    bank = _dev(bank, name="bank")
    q = _dev(q, name="q")
    n, D = bank.shape
    Q = q.shape[0]
    L = _lib.lib()
    idx = torch.empty((Q, k), device=q.device, dtype=torch.int32)
    d = torch.empty((Q, k), device=q.device, dtype=torch.float64)
    if Q == 0:
        return idx, d
    if image is not None:
        if image.bank_ptr != bank.data_ptr() or image.n_rows != n:
            raise _lib.RfError("knn_topk: the prepared image belongs to another bank")
        m = image.method
        ws = _knn_workspace(max(L.rf_knn_prepared_workspace_bytes(Q, n, k, m), 256), q.device)
        with torch.cuda.device(q.device), _timed("rf_knn_l2_topk", flops=2.0 * Q * n * 64):
            check(L.rf_knn_l2_topk_prepared(bank.data_ptr(), n, int(row_offset), image.image.data_ptr(), m, q.data_ptr(), Q, D, k,
                                            idx.data_ptr(), d.data_ptr(), ws.data_ptr(), ws.numel(), _stream(q)),
                  "rf_knn_l2_topk_prepared")
        tc = True
        _count(6)
    else:
        ws = _knn_workspace(max(L.rf_knn_workspace_bytes(Q, n, k, method), 256), q.device)
        with torch.cuda.device(q.device), _timed("rf_knn_l2_topk", flops=2.0 * Q * n * 64):
            check(L.rf_knn_l2_topk(bank.data_ptr(), n, int(row_offset), q.data_ptr(), Q, D, k, method, idx.data_ptr(),
                                   d.data_ptr(), ws.data_ptr(), ws.numel(), _stream(q)), "rf_knn_l2_topk")
        tc = method != 1
        _count(10)
    if stats:
        import ctypes
        nu, err = ctypes.c_int(-1), ctypes.c_float(-1.0)
        if tc:
            with torch.cuda.device(q.device):
                check(L.rf_knn_tc_stats(ws.data_ptr(), ctypes.byref(nu), ctypes.byref(err), _stream(q)), "rf_knn_tc_stats")
        last_knn_stats.update(n_unproven=nu.value, max_score_err=err.value)
    return idx, d


def knn_merge(parts_idx, parts_d):
    """parts_idx int32 [S,Q,k], parts_d fp64 [S,Q,k] -> merged ([Q,k], [Q,k])."""
    parts_idx = _dev(parts_idx, torch.int32, "parts_idx")
    parts_d = _dev(parts_d, torch.float64, "parts_d")
    S, Q, k = parts_idx.shape
    idx = torch.empty((Q, k), device=parts_idx.device, dtype=torch.int32)
    d = torch.empty((Q, k), device=parts_idx.device, dtype=torch.float64)
    with torch.cuda.device(parts_idx.device), _timed("rf_knn_merge"):
        check(_lib.lib().rf_knn_merge(parts_idx.data_ptr(), parts_d.data_ptr(), S, Q, k, idx.data_ptr(), d.data_ptr(),
                                      _stream(parts_idx)), "rf_knn_merge")
    _count()
    return idx, d


def knn_demote_rows(idx2k, d2k, meta, query_scene, K):
    """util/retrieval.py:93-100 -> (rows fp32 [Q,K,8], idx int32 [Q,K])."""
    idx2k = _dev(idx2k, torch.int32, "idx2k")
    d2k = _dev(d2k, torch.float64, "d2k")
    meta = _dev(meta, name="meta")
    Q, K2 = idx2k.shape
    if query_scene is not None:
        query_scene = _dev(query_scene, torch.int32, "query_scene")
    rows = torch.empty((Q, K, 8), device=idx2k.device, dtype=torch.float32)
    idx = torch.empty((Q, K), device=idx2k.device, dtype=torch.int32)
    with torch.cuda.device(idx2k.device), _timed("rf_knn_demote_rows"):
        check(_lib.lib().rf_knn_demote_rows(idx2k.data_ptr(), d2k.data_ptr(), meta.data_ptr(), _ptr(query_scene), Q, K2,
                                            K, rows.data_ptr(), idx.data_ptr(), _stream(idx2k)), "rf_knn_demote_rows")
    _count()
    return rows, idx


def compose_gather(rows, dst_extents, scene_store, n_chunks, chunk_size, trunc, ratio, prefill=False, norm_sub=0.0,
                   norm_div=0.0, patch_block=None):
    """util/retrieval.py:145-164 for non-overlapping patches.
    rows [n_chunks*P,K,8]; dst_extents int32 [P,6]; scene_store [S,sx,sy,sz] -> [n_chunks,K,cx,cy,cz].
    prefill=True initialises the output with `trunc` (:148) for scenes whose patch list does not tile the chunk.
    patch_block = (ex, ey, ez): the P destination blocks all have these extents and tile the chunk in Unfold3D's patch
    order (the caller checked) -> [n_chunks,K,P,ex,ey,ez], i.e. Unfold3D(ex, 1) of the composed volumes without the pass
    over them (rf_compose_gather_patches)."""
    rows = _dev(rows, name="rows")
    dst_extents = _dev(dst_extents, torch.int32, "dst_extents")
    scene_store = _dev(scene_store, name="scene_store")
    P = dst_extents.shape[0]
    K = rows.shape[1]
    assert rows.shape[0] == n_chunks * P
    shape = (n_chunks, K) + tuple(int(v) for v in chunk_size)
    if patch_block is not None:
        pb = tuple(int(v) for v in patch_block)
        assert P * pb[0] * pb[1] * pb[2] == shape[2] * shape[3] * shape[4], "patch blocks do not tile the chunk"
        shape = (n_chunks, K, P) + pb
    out = (torch.full(shape, float(trunc), device=rows.device, dtype=torch.float32) if prefill
           else torch.empty(shape, device=rows.device, dtype=torch.float32))
    fn = _lib.lib().rf_compose_gather if patch_block is None else _lib.lib().rf_compose_gather_patches
    with torch.cuda.device(rows.device), _timed("rf_compose_gather", nbytes=2.0 * out.numel() * 4):
        check(fn(rows.data_ptr(), dst_extents.data_ptr(), scene_store.data_ptr(),
                                           out.data_ptr(), n_chunks, P, K, scene_store.shape[0],
                                           int3(scene_store.shape[1:]), int3(chunk_size), float(trunc), float(ratio),
                                           float(norm_sub), float(norm_div), _stream(rows)), "rf_compose_gather")
    _count()
    return out


# ---------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------

def _img_array(branch):
    """branch = (wts, biases[, images]); returns the ctypes pointer array of the images or None."""
    if len(branch) < 3 or branch[2] is None:
        return None
    return ptr_array([im.data_ptr() for im in branch[2]])


def attention_fuse(x_back, x_retr, theta, phi, E, K, normalize=True, mode=0, blend=True, gumbel_noise=None, patch_grid=1,
                   out_channels_last=False, output_mapping=None):
    """model/attention.py:141-157. theta/phi: (4 wt [in,out], 4 biases[, 4 tensor-core weight images]).
    patch_grid P > 1: x_retr is the retrieval U-Net's un-folded patch batch [B*K*P^3, nf, S/P,S/P,S/P] (Fold3D's input,
    train_refinement.py:112); out_channels_last: the result is [B,S,S,S,nf], the decoder's channels-last operand
    (rf_attention_fuse_patched_fwd).  output_mapping = (Wo Wg [nf,nf], Wo bg [nf], bo [nf]): the composed g / o 1x1x1
    convolutions of attn_no_output_mapping=False (model/attention.py:56-57)."""
    _forward_only(x_back, x_retr)
    x_back = _dev(x_back, name="x_predicted")
    x_retr = _dev(x_retr, name="x_retrieved")
    B, nf, S = x_back.shape[0], x_back.shape[1], x_back.shape[2]
    P = max(1, int(patch_grid))
    assert S % P == 0 and x_retr.shape[0] == B * K * P ** 3 and x_retr.shape[1] == nf and x_retr.shape[2] == S // P
    L = _lib.lib()
    ws_bytes = L.rf_attention_workspace_bytes(B, nf, S, E, K)
    ws = torch.empty(ws_bytes, device=x_back.device, dtype=torch.uint8)
    out = (torch.empty((B, S, S, S, nf), device=x_back.device, dtype=torch.float32) if out_channels_last
           else torch.empty_like(x_back))
    if gumbel_noise is not None:
        gumbel_noise = _dev(gumbel_noise, name="gumbel_noise")
    with torch.cuda.device(x_back.device), _timed("rf_attention_fuse_fwd", nbytes=(K + 2.0) * nf * S ** 3 * 4 * B):
        check(L.rf_attention_fuse_patched_fwd(x_back.data_ptr(), x_retr.data_ptr(),
                                              ptr_array([w.data_ptr() for w in theta[0]]),
                                              ptr_array([b.data_ptr() for b in theta[1]]),
                                              ptr_array([w.data_ptr() for w in phi[0]]),
                                              ptr_array([b.data_ptr() for b in phi[1]]), _img_array(theta), _img_array(phi),
                                              _ptr(gumbel_noise), out.data_ptr(), B, nf, S, E, K, int(bool(normalize)), int(mode),
                                              int(bool(blend)), P, int(bool(out_channels_last)),
                                              None if output_mapping is None else ptr_array([t.data_ptr() for t in output_mapping]),
                                              ws.data_ptr(), ws_bytes, _stream(x_back)), "rf_attention_fuse_patched_fwd")
    n_mlp = 1 if (len(theta) > 2 and theta[2] is not None) else 4   # launches per MLP: fused chain or four linears
    _count(2 + 2 * n_mlp + 1 + (0 if out_channels_last else 1) + (0 if output_mapping is None else 1))
    return out


def attention_features(x, t, occ, theta, phi, E, normalize=True):
    """model/attention.py:132-139 -> (x_feat [R,32], p_feat [R,32], occ_any bool [R])."""
    _forward_only(x, t)
    x = _dev(x, name="x_predicted")
    t = _dev(t, name="x_target")
    B, nf, S = x.shape[0], x.shape[1], x.shape[2]
    Rp = S // E
    R = B * Rp ** 3
    occ_u8 = occ.to(torch.uint8).contiguous()
    L = _lib.lib()
    ws_bytes = L.rf_attention_workspace_bytes(B, nf, S, E, 1)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    xf = torch.empty((R, 32), device=x.device, dtype=torch.float32)
    pf = torch.empty((R, 32), device=x.device, dtype=torch.float32)
    oa = torch.empty((R,), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device), _timed("rf_attention_features"):
        check(L.rf_attention_features(x.data_ptr(), t.data_ptr(), occ_u8.data_ptr(),
                                      ptr_array([w.data_ptr() for w in theta[0]]),
                                      ptr_array([b.data_ptr() for b in theta[1]]),
                                      ptr_array([w.data_ptr() for w in phi[0]]),
                                      ptr_array([b.data_ptr() for b in phi[1]]), _img_array(theta), _img_array(phi),
                                      xf.data_ptr(), pf.data_ptr(), oa.data_ptr(), B, nf, S, E, int(bool(normalize)), ws.data_ptr(), ws_bytes,
                                      _stream(x)), "rf_attention_features")
    _count(13)
    return xf, pf, oa.bool()


# ---------------------------------------------------------------------------
# target normals, contrastive loss, evaluation metrics
# ---------------------------------------------------------------------------

def sobel_normals(target, pad_val):
    """dataset/patched_scene_dataset.py:139-146 compute_normals: [B,1,D,H,W] -> [B,3,D,H,W]."""
    _forward_only(target)
    target = _dev(target, name="target")
    assert target.dim() == 5 and target.shape[1] == 1, "compute_normals expects [B,1,D,H,W]"
    B, _, D, H, W = target.shape
    out = torch.empty((B, 3, D, H, W), device=target.device, dtype=torch.float32)
    with torch.cuda.device(target.device), _timed("rf_sobel_normals"):
        check(_lib.lib().rf_sobel_normals(target.data_ptr(), out.data_ptr(), B, D, H, W, float(pad_val), _stream(target)),
              "rf_sobel_normals")
    _count()
    return out


def occupancy_counts(pred, target):
    """The integer sums of util/metrics.py: bool [B,1,S,S,S] x2 -> int64 [B,4] = (and, or, pred, target)."""
    assert pred.shape == target.shape and pred.dtype == torch.bool and target.dtype == torch.bool
    if not pred.is_cuda:
        raise _lib.RfError("occupancy_counts: CPU tensors are not supported (no CPU fallback)")
    pred, target = pred.contiguous(), target.contiguous()
    B = pred.shape[0]
    counts = torch.empty((B, 4), device=pred.device, dtype=torch.int64)
    with torch.cuda.device(pred.device), _timed("rf_occupancy_counts"):
        check(_lib.lib().rf_occupancy_counts(pred.data_ptr(), target.data_ptr(), B, pred[0].numel(), counts.data_ptr(),
                                             _stream(pred)), "rf_occupancy_counts")
    _count()
    return counts


def chamfer_nn(a, b):
    """Nearest neighbour in b of every point of a: a [na,3], b [nb,3] fp32 -> (dist fp32 [na], idx int32 [na])."""
    a = _dev(a, name="a")
    b = _dev(b, name="b")
    assert a.dim() == 2 and a.shape[1] == 3 and b.dim() == 2 and b.shape[1] == 3
    dist = torch.empty(a.shape[0], device=a.device, dtype=torch.float32)
    idx = torch.empty(a.shape[0], device=a.device, dtype=torch.int32)
    with torch.cuda.device(a.device), _timed("rf_chamfer_nn"):
        check(_lib.lib().rf_chamfer_nn(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], dist.data_ptr(), idx.data_ptr(),
                                       _stream(a)), "rf_chamfer_nn")
    _count()
    return dist, idx


def ntxent(zis, zjs, temperature, cosine=True, iou_matrix=None, sig_scale=80.0, sig_shift=-65.0):
    """model/loss.py:48-69 NTXentLoss.forward -> scalar tensor."""
    _forward_only(zis, zjs)
    zis = _dev(zis, name="zis")
    zjs = _dev(zjs, name="zjs")
    assert zis.shape == zjs.shape and zis.dim() == 2
    N, C = zis.shape
    if iou_matrix is not None:
        iou_matrix = _dev(iou_matrix, name="iou_matrix")
        assert iou_matrix.shape == (2 * N, 2 * N)
    L = _lib.lib()
    ws = torch.empty(max(L.rf_ntxent_workspace_bytes(N), 16), device=zis.device, dtype=torch.uint8)
    loss = torch.empty(1, device=zis.device, dtype=torch.float32)
    with torch.cuda.device(zis.device), _timed("rf_ntxent_fwd"):
        check(L.rf_ntxent_fwd(zis.data_ptr(), zjs.data_ptr(), N, C, iou_matrix.data_ptr() if iou_matrix is not None else None,
                              float(temperature), float(sig_scale), float(sig_shift), 1 if cosine else 0, loss.data_ptr(),
                              ws.data_ptr(), ws.numel(), _stream(zis)), "rf_ntxent_fwd")
    _count(3 if cosine else 2)
    return loss[0]


# ---------------------------------------------------------------------------
# adjoint kernels (rf_backward.cu), used by retrieval_fuse_b200.autograd
# ---------------------------------------------------------------------------

def conv3d_wgrad(x, x2, gn, dz, ks, stride=1, pad=0):
    """Filter gradient of conv3d(): x [N,C1,D,H,W] or None, x2 [N,C2,D/2,..] or None, gn = (mu, a, beta) or None,
    dz [N,Cout,Do,Ho,Wo] -> dW [Cout, C1+C2, ks,ks,ks]."""
    ref = x if x is not None else x2
    dz = _dev(dz, name="dz")
    C1 = C2 = 0
    if x is not None:
        x = _dev(x, name="x")
        N, C1, D, H, W = x.shape
    if x2 is not None:
        x2 = _dev(x2, name="x2")
        C2 = x2.shape[1]
        if x is None:
            N, D, H, W = x2.shape[0], 2 * x2.shape[2], 2 * x2.shape[3], 2 * x2.shape[4]
    cin, cout = C1 + C2, dz.shape[1]
    dw = torch.empty((cout, cin, ks, ks, ks), device=ref.device, dtype=torch.float32)
    g = gn if gn is not None else (None, None, None)
    with torch.cuda.device(ref.device), _timed("rf_conv3d_wgrad"):
        check(_lib.lib().rf_conv3d_wgrad(_ptr(x), _ptr(x2), C2, _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), dz.data_ptr(), dw.data_ptr(), N, cin,
                                         D, H, W, cout, ks, stride, pad, _stream(ref)), "rf_conv3d_wgrad")
    _count(2)
    return dw


def act_bwd(dy, y, act, slope=0.0):
    dy, y = _dev(dy, name="dy"), _dev(y, name="y")
    if act == ACT_NONE:
        return dy
    dz = torch.empty_like(dy)
    with torch.cuda.device(dy.device), _timed("rf_act_bwd"):
        check(_lib.lib().rf_act_bwd(dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), act, float(slope), _stream(dy)), "rf_act_bwd")
    _count()
    return dz


def channel_sum(x):
    """x [N, C, ...] -> [C] sums over everything but the channel axis."""
    x = _dev(x, name="x")
    N, C = x.shape[0], x.shape[1]
    V = x.numel() // (N * C)
    out = torch.empty(C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _timed("rf_channel_sum"):
        check(_lib.lib().rf_channel_sum(x.data_ptr(), N, C, V, out.data_ptr(), _stream(x)), "rf_channel_sum")
    _count(2)
    return out


def gn_bwd(x, x2, g, mu, rstd, gamma, groups):
    """GroupNorm backward on the virtual input concat(x, up2(x2)); g = grad of the normalised input [N,C,D,H,W].
    -> (dx or None, dx2 or None, dgamma [C], dbeta [C])."""
    g = _dev(g, name="g")
    N, C, D, H, W = g.shape
    C2 = x2.shape[1] if x2 is not None else 0
    x = _dev(x, name="x") if x is not None else None
    x2 = _dev(x2, name="x2") if x2 is not None else None
    dx = torch.empty_like(x) if x is not None else None
    dx2 = torch.empty_like(x2) if x2 is not None else None
    dgamma = torch.empty(C, device=g.device, dtype=torch.float32)
    dbeta = torch.empty(C, device=g.device, dtype=torch.float32)
    L = _lib.lib()
    ws = torch.empty(L.rf_gn_bwd_workspace_bytes(N, C), device=g.device, dtype=torch.uint8)
    with torch.cuda.device(g.device), _timed("rf_gn_bwd"):
        check(L.rf_gn_bwd(_ptr(x), _ptr(x2), C2, g.data_ptr(), mu.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), N, C, D, H, W, groups,
                          _ptr(dx), _ptr(dx2), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), _stream(g)), "rf_gn_bwd")
    _count(5)
    return dx, dx2, dgamma, dbeta


def upsample2_bwd(g, C1):
    g = _dev(g, name="g")
    N, C, D, H, W = g.shape
    dx2 = torch.empty((N, C - C1, D // 2, H // 2, W // 2), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device), _timed("rf_upsample2_bwd"):
        check(_lib.lib().rf_upsample2_bwd(g.data_ptr(), N, C, C1, D, H, W, dx2.data_ptr(), _stream(g)), "rf_upsample2_bwd")
    _count()
    return dx2


def maxpool3d_2_bwd(x, dy):
    x, dy = _dev(x, name="x"), _dev(dy, name="dy")
    N, C, D, H, W = x.shape
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device), _timed("rf_maxpool3d_2_bwd"):
        check(_lib.lib().rf_maxpool3d_2_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), N, C, D, H, W, _stream(x)), "rf_maxpool3d_2_bwd")
    _count()
    return dx


def attention_epilogue(xf, pf, xu, pu, noise, rp3, K, normalize, mode, blend, sharp):
    """model/attention.py:84-113 on precomputed features: xf [R,32], pf [(b,k,r),32], xu [R,V], pu [(b,k,r),V] -> [R,V]."""
    xf, pf, xu, pu = _dev(xf, name="xf"), _dev(pf, name="pf"), _dev(xu, name="xu"), _dev(pu, name="pu")
    R, V = xu.shape
    out = torch.empty_like(xu)
    with torch.cuda.device(xu.device), _timed("rf_attention_epilogue_fwd"):
        check(_lib.lib().rf_attention_epilogue_fwd(xf.data_ptr(), pf.data_ptr(), xu.data_ptr(), pu.data_ptr(), _ptr(noise), out.data_ptr(), R,
                                                   rp3, K, V, int(bool(normalize)), int(mode), int(bool(blend)), float(sharp), _stream(xu)),
              "rf_attention_epilogue_fwd")
    _count()
    return out


def attention_epilogue_bwd(xf, pf, xu, pu, noise, dout, rp3, K, normalize, mode, blend, sharp):
    dout = _dev(dout, name="dout")
    R, V = xu.shape
    dxf, dpf, dxu, dpu = torch.empty_like(xf), torch.empty_like(pf), torch.empty_like(xu), torch.empty_like(pu)
    with torch.cuda.device(xu.device), _timed("rf_attention_epilogue_bwd"):
        check(_lib.lib().rf_attention_epilogue_bwd(xf.data_ptr(), pf.data_ptr(), xu.data_ptr(), pu.data_ptr(), _ptr(noise), dout.data_ptr(),
                                                   dxf.data_ptr(), dpf.data_ptr(), dxu.data_ptr(), dpu.data_ptr(), R, rp3, K, V,
                                                   int(bool(normalize)), int(mode), int(bool(blend)), float(sharp), _stream(xu)),
              "rf_attention_epilogue_bwd")
    _count()
    return dxf, dpf, dxu, dpu
