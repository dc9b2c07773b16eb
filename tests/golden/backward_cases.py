"""Seeded inputs shared by make_golden_backward.py (reference autograd run) and tests/test_gpu_backward.py."""
import torch

import cases as C

SEED = C.SEED
SINGLE_CONV = [(16, 0, 32, 8), (8, 16, 24, 4), (0, 16, 16, 8), (1, 0, 8, 6)]  # (C1, C2 (upsampled), Cout, extent)
REFINE_CFG = C.REFINE_CFG


def stride(n):
    """Gradient tensors are stored as a strided sample of at most ~4096 values."""
    return max(1, n // 4096)


def single_conv_tag(c1, c2, cout, S):
    return f"single_conv.{c1}.{c2}.{cout}.{S}"


def single_conv_inputs(c1, c2, cout, S):
    t = single_conv_tag(c1, c2, cout, S)
    N, cin = 3, c1 + c2
    x = C.rnd(t + ".x", (N, c1, S, S, S)) if c1 else None
    x2 = C.rnd(t + ".x2", (N, c2, S // 2, S // 2, S // 2)) if c2 else None
    gamma = 1 + 0.3 * C.rnd(t + ".gamma", (cin,))
    beta = 0.2 * C.rnd(t + ".beta", (cin,))
    gout = C.rnd(t + ".g", (N, cout, S, S, S))
    return x, x2, gamma, beta, gout


def retrieval_unet_inputs(nf):
    return C.rnd("rb.bw.x", (3, 1, 16, 16, 16)), C.rnd("rb.bw.g", (3, nf, 8, 8, 8))


def attention_tag(mode):
    return "attention." + ("gumbel" if mode else "softmax")


def attention_cfg(mode):
    return dict(nf=16, attn_patch_extent=4, K=4, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=mode,
                attn_no_output_mapping=True, attn_blend=True, attn_num_patch=4)


def attention_inputs(mode):
    nf, K, S, B = 16, 4, 8, 2
    xb = C.rnd("att.bw.x", (B, nf, S, S, S))
    xr = C.rnd("att.bw.p", (B * K, nf, S, S, S))
    xr[:, :, :4] = xb.repeat_interleave(K, 0)[:, :, :4] + 0.05 * xr[:, :, :4]  # candidates that resemble the prediction
    gout = C.rnd("att.bw.g", (B, nf, S, S, S))
    noise = None
    if mode:
        g = torch.Generator().manual_seed(5)
        noise = -torch.empty(B * (S // 2) ** 3, K).exponential_(generator=g).log()
    return xb, xr, gout, noise


def refine_inputs():
    x_in, x_re = C.refine_full_inputs()
    return x_in, x_re, C.rnd("refine.bw.g", (1, 1, 64, 64, 64))
