"""Design check for a round-2 lever: a 3x3x3 'same' convolution on channels-last data equals a
3x3x3 'same' convolution on the W-PACKED tensors

    x' = x.view(N, D, H, W / Bw, Bw * Cin)            (a pure reinterpretation of the channels-last buffer)
    y' = y.view(N, D, H, W / Bw, Bw * Cout)
    W'[(dw, co), (iw, ci), kd, kh, kw''] = W[co, ci, kd, kh, kw]   with   Bw * (kw'' - 1) + iw = dw + kw - 1

so the existing shifted-window tcgen05 kernel can run small-channel layers with Bw-times fewer GEMM rows, Bw-times
wider N (free while N <= 64: the tensor pipe is bound by the A-operand read there) and - for Cin = 1 - a full K = 8
slot instead of 7 zero channels.  This script verifies the identity on the CPU with torch (no GPU, not on the product
path) and prints the MMA-count model for the layers of the refinement forward.

    python tools/wpack_formulation.py
"""
import math

import torch


def pack_weights(w: torch.Tensor, Bw: int) -> torch.Tensor:
    """w [Cout, Cin, 3, 3, 3] -> [Bw*Cout, Bw*Cin, 3, 3, 3] (Toeplitz along w)."""
    Cout, Cin = w.shape[:2]
    wp = torch.zeros(Bw * Cout, Bw * Cin, 3, 3, 3, dtype=w.dtype)
    for dw in range(Bw):
        for kw in range(3):
            s = dw + kw - 1                      # source w offset relative to the block start
            kwp, iw = s // Bw + 1, s % Bw        # python floor division: s = -1 -> (kw'' = 0, iw = Bw - 1)
            wp[dw * Cout:(dw + 1) * Cout, iw * Cin:(iw + 1) * Cin, :, :, kwp] += w[:, :, :, :, kw]
    return wp


def check(N=2, Cin=1, Cout=8, D=6, H=5, W=16, Bw=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g, dtype=torch.float64)
    y = torch.nn.functional.conv3d(x, w, padding=1)                                     # [N, Cout, D, H, W]
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous()                                        # channels-last [N, D, H, W, Cin]
    xp = x_cl.view(N, D, H, W // Bw, Bw * Cin).permute(0, 4, 1, 2, 3)                   # packed, as NCDHW for torch
    yp = torch.nn.functional.conv3d(xp, pack_weights(w, Bw), padding=1)                 # [N, Bw*Cout, D, H, W/Bw]
    y_cl = yp.permute(0, 2, 3, 4, 1).contiguous().view(N, D, H, W, Cout)                # reinterpret back
    err = float((y_cl - y.permute(0, 2, 3, 4, 1)).abs().max())
    assert err < 1e-12, err
    return err


def mma_model(rows, Cin, Cout, Bw):
    """k-steps (K = 16) per ORIGINAL output row and (kd, kh) group, and the N of each MMA, for the kernel as it is
    (every packed tap costs its full channel range) and with all-zero k-steps skipped."""
    if Bw == 1:
        chunks = math.ceil(Cin / 8)
        steps = 2 if chunks == 1 else 3 * math.ceil(chunks / 2)       # pair mode for a single chunk
        return steps, steps, max(16, Cout)
    chunks = math.ceil(Bw * Cin / 8)
    dense = (2 if chunks == 1 else 3 * math.ceil(chunks / 2)) / Bw
    nz_chunks = math.ceil((Bw + 2) * Cin / 8)                          # only Bw + 2 source positions are non-zero
    sparse = math.ceil(nz_chunks / 2) / Bw
    return dense, sparse, Bw * Cout


if __name__ == "__main__":
    for cfg in [dict(Cin=1, Cout=8, W=16, Bw=8), dict(Cin=1, Cout=8, W=16, Bw=4), dict(Cin=8, Cout=16, W=16, Bw=4),
                dict(Cin=16, Cout=16, W=8, Bw=2), dict(Cin=3, Cout=5, W=12, Bw=4), dict(Cin=16, Cout=16, W=64, Bw=4, D=3, H=3)]:
        print("identity holds", cfg, "max err", check(**cfg))
    print("\nk-steps per output row and (kd,kh) group  [now | packed, kernel as is | packed + zero k-steps skipped], N per MMA")
    for name, Cin, Cout, Bw in [("1->8 @16^3 (fp32 FMA kernel today)", 1, 8, 8), ("8->16 @16^3", 8, 16, 4), ("16->16 @8^3", 16, 16, 4),
                                ("16->16 @64^3 (decoder)", 16, 16, 4), ("56->16 @8^3", 56, 16, 4)]:
        now = mma_model(0, Cin, Cout, 1)
        d, s, n = mma_model(0, Cin, Cout, Bw)
        print(f"  {name:36s} Bw={Bw}:  {now[0]:.2f} (N={now[2]})  |  {d:.2f}  |  {s:.2f}  (N={n})")
