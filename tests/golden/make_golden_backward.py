"""Golden GRADIENTS for SURVEY 8f.3, produced by torch autograd on the REFERENCE's own modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_backward.py      ->  tests/golden/backward.npz

The reference's `model/unet.py`, `model/attention.py`, `model/refinement.py` are imported unmodified, loaded with the
deterministic synthetic weights of `oracle.rf_oracle.synth_state_dict`, and differentiated twice: in float64 (the
yardstick) and in float32 (the reference arithmetic's own distance from the yardstick, which is what a correct fp32
implementation is allowed).  Stored per tensor: a strided sample of the fp64 gradient (as fp32), its max-abs, and the
max-abs distance of the fp32 gradient from it.  Cases = tests/golden/backward_cases.py.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import backward_cases as BC  # noqa: E402
from oracle import rf_oracle as O  # noqa: E402

import model as ref_model  # noqa: E402  (reference)
from model import attention as ref_attention  # noqa: E402
from model import unet as ref_unet  # noqa: E402

torch.set_num_threads(os.cpu_count())
OUT = {}


def record(case, name, g64, g32):
    g64 = g64.detach().double().reshape(-1)
    g32 = g32.detach().double().reshape(-1)
    OUT[f"{case}/{name}/sample"] = g64[::BC.stride(g64.numel())].float().numpy()
    OUT[f"{case}/{name}/scale"] = np.float64(g64.abs().max())
    OUT[f"{case}/{name}/noise"] = np.float64((g32 - g64).abs().max())


def run(case, build, inputs, gout, fwd):
    """build() -> module (fresh, synthetic weights loaded); fwd(module, *inputs) -> output."""
    res = {}
    for dt in (torch.float64, torch.float32):
        m = build().to(dt).train()
        ins = [None if t is None else t.to(dt).requires_grad_(True) for t in inputs]
        out = fwd(m, *ins)
        (out * gout.to(dt)).sum().backward()
        res[dt] = (m, ins, out.detach())
    m64, in64, out64 = res[torch.float64]
    m32, in32, _ = res[torch.float32]
    OUT[f"{case}/out/sample"] = out64.reshape(-1)[::BC.stride(out64.numel())].float().numpy()
    for i, (a, b) in enumerate(zip(in64, in32)):
        if a is not None and a.grad is not None:
            record(case, f"input{i}", a.grad, b.grad)
    p32 = dict(m32.named_parameters())
    for name, p in m64.named_parameters():
        if p.grad is not None:
            record(case, f"param/{name}", p.grad, p32[name].grad)
    print(case, "done:", sum(1 for k in OUT if k.startswith(case + "/") and k.endswith("/scale")), "tensors")


def load(module, shapes=None):
    sd = O.synth_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, BC.SEED)
    module.load_state_dict(sd)
    return module


def main():
    # ---- SingleConv (GroupNorm -> Conv3d -> ReLU) on concat(x, up2(x2))
    for (c1, c2, cout, S) in BC.SINGLE_CONV:
        case = BC.single_conv_tag(c1, c2, cout, S)
        x, x2, gamma, beta, gout = BC.single_conv_inputs(c1, c2, cout, S)

        def build():
            m = load(ref_unet.SingleConv(c1 + c2, cout, 3, "gcr", 8))
            with torch.no_grad():
                m.groupnorm.weight.copy_(gamma)
                m.groupnorm.bias.copy_(beta)
            return m

        def fwd(m, a, b):
            parts = ([a] if a is not None else []) + ([F.interpolate(b, scale_factor=2, mode="nearest")] if b is not None else [])
            return m(torch.cat(parts, 1))

        run(case, build, (x, x2), gout, fwd)

    # ---- RetrievalUNetBackbone
    nf = 16
    x, gout = BC.retrieval_unet_inputs(nf)
    run("retrieval_unet", lambda: load(ref_model.get_retrieval_backbone(dict(nf=nf, retrieval_fmaps=16, retrieval_num_level=4, layer_order="gcr"))),
        (x,), gout, lambda m, a: m(a))

    # ---- PatchedAttentionBlock, softmax and Gumbel mode
    for mode in (False, True):
        xb, xr, gout, noise = BC.attention_inputs(mode)

        def fwd(m, a, b):
            if mode:
                # gumbel_softmax draws its noise from the global generator: replay the recorded draw
                orig = torch.Tensor.exponential_
                state = {"n": noise}

                def fake_exponential_(self, *args, **kw):
                    # torch's gumbel_softmax: gumbels = -empty_like(logits).exponential_().log()
                    return self.copy_(torch.exp(-state["n"].to(self.dtype)))
                torch.Tensor.exponential_ = fake_exponential_
                try:
                    return m(a, b)
                finally:
                    torch.Tensor.exponential_ = orig
            return m(a, b)

        run(BC.attention_tag(mode), lambda: load(ref_model.get_attention_block(BC.attention_cfg(mode))), (xb, xr), gout, fwd)

    # ---- the inference part of forward_full (train_refinement.py:108-116) as training_step_full differentiates it
    cfg = BC.REFINE_CFG
    x_in, x_re, gout = BC.refine_inputs()

    class Full(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.unet_backbone = load(ref_model.get_unet_backbone(cfg))
            self.retrieval_backbone = load(ref_model.get_retrieval_backbone(cfg))
            self.attention = load(ref_model.get_attention_block(cfg))
            self.decoder = load(ref_model.get_decoder(cfg))
            self.unfold_shape = ref_attention.Unfold3D(16, 1)
            self.fold_features = ref_attention.Fold3D(4, 8, self.retrieval_backbone.nf)

        def forward(self, inp):
            x_back = self.unet_backbone(inp)
            retrievals = x_re.to(inp.dtype)[:, :4].reshape(4, 1, 64, 64, 64)
            x_retrieval = self.fold_features(self.retrieval_backbone(self.unfold_shape(retrievals)))
            return self.decoder(self.attention(x_back, x_retrieval))

    run("refine_full", Full, (x_in,), gout, lambda m, a: m(a))
    np.savez_compressed(os.path.join(HERE, "backward.npz"), **OUT)
    print("backward.npz", os.path.getsize(os.path.join(HERE, "backward.npz")))


if __name__ == "__main__":
    main()
