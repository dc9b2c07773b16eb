"""backward passes of the drop-in modules against gradients that torch autograd computed on the REFERENCE's
own modules (tests/golden/make_golden_backward.py -> tests/golden/backward.npz: fp64 gradients as the yardstick, plus
the distance of the reference's own fp32 gradients from them).  trainer/train_refinement.py:74-89 (training_step_full)
back-propagates through exactly these modules.

Gate, as for the forward tests: a gradient may be no further from the fp64 gradient than a few times the reference
arithmetic's own fp32 distance from it, plus 1e-4 of the gradient's max-abs."""
import os

import numpy as np
import pytest
import torch

import backward_cases as BC
import cases as C
from oracle import rf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backward.npz"))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _grad_on():
    """Other test modules switch autograd off globally at import time (inference tests); these tests need it on."""
    with torch.enable_grad():
        yield


def synth_for(module):
    return O.synth_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, BC.SEED)


def check(case, name, got, slack=4.0):
    key = f"{case}/{name}"
    assert key + "/sample" in GOLD.files, f"no golden gradient for {key}"
    got = got.detach().cpu().double().reshape(-1)
    want = torch.from_numpy(GOLD[key + "/sample"]).double()
    got = got[::BC.stride(got.numel())]
    assert got.shape == want.shape, (key, got.shape, want.shape)
    scale, noise = float(GOLD[key + "/scale"]), float(GOLD[key + "/noise"])
    err = float((got - want).abs().max())
    assert err <= slack * noise + 1e-4 * scale + 1e-30, \
        f"{key}: |ours - reference fp64| = {err:.3e}; the reference's own fp32 gradient is {noise:.3e} away; scale {scale:.3e}"
    return err / (scale + 1e-30)


def check_out(case, out, tol):
    got = out.detach().cpu().double().reshape(-1)
    got = got[::BC.stride(got.numel())]
    want = torch.from_numpy(GOLD[f"{case}/out/sample"]).double()
    err = float((got - want).abs().max())
    assert err <= tol * max(1.0, float(want.abs().max())), f"{case}: forward of the differentiable path differs by {err:.3e}"


@pytest.mark.parametrize("c1,c2,cout,S", BC.SINGLE_CONV)
def test_single_conv_backward(dev, c1, c2, cout, S):
    """GroupNorm -> Conv3d(k3, p1) -> ReLU on concat(x, nearest_up2(x2)) (model/unet.py:79-100 + Decoder join :303-306):
    input, filter, gamma and beta gradients."""
    from retrieval_fuse_b200.model.unet import SingleConv
    case = BC.single_conv_tag(c1, c2, cout, S)
    x, x2, gamma, beta, gout = BC.single_conv_inputs(c1, c2, cout, S)
    m = SingleConv(c1 + c2, cout, 3, "gcr", 8)
    m.load_state_dict(synth_for(m))
    with torch.no_grad():
        m.groupnorm.weight.copy_(gamma)
        m.groupnorm.bias.copy_(beta)
    m = m.to(dev)
    xd = None if x is None else x.to(dev).requires_grad_(True)
    x2d = None if x2 is None else x2.to(dev).requires_grad_(True)
    y = m(xd, x2d)
    (y * gout.to(dev)).sum().backward()
    check_out(case, y, 1e-4)
    if xd is not None:
        check(case, "input0", xd.grad)
    if x2d is not None:
        check(case, "input1", x2d.grad)
    for name, p in m.named_parameters():
        check(case, f"param/{name}", p.grad)


def test_retrieval_unet_backward(dev):
    """RetrievalUNetBackbone (model/refinement.py:64-73): every parameter gradient - 12 GroupNorm + conv + ReLU blocks,
    three max-pools, two nearest-upsample + concat joins."""
    from retrieval_fuse_b200.model import get_retrieval_backbone
    nf = 16
    m = get_retrieval_backbone(dict(nf=nf, retrieval_fmaps=16, retrieval_num_level=4, layer_order="gcr"))
    m.load_state_dict(synth_for(m))
    m = m.to(dev)
    x, gout = BC.retrieval_unet_inputs(nf)
    xd = x.to(dev).requires_grad_(True)
    y = m(xd)
    (y * gout.to(dev)).sum().backward()
    check_out("retrieval_unet", y, 1e-4)
    worst = check("retrieval_unet", "input0", xd.grad)
    for name, p in m.named_parameters():
        worst = max(worst, check("retrieval_unet", f"param/{name}", p.grad))
    print(f"retrieval U-Net backward: worst relative gradient error {worst:.2e}")


@pytest.mark.parametrize("mode", [False, True])
def test_patched_attention_backward(dev, mode):
    """PatchedAttentionBlock (model/attention.py:141-157): gradients of both inputs and of the theta / phi MLPs, in
    softmax mode and in hard-Gumbel (straight-through) mode with the reference's noise draw replayed."""
    from retrieval_fuse_b200.model import get_attention_block
    case = BC.attention_tag(mode)
    m = get_attention_block(BC.attention_cfg(mode))
    m.load_state_dict(synth_for(m))
    m = m.to(dev)
    xb, xr, gout, noise = BC.attention_inputs(mode)
    xbd, xrd = xb.to(dev).requires_grad_(True), xr.to(dev).requires_grad_(True)
    y = m(xbd, xrd, None if noise is None else noise.to(dev))
    (y * gout.to(dev)).sum().backward()
    check_out(case, y, 1e-3)
    check(case, "input0", xbd.grad, slack=6.0)
    check(case, "input1", xrd.grad, slack=6.0)
    for name, p in m.named_parameters():
        if "sig_" in name:  # unused in the reference's forward (:97-99): no gradient there either
            assert f"{case}/param/{name}/sample" not in GOLD.files
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        check(case, f"param/{name}", p.grad, slack=6.0)


def test_refine_training_step_backward(dev):
    """The inference part of forward_full (train_refinement.py:108-116) with autograd recording, as training_step_full
    differentiates it: a loss on pred_shape -> gradients of every parameter of the input U-Net, the retrieval U-Net,
    the attention block and the decoder."""
    from retrieval_fuse_b200.pipeline import FRONT3D_SR, RefinementPipeline
    pipe = RefinementPipeline(FRONT3D_SR, bank=None, device=dev)
    mods = dict(unet_backbone=pipe.unet_backbone, retrieval_backbone=pipe.retrieval_backbone,
                attention=pipe.patched_attention_block, decoder=pipe.decoder)
    for m in mods.values():
        m.load_state_dict(synth_for(m))
        m.train()
        for p in m.parameters():
            p.requires_grad_(True)
    x_in, x_re, gout = BC.refine_inputs()
    pred, _, _, _ = pipe.refine_train(x_in.to(dev), x_re.to(dev))
    (pred * gout.to(dev)).sum().backward()
    check_out("refine_full", pred, 2e-3)
    with torch.no_grad():  # the inference (tensor-core) path on the same weights gives the same prediction
        pred_inf = pipe.refine(x_in.to(dev), x_re.to(dev))[0]
    assert float((pred.detach() - pred_inf).abs().max()) <= 2e-3
    worst, n = 0.0, 0
    for k, m in mods.items():
        for name, p in m.named_parameters():
            if "sig_" in name:
                continue
            assert p.grad is not None, f"{k}.{name} received no gradient"
            worst = max(worst, check("refine_full", f"param/{k}.{name}", p.grad, slack=6.0))
            n += 1
    print(f"training step backward: {n} parameter tensors, worst relative gradient error {worst:.2e}")
    # an optimizer step on these gradients changes the prediction (the modules are trainable end to end)
    opt = torch.optim.SGD([p for m in mods.values() for p in m.parameters() if p.grad is not None], lr=1e-3)
    opt.step()
    with torch.no_grad():
        pred2 = pipe.refine(x_in.to(dev), x_re.to(dev))[0]
    assert float((pred2 - pred_inf).abs().max()) > 0
