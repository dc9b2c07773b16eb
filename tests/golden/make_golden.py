"""Generates tests/golden/*.npz|json by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

It imports the reference's `model/*.py` and `util/patcher.py` unmodified, loads
the deterministic synthetic weights of `oracle.rf_oracle.synth_state_dict`
into them, feeds seeded inputs and stores the outputs.  The committed files are
what pins the oracle (tests/test_oracle_golden.py) and, on the GPU box where
/root/reference does not exist, the CUDA path (tests/test_gpu_*.py).

kNN / demotion / compose have no reference-executable counterpart here
(pyflann absent) and therefore no golden file: parity unpinned for those.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

sys.path.insert(0, HERE)
from oracle import rf_oracle as O  # noqa: E402
import cases as C  # noqa: E402

import model as ref_model  # noqa: E402  (reference)
from model import attention as ref_attention  # noqa: E402
from model import refinement as ref_refinement  # noqa: E402
from model import retrieval as ref_retrieval  # noqa: E402
from util.patcher import Patcher as RefPatcher  # noqa: E402

torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count())

SEED = C.SEED


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


rnd = C.rnd


def load_synth(module, shapes_fn_result=None):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    if shapes_fn_result is not None:
        assert {k: tuple(v) for k, v in shapes_fn_result.items()} == shapes, "oracle shape table != reference state_dict"
    module.load_state_dict(O.synth_state_dict(shapes, SEED))
    module.eval()
    return shapes


def stats(t: torch.Tensor):
    d = t.double()
    return [float(d.sum()), float(d.abs().sum()), float(d.abs().max())]


def main():
    keys = {}  # state_dict key tables, for checking the product's modules off-box
    index = {}  # small json: hashes + stats

    # ---- a2-a4 fold / unfold / patcher: hashes of the reference's outputs ----
    fold = {}
    x = rnd("unfold.x1", (3, 1, 64, 64, 64))
    fold["unfold_16_1"] = sha(ref_attention.Unfold3D(16, 1)(x).numpy())
    x = rnd("unfold.x2", (2, 16, 32, 32, 32))
    fold["unfold_8_16"] = sha(ref_attention.Unfold3D(8, 16)(x).numpy())
    fold["unfold_2_16"] = sha(ref_attention.Unfold3D(2, 16)(x).numpy())
    x = rnd("unfold.x3", (2, 12, 32, 32, 32))
    fold["unfold_2_12"] = sha(ref_attention.Unfold3D(2, 12)(x).numpy())
    x = rnd("fold.x1", (2 * 64, 16, 8, 8, 8))
    fold["fold_4_8_16"] = sha(ref_attention.Fold3D(4, 8, 16)(x).contiguous().numpy())
    x = rnd("fold.x2", (1 * 4096, 16, 2, 2, 2))
    fold["fold_16_2_16"] = sha(ref_attention.Fold3D(16, 2, 16)(x).contiguous().numpy())
    x = rnd("fold.x3", (2 * 64, 1, 16, 16, 16))
    fold["fold_4_16_1"] = sha(ref_attention.Fold3D(4, 16, 1)(x).contiguous().numpy())
    x = rnd("ups.x1", (3, 1, 8, 8, 8))
    fold["padstride_4_1_2"] = sha(ref_attention.Unfold3DPadStride(4, 1, 0.37, 2)(x).numpy())
    x = rnd("ups.x2", (2, 1, 16, 16, 16))
    fold["padstride_8_2_4"] = sha(ref_attention.Unfold3DPadStride(8, 2, -1.5, 4)(x).numpy())
    x = rnd("ups.x3", (2, 1, 64, 64, 64))
    fold["padstride_32_8_16"] = sha(ref_attention.Unfold3DPadStride(32, 8, 2.25, 16)(x).numpy())
    fold["padstride_24_4_16"] = sha(ref_attention.Unfold3DPadStride(24, 4, 2.25, 16)(x).numpy())
    p = RefPatcher([16] * 3, [8] * 3, [16] * 3, 2.25, [64] * 3)
    pat = p(x)
    fold["patcher_16_8_16"] = sha(pat.numpy())
    fold["patcher_counts"] = p.get_patch_counts()
    # recompose takes [B, n_patches, k,k,k]
    rec_in = pat.reshape(2, 64, 32, 32, 32)
    fold["patcher_recompose"] = sha(p.recompose_patches(x.shape, rec_in).numpy())
    p2 = RefPatcher([2, 2, 2], [1, 1, 1], [2, 2, 2], 0.5, [8, 8, 8])
    x8 = rnd("ups.x1", (3, 1, 8, 8, 8))
    fold["patcher_2_1_2"] = sha(p2(x8).numpy())
    index["fold"] = fold

    # ---- a5-a8 patch encoders ----
    enc_cases = C.ENC_CASES
    enc_out = {}
    for cls, nf, n in enc_cases:
        m = getattr(ref_retrieval, cls)(nf, 64)
        keys[f"enc.{cls}.{nf}"] = load_synth(m, O.encoder_param_shapes(cls, nf, 64))
        x = C.encoder_input(cls, n)
        y = m(x.clone())  # clone: LeakyReLU(inplace=True) on conv outputs only, but be safe
        enc_out[f"{cls}.{nf}"] = y.reshape(n, 64).numpy()
    np.savez_compressed(os.path.join(HERE, "encoders.npz"), **enc_out)

    # ---- a13 retrieval U-Net ----
    unet_out = {}
    for nf, fm in ((16, 16), (12, 12)):
        m = ref_model.get_retrieval_backbone(dict(nf=nf, retrieval_fmaps=fm, retrieval_num_level=4, layer_order="gcr"))
        keys[f"retrieval_backbone.{nf}"] = load_synth(m, O.retrieval_backbone_shapes(nf, fm, 4))
        assert m.nf == nf
        x = C.retrieval_backbone_input(nf)
        unet_out[f"retrieval_backbone.{nf}"] = m(x).numpy()

    # ---- a15 input U-Nets ----
    for kind, cls, nf, lv, S in C.UNET_CASES:
        m = getattr(ref_refinement, cls)(nf, num_levels=lv, layer_order="gcr")
        keys[f"unet_backbone.{kind}"] = load_synth(m, O.unet_backbone_shapes(kind, nf, lv))
        x = C.unet_backbone_input(kind, S)
        y = m(x)
        assert y.shape[1:] == (nf, 32, 32, 32)
        unet_out[f"unet_backbone.{kind}"] = y[:, :, ::3, ::3, ::3].contiguous().numpy()
        index[f"unet_backbone.{kind}.stats"] = stats(y)

    # ---- a16 final decoder ----
    for nf in (16, 12):
        m = ref_refinement.Superresolution08FinalDecoder(nf, layer_order="gcr")
        keys[f"decoder.{nf}"] = load_synth(m, O.final_decoder_shapes(nf))
        x = rnd(f"dec.{nf}.x", (1, nf, 32, 32, 32))
        y = m(x)
        unet_out[f"decoder.{nf}"] = y[:, :, ::2, ::2, ::2].contiguous().numpy()
        index[f"decoder.{nf}.stats"] = stats(y)
    np.savez_compressed(os.path.join(HERE, "unets.npz"), **unet_out)

    # ---- a14 attention ----
    attn_out = {}
    for nf, K, mode in C.ATTN_CASES:
        cfg = dict(nf=nf, attn_patch_extent=4, K=K, attn_normalize=True, attn_use_switching=True,
                   attn_retrieval_mode=mode, attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16)
        m = ref_model.get_attention_block(cfg)
        tag = C.attention_tag(nf, K, mode)
        keys[tag] = load_synth(m, O.attention_shapes(nf, 2))
        B = 1
        xb, xr, occ = C.attention_inputs(nf, K, mode)
        if mode:
            # the reference draws the Gumbel noise inside gumbel_softmax: record
            # the identical draw by re-seeding (first RNG use in the forward)
            torch.manual_seed(77)
            noise = -torch.empty(B * 4096, K).exponential_().log()
            attn_out[tag + ".noise"] = noise.numpy()
            torch.manual_seed(77)
        y = m(xb, xr).contiguous()
        attn_out[tag] = y[:, :, ::2, ::2, ::2].contiguous().numpy()
        index[tag + ".stats"] = stats(y)
        if not mode and K == 4:
            xf, pf, of = m.get_features(xb, xr[:B], occ)
            attn_out[tag + ".feat_x"] = xf[::16].numpy()
            attn_out[tag + ".feat_p"] = pf[::16].numpy()
            attn_out[tag + ".feat_occ"] = of.numpy()
    np.savez_compressed(os.path.join(HERE, "attention.npz"), **attn_out)

    # ---- a17 full refine forward, BASELINE config 1 (SR 8->64, K=4, B=1) ----
    # config/super_resolution/3DFront/refinement_008_064.yaml
    cfg = C.REFINE_CFG
    unet_backbone = ref_model.get_unet_backbone(cfg)
    decoder = ref_model.get_decoder(cfg)
    retrieval_backbone = ref_model.get_retrieval_backbone(cfg)
    attn = ref_model.get_attention_block(cfg)
    for m in (unet_backbone, decoder, retrieval_backbone, attn):
        load_synth(m)
    x_in, x_re = C.refine_full_inputs()
    unfold_shape = ref_attention.Unfold3D(16, 1)
    fold_features = ref_attention.Fold3D(4, 8, retrieval_backbone.nf)
    x_back = unet_backbone(x_in)
    retrievals = x_re[:, :4].reshape(4, 1, 64, 64, 64)  # get_retrievals
    x_retr = fold_features(retrieval_backbone(unfold_shape(retrievals)))
    xa = attn(x_back, x_retr)
    pred = decoder(xa)
    full = dict(pred=pred.numpy(), x_back=x_back[:, :, ::4, ::4, ::4].contiguous().numpy(),
                x_retr=x_retr[:, :, ::4, ::4, ::4].contiguous().numpy(),
                x_attn=xa[:, :, ::4, ::4, ::4].contiguous().numpy())
    # inputs are regenerated from seeds in the tests (cases.py); store only hashes
    index["refine_full.retrieval_sha"] = sha(x_re.numpy())
    index["refine_full.input_sha"] = sha(x_in.numpy())
    index["refine_full.pred_stats"] = stats(pred)
    np.savez_compressed(os.path.join(HERE, "refine_full.npz"), **full)

    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({k: {n: list(s) for n, s in v.items()} for k, v in keys.items()}, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
