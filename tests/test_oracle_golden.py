"""Pins the CPU oracle (oracle/rf_oracle.py) against the golden vectors that
tests/golden/make_golden.py produced by running the reference's own modules."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import cases as C
from oracle import rf_oracle as O

torch.set_grad_enabled(False)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(GOLD, "index.json")))
TOL = 1e-4  # north_star: fp32 values within 1e-4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth(shapes):
    return O.synth_state_dict(shapes, C.SEED)


def test_fold_unfold_hashes():
    f = INDEX["fold"]
    x = C.rnd("unfold.x1", (3, 1, 64, 64, 64)).numpy()
    assert sha(O.unfold3d(x, 16)) == f["unfold_16_1"]
    x = C.rnd("unfold.x2", (2, 16, 32, 32, 32)).numpy()
    assert sha(O.unfold3d(x, 8)) == f["unfold_8_16"]
    assert sha(O.unfold3d(x, 2)) == f["unfold_2_16"]
    x = C.rnd("unfold.x3", (2, 12, 32, 32, 32)).numpy()
    assert sha(O.unfold3d(x, 2)) == f["unfold_2_12"]
    assert sha(O.fold3d(C.rnd("fold.x1", (128, 16, 8, 8, 8)).numpy(), 4, 8, 16)) == f["fold_4_8_16"]
    assert sha(O.fold3d(C.rnd("fold.x2", (4096, 16, 2, 2, 2)).numpy(), 16, 2, 16)) == f["fold_16_2_16"]
    assert sha(O.fold3d(C.rnd("fold.x3", (128, 1, 16, 16, 16)).numpy(), 4, 16, 1)) == f["fold_4_16_1"]
    x = C.rnd("ups.x1", (3, 1, 8, 8, 8)).numpy()
    assert sha(O.unfold3d_pad_stride(x, 4, 1, np.float32(0.37), 2)) == f["padstride_4_1_2"]
    x = C.rnd("ups.x2", (2, 1, 16, 16, 16)).numpy()
    assert sha(O.unfold3d_pad_stride(x, 8, 2, np.float32(-1.5), 4)) == f["padstride_8_2_4"]
    x = C.rnd("ups.x3", (2, 1, 64, 64, 64)).numpy()
    assert sha(O.unfold3d_pad_stride(x, 32, 8, np.float32(2.25), 16)) == f["padstride_32_8_16"]
    assert sha(O.unfold3d_pad_stride(x, 24, 4, np.float32(2.25), 16)) == f["padstride_24_4_16"]
    p = O.PatcherOracle([16] * 3, [8] * 3, [16] * 3, 2.25, [64] * 3)
    pat = p(x)
    assert sha(pat) == f["patcher_16_8_16"]
    assert p.get_patch_counts() == f["patcher_counts"]
    assert sha(p.recompose_patches(x.shape, pat.reshape(2, 64, 32, 32, 32))) == f["patcher_recompose"]
    p2 = O.PatcherOracle([2] * 3, [1] * 3, [2] * 3, 0.5, [8] * 3)
    assert sha(p2(C.rnd("ups.x1", (3, 1, 8, 8, 8)).numpy())) == f["patcher_2_1_2"]


def test_fold_is_inverse_of_unfold():
    x = C.rnd("inv.x", (2, 6, 16, 16, 16)).numpy()
    for E in (2, 4, 8):
        assert np.array_equal(O.fold3d(O.unfold3d(x, E), 16 // E, E, 6), x)


def test_chunk_patches_equals_pad_unfold():
    """a1: dataset patch extraction == Unfold3DPadStride on the normalised chunk."""
    chunk = O.synthetic_tsdf(3, 8, 0.43334)
    trunc = float(np.float16(0.43334 * 3))
    m, s = 0.81, 0.51
    a = O.chunk_patches(chunk, 2, 1, 2, trunc, m, s)
    norm = ((chunk - m) / s).astype(np.float32)
    b = O.unfold3d_pad_stride(norm[None, None], 4, 1, np.float32((np.float32(trunc) - m) / s), 2)
    assert a.shape == (64, 1, 4, 4, 4)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)


@pytest.mark.parametrize("cls,nf,n", C.ENC_CASES)
def test_encoders(cls, nf, n):
    gold = np.load(os.path.join(GOLD, "encoders.npz"))[f"{cls}.{nf}"]
    sd = synth(O.encoder_param_shapes(cls, nf, 64))
    y = O.encoder_forward(cls, sd, C.encoder_input(cls, n)).reshape(n, 64).numpy()
    np.testing.assert_allclose(y, gold, rtol=0, atol=TOL)


@pytest.mark.parametrize("nf", [16, 12])
def test_retrieval_backbone(nf):
    gold = np.load(os.path.join(GOLD, "unets.npz"))[f"retrieval_backbone.{nf}"]
    sd = synth(O.retrieval_backbone_shapes(nf, nf, 4))
    y = O.retrieval_backbone_forward(C.retrieval_backbone_input(nf), sd, nf, nf, 4).numpy()
    np.testing.assert_allclose(y, gold, rtol=0, atol=TOL)


@pytest.mark.parametrize("kind,cls,nf,lv,S", C.UNET_CASES)
def test_unet_backbone(kind, cls, nf, lv, S):
    gold = np.load(os.path.join(GOLD, "unets.npz"))[f"unet_backbone.{kind}"]
    sd = synth(O.unet_backbone_shapes(kind, nf, lv))
    y = O.unet_backbone_forward(kind, C.unet_backbone_input(kind, S), sd, nf, lv)
    np.testing.assert_allclose(y[:, :, ::3, ::3, ::3].numpy(), gold, rtol=0, atol=TOL)
    st = INDEX[f"unet_backbone.{kind}.stats"]
    assert abs(float(y.double().abs().sum()) - st[1]) <= 1e-5 * st[1]


@pytest.mark.parametrize("nf", [16, 12])
def test_final_decoder(nf):
    gold = np.load(os.path.join(GOLD, "unets.npz"))[f"decoder.{nf}"]
    sd = synth(O.final_decoder_shapes(nf))
    y = O.final_decoder_forward(C.rnd(f"dec.{nf}.x", (1, nf, 32, 32, 32)), sd, nf)
    np.testing.assert_allclose(y[:, :, ::2, ::2, ::2].numpy(), gold, rtol=0, atol=TOL)


@pytest.mark.parametrize("nf,K,mode", C.ATTN_CASES)
def test_attention(nf, K, mode):
    g = np.load(os.path.join(GOLD, "attention.npz"))
    tag = C.attention_tag(nf, K, mode)
    sd = synth(O.attention_shapes(nf, 2))
    xb, xr, occ = C.attention_inputs(nf, K, mode)
    noise = torch.from_numpy(g[tag + ".noise"]) if mode else None
    y = O.patched_attention_forward(xb, xr, sd, nf, 16, 2, K, retrieval_mode=mode, gumbel_noise=noise)
    np.testing.assert_allclose(y[:, :, ::2, ::2, ::2].numpy(), g[tag], rtol=0, atol=TOL)
    if not mode and K == 4:
        xf, pf, of = O.attention_get_features(xb, xr[:1], occ, sd, nf, 2)
        np.testing.assert_allclose(xf[::16].numpy(), g[tag + ".feat_x"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(pf[::16].numpy(), g[tag + ".feat_p"], rtol=0, atol=1e-5)
        assert np.array_equal(of.numpy(), g[tag + ".feat_occ"])


@pytest.mark.parametrize("nf,K,mode", C.ATTN_MAPPING_CASES)
def test_attention_with_output_mapping(nf, K, mode):
    """attn_no_output_mapping=False: the g / o 1x1x1 convolutions (model/attention.py:56-57,95,108) against the
    reference's own PatchedAttentionBlock (tests/golden/make_golden_attention_mapping.py)."""
    g = np.load(os.path.join(GOLD, "attention_mapping.npz"))
    tag = C.attention_tag(nf, K, mode) + ".mapped"
    sd = synth(O.attention_shapes(nf, 2, output_mapping=True))
    xb, xr, _ = C.attention_inputs(nf, K, mode)
    y = O.patched_attention_forward(xb, xr, sd, nf, 16, 2, K, retrieval_mode=mode)
    np.testing.assert_allclose(y[:, :, ::2, ::2, ::2].numpy(), g[tag], rtol=0, atol=TOL)


def test_refine_full_forward():
    """BASELINE config 1: the parity anchor."""
    g = np.load(os.path.join(GOLD, "refine_full.npz"))
    x_in, x_re = C.refine_full_inputs()
    assert sha(x_in.numpy()) == INDEX["refine_full.input_sha"]
    assert sha(x_re.numpy()) == INDEX["refine_full.retrieval_sha"]
    sds = dict(unet_backbone=synth(O.unet_backbone_shapes("sr08", 16, 4)),
               retrieval_backbone=synth(O.retrieval_backbone_shapes(16, 16, 4)),
               attention=synth(O.attention_shapes(16, 2)), decoder=synth(O.final_decoder_shapes(16)))
    cfg = dict(kind="sr08", nf=16, unet_num_level=4, retrieval_fmaps=16, retrieval_num_level=4, K=4, E=2)
    pred, x_back, x_retr, xa = O.refine_forward(x_in, x_re, sds, cfg)
    np.testing.assert_allclose(x_back[:, :, ::4, ::4, ::4].numpy(), g["x_back"], rtol=0, atol=TOL)
    np.testing.assert_allclose(x_retr[:, :, ::4, ::4, ::4].numpy(), g["x_retr"], rtol=0, atol=TOL)
    np.testing.assert_allclose(xa[:, :, ::4, ::4, ::4].numpy(), g["x_attn"], rtol=0, atol=TOL)
    np.testing.assert_allclose(pred.numpy(), g["pred"], rtol=0, atol=TOL)
    # the output must not be degenerate, or 1e-4 would be vacuous
    assert float(np.abs(g["pred"]).max()) > 0.2 and float(g["pred"].std()) > 0.05


# --------------------------------------------------------------------------- SURVEY 8f rows

def test_oracle_adjuncts_against_reference_golden():
    """NT-Xent (model/loss.py:48-69, produced by the reference's own class) and compute_normals
    (dataset/patched_scene_dataset.py:139-146, Sobel literals parsed from the reference source) golden vectors."""
    import numpy as np
    import torch
    from oracle import rf_oracle as O
    g = np.load(os.path.join(GOLD, "adjuncts.npz"))
    for tag in ("cos", "dot", "cos_iou", "cos_big"):
        temp, cosine = float(g[f"ntxent.{tag}.cfg"][0]), bool(g[f"ntxent.{tag}.cfg"][1])
        iou = torch.from_numpy(g[f"ntxent.{tag}.iou"]) if f"ntxent.{tag}.iou" in g else None
        got = float(O.ntxent_loss(torch.from_numpy(g[f"ntxent.{tag}.zis"]), torch.from_numpy(g[f"ntxent.{tag}.zjs"]), temp, cosine, iou))
        assert abs(got - float(g[f"ntxent.{tag}.loss"])) <= 1e-6 * max(1.0, abs(got)), tag
    nrm = O.compute_normals(torch.from_numpy(g["normals.target"]), float(g["normals.trunc"])).numpy()
    assert np.array_equal(nrm, g["normals.out"])
    assert float(np.abs(np.linalg.norm(nrm, axis=1)).max()) <= 1.0 + 1e-6


def test_oracle_metrics_small_cases():
    """util/metrics.py semantics on hand-checkable volumes: empty union skipped by IoU, first-index ties and the
    empty-cloud skip of Chamfer3D."""
    import numpy as np
    from oracle import rf_oracle as O
    p = np.zeros((3, 1, 4, 4, 4), dtype=bool)
    t = np.zeros((3, 1, 4, 4, 4), dtype=bool)
    p[0, 0, 0, 0, :2] = True
    t[0, 0, 0, 0, 1:4] = True          # inter 1, union 4
    p[1, 0, 1, 1, 1] = True            # target empty: union 1, inter 0
    iou_sum, iou_n, prec, rec, counts = O.occupancy_metrics(p, t)
    assert counts.tolist() == [[1, 4, 2, 3], [0, 1, 1, 0], [0, 0, 0, 0]]
    assert iou_n == 2 and abs(iou_sum - 1 / (4 + 1e-5)) < 1e-6
    assert abs(prec - (1 / (2 + 1e-5))) < 1e-6 and abs(rec - (1 / (3 + 1e-5))) < 1e-6
    d, i = O.chamfer_nn(np.array([[0, 0, 0], [5, 5, 5]], dtype=np.float32), np.array([[1, 0, 0], [0, 1, 0], [5, 5, 4]], dtype=np.float32))
    assert d.tolist() == [1.0, 1.0] and i.tolist() == [0, 2]   # tie -> first index
    cd, valid = O.chamfer_metric(p, t)
    assert valid == 1 and cd > 0


# --------------------------------------------------------------------------- properties of the oracle (hypothesis)

def test_oracle_reindex_properties():
    """Random shapes: the oracle's Unfold3DPadStride equals torch's pad + unfold (the reference's own formulation,
    model/attention.py:200-203), Fold3D inverts Unfold3D, Patcher.recompose inverts Patcher when the patches cover
    the padded volume."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 3), st.integers(1, 4), st.integers(1, 4), st.integers(0, 3), st.integers(0, 2 ** 31 - 1))
    def prop(B, C_, cnt, stride, pad, seed):
        rng = np.random.default_rng(seed)
        kern = stride + int(rng.integers(0, 3))                      # overlapping or not
        size = (cnt - 1) * stride + kern - 2 * pad
        if size < 1:
            return
        x = rng.normal(size=(B, C_, size, size, size)).astype(np.float32)
        got = O.unfold3d_pad_stride(x, kern, pad, -2.5, stride)
        xt = torch.nn.functional.pad(torch.from_numpy(x), (pad,) * 6, value=-2.5)
        w = xt.unfold(2, kern, stride).unfold(3, kern, stride).unfold(4, kern, stride)
        want = w.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(-1, C_, kern, kern, kern).numpy()
        assert np.array_equal(got.reshape(want.shape), want)
    prop()

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 3), st.integers(1, 5), st.integers(1, 4), st.sampled_from([1, 2, 3, 4]), st.integers(0, 2 ** 31 - 1))
    def inv(B, C_, R, E, seed):
        x = np.random.default_rng(seed).normal(size=(B, C_, R * E, R * E, R * E)).astype(np.float32)
        u = O.unfold3d(x, E)
        assert u.shape == (B * R ** 3, C_, E, E, E)
        assert np.array_equal(O.fold3d(u, R, E, C_), x)
    inv()


def test_row_permutations_the_inference_shortcuts_rely_on():
    """Index identities behind two GPU shortcuts, stated on the oracle's fold / unfold (pure re-indexing, CPU):
    (1) rf_attention_fuse_patched_fwd: Unfold3D(E) of Fold3D(P, s, nf)(patches) is a permutation of the rows of Unfold3D(E)
        applied to the un-folded patch batch - row (b, px, py, pz) of the former is row
        ((b * P^3 + patch) * (s/E)^3 + local) of the latter (csrc/rf_attention.cu, AttnGeo with P > 1);
    (2) rf_compose_gather_patches: destination blocks of edge 16 that tile the chunk in x-major order, stored
        contiguously, are Unfold3D(16, 1) of the composed volume (trainer/train_refinement.py:110-111)."""
    rng = np.random.default_rng(5)
    for nf, P, s, E, BK in ((16, 4, 8, 2, 3), (12, 2, 4, 2, 2), (4, 2, 8, 4, 1)):
        patches = rng.random((BK * P ** 3, nf, s, s, s)).astype(np.float32)
        vol_rows = O.unfold3d(O.fold3d(patches, P, s, nf), E)          # what the reference's attention sees
        pat_rows = O.unfold3d(patches, E)                               # what the shortcut unfolds
        Rp, ps = P * s // E, s // E
        r = np.arange(Rp ** 3)
        px, py, pz = r // (Rp * Rp), (r // Rp) % Rp, r % Rp
        mapped = (((px // ps) * P + py // ps) * P + pz // ps) * ps ** 3 + ((px % ps) * ps + py % ps) * ps + pz % ps
        for b in range(BK):
            assert np.array_equal(vol_rows[b * Rp ** 3 + r], pat_rows[b * Rp ** 3 + mapped])
    vol = rng.random((2, 1, 64, 64, 64)).astype(np.float32)
    blocks = np.stack([vol[:, 0, 16 * x:16 * x + 16, 16 * y:16 * y + 16, 16 * z:16 * z + 16]
                       for x in range(4) for y in range(4) for z in range(4)], axis=1)   # [B, 64 blocks (x-major), 16,16,16]
    assert np.array_equal(blocks.reshape(-1, 1, 16, 16, 16), O.unfold3d(vol, 16))


def test_oracle_knn_and_demotion_properties():
    """Random small banks: the C oracle equals a brute-force float64 sort under (d, id); demotion is a stable
    partition (util/retrieval.py:94-97) that keeps K entries and never reorders inside the two classes."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=30, deadline=None)
    @given(st.integers(1, 60), st.integers(1, 12), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
    def prop(N, Q, k, seed):
        rng = np.random.default_rng(seed)
        k = min(k, N)
        db = rng.integers(-2, 3, size=(N, 64)).astype(np.float32)        # small integers: many exact ties
        q = rng.integers(-2, 3, size=(Q, 64)).astype(np.float32)
        idx, dist = O.knn_exact(db, q, k)
        d = ((q[:, None, :].astype(np.float64) - db[None].astype(np.float64)) ** 2).sum(-1)
        for i in range(Q):
            order = sorted(range(N), key=lambda j: (d[i, j], j))[:k]
            assert idx[i].tolist() == order
            assert np.array_equal(dist[i], d[i, order].astype(np.float32))
        if k >= 2:
            K = k // 2
            row_scene = rng.integers(0, 3, size=N)
            qs = rng.integers(-1, 3, size=Q)
            di, dd = O.demote_same_scene(idx, dist, row_scene, qs, K)
            for i in range(Q):
                other = [j for j in idx[i] if row_scene[j] != qs[i]]
                same = [j for j in idx[i] if row_scene[j] == qs[i]]
                assert di[i].tolist() == (other + same)[:K]
    prop()


def test_evaluation_metrics_match_reference_classes():
    """IoU / Chamfer3D / Precision / Recall as the reference's own util/metrics.py classes compute them
    (tests/golden/make_golden_adjuncts.py ran them with a stubbed torchmetrics base and the submodule's pure-torch
    distChamfer), and the nearest-neighbour search on voxel coordinates bit-exact."""
    g = np.load(os.path.join(GOLD, "adjuncts.npz"))
    p, t = g["metrics.pred"], g["metrics.target"]
    iou_sum, iou_n, prec, rec, _ = O.occupancy_metrics(p, t)
    cd, valid = O.chamfer_metric(p, t)
    mine = np.array([iou_sum / iou_n, cd / valid, prec / p.shape[0], rec / p.shape[0]])
    np.testing.assert_allclose(mine, g["metrics.values"], rtol=1e-5, atol=1e-7)
    a, b = np.argwhere(t[0, 0]).astype(np.float32), np.argwhere(p[0, 0]).astype(np.float32)
    d1, i1 = O.chamfer_nn(a, b)
    d2, i2 = O.chamfer_nn(b, a)
    assert np.array_equal(d1, g["chamfer.d1"]) and np.array_equal(d2, g["chamfer.d2"])
    assert np.array_equal(i1, g["chamfer.i1"]) and np.array_equal(i2, g["chamfer.i2"])
