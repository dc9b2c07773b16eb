"""CPU-side checks: the C-ABI library loads and exports every symbol that
include/rf_b200.h declares, the drop-in modules expose the reference's
state_dict keys, host-side helpers behave like the reference's, and the
product path refuses to run without CUDA (no fallback)."""
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import rf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from retrieval_fuse_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "librf_b200.so not built: run __graft_entry__.build()"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/rf_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES.keys()) == declared, "ctypes prototype table out of sync with the header"
    L = _lib.lib()
    assert L.rf_version() >= 100
    assert L.rf_attention_workspace_bytes(1, 16, 32, 2, 4) > 0
    assert L.rf_knn_workspace_bytes(64, 131073, 8, 1) > 0


def test_no_cpu_fallback():
    from retrieval_fuse_b200 import _lib, ops
    with pytest.raises(_lib.RfError):
        ops.unfold3d(torch.zeros(1, 1, 4, 4, 4), 2)
    with pytest.raises(_lib.RfError):
        ops.knn_topk(torch.zeros(8, 64), torch.zeros(2, 64), 1)


def test_forward_only_is_loud():
    from retrieval_fuse_b200 import ops
    x = torch.zeros(1, 1, 4, 4, 4, requires_grad=True)
    with torch.enable_grad(), pytest.raises(NotImplementedError):
        ops.unfold3d(x, 2)


def test_state_dict_keys_match_reference():
    import retrieval_fuse_b200.model as M
    from retrieval_fuse_b200.model import retrieval as R
    keys = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))

    def chk(tag, m):
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == keys[tag], tag

    for tag in keys:
        if tag.startswith("enc."):
            _, cls, nf = tag.split(".")
            chk(tag, getattr(R, cls)(int(nf), 64))
    rb = M.get_retrieval_backbone(dict(nf=16, retrieval_fmaps=16, retrieval_num_level=4, layer_order="gcr"))
    assert rb.nf == 16
    chk("retrieval_backbone.16", rb)
    chk("retrieval_backbone.12", M.get_retrieval_backbone(dict(nf=12, retrieval_fmaps=12, retrieval_num_level=4, layer_order="gcr")))
    cfg = dict(task="superresolution", nf=16, unet_num_level=4, layer_order="gcr", dataset_train=dict(input_chunk_size=8))
    chk("unet_backbone.sr08", M.get_unet_backbone(cfg))
    cfg["dataset_train"]["input_chunk_size"] = 16
    chk("unet_backbone.sr16", M.get_unet_backbone(cfg))
    chk("unet_backbone.surface", M.get_unet_backbone(dict(task="surface_reconstruction", nf=12, unet_num_level=5, layer_order="gcr")))
    chk("decoder.16", M.get_decoder(dict(nf=16, layer_order="gcr")))
    chk("decoder.12", M.get_decoder(dict(nf=12, layer_order="gcr")))
    for tag in ("attention.16.4.softmax", "attention.12.8.softmax", "attention.16.4.gumbel"):
        _, nf, K, mode = tag.split(".")
        chk(tag, M.get_attention_block(dict(nf=int(nf), attn_patch_extent=4, K=int(K), attn_normalize=True,
                                            attn_use_switching=True, attn_retrieval_mode=mode == "gumbel",
                                            attn_no_output_mapping=True, attn_blend=True, attn_num_patch=16)))
    # attn_no_output_mapping=False: the g / o convolutions' keys (tests/golden/make_golden_attention_mapping.py asserts that
    # this oracle table equals the reference module's state_dict)
    mapped = M.get_attention_block(dict(nf=12, attn_patch_extent=4, K=8, attn_normalize=True, attn_use_switching=True,
                                        attn_retrieval_mode=False, attn_no_output_mapping=False, attn_blend=True, attn_num_patch=16))
    assert ({k: tuple(v.shape) for k, v in mapped.state_dict().items()} ==
            {k: tuple(v) for k, v in O.attention_shapes(12, 2, output_mapping=True).items()})
    fi, ft = M.get_retrieval_networks(dict(network_input="2+1", network_target="16+8", nf_input=32, nf_target=8, latent_dim=64))
    assert type(fi).__name__ == "Patch04" and type(ft).__name__ == "Patch32"
    fi, ft = M.get_retrieval_networks(dict(network_input="pc_32+8", network_target="16+4", nf_input=10, nf_target=12, latent_dim=64))
    assert type(fi).__name__ == "PCPatch48" and type(ft).__name__ == "Patch24"


def test_unsupported_configs_fail_loudly():
    import retrieval_fuse_b200.model as M
    # attn_no_output_mapping=False is supported (softmax mode); with the Gumbel mode the reference's own forward raises
    # (model/attention.py:103) - the module constructs, its composed output mapping is refused
    blk = M.get_attention_block(dict(nf=16, attn_patch_extent=4, K=4, attn_normalize=True, attn_use_switching=True,
                                     attn_retrieval_mode=True, attn_no_output_mapping=False, attn_blend=True, attn_num_patch=16))
    with pytest.raises(ValueError):
        blk.attention_blocks_layer.output_mapping()
    with pytest.raises(NotImplementedError):
        M.get_retrieval_backbone(dict(nf=16, retrieval_fmaps=16, retrieval_num_level=4, layer_order="cbr"))


def test_scene_access_matches_oracle():
    from retrieval_fuse_b200.dataset.patched_scene_dataset import PatchedSceneDataset
    from retrieval_fuse_b200.dataset.scene import InMemorySceneHandler, SceneHandler, point_cloud_to_grid
    from retrieval_fuse_b200.pipeline import FRONT3D_SR, MATTERPORT_SR16
    for cfg, s_in in ((FRONT3D_SR, 8), (MATTERPORT_SR16, 16)):
        dc = dict(cfg["dataset"], occupancy_threshold=-1)
        tg = {"a": O.synthetic_tsdf(1, 64, dc["voxel_size_target"])}
        inp = {"a": O.downsample_tsdf(tg["a"] / dc["voxel_size_target"] * dc["voxel_size_input"], 64 // s_in, dc["voxel_size_input"])}
        sh = InMemorySceneHandler("superresolution", dc, inp, tg)
        ei, et = sh.get_scene_patches("a")
        assert np.array_equal(et, O.get_extents_for_size([64] * 3, 16, 8, 16))
        assert np.array_equal(ei, O.get_extents_for_size([s_in] * 3, dc["patch_size_input"], dc["patch_context_input"], s_in // 4))
        ds = PatchedSceneDataset("val", dc, sh)
        want_in = O.chunk_patches(inp["a"], dc["patch_size_input"], dc["patch_context_input"], s_in // 4,
                                  O.f16_trunc(dc["voxel_size_input"]), dc["input_mean"], dc["input_std"])
        want_tg = O.chunk_patches(tg["a"], 16, 8, 16, O.f16_trunc(dc["voxel_size_target"]), dc["target_mean"], dc["target_std"])
        got_in = np.stack([ds[i]["input"] for i in range(64)])
        got_tg = np.stack([ds[i]["target"] for i in range(64)])
        assert np.array_equal(got_in.astype(np.float32), want_in) and np.array_equal(got_tg.astype(np.float32), want_tg)
        name = ds[5]["name"]
        assert SceneHandler.get_extent_from_name(name) == ("a", list(et[5]))
        assert ds.unpad(*et[5]) == [et[5][0], et[5][1] - 16, et[5][2], et[5][3] - 16, et[5][4], et[5][5] - 16]
    pts = np.random.default_rng(0).random((500, 3)) * 64
    assert np.array_equal(point_cloud_to_grid(pts, 128, 2.0, 8), O.point_cloud_to_grid(pts, 128, 2.0, 8))


def test_patcher_host_helpers():
    from retrieval_fuse_b200.util.patcher import Patcher
    p = Patcher([16] * 3, [8] * 3, [16] * 3, 2.25, [64] * 3)
    o = O.PatcherOracle([16] * 3, [8] * 3, [16] * 3, 2.25, [64] * 3)
    assert p.get_patch_counts() == o.get_patch_counts() == [4, 4, 4]
    assert p.get_patch_extents() == o.get_patch_extents() and p.get_patch_ratio() == o.get_patch_ratio()
    assert p.get_stride_ratio() == o.get_stride_ratio()


def test_oracle_knn_c_matches_numpy_and_handles_ties():
    rng = np.random.default_rng(0)
    db = rng.normal(size=(3000, 64)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    db[100] = db[7]
    db[2900] = db[7]
    q = rng.normal(size=(50, 64)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[3] = db[7]
    i1, d1 = O.knn_exact(db, q, 8)
    i2, d2 = O.knn_exact(db, q, 8, force_numpy=True)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert list(i1[3, :3]) == [7, 100, 2900] and np.all(d1[3, :3] == 0)
    assert np.all(np.diff(d1.astype(np.float64), axis=1) >= 0)


def test_oracle_demotion_rule():
    idx = np.array([[4, 9, 2, 7, 1, 5]], dtype=np.int32)
    d = np.arange(6, dtype=np.float32)[None]
    row_scene = np.array([0, 3, 3, 0, 1, 2, 0, 3, 0, 1])
    oi, od = O.demote_same_scene(idx, d, row_scene, np.array([3]), 3)  # rows 2, 7, 1 belong to scene 3
    assert list(oi[0]) == [4, 9, 5] and list(od[0]) == [0, 1, 5]
    oi, _ = O.demote_same_scene(idx, d, row_scene, np.array([-1]), 3)
    assert list(oi[0]) == [4, 9, 2]
    oi, _ = O.demote_same_scene(idx[:, :4], d[:, :4], np.full(10, 3), np.array([3]), 2)  # everything same-scene: order kept
    assert list(oi[0]) == [4, 9]


def test_fixed_divisor_division_is_ieee(tmp_path):
    """rf_div_rn_fixed (the normalisation's division in the re-indexing kernels) against the host's IEEE division:
    tools/check_fixed_div.c runs the same 5-step FMA sequence on 5e7 random (a, b) pairs of the guarded ranges."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_fixed_div.c")
    exe = str(tmp_path / "check_fixed_div")
    flags = ["-O2", "-ffp-contract=off"]
    if "fma" in open("/proc/cpuinfo").read():
        flags.append("-mfma")  # hardware fmaf; glibc's software fmaf is correctly rounded as well, only slower
    subprocess.check_call(["gcc"] + flags + ["-o", exe, src, "-lm"])
    out = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and " bad 0" in out.stdout, out.stdout


def test_wpack_weight_expansion_identity():
    """The experimental W-packing of small-channel 3x3x3 layers (model.unet.W_PACK, off by default): conv3d on the
    packed view with the expanded weights equals the original convolution (CPU, float64)."""
    from retrieval_fuse_b200.model.unet import W_PACK, wpack_weights
    assert W_PACK == {}, "the experimental path must stay off by default"
    g = torch.Generator().manual_seed(3)
    for (cin, cout, W, Bw) in [(1, 8, 16, 8), (8, 16, 16, 4), (3, 5, 12, 2)]:
        x = torch.randn(2, cin, 5, 6, W, generator=g, dtype=torch.float64)
        w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
        want = torch.nn.functional.conv3d(x, w, padding=1).permute(0, 2, 3, 4, 1)
        xp = x.permute(0, 2, 3, 4, 1).contiguous().view(2, 5, 6, W // Bw, Bw * cin).permute(0, 4, 1, 2, 3)
        yp = torch.nn.functional.conv3d(xp, wpack_weights(w, Bw), padding=1)
        got = yp.permute(0, 2, 3, 4, 1).contiguous().view(2, 5, 6, W, cout)
        assert float((got - want).abs().max()) < 1e-12
