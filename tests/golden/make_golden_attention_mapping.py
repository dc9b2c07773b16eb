"""Generates tests/golden/attention_mapping.npz by running the REFERENCE's PatchedAttentionBlock with
attn_no_output_mapping=False (the g / o 1x1x1 convolutions of model/attention.py:5-15,56-57,95,108).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_attention_mapping.py

Same recipe as make_golden.py: the reference's model/attention.py unmodified, the deterministic synthetic weights of
oracle.rf_oracle.synth_state_dict, seeded inputs; the stored outputs pin the oracle (tests/test_oracle_golden.py) and
the CUDA path (tests/test_gpu_parity.py)."""
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

torch.set_grad_enabled(False)


def main():
    out = {}
    for nf, K, mode in C.ATTN_MAPPING_CASES:
        cfg = dict(nf=nf, attn_patch_extent=4, K=K, attn_normalize=True, attn_use_switching=True, attn_retrieval_mode=mode,
                   attn_no_output_mapping=False, attn_blend=True, attn_num_patch=16)
        m = ref_model.get_attention_block(cfg)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert {k: tuple(v) for k, v in O.attention_shapes(nf, 2, output_mapping=True).items()} == shapes, \
            "oracle shape table != reference state_dict"
        m.load_state_dict(O.synth_state_dict(shapes, C.SEED))
        m.eval()
        tag = C.attention_tag(nf, K, mode) + ".mapped"
        xb, xr, _ = C.attention_inputs(nf, K, mode)
        if mode:  # the reference draws the Gumbel noise inside gumbel_softmax: record the identical draw by re-seeding
            torch.manual_seed(77)
            out[tag + ".noise"] = (-torch.empty(4096, K).exponential_().log()).numpy()
            torch.manual_seed(77)
        y = m(xb, xr).contiguous()
        out[tag] = y[:, :, ::2, ::2, ::2].contiguous().numpy()
        print(tag, float(y.abs().max()), float((y - xb).abs().max()))
    np.savez_compressed(os.path.join(HERE, "attention_mapping.npz"), **out)


if __name__ == "__main__":
    main()
