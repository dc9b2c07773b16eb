"""GPU diagnostic: per-stage error of the refine forward against an fp64 evaluation of the
same network, for the tensor-core path and the fp32 FMA path (and the reference's own fp32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
import cases as C
from oracle import rf_oracle as O
from retrieval_fuse_b200.pipeline import FRONT3D_SR, RefinementPipeline
from retrieval_fuse_b200.model import unet as U

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
x_in, x_re = C.refine_full_inputs()
syn = lambda sh: O.synth_state_dict(sh, C.SEED)
sds = dict(unet_backbone=syn(O.unet_backbone_shapes("sr08", 16, 4)), retrieval_backbone=syn(O.retrieval_backbone_shapes(16, 16, 4)),
           attention=syn(O.attention_shapes(16, 2)), decoder=syn(O.final_decoder_shapes(16)))
cfg = dict(kind="sr08", nf=16, unet_num_level=4, retrieval_fmaps=16, retrieval_num_level=4, K=4, E=2)
sd64 = {k: {n: v.double() for n, v in d.items()} for k, d in sds.items()}
r64 = O.refine_forward(x_in.double(), x_re.double(), sd64, cfg)
r32 = O.refine_forward(x_in, x_re, sds, cfg)
names = ["pred", "x_back", "x_retr", "x_attn"]
print("reference fp32 (torch CPU) vs fp64:", {n: f"{float((a.double() - b).abs().max()):.2e}" for n, a, b in zip(names, r32, r64)})
pipe = RefinementPipeline(FRONT3D_SR, bank=None, device=dev)
pipe.unet_backbone.load_state_dict(sds["unet_backbone"]); pipe.retrieval_backbone.load_state_dict(sds["retrieval_backbone"])
pipe.patched_attention_block.load_state_dict(sds["attention"]); pipe.decoder.load_state_dict(sds["decoder"])
for tc in (True, False):
    U.USE_TENSOR_CORES = tc
    pipe.patched_attention_block.attention_blocks_layer.use_tensor_cores = tc
    out = pipe.refine(x_in.to(dev), x_re.to(dev))
    print(f"ours tensor_cores={tc} vs fp64:", {n: f"{float((a.cpu().double() - b).abs().max()):.2e}" for n, a, b in zip(names, out, r64)})
    # isolate the attention and the decoder: feed them the fp64-exact inputs
    xa = pipe.patched_attention_block(r64[1].float().to(dev), r64[2].float().to(dev))
    pd = pipe.decoder(r64[3].float().to(dev))
    print(f"   attention alone (exact inputs): {float((xa.cpu().double() - r64[3]).abs().max()):.2e}   decoder alone: {float((pd.cpu().double() - r64[0]).abs().max()):.2e}")
xa32 = O.patched_attention_forward(r64[1].float(), r64[2].float(), sds["attention"], 16, 16, 2, 4)
print(f"reference fp32 attention alone (exact inputs): {float((xa32.double() - r64[3]).abs().max()):.2e}")
