"""Drop-in for the reference's `model` package: the string -> class factories
of model/__init__.py:6-61 with identical signatures and config keys."""
from .attention import AttentionBlock, PatchedAttentionBlock, Fold3D, Unfold3D, Unfold3DPadStride
from .refinement import (Superresolution08UNetBackbone, SurfaceReconstructionUNetBackbone,
                         Superresolution08FinalDecoder, RetrievalUNetBackbone, Superresolution16UNetBackbone)
from .retrieval import (Patch04, Patch08, Patch16, Patch24, Patch32, PCPatch32, PCPatch48, PCPatch64, Patch12,
                        PatchNorm08, PatchNorm32, Patch24V2, Patch04V2, Patch05)

_INPUT_NETS = {"2+1": Patch04, "2+1V2": Patch04V2, "4+2": Patch08, "4+2N": PatchNorm08, "16+4": Patch24,
               "pc_16+8": PCPatch32, "pc_32+8": PCPatch48, "pc_32+16": PCPatch64}
_TARGET_NETS = {"pc_32+16": PCPatch64, "8+2": Patch12, "8+4": Patch16, "16+4": Patch24, "16+4V2": Patch24V2,
                "16+8": Patch32, "16+8N": PatchNorm32}


def get_retrieval_networks(model_config):
    """model/__init__.py:6-38. Unknown keys yield None, as in the reference."""
    cin = _INPUT_NETS.get(model_config["network_input"])
    ctg = _TARGET_NETS.get(model_config["network_target"])
    fenc_input = cin(model_config["nf_input"], model_config["latent_dim"]) if cin else None
    fenc_target = ctg(model_config["nf_target"], model_config["latent_dim"]) if ctg else None
    return fenc_input, fenc_target


def get_unet_backbone(config):
    """model/__init__.py:41-48."""
    if config["task"] == "superresolution":
        size = config["dataset_train"]["input_chunk_size"]
        if size == 8:
            return Superresolution08UNetBackbone(config["nf"], num_levels=config["unet_num_level"], layer_order=config["layer_order"])
        if size == 16:
            return Superresolution16UNetBackbone(config["nf"], num_levels=config["unet_num_level"], layer_order=config["layer_order"])
    if config["task"] == "surface_reconstruction":
        return SurfaceReconstructionUNetBackbone(config["nf"], num_levels=config["unet_num_level"], layer_order=config["layer_order"])
    return None


def get_decoder(config):
    """model/__init__.py:51-52."""
    return Superresolution08FinalDecoder(config["nf"], layer_order=config["layer_order"])


def get_retrieval_backbone(config):
    """model/__init__.py:55-56."""
    return RetrievalUNetBackbone(nf=config["nf"], f_maps=config["retrieval_fmaps"], num_levels=config["retrieval_num_level"],
                                 layer_order=config["layer_order"])


def get_attention_block(config):
    """model/__init__.py:59-61."""
    attention_block = AttentionBlock(config["nf"], config["attn_patch_extent"] // 2, config["K"], config["attn_normalize"],
                                     config["attn_use_switching"], config["attn_retrieval_mode"],
                                     config["attn_no_output_mapping"], config["attn_blend"])
    return PatchedAttentionBlock(config["nf"], config["attn_num_patch"], config["attn_patch_extent"] // 2, config["K"],
                                 attention_block)
