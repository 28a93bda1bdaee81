"""Oracle (test infrastructure): state-dict key maps  oracle (diffusers names) -> the reference's in-tree LDM modules.

diffusers' UNet / AutoencoderKL blocks are key-renamed ports of the LDM blocks the reference carries under
``instructany2pix/llm/model/vae/modules`` (blocks.py: ResnetBlock :83-142, AttnBlock :151-203, Encoder :369-460, Decoder
:463-570; attention.py: GEGLU :37-44, FeedForward :47-64, CrossAttention :152-193, BasicTransformerBlock :196-215,
SpatialTransformer :218-260).  ``oracle/gen_golden.py::gen_ldm`` draws name-seeded weights for the ORACLE module, renames
them with these maps, loads them into the reference module and records the reference's output; the test then only needs the
oracle module + the same name-seeded weights.  Differences that are NOT renames are handled here:
1x1 convs <-> Linear weights ([C,C,1,1] <-> [C,C]), the reversed order of ``Decoder.up``.
"""
from __future__ import annotations


def resnet_to_ldm(sd):
    """oracle ResnetBlock2D / VAE _Res keys -> LDM ResnetBlock keys"""
    ren = {"time_emb_proj": "temb_proj", "conv_shortcut": "nin_shortcut"}
    out = {}
    for k, v in sd.items():
        head, _, tail = k.rpartition(".")
        parts = head.split(".")
        parts[-1] = ren.get(parts[-1], parts[-1])
        out[".".join(parts) + "." + tail] = v
    return out


def vae_attn_to_ldm(sd, prefix=""):
    """oracle VAE _Attn (Linear q/k/v/out over tokens) -> LDM AttnBlock (1x1 convs over the map)"""
    ren = {"group_norm": "norm", "to_q": "q", "to_k": "k", "to_v": "v", "to_out.0": "proj_out"}
    out = {}
    for k, v in sd.items():
        head, _, tail = k.rpartition(".")
        new = ren[head]
        if new != "norm" and tail == "weight":
            v = v[:, :, None, None]
        out[f"{prefix}{new}.{tail}"] = v
    return out


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _put(out, prefix, sd):
    for k, v in sd.items():
        out[prefix + k] = v


def _mid_and_ends(sd, out):
    for k in ("conv_in", "conv_out"):
        _put(out, k + ".", _sub(sd, k + "."))
    _put(out, "norm_out.", _sub(sd, "conv_norm_out."))
    _put(out, "mid.block_1.", resnet_to_ldm(_sub(sd, "mid_res0.")))
    _put(out, "mid.block_2.", resnet_to_ldm(_sub(sd, "mid_res1.")))
    out.update(vae_attn_to_ldm(_sub(sd, "mid_attn."), "mid.attn_1."))


def vae_decoder_to_ldm(sd, n_levels):
    """OracleVAEDecoder (without post_quant_conv) -> LDM Decoder.  ``Decoder.up[i]`` is indexed by resolution level and
    executed from the last one down (blocks.py:524-545, :558-564); the oracle's up_blocks are in execution order."""
    out = {}
    _mid_and_ends(sd, out)
    for i in range(n_levels):
        src, dst = f"up_blocks.{i}.", f"up.{n_levels - 1 - i}."
        blk = _sub(sd, src)
        _put(out, dst, resnet_to_ldm({k.replace("resnets.", "block."): v for k, v in blk.items() if k.startswith("resnets.")}))
        _put(out, dst + "upsample.conv.", _sub(blk, "upsamplers.0.conv."))
    return out


def vae_encoder_to_ldm(sd, n_levels):
    """OracleVAEEncoder (without quant_conv) -> LDM Encoder"""
    out = {}
    _mid_and_ends(sd, out)
    for i in range(n_levels):
        src, dst = f"down_blocks.{i}.", f"down.{i}."
        blk = _sub(sd, src)
        _put(out, dst, resnet_to_ldm({k.replace("resnets.", "block."): v for k, v in blk.items() if k.startswith("resnets.")}))
        _put(out, dst + "downsample.conv.", _sub(blk, "downsamplers.0.conv."))
    return out


def transformer2d_to_ldm(sd):
    """oracle Transformer2DModel (Linear proj_in/out, SDXL ``use_linear_projection``) -> LDM SpatialTransformer (1x1 convs);
    everything inside ``transformer_blocks`` carries identical names in both."""
    out = {}
    for k, v in sd.items():
        if k in ("proj_in.weight", "proj_out.weight"):
            v = v[:, :, None, None]
        out[k] = v
    return out
