"""Corroboration of the timm restatement (the only un-pinned half of the oracle): HuggingFace
transformers ships an independent port of the same network (BiT ResNetV2 backbone + ViT-B + DPT neck
and head, `DPTForDepthEstimation(is_hybrid=True)`).  With the seeded weights remapped key by key the
two implementations must agree to fp32 round-off; the HF port standardises weights as
(w-mean)/sqrt(var+eps) while timm 0.4.12 uses (w-mean)/(std+eps) — a <= 5e-6 relative difference."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def remap(sd):
    out = {}
    P = "pretrained.model."
    out["dpt.embeddings.cls_token"] = sd[P + "cls_token"]
    out["dpt.embeddings.position_embeddings"] = sd[P + "pos_embed"]
    bb, hb = P + "patch_embed.backbone.", "dpt.embeddings.backbone.bit."
    out[hb + "embedder.convolution.weight"] = sd[bb + "stem.conv.weight"]
    out[hb + "embedder.norm.weight"] = sd[bb + "stem.norm.weight"]
    out[hb + "embedder.norm.bias"] = sd[bb + "stem.norm.bias"]
    for k, v in sd.items():
        if k.startswith(bb + "stages."):
            out[hb + "encoder." + k[len(bb):].replace(".blocks.", ".layers.")] = v
    out["dpt.embeddings.projection.weight"] = sd[P + "patch_embed.proj.weight"]
    out["dpt.embeddings.projection.bias"] = sd[P + "patch_embed.proj.bias"]
    for i in range(12):
        p, h = f"{P}blocks.{i}.", f"dpt.encoder.layer.{i}."
        for j, name in enumerate(("query", "key", "value")):
            out[f"{h}attention.attention.{name}.weight"] = sd[p + "attn.qkv.weight"][j * 768:(j + 1) * 768]
            out[f"{h}attention.attention.{name}.bias"] = sd[p + "attn.qkv.bias"][j * 768:(j + 1) * 768]
        for a, b in (("attn.proj", "attention.output.dense"), ("mlp.fc1", "intermediate.dense"),
                     ("mlp.fc2", "output.dense"), ("norm1", "layernorm_before"), ("norm2", "layernorm_after")):
            out[h + b + ".weight"] = sd[p + a + ".weight"]
            out[h + b + ".bias"] = sd[p + a + ".bias"]
    out["dpt.layernorm.weight"], out["dpt.layernorm.bias"] = sd[P + "norm.weight"], sd[P + "norm.bias"]
    for n in (3, 4):
        p = f"pretrained.act_postprocess{n}."
        for s in ("weight", "bias"):
            out[f"neck.reassemble_stage.readout_projects.{n - 1}.0.{s}"] = sd[p + "0.project.0." + s]
            out[f"neck.reassemble_stage.layers.{n - 1}.projection.{s}"] = sd[p + "3." + s]
    for s in ("weight", "bias"):
        out[f"neck.reassemble_stage.layers.3.resize.{s}"] = sd["pretrained.act_postprocess4.4." + s]
    for n in (1, 2, 3, 4):
        out[f"neck.convs.{n - 1}.weight"] = sd[f"scratch.layer{n}_rn.weight"]
        p, h = f"scratch.refinenet{n}.", f"neck.fusion_stage.layers.{4 - n}."
        for s in ("weight", "bias"):
            out[h + "projection." + s] = sd[p + "out_conv." + s]
            for u in (1, 2):
                for cv in (1, 2):
                    out[f"{h}residual_layer{u}.convolution{cv}.{s}"] = sd[f"{p}resConfUnit{u}.conv{cv}.{s}"]
    for i in (0, 2, 4):
        for s in ("weight", "bias"):
            out[f"head.head.{i}.{s}"] = sd[f"scratch.output_conv.{i}.{s}"]
    return out


def test_oracle_with_timm_shim_matches_hf_dpt_hybrid_port():
    from transformers import DPTConfig, DPTForDepthEstimation
    from oracle import dpt_oracle, make_golden, weights
    cfg = DPTConfig(is_hybrid=True, image_size=384, patch_size=16, hidden_size=768, num_hidden_layers=12,
                    num_attention_heads=12, intermediate_size=3072, backbone_out_indices=[2, 5, 8, 11],
                    readout_type="project", neck_hidden_sizes=[256, 512, 768, 768], fusion_hidden_size=256,
                    backbone_featmap_shape=[1, 1024, 24, 24], neck_ignore_stages=[0, 1], qkv_bias=True,
                    hidden_act="gelu", layer_norm_eps=1e-6, hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0)
    hf = DPTForDepthEstimation(cfg).eval()
    sd = weights.make_state_dict(0, 1)
    mapped = remap(sd)
    missing, unexpected = hf.load_state_dict(mapped, strict=False)
    assert not unexpected, unexpected
    assert all("num_batches_tracked" in k or "running_" in k for k in missing), missing
    # every oracle tensor except the dead ImageNet classifier must have been consumed
    assert sum(v.numel() for v in mapped.values()) == sum(
        v.numel() for k, v in sd.items() if not k.startswith("pretrained.model.head."))
    x = make_golden.golden_input(1)
    with torch.no_grad():
        y_hf = hf(pixel_values=x).predicted_depth
        y = dpt_oracle.forward_fp32(sd, x)
    err = float((y - y_hf).norm() / y_hf.norm())
    assert tuple(y_hf.shape) == tuple(y.shape)
    assert err < 2e-4, err


def remap_plain_vit(sd, depth, dim):
    """reference DPT with a plain ViT encoder (vitl16_384 / vitb16_384) -> HF DPTForDepthEstimation(is_hybrid=False)."""
    out = {}
    P = "pretrained.model."
    out["dpt.embeddings.cls_token"] = sd[P + "cls_token"]
    out["dpt.embeddings.position_embeddings"] = sd[P + "pos_embed"]
    out["dpt.embeddings.patch_embeddings.projection.weight"] = sd[P + "patch_embed.proj.weight"]
    out["dpt.embeddings.patch_embeddings.projection.bias"] = sd[P + "patch_embed.proj.bias"]
    for i in range(depth):
        p, h = f"{P}blocks.{i}.", f"dpt.encoder.layer.{i}."
        for j, name in enumerate(("query", "key", "value")):
            out[f"{h}attention.attention.{name}.weight"] = sd[p + "attn.qkv.weight"][j * dim:(j + 1) * dim]
            out[f"{h}attention.attention.{name}.bias"] = sd[p + "attn.qkv.bias"][j * dim:(j + 1) * dim]
        for a, b in (("attn.proj", "attention.output.dense"), ("mlp.fc1", "intermediate.dense"),
                     ("mlp.fc2", "output.dense"), ("norm1", "layernorm_before"), ("norm2", "layernorm_after")):
            out[h + b + ".weight"] = sd[p + a + ".weight"]
            out[h + b + ".bias"] = sd[p + a + ".bias"]
    out["dpt.layernorm.weight"], out["dpt.layernorm.bias"] = sd[P + "norm.weight"], sd[P + "norm.bias"]
    for n in (1, 2, 3, 4):
        p = f"pretrained.act_postprocess{n}."
        for s in ("weight", "bias"):
            out[f"neck.reassemble_stage.readout_projects.{n - 1}.0.{s}"] = sd[p + "0.project.0." + s]
            out[f"neck.reassemble_stage.layers.{n - 1}.projection.{s}"] = sd[p + "3." + s]
            if n != 3:
                out[f"neck.reassemble_stage.layers.{n - 1}.resize.{s}"] = sd[p + "4." + s]
    for n in (1, 2, 3, 4):
        out[f"neck.convs.{n - 1}.weight"] = sd[f"scratch.layer{n}_rn.weight"]
        p, h = f"scratch.refinenet{n}.", f"neck.fusion_stage.layers.{4 - n}."
        for s in ("weight", "bias"):
            out[h + "projection." + s] = sd[p + "out_conv." + s]
            for u in (1, 2):
                for cv in (1, 2):
                    out[f"{h}residual_layer{u}.convolution{cv}.{s}"] = sd[f"{p}resConfUnit{u}.conv{cv}.{s}"]
    for i in (0, 2, 4):
        for s in ("weight", "bias"):
            out[f"head.head.{i}.{s}"] = sd[f"scratch.output_conv.{i}.{s}"]
    return out


@pytest.mark.parametrize("backbone,golden,dim,depth,heads,hooks,neck", [
    ("vitl16_384", "dpt_large_fp32_seed0_c1.pt", 1024, 24, 16, [5, 11, 17, 23], [256, 512, 1024, 1024]),
    ("vitb16_384", "dpt_vitb16_fp32_seed0_c1.pt", 768, 12, 12, [2, 5, 8, 11], [96, 192, 384, 768]),
])
def test_plain_vit_goldens_match_hf_dpt_port(backbone, golden, dim, depth, heads, hooks, neck):
    """The golden vectors of the plain-ViT DPTs were produced by the unmodified reference class on the timm
    SHIM's vit_{large,base}_patch16_384.  HuggingFace's independent DPT port (the published DPT-Large layout) with
    the same seeded weights must reproduce the recorded output: corroborates the shim half of those goldens."""
    from pathlib import Path
    from transformers import DPTConfig, DPTForDepthEstimation
    from omnidata_b200 import synthetic
    from oracle import make_golden
    rec = torch.load(Path(__file__).parent / "golden" / golden)
    spec = [(k, tuple(s)) for k, s in rec["spec"]]
    sd = synthetic.make_state_dict(0, 1, spec=spec)
    cfg = DPTConfig(is_hybrid=False, image_size=384, patch_size=16, hidden_size=dim, num_hidden_layers=depth,
                    num_attention_heads=heads, intermediate_size=4 * dim, backbone_out_indices=hooks,
                    readout_type="project", neck_hidden_sizes=neck, reassemble_factors=[4, 2, 1, 0.5],
                    fusion_hidden_size=256, qkv_bias=True, hidden_act="gelu", layer_norm_eps=1e-6,
                    hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = DPTForDepthEstimation(cfg).eval()
    mapped = remap_plain_vit(sd, depth, dim)
    missing, unexpected = hf.load_state_dict(mapped, strict=False)
    assert not unexpected, unexpected
    assert all("num_batches_tracked" in k or "running_" in k for k in missing), missing
    assert sum(v.numel() for v in mapped.values()) == sum(
        v.numel() for k, v in sd.items() if not k.startswith("pretrained.model.head."))
    with torch.no_grad():
        y_hf = hf(pixel_values=make_golden.golden_input(1)).predicted_depth
    ref = rec["output_sub8"]
    err = float((y_hf[..., ::8, ::8] - ref).norm() / ref.norm())
    assert err < 2e-4, err
