"""TEST INFRASTRUCTURE — deterministic synthetic weights in the reference state_dict layout.

No pretrained checkpoint is reachable offline (zenodo links in tools/download_*_models.sh), so
parity and benchmarks use seeded random weights, scaled so that activations stay O(1) through the
ResNetV2 stem, 12 ViT blocks and the RefineNet decoder, and so that the final ReLU does not zero
the output (SURVEY.md §7 hard part 2).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def state_dict_spec(num_channels: int = 1):
    """(key, shape) list of DPTDepthModel(backbone='vitb_rn50_384') — SURVEY.md Appendix B."""
    spec = []
    add = lambda k, *s: spec.append((k, tuple(s)))
    P = "pretrained.model."
    add(P + "cls_token", 1, 1, 768)
    add(P + "pos_embed", 1, 577, 768)
    bb = P + "patch_embed.backbone."
    add(bb + "stem.conv.weight", 64, 3, 7, 7)
    add(bb + "stem.norm.weight", 64)
    add(bb + "stem.norm.bias", 64)
    cin = 64
    for s, (cout, depth) in enumerate(zip((256, 512, 1024), (3, 4, 9))):
        mid = cout // 4
        for b in range(depth):
            p = f"{bb}stages.{s}.blocks.{b}."
            if b == 0:
                add(p + "downsample.conv.weight", cout, cin, 1, 1)
                add(p + "downsample.norm.weight", cout)
                add(p + "downsample.norm.bias", cout)
            add(p + "conv1.weight", mid, cin if b == 0 else cout, 1, 1)
            add(p + "norm1.weight", mid); add(p + "norm1.bias", mid)
            add(p + "conv2.weight", mid, mid, 3, 3)
            add(p + "norm2.weight", mid); add(p + "norm2.bias", mid)
            add(p + "conv3.weight", cout, mid, 1, 1)
            add(p + "norm3.weight", cout); add(p + "norm3.bias", cout)
        cin = cout
    add(P + "patch_embed.proj.weight", 768, 1024, 1, 1)
    add(P + "patch_embed.proj.bias", 768)
    for i in range(12):
        p = f"{P}blocks.{i}."
        add(p + "norm1.weight", 768); add(p + "norm1.bias", 768)
        add(p + "attn.qkv.weight", 2304, 768); add(p + "attn.qkv.bias", 2304)
        add(p + "attn.proj.weight", 768, 768); add(p + "attn.proj.bias", 768)
        add(p + "norm2.weight", 768); add(p + "norm2.bias", 768)
        add(p + "mlp.fc1.weight", 3072, 768); add(p + "mlp.fc1.bias", 3072)
        add(p + "mlp.fc2.weight", 768, 3072); add(p + "mlp.fc2.bias", 768)
    add(P + "norm.weight", 768); add(P + "norm.bias", 768)
    add(P + "head.weight", 1000, 768); add(P + "head.bias", 1000)
    for n in (3, 4):
        p = f"pretrained.act_postprocess{n}."
        add(p + "0.project.0.weight", 768, 1536); add(p + "0.project.0.bias", 768)
        add(p + "3.weight", 768, 768, 1, 1); add(p + "3.bias", 768)
        if n == 4:
            add(p + "4.weight", 768, 768, 3, 3); add(p + "4.bias", 768)
    for n, c in zip((1, 2, 3, 4), (256, 512, 768, 768)):
        add(f"scratch.layer{n}_rn.weight", 256, c, 3, 3)
    for n in (1, 2, 3, 4):
        p = f"scratch.refinenet{n}."
        add(p + "out_conv.weight", 256, 256, 1, 1); add(p + "out_conv.bias", 256)
        for u in (1, 2):
            for cv in (1, 2):
                add(f"{p}resConfUnit{u}.conv{cv}.weight", 256, 256, 3, 3)
                add(f"{p}resConfUnit{u}.conv{cv}.bias", 256)
    add("scratch.output_conv.0.weight", 128, 256, 3, 3); add("scratch.output_conv.0.bias", 128)
    add("scratch.output_conv.2.weight", 32, 128, 3, 3); add("scratch.output_conv.2.bias", 32)
    add("scratch.output_conv.4.weight", num_channels, 32, 1, 1); add("scratch.output_conv.4.bias", num_channels)
    return spec


def make_state_dict(seed: int = 0, num_channels: int = 1):
    """Seeded synthetic checkpoint over THIS file's (reference-pinned) key/shape table; the RNG recipe
    is shared with the product's omnidata_b200/synthetic.py so both sides see identical weights."""
    from omnidata_b200.synthetic import make_state_dict as gen
    return gen(seed, num_channels, spec=state_dict_spec(num_channels))
