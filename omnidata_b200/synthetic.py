"""Seeded synthetic weights in the reference state_dict layout (random-init weights of the
DPT-Hybrid-384 architecture; no checkpoint is reachable offline).  Scaled so that activations stay
O(1) through the ResNetV2 stem, 12 ViT blocks and the RefineNet decoder, residual branches are
perturbations of their shortcuts (as in a trained network) and the final ReLU stays alive."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def make_state_dict(seed: int = 0, num_channels: int = 1, spec=None) -> "OrderedDict[str, torch.Tensor]":
    if spec is None:
        from .model import state_dict_spec
        spec = state_dict_spec(num_channels)
    g = torch.Generator(device="cpu").manual_seed(1234 + seed)
    # the plain-ViT encoders (DPT-Large): identified by their ConvTranspose reassemble layer
    plain_vit = any(k == "pretrained.act_postprocess1.4.weight" for k, _ in spec)
    sd = OrderedDict()
    for key, shape in spec:
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("cls_token") or key.endswith("pos_embed"):
            t = torch.randn(shape, generator=g) * 0.5
        elif ".norm" in key and leaf == "weight" and len(shape) == 1:      # GN / LN gain
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            if ".norm3." in key:
                # trained residual nets keep the bottleneck branch a perturbation of the shortcut;
                # unit gains here make a random network chaotic (rounding noise x10 per stage)
                t = 0.25 * t
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
            if key == "scratch.output_conv.4.bias":
                t = t + 0.4            # keep the final ReLU alive on most pixels
            if key == "scratch.output_conv.2.bias":
                t = t + 0.2
        else:                                                               # conv / linear weight
            fan_in = math.prod(shape[1:])
            if key in ("pretrained.act_postprocess1.4.weight", "pretrained.act_postprocess2.4.weight") \
                    and len(shape) == 4 and shape[2] in (2, 4):
                fan_in = shape[0]      # ConvTranspose2d [in, out, k, k] with stride k: one input pixel per output pixel
            gain = 1.0
            if ".attn.proj." in key or ".mlp.fc2." in key:
                gain = 0.25            # residual branches: keep the token stream well conditioned
            if "resConfUnit" in key and ".conv2." in key:
                gain = 0.35
            if ".mlp.fc1." in key or ".attn.qkv." in key:
                gain = 1.0
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
            if key.startswith("scratch.output_conv.4."):
                t = t * 2.0
                if plain_vit:
                    t = 0.25 * (t.abs() - 0.6 * t.abs().mean())   # mostly positive read-out: the final ReLU stays alive
        sd[key] = t.float()
    return sd
