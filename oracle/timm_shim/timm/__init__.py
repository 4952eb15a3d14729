"""TEST INFRASTRUCTURE — a stand-in for `timm` (pinned ==0.4.12 by the reference,
omnidata_tools/torch/requirements.txt:15) exposing only what the reference touches:

    timm.create_model("vit_base_resnet50_384", pretrained=...)     (modules/midas/vit.py:483)

timm is a third-party dependency that is NOT vendored under /root/reference and cannot be
installed here (no network).  This file restates the published architecture of timm 0.4.12's
`vit_base_resnet50_384` (VisionTransformer + HybridEmbed + ResNetV2(3,4,9) with StdConv2dSame /
GroupNormAct / non-pre-activation Bottleneck) in plain torch so that the reference's own
`modules.midas.dpt_depth.DPTDepthModel` can be imported UNMODIFIED in this container and used as
the parity oracle.  Parameter names reproduce timm's state_dict layout.
`pretrained` is ignored (no ImageNet weights offline).  Corroboration: HuggingFace transformers'
BiT/DPT-hybrid port of the same network (tests/test_oracle_cpu.py cross-checks numerically).
Parity at this boundary is otherwise UNPINNED: the reference ships no test or golden tensor for it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything under oracle/.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__version__ = "0.4.12-shim"


def _same_pad(size: int, k: int, s: int) -> int:
    """TF 'SAME' total padding for one spatial dim (timm.models.layers.padding.get_same_padding)."""
    return max((math.ceil(size / s) - 1) * s + (k - 1) + 1 - size, 0)


def _pad_same(x, k, s, value=0.0):
    ph, pw = _same_pad(x.shape[-2], k, s), _same_pad(x.shape[-1], k, s)
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)
    return x


class StdConv2dSame(nn.Conv2d):
    """Weight-standardised conv with TF-SAME padding, no bias (timm.models.layers.std_conv)."""

    def __init__(self, cin, cout, kernel_size, stride=1, eps=1e-8):
        super().__init__(cin, cout, kernel_size, stride=stride, padding=0, bias=False)
        self.eps = eps

    def standardized_weight(self):
        std, mean = torch.std_mean(self.weight, dim=[1, 2, 3], keepdim=True, unbiased=False)
        return (self.weight - mean) / (std + self.eps)

    def forward(self, x):
        x = _pad_same(x, self.kernel_size[0], self.stride[0])
        return F.conv2d(x, self.standardized_weight(), None, self.stride, 0)


class GroupNormAct(nn.GroupNorm):
    def __init__(self, channels, apply_act=True):
        super().__init__(32, channels, eps=1e-5, affine=True)
        self.apply_act = apply_act

    def forward(self, x):
        x = F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return F.relu(x) if self.apply_act else x


class MaxPool2dSame(nn.Module):
    def forward(self, x):
        return F.max_pool2d(_pad_same(x, 3, 2, value=float("-inf")), 3, 2)


class DownsampleConv(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = StdConv2dSame(cin, cout, 1, stride=stride)
        self.norm = GroupNormAct(cout, apply_act=False)

    def forward(self, x):
        return self.norm(self.conv(x))


class Bottleneck(nn.Module):
    """Non-pre-activation bottleneck (timm.models.resnetv2.Bottleneck), ratio 0.25."""

    def __init__(self, cin, cout, stride, project):
        super().__init__()
        mid = cout // 4
        self.downsample = DownsampleConv(cin, cout, stride) if project else None
        self.conv1 = StdConv2dSame(cin, mid, 1)
        self.norm1 = GroupNormAct(mid)
        self.conv2 = StdConv2dSame(mid, mid, 3, stride=stride)
        self.norm2 = GroupNormAct(mid)
        self.conv3 = StdConv2dSame(mid, cout, 1)
        self.norm3 = GroupNormAct(cout, apply_act=False)

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        y = self.norm1(self.conv1(x))
        y = self.norm2(self.conv2(y))
        y = self.norm3(self.conv3(y))
        return F.relu(y + shortcut)


class ResNetStage(nn.Module):
    def __init__(self, cin, cout, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(*[
            Bottleneck(cin if i == 0 else cout, cout, stride if i == 0 else 1, project=(i == 0))
            for i in range(depth)])

    def forward(self, x):
        return self.blocks(x)


class ResNetV2(nn.Module):
    def __init__(self, layers=(3, 4, 9)):
        super().__init__()
        self.stem = nn.Sequential()
        self.stem.add_module("conv", StdConv2dSame(3, 64, 7, stride=2))
        self.stem.add_module("norm", GroupNormAct(64))
        self.stem.add_module("pool", MaxPool2dSame())
        widths, cin = (256, 512, 1024), 64
        stages = []
        for i, (w, d) in enumerate(zip(widths, layers)):
            stages.append(ResNetStage(cin, w, 1 if i == 0 else 2, d))
            cin = w
        self.stages = nn.Sequential(*stages)
        self.norm = nn.Identity()
        self.num_features = cin

    def forward(self, x):
        return self.norm(self.stages(self.stem(x)))


class HybridEmbed(nn.Module):
    def __init__(self, backbone, embed_dim):
        super().__init__()
        self.backbone = backbone
        self.proj = nn.Conv2d(backbone.num_features, embed_dim, 1)
        self.num_patches = 24 * 24

    def forward(self, x):
        return self.proj(self.backbone(x)).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop = nn.Dropout(0.0)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    """timm 0.4.12 layers/patch_embed.py: non-overlapping 16x16 patches by a strided convolution."""

    def __init__(self, img_size=384, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, embed_dim=768, depth=12, heads=12, num_classes=1000, hybrid=True):
        super().__init__()
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = HybridEmbed(ResNetV2((3, 4, 9)), embed_dim) if hybrid else PatchEmbed(384, 16, 3, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.ModuleList([Block(embed_dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        return self.head(self.norm(x)[:, 0])


def create_model(name, pretrained=False, **kwargs):
    if name == "vit_base_resnet50_384":
        return VisionTransformer()
    if name == "vit_large_patch16_384":     # timm 0.4.12 vision_transformer.py: patch 16, dim 1024, depth 24, heads 16
        return VisionTransformer(embed_dim=1024, depth=24, heads=16, hybrid=False)
    if name == "vit_base_patch16_384":      # patch 16, dim 768, depth 12, heads 12
        return VisionTransformer(embed_dim=768, depth=12, heads=12, hybrid=False)
    raise RuntimeError(f"timm shim: {name!r} is not restated")
