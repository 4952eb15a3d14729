"""TEST INFRASTRUCTURE ONLY — plain-PyTorch restatement of the reference's 3-D refocus augmentation
(omnidata_tools/torch/data/refocus_augmentation.py), usable on the GPU box where /root/reference is absent.
Pinned against the UNMODIFIED reference module in the build container (tests/test_refocus_cpu.py) and by golden
values (tests/golden/refocus_seed0.pt)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def compute_quantiles(depth, n_quantiles: int, eps: float = 0.0001):
    """:82-87 and :186-188 -> quantile_vals [B, n_quantiles + 1]."""
    quantiles = torch.arange(0, n_quantiles + 1, device=depth.device) / n_quantiles
    qv = torch.quantile(depth.reshape(depth.shape[0], -1), quantiles, dim=1)
    qv[0] -= eps
    qv[-1] += eps
    return qv.permute(1, 0).contiguous()


def blur_radii(quantile_vals, apertures, focus_dists):
    """compute_circle_of_confusion_no_magnification (:76-78)."""
    return apertures * torch.abs(quantile_vals - focus_dists) / quantile_vals


def separable_gaussian(img, r: float, cutoff: int):
    """:30-58 for one image [1,C,H,W]."""
    if r < 1e-1:
        return img
    n = torch.arange(0, cutoff) - (cutoff - 1.0) / 2.0
    fil = torch.exp(-n ** 2 / (2 * r * r)) if cutoff > 1 else torch.ones(1)
    filsum = fil.sum()
    c = img.shape[1]
    k = torch.stack([fil] * c, 0)
    half = cutoff // 2
    x = F.pad(img, (half, half, half, half), "replicate")
    x = F.conv2d(x, k.unsqueeze(1).unsqueeze(-2), groups=c) / filsum
    x = F.conv2d(x, k.unsqueeze(1).unsqueeze(-1), groups=c) / filsum
    return x


def refocus_image(rgb, depth, focus_dists, apertures, quantile_vals, return_segments: bool = False):
    """refocus_image (:144-157): rgb [B,3,H,W], depth [B,1,H,W], focus_dists / apertures [B,1], quantile_vals [B,Q+1]."""
    B = rgb.shape[0]
    depth_flat = depth.reshape(B, -1)
    idx = torch.searchsorted(quantile_vals, depth_flat)
    left = idx - 1
    qr = torch.gather(quantile_vals, 1, idx).reshape(depth.shape)
    ql = torch.gather(quantile_vals, 1, left).reshape(depth.shape)
    dist = qr - ql
    d_right, d_left = (qr - depth) / dist, (depth - ql) / dist
    radii = blur_radii(quantile_vals, apertures, focus_dists)
    stack = []
    for b in range(B):
        levels = []
        for r in radii[b]:
            cutoff = int(r * 3)
            if cutoff % 2 == 0:
                cutoff += 1
            levels.append(separable_gaussian(rgb[b:b + 1], float(r), cutoff))
        stack.append(torch.stack(levels, 1))
    stack = torch.cat(stack, 0)                                          # [B, L, 3, H, W]
    sl, sr = 1 - d_left ** 2, 1 - d_right ** 2                           # [B,1,H,W]
    li, ri = left.reshape(depth.shape), idx.reshape(depth.shape)
    a = torch.gather(stack, 1, li.unsqueeze(1).expand(-1, 1, 3, -1, -1)).squeeze(1)
    e = torch.gather(stack, 1, ri.unsqueeze(1).expand(-1, 1, 3, -1, -1)).squeeze(1)
    out = (sl * a + sr * e) / (sl + sr)
    return (out, li) if return_segments else out


def refocus_inputs(seed: int = 0, batch: int = 2, size: int = 128):
    """Seeded smooth RGB + depth maps and the random draws of RefocusImageAugmentation (:190-199) made explicit."""
    g = torch.Generator().manual_seed(300 + seed)
    yy, xx = torch.meshgrid(torch.arange(size).float(), torch.arange(size).float(), indexing="ij")
    rgb = torch.stack([torch.stack([0.5 + 0.4 * torch.sin(xx / (5 + c + b)) * torch.cos(yy / (7 + b)) for c in range(3)])
                       for b in range(batch)]) + 0.05 * torch.rand(batch, 3, size, size, generator=g)
    depth = 0.2 + 0.7 * torch.rand(batch, 1, 1, 1, generator=g) * (xx + yy)[None, None] / (2 * size) \
        + 0.1 * torch.rand(batch, 1, size, size, generator=g)
    n_q = 10
    focus_idx = torch.randint(1, n_q, (batch,), generator=g)
    # wide apertures: blur radii up to ~20 px, so that the long Gaussian kernels and the replicate padding matter
    apertures = torch.exp(torch.rand(batch, 1, generator=g) * (torch.log(torch.tensor(40.0)) - torch.log(torch.tensor(8.0)))
                          + torch.log(torch.tensor(8.0)))
    return rgb.clamp(0, 1), depth, n_q, focus_idx, apertures
