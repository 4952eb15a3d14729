"""3-D refocus augmentation on the device: mirror of the reference's
`data/refocus_augmentation.py::RefocusImageAugmentation` (used by demo_refocus.py:50-69).

    aug = RefocusImageAugmentation(n_quantiles=10, aperture_min=0.001, aperture_max=6)
    refocused = aug(rgb, depth)          # rgb [B,3,H,W], depth [B,1,H,W], CUDA fp32

The random draws (focus quantile, log-uniform aperture) use the same torch calls in the same order as the
reference (:190-199); `refocus_image` is the deterministic part with the draws passed in."""
from __future__ import annotations

import torch

from . import _capi
from ._capi import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _capi.OdbError(f"{name}: CUDA tensor required (no CPU path)")
    return t.detach().float().contiguous()


@_capi.on_tensor_device
def compute_quantiles(depth: torch.Tensor, n_quantiles: int, eps: float = 0.0001) -> torch.Tensor:
    """compute_quantiles (:82-87) + the permute of :188 -> quantile_vals [B, n_quantiles + 1]."""
    d = _f32c(depth, "depth")
    b = d.shape[0]
    h, w = d.shape[-2:]
    qv = torch.empty(b, n_quantiles + 1, device=d.device, dtype=torch.float32)
    check(lib().odb_refocus_quantiles(d.data_ptr(), b, h, w, n_quantiles, eps, qv.data_ptr(), _stream()),
          "odb_refocus_quantiles")
    return qv


@_capi.on_tensor_device
def refocus_image(rgb, depth, focus_distance, aperture_size, quantile_vals, return_segments: bool = False):
    """refocus_image (:144-157)."""
    x, d, qv = _f32c(rgb, "rgb"), _f32c(depth, "depth"), _f32c(quantile_vals, "quantile_vals")
    b, c, h, w = x.shape
    if c != 3 or d.numel() != b * h * w:
        raise _capi.OdbError("refocus_image: rgb [B,3,H,W], depth [B,1,H,W]")
    levels = qv.shape[1]
    # compute_circle_of_confusion_no_magnification (:76-78): B x levels numbers
    radii = (aperture_size.to(qv) * torch.abs(qv - focus_distance.to(qv)) / qv).contiguous()
    tmp = torch.empty(b, levels, 3, h, w, device=x.device, dtype=torch.float32)
    stack = torch.empty_like(tmp)
    out = torch.empty_like(x)
    seg = torch.empty(b, 1, h, w, device=x.device, dtype=torch.int32) if return_segments else None
    check(lib().odb_refocus_compose(x.data_ptr(), d.data_ptr(), qv.data_ptr(), radii.data_ptr(), b, h, w, levels,
                                    tmp.data_ptr(), stack.data_ptr(), out.data_ptr(),
                                    None if seg is None else seg.data_ptr(), _stream()), "odb_refocus_compose")
    return (out, seg.long()) if return_segments else out


def RefocusImageAugmentation(n_quantiles, aperture_min, aperture_max, return_segments=False):
    """Same signature and random-draw sequence as the reference factory (:163-203)."""
    def refocus_image_(rgb, depth):
        with torch.no_grad():
            device = depth.device
            quantile_vals = compute_quantiles(depth, n_quantiles, eps=0.0001)
            focus_dist_idxs = torch.randint(low=1, high=n_quantiles, size=(rgb.shape[0],), device=device)
            focus_dists = torch.gather(quantile_vals, 1, focus_dist_idxs.unsqueeze(1))
            log_min = torch.log(torch.tensor(aperture_min, device=device))
            log_max = torch.log(torch.tensor(aperture_max, device=device))
            apertures = torch.exp(torch.rand(size=(rgb.shape[0], 1), device=device) * (log_max - log_min) + log_min)
            return refocus_image(rgb, depth, focus_dists, apertures, quantile_vals, return_segments)
    return refocus_image_
