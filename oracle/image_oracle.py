"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the reference's image pre- / post-processing, evaluated
with the reference's own third-party stack (Pillow + torchvision, both present in the image).

Follows omnidata_tools/torch/demo.py:
  trans_totensor   :74-76 (normal), :92-95 (depth)
  depth post       :140-145   (clamp, bicubic to 512, clamp, 1 - x; the colormap is presentation)
  normal post      :140, :150 (clamp, ToPILImage)
Parity is pinned: these ARE the library calls the reference makes, on the same inputs.
"""
from __future__ import annotations

import numpy as np
import PIL
import torch
import torch.nn.functional as F
from PIL import Image
from torchvision import transforms


def trans_totensor(task: str, image_size: int = 384):
    if task == "normal":
        # get_transform('rgb', image_size=None) is transforms.ToTensor() for 8-bit RGB (data/transforms.py)
        return transforms.Compose([transforms.Resize(image_size, interpolation=PIL.Image.BILINEAR),
                                   transforms.CenterCrop(image_size), transforms.ToTensor()])
    return transforms.Compose([transforms.Resize(image_size, interpolation=PIL.Image.BILINEAR),
                               transforms.CenterCrop(image_size), transforms.ToTensor(),
                               transforms.Normalize(mean=0.5, std=0.5)])


def reference_input_tensor(img: Image.Image, task: str) -> torch.Tensor:
    """demo.py:131-138: trans_totensor(img)[:3], single channel repeated -> [3, 384, 384]."""
    t = trans_totensor(task)(img)[:3].unsqueeze(0)
    if t.shape[1] == 1:
        t = t.repeat_interleave(3, 1)
    return t[0]


def reference_depth_post(output: torch.Tensor) -> torch.Tensor:
    """output: model(img) [1, 384, 384] -> [1, 512, 512] as handed to plt.imsave (demo.py:140-145)."""
    output = output.clamp(min=0, max=1)
    output = F.interpolate(output.unsqueeze(0), (512, 512), mode="bicubic").squeeze(0)
    output = output.clamp(0, 1)
    return 1 - output


def reference_normal_post(output: torch.Tensor) -> np.ndarray:
    """output: model(img)[0] [3, 384, 384] -> the uint8 HWC array ToPILImage saves (demo.py:140,150)."""
    return np.asarray(transforms.ToPILImage()(output.clamp(min=0, max=1)))


def synthetic_image(w: int, h: int, seed: int = 0, channels: int = 3) -> Image.Image:
    """Smooth structure plus noise: exercises both the antialiasing filter and the rounding."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    planes = []
    for c in range(channels):
        a = 127.5 + 90 * np.sin(xx / (7.0 + 3 * c) + c) * np.cos(yy / (11.0 + 2 * c)) + rng.randn(h, w) * 25
        planes.append(np.clip(a, 0, 255).astype(np.uint8))
    arr = planes[0] if channels == 1 else np.stack(planes, -1)
    return Image.fromarray(arr, mode="L" if channels == 1 else "RGB")
