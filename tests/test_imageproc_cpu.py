"""Host side of the device pre-processing: the restated Pillow coefficient tables reproduce Pillow's
resize bit for bit (numpy evaluation of exactly the arithmetic the kernels perform), on the geometries
demo.py produces (landscape / portrait / up-scaling / 8-bit grey)."""
import numpy as np
import pytest
import torch
from PIL import Image

from omnidata_b200 import imageproc as ip
from oracle import image_oracle as io_


def emulate(img: np.ndarray, size: int = 384) -> np.ndarray:
    """What the two kernels compute, in numpy int64 (same tables, same rounding, same 8-bit intermediate)."""
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    nw, nh = ip.resized_size(w, h, size)
    left, top = ip.center_crop_offset(nw, size), ip.center_crop_offset(nh, size)
    bh, kh, _ = ip.pil_bilinear_coeffs(w, nw)
    bv, kv, _ = ip.pil_bilinear_coeffs(h, nh)
    bh, kh, bv, kv = bh[left:left + size], kh[left:left + size], bv[top:top + size], kv[top:top + size]
    row0 = int(bv[:, 0].min())
    nrows = int((bv[:, 0] + bv[:, 1]).max()) - row0
    tmp = np.zeros((nrows, size, c), dtype=np.uint8)
    src = img.astype(np.int64)
    for x in range(size):
        x0, n = bh[x]
        acc = (1 << 21) + np.tensordot(src[row0:row0 + nrows, x0:x0 + n, :], kh[x, :n].astype(np.int64), axes=([1], [0]))
        tmp[:, x, :] = np.clip(acc >> 22, 0, 255)
    out = np.zeros((size, size, c), dtype=np.uint8)
    t64 = tmp.astype(np.int64)
    for y in range(size):
        y0, n = bv[y]
        acc = (1 << 21) + np.tensordot(kv[y, :n].astype(np.int64), t64[y0 - row0:y0 - row0 + n], axes=([0], [0]))
        out[y] = np.clip(acc >> 22, 0, 255)
    return out


@pytest.mark.parametrize("w,h,ch", [(640, 480, 3), (480, 640, 3), (1000, 751, 3), (384, 384, 3), (200, 150, 3),
                                    (517, 389, 1), (1920, 1080, 3), (385, 900, 3)])
def test_restated_pillow_resize_matches_pillow(w, h, ch):
    img = io_.synthetic_image(w, h, seed=w + h, channels=ch)
    nw, nh = ip.resized_size(w, h, 384)
    ref = img.resize((nw, nh), Image.BILINEAR)
    left, top = ip.center_crop_offset(nw, 384), ip.center_crop_offset(nh, 384)
    ref = np.asarray(ref.crop((left, top, left + 384, top + 384)))
    got = emulate(np.asarray(img))
    if ch == 1:
        got = got[..., 0]
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("task", ["depth", "normal"])
def test_reference_compose_equals_resize_crop_scale(task):
    """The oracle's Compose (the reference's literal transforms) == Pillow resize + crop + /255 [+ normalise]."""
    img = io_.synthetic_image(700, 500, seed=3)
    t = io_.reference_input_tensor(img, task)
    u8 = emulate(np.asarray(img))
    v = torch.from_numpy(u8).permute(2, 0, 1).float().div(255)
    if task == "depth":
        v = (v - 0.5) / 0.5
    assert torch.equal(t, v)


def test_restated_pillow_resize_random_geometries():
    """Seeded sweep over odd sizes / aspect ratios (incl. up-scaling from tiny images and panoramas)."""
    rng = np.random.RandomState(7)
    sizes = [(int(rng.randint(40, 900)), int(rng.randint(40, 900))) for _ in range(10)] + [(2000, 97), (61, 1500)]
    for w, h in sizes:
        img = io_.synthetic_image(w, h, seed=w * 31 + h)
        nw, nh = ip.resized_size(w, h, 384)
        ref = img.resize((nw, nh), Image.BILINEAR)
        left, top = ip.center_crop_offset(nw, 384), ip.center_crop_offset(nh, 384)
        ref = np.asarray(ref.crop((left, top, left + 384, top + 384)))
        assert np.array_equal(emulate(np.asarray(img)), ref), (w, h)
