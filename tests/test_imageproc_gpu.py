"""Device pre- / post-processing against the reference's own transforms (Pillow + torchvision), bit-exact
where the reference is integer / single-rounding arithmetic."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup(lib_built):
    yield


@pytest.mark.parametrize("w,h,ch", [(640, 480, 3), (480, 640, 3), (1000, 751, 3), (384, 384, 3), (200, 150, 3),
                                    (517, 389, 1), (1920, 1080, 3), (385, 900, 3), (4000, 3000, 3)])
@pytest.mark.parametrize("task", ["depth", "normal"])
def test_device_preprocess_bit_exact(w, h, ch, task):
    from omnidata_b200.imageproc import DevicePreprocessor
    from oracle import image_oracle as io_
    img = io_.synthetic_image(w, h, seed=w + h, channels=ch)
    ref = io_.reference_input_tensor(img, task)
    pre = DevicePreprocessor(task)
    u8 = torch.empty(384, 384, ch, device="cuda", dtype=torch.uint8)
    got = pre(np.asarray(img), out_u8=u8)
    torch.cuda.synchronize()
    assert got.shape == (3, 384, 384)
    assert torch.equal(got.cpu(), ref), f"max abs diff {(got.cpu() - ref).abs().max():.3e}"
    # second call re-uses the cached tables; device-resident input
    got2 = pre(torch.from_numpy(np.ascontiguousarray(np.asarray(img))).cuda())
    assert torch.equal(got2, got)


def test_device_depth_post():
    from omnidata_b200.imageproc import bicubic_resize
    from oracle import image_oracle as io_
    g = torch.Generator().manual_seed(0)
    out = torch.rand(1, 384, 384, generator=g) * 1.4 - 0.2          # exercises both clamps
    ref = io_.reference_depth_post(out)
    got = bicubic_resize(out.cuda(), (512, 512), clamp_in=True, clamp_out=True, invert=True)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert float((got.cpu() - ref).abs().max()) < 2e-6               # fp32, different summation order


def test_device_normal_post():
    from omnidata_b200.imageproc import to_uint8_hwc
    from oracle import image_oracle as io_
    g = torch.Generator().manual_seed(1)
    out = torch.rand(3, 384, 384, generator=g) * 1.2 - 0.1
    out[0, 0, :8] = torch.tensor([0.0, 1.0, 0.5, 1 / 255, 254.999 / 255, 0.999999, 2.0, -1.0])
    ref = io_.reference_normal_post(out)
    got = to_uint8_hwc(out.cuda(), clamp01=True)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), ref)
