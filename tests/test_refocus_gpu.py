"""Refocus kernels (quantiles by radix select, separable Gaussian blur stack, composite) against the oracle
restatement (== the unmodified reference module, tests/test_refocus_cpu.py) and the reference golden values."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


@pytest.mark.parametrize("seed,batch,size", [(0, 2, 128), (2, 3, 96)])
def test_refocus_matches_oracle(lib_built, seed, batch, size):
    from omnidata_b200 import refocus
    from oracle import refocus_oracle as ro
    rgb, depth, n_q, fidx, ap = ro.refocus_inputs(seed, batch, size)
    qv_ref = ro.compute_quantiles(depth, n_q)
    qv = refocus.compute_quantiles(depth.cuda(), n_q)
    assert torch.allclose(qv.cpu(), qv_ref, rtol=0, atol=1e-7)              # exact order statistics, one lerp
    focus = torch.gather(qv_ref, 1, fidx.unsqueeze(1))
    ref, seg_ref = ro.refocus_image(rgb, depth, focus, ap, qv_ref, True)
    out, seg = refocus.refocus_image(rgb.cuda(), depth.cuda(), focus.cuda(), ap.cuda(), qv_ref.cuda(), True)
    torch.cuda.synchronize()
    assert torch.equal(seg.cpu(), seg_ref)
    assert float((out.cpu() - ref).abs().max()) <= 5e-6
    if seed == 0 and batch == 2 and size == 128:
        rec = torch.load(GOLDEN / "refocus_seed0.pt")                        # values of the UNMODIFIED reference
        assert float((out.cpu()[:, :, ::4, ::4] - rec["out_sub4"]).abs().max()) <= 5e-6


def test_refocus_augmentation_factory(lib_built):
    """demo_refocus.py:50-69 usage: the factory draws its focus plane / aperture and returns an image of the same shape."""
    from omnidata_b200 import refocus
    from oracle import refocus_oracle as ro
    rgb, depth, n_q, _, _ = ro.refocus_inputs(1, 2, 128)
    aug = refocus.RefocusImageAugmentation(10, 0.5, 6.0)
    torch.manual_seed(0)
    a = aug(rgb.cuda(), depth.cuda())
    torch.manual_seed(0)
    b = aug(rgb.cuda(), depth.cuda())
    assert a.shape == rgb.shape and torch.equal(a, b) and bool(torch.isfinite(a).all())
    assert 0.0 <= float(a.min()) and float(a.max()) <= 1.0 + 1e-5
