"""BASELINE.json configs[0] — "demo.py --task normal, one 384x384 RGB image, CPU PyTorch reference forward
(plumbing, no GPU)": the whole demo pipeline (reference transforms -> network -> clamp -> ToPILImage) evaluated with
the oracle restatement, and in the build container compared with the UNMODIFIED reference module fed by the same
reference transforms.  This is the CPU-runnable end-to-end case the GPU entry-point tests mirror."""
import numpy as np
import pytest
import torch

from oracle import dpt_oracle, image_oracle as io_, reference_loader, weights


def _pipeline(forward):
    img = io_.synthetic_image(512, 384, seed=11)
    x = io_.reference_input_tensor(img, "normal").unsqueeze(0)            # demo.py:74-76,131-138
    assert x.shape == (1, 3, 384, 384) and 0.0 <= float(x.min()) and float(x.max()) <= 1.0
    with torch.no_grad():
        out = forward(x)                                                   # [1,3,384,384]
    return io_.reference_normal_post(out[0])                               # demo.py:140,150 -> uint8 HWC


def test_config0_oracle_pipeline_runs_and_is_deterministic():
    sd = weights.make_state_dict(0, 3)
    a = _pipeline(lambda x: dpt_oracle.forward_fp32(sd, x))
    b = _pipeline(lambda x: dpt_oracle.forward_fp32(sd, x))
    assert a.shape == (384, 384, 3) and a.dtype == np.uint8 and np.array_equal(a, b)
    assert a.max() > 0                                                     # the seeded head is alive


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not on this box")
def test_config0_oracle_equals_unmodified_reference_module():
    sd = weights.make_state_dict(0, 3)
    ref = reference_loader.load_reference_dpt(3).eval()
    ref.load_state_dict(sd, strict=True)
    a = _pipeline(lambda x: dpt_oracle.forward_fp32(sd, x))
    b = _pipeline(lambda x: ref(x))
    assert np.array_equal(a, b)
