"""The refocus oracle against the UNMODIFIED reference module (build container) and the committed golden values."""
from pathlib import Path

import pytest
import torch

from oracle import reference_loader, refocus_oracle as ro

GOLDEN = Path(__file__).parent / "golden"


def _case(seed):
    rgb, depth, n_q, fidx, ap = ro.refocus_inputs(seed)
    qv = ro.compute_quantiles(depth, n_q)
    focus = torch.gather(qv, 1, fidx.unsqueeze(1))
    return rgb, depth, n_q, qv, focus, ap


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not on this box")
def test_refocus_oracle_equals_unmodified_reference():
    ref = reference_loader.load_reference_refocus()
    for seed in (0, 1):
        rgb, depth, n_q, qv, focus, ap = _case(seed)
        quantiles = torch.arange(0, n_q + 1) / n_q
        assert torch.equal(ref.compute_quantiles(depth, quantiles, eps=0.0001)[1].permute(1, 0), qv)
        out_ref, seg_ref = ref.refocus_image(rgb, depth, focus, ap, qv, True)
        out, seg = ro.refocus_image(rgb, depth, focus, ap, qv, True)
        assert torch.equal(seg, seg_ref)
        assert float((out - out_ref).abs().max()) <= 1e-6
        assert float((out_ref - rgb).abs().mean()) > 5e-3          # the inputs really get blurred


def test_refocus_oracle_golden():
    rec = torch.load(GOLDEN / "refocus_seed0.pt")
    rgb, depth, n_q, qv, focus, ap = _case(0)
    out = ro.refocus_image(rgb, depth, focus, ap, qv)
    assert torch.allclose(qv, rec["quantile_vals"], rtol=0, atol=1e-7)
    assert float((out[:, :, ::4, ::4] - rec["out_sub4"]).abs().max()) <= 2e-6
