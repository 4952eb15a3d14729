"""The loss restatement (oracle/loss_oracle.py) against the UNMODIFIED reference loss modules (build
container only) and against the committed golden values (any box)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import loss_oracle, reference_loader

GOLDEN = Path(__file__).parent / "golden"


def test_loss_oracle_matches_golden():
    rec = torch.load(GOLDEN / "losses_seed0.pt")
    pred, gt, mf = loss_oracle.loss_inputs(0)
    mask = loss_oracle.make_valid_mask(mf)
    total, ssi, reg = loss_oracle.midas_loss(pred, gt, mask)
    assert abs(float(total) - rec["midas_total"]) <= 2e-5 * abs(rec["midas_total"])
    assert abs(float(ssi) - rec["midas_ssi"]) <= 2e-5 * abs(rec["midas_ssi"])
    assert abs(float(reg) - rec["midas_reg"]) <= 2e-5 * abs(rec["midas_reg"])
    np.random.seed(0)
    pts = loss_oracle.vnl_select_index(384, 384)
    vn = loss_oracle.vnl_loss(pred, gt, pts)
    assert abs(float(vn) - rec["vnl"]) <= 2e-5 * abs(rec["vnl"])
    assert int(mask.sum()) == rec["mask_valid_count"]


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not on this box")
def test_loss_oracle_equals_unmodified_reference():
    MidasLoss, VNL_Loss = reference_loader.load_reference_losses()
    for seed in (1, 2):
        pred, gt, mf = loss_oracle.loss_inputs(seed)
        mask = loss_oracle.make_valid_mask(mf)
        ref = MidasLoss(alpha=0.1, scales=4, reduction="image-based")(pred, gt, mask)
        mine = loss_oracle.midas_loss(pred, gt, mask)
        for a, b in zip(ref, mine):
            assert abs(float(a) - float(b)) <= 1e-6 * abs(float(a))
        np.random.seed(seed)
        ref_v = VNL_Loss(1.0, 1.0, (384, 384))(pred, gt)
        np.random.seed(seed)
        pts = loss_oracle.vnl_select_index(384, 384)
        mine_v = loss_oracle.vnl_loss(pred, gt, pts)
        assert abs(float(ref_v) - float(mine_v)) <= 1e-6 * abs(float(ref_v))


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not on this box")
def test_normal_loss_oracle_equals_unmodified_reference():
    """masked_l1_loss / masked_cosine_angular_loss restatement == the reference functions (train_normal.py:247-258 mix)."""
    ref_l1, ref_cos = reference_loader.load_reference_masked_losses()
    for seed in (0, 1):
        pred, gt, mf = loss_oracle.normal_loss_inputs(seed)
        p = torch.clamp(pred, 0, 1)
        mask = loss_oracle.make_valid_mask(mf).repeat_interleave(3, 1)
        a, b = ref_l1(p.clone(), gt.clone(), mask), ref_cos(p.clone(), gt.clone(), mask)
        tot, l1, cos = loss_oracle.normal_step(pred, gt, mf)
        assert abs(float(a) - float(l1)) <= 1e-6 * abs(float(a))
        assert abs(float(b) - float(cos)) <= 1e-6 * abs(float(b))
        assert abs(float(b + 10 * a) - float(tot)) <= 1e-6 * abs(float(tot))


def test_normal_loss_oracle_golden():
    """Values of the unmodified reference functions on the seeded inputs (made in the build container)."""
    rec = torch.load(GOLDEN / "normal_losses_seed0.pt")
    tot, l1, cos = loss_oracle.normal_step(*loss_oracle.normal_loss_inputs(0))
    assert abs(float(l1) - rec["l1"]) <= 2e-6 * abs(rec["l1"])
    assert abs(float(cos) - rec["cos"]) <= 2e-6 * abs(rec["cos"])


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not on this box")
def test_loss_oracle_gradients_equal_unmodified_reference():
    """autograd through the oracle restatements == autograd through the reference modules (the GPU backward
    kernels are tested against the former on the GPU box)."""
    MidasLoss, VNL_Loss = reference_loader.load_reference_losses()
    pred, gt, mf = loss_oracle.loss_inputs(0)
    mask = loss_oracle.make_valid_mask(mf)
    p1 = pred.clone().requires_grad_(True)
    MidasLoss(alpha=0.1, scales=4, reduction="image-based")(p1, gt, mask)[0].backward()
    p2 = pred.clone().requires_grad_(True)
    loss_oracle.midas_loss(p2, gt, mask)[0].backward()
    assert float((p1.grad - p2.grad).norm() / p1.grad.norm()) <= 1e-6
    np.random.seed(0)
    p3 = pred.clone().requires_grad_(True)
    VNL_Loss(1.0, 1.0, (384, 384))(p3, gt).backward()
    np.random.seed(0)
    pts = loss_oracle.vnl_select_index(384, 384)
    p4 = pred.clone().requires_grad_(True)
    loss_oracle.vnl_loss(p4, gt, pts).backward()
    assert float((p3.grad - p4.grad).norm() / p3.grad.norm()) <= 1e-6
