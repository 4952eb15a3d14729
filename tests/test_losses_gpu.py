"""Loss forward kernels (through the C ABI) against the oracle restatement and the reference golden values.
fp32 path: tolerance 1e-5 relative (BASELINE north star: 1e-5 in fp32)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
TOL = 1e-5


def close(a, b, tol=TOL):
    return abs(float(a) - float(b)) <= tol * max(abs(float(b)), 1e-6)


@pytest.mark.parametrize("seed,batch", [(0, 2), (3, 5)])
def test_losses_match_oracle_and_golden(lib_built, seed, batch):
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.loss_inputs(seed, batch)
    mask_ref = loss_oracle.make_valid_mask(mf)
    mask = losses.make_valid_mask(mf.cuda())
    assert torch.equal(mask.cpu(), mask_ref)                              # bit-exact
    tot, ssi, reg = losses.MidasLoss(alpha=0.1, scales=4)(pred.cuda(), gt.cuda(), mask)
    rt, rs, rr = loss_oracle.midas_loss(pred, gt, mask_ref)
    assert close(ssi, rs) and close(reg, rr) and close(tot, rt), (float(ssi), float(rs), float(reg), float(rr))
    np.random.seed(seed)
    vnl = losses.VNL_Loss(1.0, 1.0, (384, 384))
    v = vnl(pred.cuda(), gt.cuda())
    np.random.seed(seed)
    pts = loss_oracle.vnl_select_index(384, 384)
    rv = loss_oracle.vnl_loss(pred, gt, pts)
    assert close(v, rv), (float(v), float(rv))
    if seed == 0 and batch == 2:
        rec = torch.load(GOLDEN / "losses_seed0.pt")                      # values of the UNMODIFIED reference
        assert close(tot, rec["midas_total"], 2e-5) and close(v, rec["vnl"], 2e-5)
    # deterministic
    tot2, _, _ = losses.MidasLoss()(pred.cuda(), gt.cuda(), mask)
    assert float(tot2) == float(tot)


def test_shared_step_loss_mix(lib_built):
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.loss_inputs(7, 2)
    pred = pred * 1.3 - 0.1                                              # exercise the clamp
    np.random.seed(7)
    out = losses.depth_step_losses(pred.cuda(), gt.cuda(), mf.cuda(), losses.MidasLoss(), losses.VNL_Loss(1.0, 1.0, (384, 384)))
    pc = pred.clamp(0, 1)
    mask = loss_oracle.make_valid_mask(mf)
    _, ssi, reg = loss_oracle.midas_loss(pc, gt, mask)
    np.random.seed(7)
    vn = loss_oracle.vnl_loss(pc, gt, loss_oracle.vnl_select_index(384, 384))
    assert close(out["depth_loss"], ssi + 0.1 * reg + 10 * vn)
    early = losses.depth_step_losses(pred.cuda(), gt.cuda(), mf.cuda(), losses.MidasLoss(),
                                     losses.VNL_Loss(1.0, 1.0, (384, 384)), global_step=10)
    assert close(early["depth_loss"], ssi) and early["vn_loss"] == 0


@pytest.mark.parametrize("seed,batch", [(0, 2), (4, 3)])
def test_normal_losses(lib_built, seed, batch):
    """masked_l1 + masked cosine-angular losses of train_normal.py in one kernel pass vs the oracle / reference golden."""
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.normal_loss_inputs(seed, batch)
    rt, rl1, rcos = loss_oracle.normal_step(pred, gt, mf)
    out = losses.normal_step_losses(pred.cuda(), gt.cuda(), mf.cuda())
    assert close(out["l1_loss"], rl1) and close(out["cos_loss"], rcos) and close(out["normal_loss"], rt)
    if seed == 0 and batch == 2:
        rec = torch.load(GOLDEN / "normal_losses_seed0.pt")
        assert close(out["l1_loss"], rec["l1"], 2e-5) and close(out["cos_loss"], rec["cos"], 2e-5)
    # the reference's own calling convention (mask repeated over the channels, pre-clamped prediction)
    mask3 = losses.make_valid_mask(mf.cuda()).repeat_interleave(3, 1)
    tot2, l12, cos2 = losses.normal_losses(pred.cuda().clamp(0, 1), gt.cuda(), mask3)
    assert float(l12) == float(out["l1_loss"]) and float(cos2) == float(out["cos_loss"])


@pytest.mark.parametrize("seed,batch", [(0, 2), (5, 3)])
def test_midas_loss_backward(lib_built, seed, batch):
    """odb_midas_loss_bwd (through torch.autograd) against float64 autograd of the oracle restatement, whose
    gradient equals the unmodified reference module's (tests/test_losses_cpu.py)."""
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.loss_inputs(seed, batch)
    mask = loss_oracle.make_valid_mask(mf)
    for weights in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 0.3, -0.2)):
        p64 = pred.double().requires_grad_(True)
        tot, ssi, reg = loss_oracle.midas_loss(p64, gt.double(), mask)
        (weights[0] * tot + weights[1] * ssi + weights[2] * reg).backward()
        ref = p64.grad.float()
        p = pred.cuda().requires_grad_(True)
        t2, s2, r2 = losses.MidasLoss(alpha=0.1, scales=4)(p, gt.cuda(), mask.cuda())
        (weights[0] * t2 + weights[1] * s2 + weights[2] * r2).backward()
        got = p.grad.cpu()
        err = float((got - ref).norm() / ref.norm())
        assert err <= 1e-5, (weights, err)
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    # invalid pixels get no gradient; deterministic
    assert float(got[~mask].abs().max()) == 0.0
    p2 = pred.cuda().requires_grad_(True)
    t3, s3, r3 = losses.MidasLoss(alpha=0.1, scales=4)(p2, gt.cuda(), mask.cuda())
    (1.0 * t3 + 0.3 * s3 - 0.2 * r3).backward()
    assert torch.equal(p2.grad.cpu(), got)


@pytest.mark.parametrize("seed,batch", [(0, 2), (6, 3)])
def test_vnl_loss_backward_and_train_step_gradient(lib_built, seed, batch):
    """odb_vnl_loss_bwd against fp32 autograd of the oracle (same fp32 mask decisions), and the gradient of the whole
    train_depth.py loss mix (clamp -> ssi + 0.1 reg + 10 vn) with respect to the raw network output."""
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.loss_inputs(seed, batch)
    np.random.seed(seed)
    pts = loss_oracle.vnl_select_index(384, 384)
    p32 = pred.clone().requires_grad_(True)
    loss_oracle.vnl_loss(p32, gt, pts).backward()
    ref = p32.grad
    p = pred.cuda().requires_grad_(True)
    np.random.seed(seed)
    vnl = losses.VNL_Loss(1.0, 1.0, (384, 384))
    v = vnl(p, gt.cuda())
    v.backward()
    got = p.grad.cpu()
    assert float((got - ref).norm() / ref.norm()) <= 2e-4
    p2 = pred.cuda().requires_grad_(True)
    np.random.seed(seed)
    vnl(p2, gt.cuda()).backward()
    assert torch.equal(p2.grad.cpu(), got)                               # fixed-point scatter: bit-reproducible
    # ---- the train-step mix, through the clamp (train_depth.py:263-279)
    raw = (pred * 1.3 - 0.1)
    r32 = raw.clone().requires_grad_(True)
    dp = torch.clamp(r32, 0, 1)
    mask = loss_oracle.make_valid_mask(mf)
    _, ssi, reg = loss_oracle.midas_loss(dp, gt, mask)
    vn = loss_oracle.vnl_loss(dp, gt, pts)
    (ssi + 0.1 * reg + 10 * vn).backward()
    r = raw.cuda().requires_grad_(True)
    np.random.seed(seed)
    out = losses.depth_step_losses(r, gt.cuda(), mf.cuda(), losses.MidasLoss(), losses.VNL_Loss(1.0, 1.0, (384, 384)))
    out["depth_loss"].backward()
    g = r.grad.cpu()
    assert float((g - r32.grad).norm() / r32.grad.norm()) <= 2e-4


def test_normal_losses_backward(lib_built):
    """odb_normal_loss_bwd against autograd of the oracle (== the reference functions, tests/test_losses_cpu.py)."""
    from omnidata_b200 import losses
    from oracle import loss_oracle
    pred, gt, mf = loss_oracle.normal_loss_inputs(2, 2)
    for weights in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)):
        p64 = pred.double().requires_grad_(True)
        tot, l1, cos = loss_oracle.normal_step(p64, gt.double(), mf.double())
        (weights[0] * tot + weights[1] * l1 + weights[2] * cos).backward()
        ref = p64.grad.float()
        p = pred.cuda().requires_grad_(True)
        out = losses.normal_step_losses(p, gt.cuda(), mf.cuda())
        (weights[0] * out["normal_loss"] + weights[1] * out["l1_loss"] + weights[2] * out["cos_loss"]).backward()
        got = p.grad.cpu()
        assert float((got - ref).norm() / ref.norm()) <= 1e-5, weights
