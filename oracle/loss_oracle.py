"""TEST INFRASTRUCTURE — functional CPU restatement (plain PyTorch fp32) of the reference's depth-training
losses, usable on the GPU box where /root/reference is absent.  Validated in the build container against
the UNMODIFIED reference modules (tests/test_losses_cpu.py) — citations into omnidata_tools/torch/.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def make_valid_mask(mask_float: torch.Tensor, max_pool_size: int = 4) -> torch.Tensor:
    """train_depth.py:215-242 (4-D input)."""
    h, w = mask_float.shape[2], mask_float.shape[3]
    m = F.max_pool2d(1 - mask_float, kernel_size=max_pool_size)
    m = F.interpolate(m, (h, w), mode="nearest")
    return m == 0


def _shift_scale(x, mask):
    """masked_shift_and_scale for one tensor (losses/midas_loss.py:33-56)."""
    xn = x.clone()
    xn[~mask] = float("nan")
    n1 = mask.flatten(2).sum(-1, keepdim=True) + 1
    t = xn.flatten(2).nanmedian(-1, keepdim=True)[0].unsqueeze(-1)
    t[torch.isnan(t)] = 0
    d = torch.abs(x - t)
    d[~mask] = 0
    s = (d.flatten(2).sum(-1, keepdim=True) / n1).unsqueeze(-1)
    return (x - t) / (s + 1e-6)


def midas_loss(prediction, target, mask, alpha: float = 0.1, scales: int = 4):
    """MidasLoss(alpha, scales, 'image-based').forward (losses/midas_loss.py:137-157) -> (total, ssi, reg)."""
    pa, ga = _shift_scale(prediction, mask), _shift_scale(target, mask)
    e = torch.abs(pa - ga)
    e[~mask] = 0
    ssi = e.sum() / mask.sum()                                           # masked_l1_loss
    pi, ti, m = 1 / (prediction.squeeze(1) + 1e-6), 1 / (target.squeeze(1) + 1e-6), mask.squeeze(1)
    a00, a01, a11 = (m * pi * pi).sum((1, 2)), (m * pi).sum((1, 2)), m.sum((1, 2))
    b0, b1 = (m * pi * ti).sum((1, 2)), (m * ti).sum((1, 2))
    det = a00 * a11 - a01 * a01                                          # compute_scale_and_shift :10-30
    x0, x1 = torch.zeros_like(b0), torch.zeros_like(b1)
    ok = det != 0
    x0[ok] = (a11[ok] * b0[ok] - a01[ok] * b1[ok]) / (det[ok] + 1e-6)
    x1[ok] = (-a01[ok] * b0[ok] + a00[ok] * b1[ok]) / (det[ok] + 1e-6)
    pssi = x0.view(-1, 1, 1) * pi + x1.view(-1, 1, 1)
    reg = 0
    for s in range(scales):                                              # GradientMatchingTerm :114-134
        st = 2 ** s
        p_, t_, m_ = pssi[:, ::st, ::st], ti[:, ::st, ::st], m[:, ::st, ::st]
        M = m_.sum((1, 2))
        d = m_ * (p_ - t_)
        gx = torch.abs(d[:, :, 1:] - d[:, :, :-1]) * (m_[:, :, 1:] * m_[:, :, :-1])
        gy = torch.abs(d[:, 1:, :] - d[:, :-1, :]) * (m_[:, 1:, :] * m_[:, :-1, :])
        il = gx.sum((1, 2)) + gy.sum((1, 2))
        nz = M != 0
        il[nz] = il[nz] / M[nz]                                          # reduction_image_based :71-79
        reg = reg + il.mean()
    return ssi + alpha * reg, ssi, reg


def vnl_select_index(h: int, w: int, sample_ratio: float = 0.15):
    """VNL_Loss.select_index (losses/virtual_normal_loss.py:52-72): flat indices of three point sets."""
    num = h * w
    pts = []
    for _ in range(3):
        p = np.random.choice(num, int(num * sample_ratio), replace=True)
        np.random.shuffle(p)
        pts.append(p.astype(np.int64))
    return pts


def vnl_loss(first, second, points, fx: float = 1.0, fy: float = 1.0, delta_z: float = 1e-4, select: bool = True):
    """VNL_Loss.forward(first, second) (losses/virtual_normal_loss.py:151-194) for given triplets."""
    B, _, H, W = first.shape
    u = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(1, H, W) - float(W // 2)
    v = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(1, H, W) - float(H // 2)

    def xyz(d):                                                          # transfer_xyz :44-50
        return torch.cat([u * torch.abs(d) / fx, v * torch.abs(d) / fy, d], 1).permute(0, 2, 3, 1)

    def groups(pw):                                                      # form_pw_groups :74-93
        return torch.stack([pw[:, torch.as_tensor(p // W), torch.as_tensor(p % W), :] for p in points], 3)

    g, d = groups(xyz(first)), groups(xyz(second))                       # [B,N,3(xyz),3(pts)]
    diff = torch.stack([g[..., 1] - g[..., 0], g[..., 2] - g[..., 0], g[..., 2] - g[..., 1]], 3)
    q = diff.reshape(-1, 3, 3).permute(0, 2, 1)                          # filter_mask :95-128
    qn = q.norm(2, dim=2)
    nm = torch.bmm(qn.view(-1, 3, 1), qn.view(-1, 1, 3))
    en = (torch.bmm(q, diff.reshape(-1, 3, 3)) / (nm + 1e-8)).view(-1, 9)
    mask_cos = (torch.sum((en > 0.867) + (en < -0.867), 1) > 3).view(B, -1)
    mask_pad = torch.sum(g[:, :, 2, :] > delta_z, 2) == 3
    near = [torch.sum(torch.abs(diff[:, :, c, :]) < 0.005, 2) > 0 for c in range(3)]
    mask = mask_pad & ~((near[0] & near[1] & near[2]) | mask_cos)
    d = d.clone()
    d[d[:, :, 2, :] == 0] = 0.0001                                       # :144 (indexes the coordinate axis)
    mb = mask.repeat(1, 9).reshape(B, 3, 3, -1).permute(0, 3, 1, 2)
    gp, dp = g[mb].reshape(1, -1, 3, 3), d[mb].reshape(1, -1, 3, 3)
    gn = torch.cross(gp[..., 1] - gp[..., 0], gp[..., 2] - gp[..., 0], dim=2)
    dn = torch.cross(dp[..., 1] - dp[..., 0], dp[..., 2] - dp[..., 0], dim=2)
    gl, dl = gn.norm(2, dim=2, keepdim=True), dn.norm(2, dim=2, keepdim=True)
    gl = gl + (gl == 0).float() * 0.01
    dl = dl + (dl == 0).float() * 0.01
    loss = torch.abs(gn / gl - dn / dl).sum(2).sum(0)
    if select:
        loss = torch.sort(loss)[0][int(loss.numel() * 0.25):]
    return loss.mean()


def loss_inputs(seed: int = 0, batch: int = 2, size: int = 384):
    """Seeded, well-conditioned train-step tensors (depths in [0.05, 1], a few exact zeros in the
    prediction as a clamped ReLU output has, ~3 % invalid pixels)."""
    g = torch.Generator().manual_seed(100 + seed)
    pred = 0.05 + 0.95 * torch.rand(batch, 1, size, size, generator=g)
    gt = 0.05 + 0.95 * torch.rand(batch, 1, size, size, generator=g)
    pred[torch.rand(batch, 1, size, size, generator=g) < 0.01] = 0.0
    mask_float = (torch.rand(batch, 1, size, size, generator=g) > 0.03).float()
    return pred, gt, mask_float


def masked_l1_loss(preds, target, mask_valid):
    """losses/masked_losses.py:4-7."""
    e = (preds - target).abs()
    e = e * mask_valid
    return e.sum() / mask_valid.sum()


def masked_cosine_angular_loss(preds, target, mask_valid):
    """losses/masked_losses.py:14-23."""
    preds = (2 * preds - 1).clamp(-1, 1)
    target = (2 * target - 1).clamp(-1, 1)
    mv = mask_valid[:, 0, :, :].bool()
    preds = preds.permute(0, 2, 3, 1)[mv, :]
    target = target.permute(0, 2, 3, 1)[mv, :]
    pn = F.normalize(preds, p=2, dim=1)
    tn = F.normalize(target, p=2, dim=1)
    return torch.mean(-torch.sum(pn * tn, dim=1))


def normal_step(normal_preds, normal_gt, mask_float):
    """train_normal.py:247-265 -> (normal_loss, l1_loss, cos_loss)."""
    normal_preds = torch.clamp(normal_preds, 0, 1)
    mask_valid = make_valid_mask(mask_float).repeat_interleave(3, 1)
    l1 = masked_l1_loss(normal_preds, normal_gt, mask_valid)
    cos = masked_cosine_angular_loss(normal_preds, normal_gt, mask_valid)
    return cos + 10 * l1, l1, cos


def normal_loss_inputs(seed: int = 0, batch: int = 2, size: int = 384):
    """Seeded normal-map-like tensors in [0, 1] (prediction slightly outside to exercise the clamp)."""
    g = torch.Generator().manual_seed(200 + seed)
    gt = torch.rand(batch, 3, size, size, generator=g)
    # a prediction correlated with the target (mean cosine ~0.9: a well-conditioned loss value), 10 % outside [0, 1]
    pred = (gt + 0.25 * torch.randn(batch, 3, size, size, generator=g)) * 1.2 - 0.1
    mask_float = (torch.rand(batch, 1, size, size, generator=g) > 0.03).float()
    return pred, gt, mask_float
