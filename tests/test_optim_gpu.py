"""Fused clip_grad_norm_ + Adam step (row a21) against torch.optim.Adam / torch.nn.utils.clip_grad_norm_."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,max_norm", [(1000003, 10.0), (4096, 0.5), (123147 * 10 + 1, None)])
def test_flat_adam_matches_torch(lib_built, n, max_norm):
    from omnidata_b200.optim import FlatAdam
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    ref = torch.optim.Adam([ref_p], lr=1e-5)                       # train_depth.py:381-383
    mine_p = p0.clone().cuda()
    opt = FlatAdam(mine_p, lr=1e-5)
    for step in range(4):
        grad = torch.randn(n, generator=g) * (3.0 if step % 2 else 0.01)
        ref_p.grad = grad.clone()
        if max_norm is not None:
            ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], max_norm)   # what PL's gradient_clip_val does
        ref.step()
        norm = opt.step(grad.cuda(), max_norm=max_norm)
        torch.cuda.synchronize()
        if max_norm is not None:
            exact = float(grad.double().norm())                        # torch's fp32 reduction is itself ~1e-5 off
            assert abs(float(norm) - exact) <= 1e-6 * exact
            assert abs(float(ref_norm) - exact) <= 1e-4 * exact
        st = ref.state[ref_p]
        # (when the clip is active torch's coefficient carries the ~1e-5 error of its fp32 norm)
        rt = 1e-5 if max_norm is None else 1e-4
        # (m = 0.9 m + 0.1 g cancels for some elements: absolute tolerance relative to the typical magnitude)
        assert torch.allclose(opt.exp_avg.cpu(), st["exp_avg"], rtol=rt, atol=rt * float(st["exp_avg"].abs().max()))
        assert torch.allclose(opt.exp_avg_sq.cpu(), st["exp_avg_sq"], rtol=2 * rt, atol=1e-20)
        # the parameters themselves: the lr-sized update is below fp32 resolution of p, so agreement is
        # limited by one rounding of p per step (ulp(4) = 4.8e-7)
        assert float((mine_p.cpu() - ref_p.detach()).abs().max()) <= 5e-7 * (step + 1)


def test_flatten_parameters_views(lib_built):
    from omnidata_b200.optim import FlatAdam, flatten_parameters
    m = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3)).cuda()
    before = [p.detach().clone() for p in m.parameters()]
    flat = flatten_parameters(m)
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p.detach(), b)
    opt = FlatAdam(flat, lr=1e-2)
    opt.step(torch.ones_like(flat), max_norm=None)
    torch.cuda.synchronize()
    for p, b in zip(m.parameters(), before):                        # first Adam step: p -= lr * sign(g)
        assert torch.allclose(p.detach(), b - 1e-2, atol=1e-6)
