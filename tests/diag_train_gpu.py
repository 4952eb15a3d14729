"""Diagnostics (not a pytest): per-tensor gradient error of the TrainEngine against float64 autograd of the oracle."""
import inspect
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import dpt_oracle, make_golden, weights  # noqa

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"

code = inspect.getsource(dpt_oracle.forward_fp32).replace(".float()", ".double()").replace("def forward_fp32", "def forward_fp64")
ns = dict(dpt_oracle.__dict__)
exec(code, ns)

sd = weights.make_state_dict(0, 1)
x = make_golden.golden_input(1, seed=0)
g = torch.Generator(device="cpu").manual_seed(123)
R = torch.randn(1, 384, 384, generator=g).to(dev)
if len(sys.argv) > 2 and sys.argv[2] == "loss":
    # R := d(ssi + 0.1 reg + 10 vn)/d(pred) at the fp32 forward's output for a perturbed-prediction target (the bench's
    # synthetic data): the linear functional sum(y * R) then has the train step's gradient
    import numpy as np
    from omnidata_b200 import losses
    from omnidata_b200.model import DPTDepthModel as _M
    x = torch.cat([make_golden.golden_input(1, seed=0), make_golden.golden_input(1, seed=7)])
    m0 = _M(); m0.load_state_dict(sd); m0 = m0.to(dev).eval(); m0.precision = "fp32"
    with torch.no_grad():
        p0 = m0(x.to(dev)).float().unsqueeze(1)
    noise = torch.rand(2, 1, 384, 384, generator=g).to(dev)
    gt = (p0 * (0.8 + 0.4 * noise) + 0.05 * torch.rand(2, 1, 384, 384, generator=g).to(dev)).clamp(0, 1)
    mask = (torch.rand(2, 1, 384, 384, generator=g) > 0.1).float().to(dev)
    np.random.seed(7)
    fn = losses.DepthStepLoss((384, 384))
    _, dpred = fn(p0, gt, mask, full_mix=True)
    R = dpred[:, 0].clone()
    print("real-loss gradient: |R| =", float(R.norm()), " max", float(R.abs().max()))
    del m0

leaves = {k: v.to(dev).double().requires_grad_(True) for k, v in sd.items()}
y64 = ns["forward_fp64"](leaves, x.to(dev).double())
grads = torch.autograd.grad((y64 * R.double()).sum(), list(leaves.values()), allow_unused=True)
g_ref = {k: (gg if gg is not None else torch.zeros_like(leaves[k])) for k, gg in zip(leaves, grads)}

# the reference's own fp32 autograd (torch library kernels) against the same float64 truth
l32 = {k: v.to(dev).float().requires_grad_(True) for k, v in sd.items()}
y32 = dpt_oracle.forward_fp32(l32, x.to(dev))
gr32 = torch.autograd.grad((y32 * R).sum(), list(l32.values()), allow_unused=True)
g_ref32 = {k: (gg if gg is not None else torch.zeros_like(l32[k])) for k, gg in zip(l32, gr32)}

# stock torch.autocast(bfloat16) training of the same network: the bf16 yardstick
lac = {k: v.to(dev).float().requires_grad_(True) for k, v in sd.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    yac = dpt_oracle.forward_fp32(lac, x.to(dev))
grac = torch.autograd.grad((yac.float() * R).sum(), list(lac.values()), allow_unused=True)
g_refac = {k: (gg if gg is not None else torch.zeros_like(lac[k])) for k, gg in zip(lac, grac)}

from omnidata_b200.model import DPTDepthModel  # noqa
model = DPTDepthModel()
model.load_state_dict(sd, strict=True)
model = model.to(dev).train()
model.precision = precision
y = model(x.to(dev))
(y * R).sum().backward()
torch.cuda.synchronize()
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
print("forward rel", rel(y.detach(), y64.detach()))
rows = []
for k, p in model.named_parameters():
    gr = g_ref[k]
    if float(gr.norm()) == 0:
        rows.append((0.0 if float(p.grad.norm()) == 0 else 9e9, k, 0.0, 0.0, 0.0))
        continue
    rows.append((rel(p.grad, gr), k, float(gr.norm()), rel(g_ref32[k], gr), rel(g_refac[k], gr)))
rows.sort(reverse=True)
import math
tot = lambda idx: math.sqrt(sum((r[idx] * r[2]) ** 2 for r in rows) / sum(r[2] ** 2 for r in rows))
print(f"global rel-L2 vs float64: engine {tot(0):.3e}   torch fp32 autograd {tot(3):.3e}   torch autocast bf16 {tot(4):.3e}")
for e, k, n, e32, eac in rows[:25]:
    print(f"{e:.3e}  (torch fp32: {e32:.3e}, autocast: {eac:.3e})  |g|={n:.3e}  {k}")
import re
for pat in ("blocks.0.attn.qkv.weight", "blocks.11.mlp.fc2.weight", "blocks.5.norm1.weight", "refinenet1.resConfUnit2.conv1.weight",
            "output_conv.0.weight", "layer1_rn.weight", "act_postprocess3.3.weight", "patch_embed.proj.weight", "stem.conv.weight",
            "stages.2.blocks.8.conv3.weight", "stages.2.blocks.8.norm3.bias", "stages.2.blocks.8.conv1.weight"):
    for e, k, n, e32, eac in rows:
        if k.endswith(pat):
            print(f"   {k}: engine {e:.3e} autocast {eac:.3e}")
rows = [(e, k, n) for e, k, n, _, _ in rows]
print("...")
for e, k, n in rows[-5:]:
    print(f"{e:.3e}  |g|={n:.3e}  {k}")
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"diag_train_{precision}.json").write_text(json.dumps([[e, k, n] for e, k, n in rows]))
