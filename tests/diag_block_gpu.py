"""Diagnostics: one ResNetV2 bottleneck's backward, intermediate by intermediate, against float64 autograd."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200 import bwd, ops
from omnidata_b200.model import DPTDepthModel
from omnidata_b200.train import TrainEngine
from oracle import make_golden, weights
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
sd = weights.make_state_dict(0, 1)
model = DPTDepthModel(); model.load_state_dict(sd); model = model.to(dev)
eng = TrainEngine(model, "fp32")
x = make_golden.golden_input(1, seed=0).to(dev)
eng.forward(x)
rec = [r for r in eng.saved["blocks"] if r["tag"] == "s2b8"][0]
p = rec["p"]
P = eng.P
nchw = lambda t: t.double().permute(0, 3, 1, 2)
def std(w):
    s, m = torch.std_mean(w, dim=[1, 2, 3], keepdim=True, unbiased=False)
    return (w - m) / (s + 1e-8)
t_in = nchw(rec["t_in"]).requires_grad_(True)
w1, w2, w3 = (P[p + f"conv{i}.weight"].double().requires_grad_(True) for i in (1, 2, 3))
gs = [P[p + f"norm{i}.weight"].double().requires_grad_(True) for i in (1, 2, 3)]
bs = [P[p + f"norm{i}.bias"].double().requires_grad_(True) for i in (1, 2, 3)]
y1 = F.conv2d(t_in, std(w1)); y1.retain_grad()
a1 = F.relu(F.group_norm(y1, 32, gs[0], bs[0], 1e-5)); a1.retain_grad()
y2 = F.conv2d(a1, std(w2), padding=1); y2.retain_grad()
a2 = F.relu(F.group_norm(y2, 32, gs[1], bs[1], 1e-5)); a2.retain_grad()
y3 = F.conv2d(a2, std(w3)); y3.retain_grad()
out = F.relu(F.group_norm(y3, 32, gs[2], bs[2], 1e-5) + t_in)
print("fwd y1", rel(rec["y1"], y1.permute(0, 2, 3, 1)), "a1", rel(rec["a1"], a1.permute(0, 2, 3, 1)), "y2", rel(rec["y2"], y2.permute(0, 2, 3, 1)),
      "a2", rel(rec["a2"], a2.permute(0, 2, 3, 1)), "y3", rel(rec["y3"], y3.permute(0, 2, 3, 1)), "out", rel(rec["out"], out.permute(0, 2, 3, 1)))
print("mask mismatches a1", int(((rec["a1"] > 0) != (a1.permute(0, 2, 3, 1) > 0)).sum()), "a2", int(((rec["a2"] > 0) != (a2.permute(0, 2, 3, 1) > 0)).sum()),
      "out", int(((rec["out"] > 0) != (out.permute(0, 2, 3, 1) > 0)).sum()))
g = torch.Generator(device="cpu").manual_seed(5)
d_out = torch.randn(rec["out"].shape, generator=g).to(dev)
out.backward(nchw(d_out))
# engine pieces
Wt = eng.W
gbuf = torch.empty_like(d_out); bwd.mask_add(gbuf, d_out, mask=rec["out"])
dy3 = torch.empty_like(rec["y3"]); dg = torch.empty(1024, device=dev); db = torch.empty(1024, device=dev)
bwd.groupnorm_bwd(gbuf, rec["y3"], rec["st3"], P[p + "norm3.weight"], dy3, dg, db)
print("dy3", rel(dy3, y3.grad.permute(0, 2, 3, 1)), "dg3", rel(dg, gs[2].grad), "db3", rel(db, bs[2].grad))
da2 = torch.empty_like(rec["a2"]); ops.conv1x1(dy3, Wt["s2b8.w3"][1], da2)
print("da2", rel(da2, a2.grad.permute(0, 2, 3, 1)))
dy2 = torch.empty_like(rec["y2"]); dg2 = torch.empty(256, device=dev); db2 = torch.empty(256, device=dev)
bwd.groupnorm_bwd(da2, rec["y2"], rec["st2"], P[p + "norm2.weight"], dy2, dg2, db2, mask=rec["a2"])
print("dy2", rel(dy2, y2.grad.permute(0, 2, 3, 1)), "dg2", rel(dg2, gs[1].grad), "db2", rel(db2, bs[1].grad))
# same with the exact (fp64-derived) incoming gradient
da2x = a2.grad.permute(0, 2, 3, 1).float().contiguous()
bwd.groupnorm_bwd(da2x, rec["y2"], rec["st2"], P[p + "norm2.weight"], dy2, dg2, db2, mask=rec["a2"])
print("with exact da2: dy2", rel(dy2, y2.grad.permute(0, 2, 3, 1)), "dg2", rel(dg2, gs[1].grad), "db2", rel(db2, bs[1].grad))
st = rec["st2"]
yy = y2.detach().reshape(1, 32, -1)
print("stats mean err", float((st[0, :, 0].double() - yy.mean(2)[0]).abs().max()), "rstd rel", rel(st[0, :, 1], 1 / torch.sqrt(yy.var(2, unbiased=False)[0] + 1e-5)))
