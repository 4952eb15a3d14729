"""Diagnostics: loss trajectory of the train step (bf16 / fp32 engine) next to torch autograd + torch.optim.Adam over the
reference arithmetic, same data, same VNL indices: python tests/diag_dynamics_gpu.py [steps] [batch] [lr]"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200 import losses, synthetic
from omnidata_b200.model import DPTDepthModel
from omnidata_b200.train import DepthTrainStep
from oracle import dpt_oracle
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
g = torch.Generator().manual_seed(2000)
rgb = (torch.rand(B, 3, 384, 384, generator=g) * 2 - 1).to(dev)
sd = synthetic.make_state_dict(0, 1)
m0 = DPTDepthModel(); m0.load_state_dict(sd); m0 = m0.to(dev).eval()
with torch.no_grad():
    p0 = m0(rgb).float().unsqueeze(1)
noise = torch.rand(B, 1, 384, 384, generator=g).to(dev)
gt = (p0 * (0.8 + 0.4 * noise) + 0.05 * torch.rand(B, 1, 384, 384, generator=g).to(dev)).clamp(0, 1)
mask = (torch.rand(B, 1, 384, 384, generator=g) > 0.1).float().to(dev)
np.random.seed(7)
vnl_ = losses.VNL_Loss(1.0, 1.0, (384, 384))
pts = [vnl_.select_index() for _ in range(steps)]
print("alive fraction of the initial prediction:", float((p0 > 0).float().mean()), "gt mean", float(gt.mean()))

for prec in ("bf16", "fp32"):
    model = DPTDepthModel(); model.load_state_dict(sd); model = model.to(dev).train()
    st = DepthTrainStep(model, lr=lr, clip=10.0, precision=prec)
    for i in range(steps):
        r = st.step(rgb, gt, mask, points=pts[i], full_mix=True).cpu()
        print(prec, i, " ".join(f"{float(v):.5f}" for v in r))

# torch reference: autograd over the oracle + torch.optim.Adam + clip_grad_norm_
leaves = {k: v.to(dev).float().requires_grad_(True) for k, v in sd.items()}
opt = torch.optim.Adam(list(leaves.values()), lr=lr)
midas = losses.MidasLoss(0.1, 4)
for i in range(steps):
    opt.zero_grad()
    y = dpt_oracle.forward_fp32(leaves, rgb).unsqueeze(1)
    pc = torch.clamp(y, 0, 1)
    mv = losses.make_valid_mask(mask)
    _, ssi, reg = midas(pc, gt, mv)
    vn = vnl_(pc, gt, points=pts[i])
    loss = ssi + 0.1 * reg + 10 * vn
    loss.backward()
    for p in leaves.values():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    nrm = torch.nn.utils.clip_grad_norm_(list(leaves.values()), 10.0)
    opt.step()
    print("torch", i, f"{float(loss):.5f} {float(ssi):.5f} {float(reg):.5f} {float(vn):.5f} {float(nrm):.5f}")
