"""Per-launch timing of one train step (CUDA events around every C-ABI call): python profiles/train_layers.py [batch]"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200 import ops, synthetic
from omnidata_b200.model import DPTDepthModel
from omnidata_b200.train import DepthTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model = DPTDepthModel(); model.load_state_dict(synthetic.make_state_dict(0, 1)); model = model.to(dev).train()
step = DepthTrainStep(model, precision="bf16")
g = torch.Generator().manual_seed(0)
rgb = (torch.rand(B, 3, 384, 384, generator=g) * 2 - 1).to(dev)
gt = torch.rand(B, 1, 384, 384, generator=g).to(dev)
mask = (torch.rand(B, 1, 384, 384, generator=g) > 0.1).float().to(dev)
np.random.seed(0)
for _ in range(2):
    step.step(rgb, gt, mask, full_mix=True)
torch.cuda.synchronize()
with ops.LaunchTimer() as lt:
    step.step(rgb, gt, mask, full_mix=True)
recs = lt.results()
tot = sum(t for _, _, t in recs)
print(f"{len(recs)} timed calls, {tot:.2f} ms")
rows = []
for i, (name, info, t) in enumerate(recs):
    fl = 2.0 * info["m"] * info["n"] * info["k"] if "m" in info else info.get("flops", 0.0)
    rows.append((t, i, name, info, fl))
rows.sort(reverse=True)
for t, i, name, info, fl in rows[:70]:
    desc = {k: v for k, v in info.items() if k in ("m", "n", "k", "taps", "w", "h", "wgrad")}
    print(f"{t*1e3:8.1f} us  #{i:4d} {name:22s} {fl / (t * 1e-3) / 1e12 if fl else 0:7.1f} TF/s  {desc}")
