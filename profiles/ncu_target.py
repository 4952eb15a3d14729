"""Profiling target: one warm-up forward, then ONE forward (batch 32, configs[1]) bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly one step.

  # launch list (share of the step per kernel)
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python profiles/ncu_target.py
  # full capture of ViT block 0's qkv / proj / fc1 / fc2 GEMMs (conv_gemm launches 53..56 of the step)
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:conv_gemm -s 53 -c 4 -o gpurun_out/vit_gemm python profiles/ncu_target.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import synthetic  # noqa: E402
from omnidata_b200.model import DPTDepthModel  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = DPTDepthModel()
model.load_state_dict(synthetic.make_state_dict(0, 1))
model = model.cuda().eval()
x = torch.rand(batch, 3, 384, 384, device="cuda") * 2 - 1
with torch.no_grad():
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
