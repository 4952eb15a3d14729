"""Per-launch device times of one forward (eager, CUDA events around every C-ABI call): which layer
costs what.  Diagnostics, not a bench (event pairs add ~2 us per launch and defeat PDL overlap).

  python profiles/layer_times.py [batch] > gpurun_out/layers.txt
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import ops, synthetic  # noqa: E402
from omnidata_b200.model import DPTDepthModel  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    model = DPTDepthModel()
    model.load_state_dict(synthetic.make_state_dict(0, 1))
    model = model.cuda().eval()
    model.use_cuda_graph = False
    x = torch.rand(batch, 3, 384, 384, device="cuda") * 2 - 1
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        with ops.LaunchTimer() as lt:
            for _ in range(3):
                model(x)
        recs = lt.results()
    n = len(recs) // 3
    total = 0.0
    rows = []
    for i in range(n):
        name, info, _ = recs[i]
        t = min(recs[i][2], recs[n + i][2], recs[2 * n + i][2]) * 1e3
        total += t
        desc = ""
        if name == "odb_conv_gemm":
            fl = 2.0 * info["m"] * info["n"] * info["k"]
            desc = (f"m={info['m']:8d} n={info['n']:5d} k={info['k']:5d} taps={info['taps']} "
                    f"{info['w']}x{info['h']}  {fl / t / 1e6:7.0f} TFLOP/s")
        elif "flops" in info:
            desc = f"{info['flops'] / t / 1e6:7.0f} TFLOP/s"
        elif "bytes" in info:
            desc = f"{info['bytes'] / t / 1e3:7.0f} GB/s"
        rows.append((i, name, t, desc))
    for i, name, t, desc in rows:
        print(f"{i:4d} {name:28s} {t:8.1f} us  {desc}")
    print(f"total {total / 1e3:.3f} ms over {n} launches")
    agg = {}
    for _, name, t, _ in rows:
        agg[name] = agg.get(name, 0.0) + t
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {v / 1e3:8.3f} ms  {100 * v / total:5.1f} %")


if __name__ == "__main__":
    main()
