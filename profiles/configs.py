"""The other BASELINE.json configurations on one GPU (not bench lines — bench.py measures configs[1]):
  configs[2]  depth AND normal networks on the same 64 images (two full DPT-Hybrid forwards per image)
  configs[3]  at N = 1: the per-GPU share of the 512-image batch as chunks of 64 (and the batch-size sweep)
CUDA events, CUDA-graph replay, inputs resident in HBM; prints one JSON line per configuration.

  python profiles/configs.py > gpurun_out/configs.json
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import synthetic  # noqa: E402
from omnidata_b200.model import DPTDepthModel  # noqa: E402


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = torch.device("cuda:0")
    depth = DPTDepthModel()
    depth.load_state_dict(synthetic.make_state_dict(0, 1))
    depth = depth.to(dev).eval()
    depth.use_cuda_graph = True
    normal = DPTDepthModel(num_channels=3)
    normal.load_state_dict(synthetic.make_state_dict(1, 3))
    normal = normal.to(dev).eval()
    normal.use_cuda_graph = True
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        x64 = (torch.rand(64, 3, 384, 384, generator=g) * 2 - 1).to(dev)
        xn64 = (x64 + 1) / 2

        def both():
            depth(x64)
            normal(xn64)
        ms = timed(both)
        print(json.dumps({"config": "configs[2]: depth + normal networks, bf16, batch 64, 1 GPU", "ms_per_step": round(ms, 3),
                          "images_per_s": round(64 / ms * 1e3, 1), "network_forwards_per_s": round(128 / ms * 1e3, 1),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
        for b in (1, 8, 64, 128):
            xb = (torch.rand(b, 3, 384, 384, generator=g) * 2 - 1).to(dev)
            ms = timed(lambda: depth(xb), steps=10 if b > 8 else 50)
            print(json.dumps({"config": f"depth, bf16, batch {b}, 1 GPU" + (" (configs[3] chunk size: 512 = 8 x 64)" if b == 64 else ""),
                              "ms_per_step": round(ms, 3), "images_per_s": round(b / ms * 1e3, 1),
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))


if __name__ == "__main__":
    main()
