"""Micro-benchmark of the backward contractions (tcgen05 bgemm): python profiles/bgemm_micro.py [wgrad|att|all]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200 import bwd, ops
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def timeit(fn, flops, name):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms*1e3:9.1f} us  {flops / (ms * 1e-3) / 1e12:7.1f} TFLOP/s")


if what in ("wgrad", "all"):
    for (b, h, w, c, n, taps) in [(16, 96, 96, 256, 256, 9), (16, 192, 192, 256, 128, 9), (16, 384, 384, 128, 64, 9),
                                  (1, 1, 9232, 768, 3072, 1), (1, 1, 9232, 3072, 768, 1), (1, 1, 9232, 768, 2304, 1),
                                  (16, 96, 96, 64, 64, 9), (16, 96, 96, 64, 256, 1), (16, 24, 24, 768, 768, 1)]:
        x = torch.randn(b, h, w, c, device=dev).to(torch.bfloat16)
        dy = torch.randn(b, h, w, n, device=dev).to(torch.bfloat16)
        gp = torch.empty(n, taps * c, device=dev)
        tp = bwd.TAPS_3X3 if taps == 9 else bwd.TAPS_1
        timeit(lambda: bwd.conv_wgrad([x], tp, dy, gp), 2.0 * b * h * w * c * n * taps, f"wgrad b{b} {h}x{w} c{c} n{n} taps{taps}")
if what in ("att", "all"):
    b, n = 16, 577
    qkv = torch.randn(b, n, 2304, device=dev).to(torch.bfloat16)
    out = torch.empty(b, n, 768, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(b, 12, n, device=dev)
    ops.attention(qkv, out, lse=lse)
    d_o = torch.randn(b, n, 768, device=dev).to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    timeit(lambda: bwd.attention_bwd(qkv, out, d_o, lse, dqkv), 10.0 * b * 12 * n * n * 64, "attention_bwd b16")
