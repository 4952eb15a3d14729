"""Micro-benchmarks (not a pytest): A/B timing of conv_gemm modes on the layer shapes of the network."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import ops as o  # noqa


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    B = 32
    g = torch.Generator().manual_seed(0)
    # head tail: conv3x3 128 -> 32 + fused 1x1, 384x384
    x = torch.randn(B, 384, 384, 128, generator=g).to(dev).to(torch.bfloat16)
    w = o.pack_conv_weight(torch.randn(32, 128, 3, 3, generator=g).to(dev) * 0.03)
    bias = torch.zeros(32, device=dev)
    hw, hb = torch.randn(1, 32, device=dev), torch.zeros(1, device=dev)
    hout = torch.empty(B, 1, 384, 384, device=dev)
    for halo in (-1, 1):
        for tile in (None, (32, 4), (128, 1)):
            if halo == 1 and tile is not None:
                continue
            t = timeit(lambda: o.conv3x3(x, w, None, bias=bias, head=(hw, hb, hout, True), halo=halo, tile=tile))
            print(f"head2 128->32 384^2  halo={halo:2d} tile={tile}: {t:8.1f} us  ({B * 384 * 384 * 32 * 1152 * 2 / t / 1e6:.0f} TFLOP/s)")
    del x, hout
    shapes = [("head0 256->128 192^2", 192, 256, 128), ("rn1 256->256 96^2", 96, 256, 256), ("s0.c2 64->64 96^2", 96, 64, 64),
              ("s1.c2 128->128 48^2", 48, 128, 128), ("s2.c2 256->256 24^2", 24, 256, 256), ("ff2 256->256 48^2", 48, 256, 256)]
    for name, hw_, c, n in shapes:
        x = torch.randn(B, hw_, hw_, c, generator=g).to(dev).to(torch.bfloat16)
        w = o.pack_conv_weight(torch.randn(n, c, 3, 3, generator=g).to(dev) * 0.03)
        out = torch.empty(B, hw_, hw_, n, device=dev, dtype=torch.bfloat16)
        for halo in (-1, 1):
            for pair in ((-1, 1) if n == 256 else (-1,)):
                try:
                    t = timeit(lambda: o.conv3x3(x, w, out, halo=halo, cta_pair=pair, block_n=256 if pair == 1 else 0))
                except Exception as e:  # noqa
                    print(name, halo, pair, "ERR", str(e)[:80])
                    continue
                print(f"{name:24s} halo={halo:2d} pair={pair:2d}: {t:8.1f} us  ({B * hw_ * hw_ * n * c * 18 / t / 1e6:.0f} TFLOP/s)")


if __name__ == "__main__":
    main()
