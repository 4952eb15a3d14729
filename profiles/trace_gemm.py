"""In-kernel timeline of the tcgen05 GEMM on the four ViT-block shapes (diagnostics, not a bench).

Every CTA stamps %globaltimer at its pipeline events (odb_debug_conv_trace, include/omnidata_b200.h);
this script prints, per launch: kernel span, dependency-wait, and for the busiest CTA the per-tile
MMA-issue window, the accumulator-ready time and the epilogue window.

  python profiles/trace_gemm.py [batch]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import _capi, ops as o  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    rows = B * 577
    g = torch.Generator().manual_seed(0)

    def rnd(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).to(dev)

    x768 = rnd(rows, 768).to(torch.bfloat16)
    x3072 = rnd(rows, 3072).to(torch.bfloat16)
    res = rnd(rows, 768).to(torch.bfloat16)
    shapes = {
        "qkv  768->2304 bias": (x768, 2304, dict()),
        "proj 768->768  bias+res": (x768, 768, dict(residual=res)),
        "fc1  768->3072 bias+gelu": (x768, 3072, dict(act=o.ACT_GELU)),
        "fc2  3072->768 bias+res": (x3072, 768, dict(residual=res)),
    }
    lib = _capi.lib()
    slots = lib.odb_debug_conv_trace(None)
    trace = torch.zeros(160 * slots, dtype=torch.int64, device=dev)
    for name, (x, n, kw) in shapes.items():
        k = x.shape[1]
        w = rnd(n, k, scale=0.03).to(torch.bfloat16)
        bias = rnd(n)
        out = torch.empty(rows, n, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            o.linear(x, w, out, bias=bias, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            o.linear(x, w, out, bias=bias, **kw)
        e1.record()
        torch.cuda.synchronize()
        warm_us = e0.elapsed_time(e1) * 100
        trace.zero_()
        lib.odb_debug_conv_trace(trace.data_ptr())
        o.linear(x, w, out, bias=bias, **kw)
        torch.cuda.synchronize()
        lib.odb_debug_conv_trace(None)
        t = trace.view(160, slots).cpu()
        used = t[:, 0] > 0
        t = t[used]
        t0 = int(t[:, 0].min())
        span = (int(t[:, 2].max()) - t0) / 1e3
        flops = 2.0 * rows * n * k
        print(f"\n== {name}: warm back-to-back {warm_us:.1f} us/launch ({flops / warm_us / 1e6:.0f} TFLOP/s); "
              f"traced span {span:.1f} us over {t.shape[0]} CTAs")
        pro = (t[:, 0] - t0).float() / 1e3
        dep = (t[:, 1] - t[:, 0]).float() / 1e3
        end = (t[:, 2] - t0).float() / 1e3
        print(f"   prologue-done spread {pro.min():.2f}..{pro.max():.2f} us, dep-wait {dep.median():.2f} us, "
              f"CTA end {end.min():.1f}..{end.max():.1f} us")
        # leader CTA 0 (MMA issuer stamps live in the leader of a pair) and its epilogue
        for cta in (0, 1):
            r = t[cta]
            line = []
            for i in range(24):
                base = 8 + 5 * i
                ev = [int(v) for v in r[base:base + 5]]
                if ev[3] == 0 and ev[0] == 0:
                    break
                f = [(v - t0) / 1e3 if v else float("nan") for v in ev]
                line.append(f"     tile {i:2d}: mma start {f[0]:7.2f} first-full {f[1]:7.2f} commit {f[2]:7.2f} | "
                            f"acc ready {f[3]:7.2f} epi end {f[4]:7.2f}  (epi {f[4] - f[3]:5.2f} us)")
            print(f"   CTA {cta}:")
            print("\n".join(line))
            ck = [int(v) for v in r[104:128]]
            if ck[0]:
                base = int(r[8 + 5 * 1 + 3])   # accumulator-ready stamp of tile 1
                print("     tile 1 chunks (us after acc ready): tmem-ld done | math+pack done | smem written+fenced | "
                      "prev store read | barrier passed | store issued")
                for c in range(4):
                    if ck[6 * c]:
                        print("       chunk %d: " % c + " ".join(f"{(v - base) / 1e3:6.2f}" for v in ck[6 * c:6 * c + 6]))


def gaps(B=32):
    """Idle time between consecutive launches of one ViT block's four GEMMs (eager stream and CUDA graph)."""
    dev = torch.device("cuda:0")
    rows = B * 577
    g = torch.Generator().manual_seed(1)
    lib = _capi.lib()
    slots = lib.odb_debug_conv_trace(None)
    x = (torch.randn(rows, 768, generator=g)).to(dev).to(torch.bfloat16)
    res = torch.randn(rows, 768, generator=g).to(dev).to(torch.bfloat16)
    ws = {n: (torch.randn(n, k, generator=g) * 0.03).to(dev).to(torch.bfloat16) for n, k in ((2304, 768), (3072, 768))}
    w_proj = (torch.randn(768, 768, generator=g) * 0.03).to(dev).to(torch.bfloat16)
    w_fc2 = (torch.randn(768, 3072, generator=g) * 0.03).to(dev).to(torch.bfloat16)
    bias = {n: torch.randn(n, generator=g).to(dev) for n in (768, 2304, 3072)}
    qkv = torch.empty(rows, 2304, device=dev, dtype=torch.bfloat16)
    h1 = torch.empty(rows, 768, device=dev, dtype=torch.bfloat16)
    mlp = torch.empty(rows, 3072, device=dev, dtype=torch.bfloat16)
    h2 = torch.empty(rows, 768, device=dev, dtype=torch.bfloat16)
    traces = [torch.zeros(160 * slots, dtype=torch.int64, device=dev) for _ in range(8)]

    def block(tr):
        it = iter(tr)
        for _ in range(2):
            lib.odb_debug_conv_trace(next(it).data_ptr()); o.linear(x, ws[2304], qkv, bias=bias[2304])
            lib.odb_debug_conv_trace(next(it).data_ptr()); o.linear(x, w_proj, h1, bias=bias[768], residual=res)
            lib.odb_debug_conv_trace(next(it).data_ptr()); o.linear(h1, ws[3072], mlp, bias=bias[3072], act=o.ACT_GELU)
            lib.odb_debug_conv_trace(next(it).data_ptr()); o.linear(mlp, w_fc2, h2, bias=bias[768], residual=res)
        lib.odb_debug_conv_trace(None)

    def report(tag):
        spans = []
        for tr in traces:
            t = tr.view(160, slots).cpu()
            t = t[t[:, 0] > 0]
            spans.append((int(t[:, 0].min()), int(t[:, 1].max()), int(t[:, 2].min()), int(t[:, 2].max())))
        t00 = spans[0][0]
        print(f"\n== launch gaps, {tag}: (first prologue done, last dep-wait done, first CTA end, last CTA end) us")
        for i, sp in enumerate(spans):
            gap = (sp[0] - spans[i - 1][3]) / 1e3 if i else float("nan")
            print(f"   launch {i}: " + " ".join(f"{(v - t00) / 1e3:8.2f}" for v in sp) + f"   gap after previous end {gap:6.2f} us")
        print(f"   total {(spans[-1][3] - t00) / 1e3:.1f} us, sum of gaps {sum((spans[i][0] - spans[i - 1][3]) for i in range(1, 8)) / 1e3:.1f} us")

    for _ in range(2):
        block(traces)
    torch.cuda.synchronize()
    for tr in traces:
        tr.zero_()
    block(traces)
    torch.cuda.synchronize()
    report("eager stream")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        block(traces)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            block(traces)
    for tr in traces:
        tr.zero_()
    gr.replay()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    report("CUDA graph replay")


def attn(B=32):
    """Timeline of the softmax warps of the tcgen05 attention kernel (CTA 0, first q tiles)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(B, 577, 2304, generator=g).to(dev).to(torch.bfloat16)
    out = torch.empty(B, 577, 768, device=dev, dtype=torch.bfloat16)
    lib = _capi.lib()
    slots = lib.odb_debug_conv_trace(None)
    trace = torch.zeros(160 * slots, dtype=torch.int64, device=dev)
    for _ in range(3):
        o.attention(qkv, out, impl="tc")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o.attention(qkv, out, impl="tc")
    e1.record()
    torch.cuda.synchronize()
    lib.odb_debug_conv_trace(trace.data_ptr())
    o.attention(qkv, out, impl="tc")
    torch.cuda.synchronize()
    lib.odb_debug_conv_trace(None)
    t = trace.view(160, slots).cpu()
    print(f"\n== attention tc: warm {e0.elapsed_time(e1) * 100:.1f} us/launch")
    names = ["S0 seen", "max xchg", "P0", "P1", "P2", "P3", "P4", "sum xchg", "O ready", "stored"]
    for cta in (0, 77):
        r = t[cta]
        t0 = int(r[8])
        print(f"   CTA {cta}: " + " ".join(f"{n:>9s}" for n in names))
        for i in range(8):
            ev = [int(v) for v in r[8 + 12 * i: 8 + 12 * i + 10]]
            if ev[0] == 0:
                break
            print(f"     qt {i}: " + " ".join(f"{(v - t0) / 1e3:9.2f}" for v in ev))


if __name__ == "__main__":
    attn()
    main()
