"""Diagnostics (not a pytest): actual error levels of each kernel against fp64, next to the error of
ideal bf16 rounding, and a fine-grained tap walk of the model against the bf16-rounding oracle."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import ops as o  # noqa
from omnidata_b200.model import DPTDepthModel  # noqa

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def report(name, out, ref64):
    ideal = ref64.to(torch.bfloat16)
    print(f"{name:42s} err_vs_fp64 {rel(out, ref64):.3e}  ideal_rounding {rel(ideal, ref64):.3e}  "
          f"mismatch_vs_ideal {rel(out, ideal):.3e}  frac_equal {float((out == ideal).float().mean()):.5f}")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def kernels():
    for (m, k, n) in [(4096, 64, 64), (4096, 768, 768), (4096, 3072, 768), (18464, 768, 2304)]:
        x = rnd(m, k).to(torch.bfloat16); w = rnd(n, k, scale=k ** -0.5).to(torch.bfloat16)
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        o.linear(x, w, out)
        report(f"linear {m}x{k}x{n}", out, x.double() @ w.double().t())
    x = rnd(2, 96, 96, 64).to(torch.bfloat16); w = rnd(64, 64, 3, 3, scale=(576) ** -0.5).to(torch.bfloat16)
    out = torch.empty(2, 96, 96, 64, device=dev, dtype=torch.bfloat16)
    o.conv3x3(x, o.pack_conv_weight(w), out)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), padding=1).permute(0, 2, 3, 1)
    report("conv3x3 96x96x64", out, ref)
    for c, hw in [(64, 9216), (256, 9216), (1024, 576)]:
        xx = (rnd(2, hw, c) * 2 + 0.3).to(torch.bfloat16)
        g, bt = rnd(c) * 0.1 + 1, rnd(c) * 0.1
        st = torch.empty(2, 32, 2, device=dev); o.groupnorm_stats(xx, st)
        out = torch.empty_like(xx); o.groupnorm_apply(xx, st, g, bt, out, relu=True)
        ref = F.relu(F.group_norm(xx.double().transpose(1, 2), 32, g.double(), bt.double(), 1e-5)).transpose(1, 2)
        report(f"groupnorm c{c}", out, ref)
    xx = (rnd(2308, 768) * 2 + 0.3).to(torch.bfloat16); g, bt = rnd(768) * 0.1 + 1, rnd(768) * 0.1
    out = torch.empty_like(xx); o.layernorm(xx, g, bt, out)
    report("layernorm", out, F.layer_norm(xx.double(), (768,), g.double(), bt.double(), 1e-6))
    qkv = rnd(2, 577, 2304); qkv[..., :1536] *= 2; qkv = qkv.to(torch.bfloat16)
    out = torch.empty(2, 577, 768, device=dev, dtype=torch.bfloat16); o.attention(qkv, out)
    q, k, v = qkv.double().view(2, 577, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(2, 577, 768)
    report("attention", out, ref)
    z = rnd(2, 48, 48, 256).to(torch.bfloat16); res = rnd(2, 96, 96, 256, seed=3).to(torch.bfloat16)
    out = torch.empty_like(res); o.upsample2x_add(z, out, res=res)
    ref = F.interpolate(z.double().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True).permute(0, 2, 3, 1) + res.double()
    report("upsample2x_add", out, ref)


def model_walk():
    from oracle import dpt_oracle, make_golden, weights
    sd = weights.make_state_dict(0, 1)
    model = DPTDepthModel(); model.load_state_dict(sd); model = model.to(dev).eval(); model.keep_taps = True
    x = torch.cat([make_golden.golden_input(1, seed=0), make_golden.golden_input(1, seed=7)])
    t16 = {}
    with torch.no_grad():
        y16 = dpt_oracle.forward_bf16(sd, x, t16)
        y = model(x.to(dev))
        taps2 = {k: v.float().cpu() for k, v in model.taps.items()}
        y_again = model(x.to(dev))
        taps_again = {k: v.float().cpu() for k, v in model.taps.items()}
        y1 = model(x[1:].to(dev))
        taps1 = {k: v.float().cpu() for k, v in model.taps.items()}
    torch.cuda.synchronize()

    def nchw(k, t):
        return t if k.startswith("tokens") else t.permute(0, 3, 1, 2)
    print(f"{'tap':16s} {'vs bf16-oracle':>15s} {'run-to-run':>12s} {'B=1 vs B=2[1]':>14s}")
    for k in t16:
        if k not in taps2:
            continue
        a = nchw(k, taps2[k])
        print(f"{k:16s} {rel(a, t16[k]):15.3e} {rel(taps_again[k], taps2[k]):12.3e} "
              f"{rel(taps1[k], taps2[k][1:]):14.3e}")
    print("output vs bf16-oracle", rel(y.float().cpu(), y16), "run-to-run", rel(y_again, y), "B1 vs B2", rel(y1, y[1:]))


if __name__ == "__main__":
    kernels()
    model_walk()
