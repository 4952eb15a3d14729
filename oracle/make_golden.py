"""TEST INFRASTRUCTURE — generates tests/golden/*.pt|json from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
What it pins:
  * state_dict_keys.json — key/shape/order of the reference DPTDepthModel.state_dict()
    (instantiated from the reference class; timm half from oracle/timm_shim);
  * dpt_fp32_seed0_c{1,3}.pt — for seeded weights (oracle/weights.py) and a seeded input, the
    reference module's forward output (8x-subsampled) and, per tap, mean / rms / 256 sampled values
    at fixed indices.  Taps come from the reference's own `pretrained.activations` hooks
    (modules/midas/vit.py:158-165) and from forward hooks on scratch.* modules.
  * losses_seed0.pt — reference MidasLoss / VNL_Loss values on seeded inputs (config 5 inputs).
  * normal_losses_seed0.pt — reference masked L1 / cosine-angular losses (train_normal.py loss pair).
The fixtures are small (< 300 KiB in total) so that they can be committed.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from . import reference_loader as rl
from . import weights

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"
N_SAMPLES = 256


def sample_indices(numel: int, name: str) -> torch.Tensor:
    seed = sum(ord(c) for c in name) * 7919 + numel
    g = torch.Generator().manual_seed(seed % (2 ** 31))
    return torch.randint(0, numel, (N_SAMPLES,), generator=g)


def summarize(name: str, t: torch.Tensor) -> dict:
    flat = t.detach().float().reshape(-1)
    idx = sample_indices(flat.numel(), name)
    return {"shape": list(t.shape), "mean": float(flat.mean()), "rms": float(flat.pow(2).mean().sqrt()),
            "samples": flat[idx].clone()}


def golden_input(batch: int, seed: int = 0, size: int = 384) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, size, size, generator=g) * 2 - 1


def run_reference(num_channels: int):
    model = rl.load_reference_dpt(num_channels).eval()
    sd = weights.make_state_dict(0, num_channels)
    model.load_state_dict(sd, strict=True)
    taps = {}

    def grab(name):
        def hook(mod, inp, out):
            taps[name] = out.detach().clone()
        return hook
    for n in (1, 2, 3, 4):
        getattr(model.scratch, f"layer{n}_rn").register_forward_hook(grab(f"layer_{n}_rn"))
        getattr(model.scratch, f"refinenet{n}").register_forward_hook(grab(f"path_{n}"))
    model.scratch.output_conv[4].register_forward_hook(grab("head_pre_relu"))
    model.pretrained.act_postprocess3.register_forward_hook(grab("layer_3"))
    model.pretrained.act_postprocess4.register_forward_hook(grab("layer_4"))
    x = golden_input(1)
    with torch.no_grad():
        y = model(x)
    acts = model.pretrained.activations
    taps["layer_1"], taps["layer_2"] = acts["1"].detach(), acts["2"].detach()
    taps["tokens_8"], taps["tokens_11"] = acts["3"].detach(), acts["4"].detach()
    return y, taps


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    ref = rl.load_reference_dpt(1)
    keys = [[k, list(v.shape)] for k, v in ref.state_dict().items()]
    (GOLDEN / "state_dict_keys.json").write_text(json.dumps(keys, indent=0))
    for c in (1, 3):
        y, taps = run_reference(c)
        # layer_3/4 hooks fire on the Sequential slices too; keep the full-module outputs only
        rec = {"output_sub8": y[..., ::8, ::8].clone(), "output_mean": float(y.mean()),
               "taps": {k: summarize(k, v) for k, v in sorted(taps.items())}}
        torch.save(rec, GOLDEN / f"dpt_fp32_seed0_c{c}.pt")
        print(f"c={c}: output mean {rec['output_mean']:.6f}, taps {sorted(taps)}")
    # ---- losses (reference modules, unmodified) on the seeded train-step tensors of loss_oracle.loss_inputs
    from . import loss_oracle
    MidasLoss, VNL_Loss = rl.load_reference_losses()
    pred, gt, mf = loss_oracle.loss_inputs(0)
    mask = loss_oracle.make_valid_mask(mf)          # train_depth.py cannot be imported (PL, kornia): restated
    total, ssi, reg = MidasLoss(alpha=0.1, scales=4, reduction="image-based")(pred, gt, mask)
    np.random.seed(0)
    vnl = VNL_Loss(1.0, 1.0, (384, 384))(pred, gt)
    torch.save({"midas_total": float(total), "midas_ssi": float(ssi), "midas_reg": float(reg), "vnl": float(vnl),
                "mask_valid_count": int(mask.sum())}, GOLDEN / "losses_seed0.pt")
    print("losses", float(total), float(ssi), float(reg), float(vnl))
    make_normal_loss_golden()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run_reference_large(dtype=torch.float32, n_samples: int = 4096, backbone: str = "vitl16_384"):
    """DPT-Large (backbone 'vitl16_384', demo.py:81) / plain ViT-B ('vitb16_384') — the UNMODIFIED reference class
    on the timm shim's plain ViTs; seeded weights over the reference's own key/shape table."""
    from omnidata_b200.synthetic import make_state_dict as gen
    model = rl.load_reference_dpt(1, backbone).eval()
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(gen(0, 1, spec=spec), strict=True)
    model = model.to(dtype)
    taps = {}

    def grab(name):
        def hook(mod, inp, out):
            taps[name] = out.detach().float().clone()
        return hook
    for n in (1, 2, 3, 4):
        getattr(model.scratch, f"layer{n}_rn").register_forward_hook(grab(f"layer_{n}_rn"))
        getattr(model.scratch, f"refinenet{n}").register_forward_hook(grab(f"path_{n}"))
        getattr(model.pretrained, f"act_postprocess{n}").register_forward_hook(grab(f"layer_{n}"))
    model.scratch.output_conv[4].register_forward_hook(grab("head_pre_relu"))
    x = golden_input(1)
    with torch.no_grad():
        y = model(x.to(dtype)).float()
    acts = model.pretrained.activations
    for n, hk in zip("1234", (5, 11, 17, 23) if backbone == "vitl16_384" else (2, 5, 8, 11)):
        taps[f"tokens_{hk}"] = acts[n].detach().float().clone()
    return y, taps, spec


def make_large_golden(n_samples: int = 4096, backbone: str = "vitl16_384", fname: str = "dpt_large_fp32_seed0_c1.pt"):
    """dpt_large_fp32_seed0_c1.pt / dpt_vitb16_fp32_seed0_c1.pt — reference forward (fp32) of the plain-ViT DPTs +
    the drift the same reference module shows when run entirely in bf16 on the CPU (the yardstick for the bf16
    kernels)."""
    global N_SAMPLES
    y, taps, spec = run_reference_large(torch.float32, backbone=backbone)
    yb, tapsb, _ = run_reference_large(torch.bfloat16, backbone=backbone)
    old = N_SAMPLES
    N_SAMPLES = n_samples
    try:
        rec = {"output_sub8": y[..., ::8, ::8].clone(), "output_mean": float(y.mean()),
               "taps": {k: summarize(k, v) for k, v in sorted(taps.items())},
               "bf16_drift": {k: rel_l2(tapsb[k], taps[k]) for k in sorted(taps)},
               "bf16_output_drift": rel_l2(yb, y),
               "spec": [[k, list(sh)] for k, sh in spec]}
    finally:
        N_SAMPLES = old
    torch.save(rec, GOLDEN / fname)
    print(backbone + ": output mean %.6f, bf16 drift (output) %.3e" % (rec["output_mean"], rec["bf16_output_drift"]))
    for k, v in rec["bf16_drift"].items():
        print(f"   {k:14s} rms {rec['taps'][k]['rms']:.4f}  pure-bf16 drift {v:.3e}")


def make_normal_loss_golden():
    """normal_losses_seed0.pt — reference masked_l1_loss / masked_cosine_angular_loss (losses/masked_losses.py,
    unmodified) as train_normal.py:247-258 calls them, on loss_oracle.normal_loss_inputs(0)."""
    from . import loss_oracle
    ref_l1, ref_cos = rl.load_reference_masked_losses()
    pred, gt, mf = loss_oracle.normal_loss_inputs(0)
    p = torch.clamp(pred, 0, 1)
    mask = loss_oracle.make_valid_mask(mf).repeat_interleave(3, 1)
    l1, cos = float(ref_l1(p.clone(), gt.clone(), mask)), float(ref_cos(p.clone(), gt.clone(), mask))
    torch.save({"l1": l1, "cos": cos}, GOLDEN / "normal_losses_seed0.pt")
    print("normal losses", l1, cos)


if __name__ == "__main__":
    main()
