"""Diagnostics (not a pytest): per-tap rel-L2 of the bf16 production path against the fp32 oracle and the
bf16-rounding oracle, next to stock torch autocast(bf16) of the same network on the same GPU.  Writes
gpurun_out/diag_taps.json — the source of the absolute ceilings in tests/test_model_gpu.py."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200.model import DPTDepthModel  # noqa
from oracle import dpt_oracle, make_golden, weights  # noqa

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
TAPS = ["layer_1", "layer_2", "tokens_8", "tokens_11", "layer_3", "layer_4", "layer_1_rn", "layer_2_rn", "layer_3_rn",
        "layer_4_rn", "path_4", "path_3", "path_2", "path_1"]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nchw(k, t):
    t = t.float().cpu()
    return t if (k.startswith("tokens") or t.dim() != 4) else t.permute(0, 3, 1, 2)


out = {}
for c in (1, 3):
    sd = weights.make_state_dict(0, c)
    model = DPTDepthModel(num_channels=c); model.load_state_dict(sd); model = model.to(dev).eval(); model.keep_taps = True
    x = torch.cat([make_golden.golden_input(1, seed=0), make_golden.golden_input(1, seed=7)])
    t32, t16, tac = {}, {}, {}
    with torch.no_grad():
        y = model(x.to(dev)).float().cpu()
        got = {k: nchw(k, v) for k, v in model.taps.items()}
        y32 = dpt_oracle.forward_fp32(sd, x, t32)
        y16 = dpt_oracle.forward_bf16(sd, x, t16)
        sdg = {k: v.to(dev) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yac = dpt_oracle.forward_fp32(sdg, x.to(dev), tac).float().cpu()
    rec = {}
    for k in TAPS + ["stem_conv", "stem_pool"]:
        if k not in got:
            continue
        rec[k] = {"vs_fp32": rel(got[k], t32[k]) if k in t32 else None, "vs_bf16_oracle": rel(got[k], t16[k]),
                  "bf16_oracle_vs_fp32": rel(t16[k], t32[k]) if k in t32 else None,
                  "autocast_vs_fp32": rel(tac[k], t32[k]) if k in tac and k in t32 else None}
    rec["output"] = {"vs_fp32": rel(y, y32), "vs_bf16_oracle": rel(y, y16), "bf16_oracle_vs_fp32": rel(y16, y32),
                     "autocast_vs_fp32": rel(yac, y32)}
    rec["head_pre_relu"] = {"bf16_oracle_vs_fp32": rel(t16["head_pre_relu"], t32["head_pre_relu"]),
                            "autocast_vs_fp32": rel(tac["head_pre_relu"], t32["head_pre_relu"])}
    out[f"c{c}"] = rec
    for k, v in rec.items():
        print(c, k, {a: (f"{b:.3e}" if b is not None else None) for a, b in v.items()})

# ---- golden-vector comparisons (what tests/test_model_gpu.py / test_model_large_gpu.py assert): sampled values
GOLDEN = ROOT / "tests" / "golden"
for c in (1, 3):
    sd = weights.make_state_dict(0, c)
    model = DPTDepthModel(num_channels=c); model.load_state_dict(sd); model = model.to(dev).eval(); model.keep_taps = True
    with torch.no_grad():
        y = model(make_golden.golden_input(1, seed=0).to(dev)).float().cpu()
    got = {k: nchw(k, v) for k, v in model.taps.items()}
    rec = torch.load(GOLDEN / f"dpt_fp32_seed0_c{c}.pt")
    g = {"output_sub8": rel(y[..., ::8, ::8], rec["output_sub8"])}
    for name, gv in rec["taps"].items():
        if name in got:
            t = got[name].reshape(-1)
            g[name] = rel(t[make_golden.sample_indices(t.numel(), name)], gv["samples"])
    out[f"c{c}"]["golden"] = g
    print(c, "golden", {k: f"{v:.3e}" for k, v in g.items()})

from omnidata_b200 import synthetic  # noqa
from omnidata_b200.model import state_dict_spec  # noqa
for backbone, fn in (("vitl16_384", "dpt_large_fp32_seed0_c1.pt"), ("vitb16_384", "dpt_vitb16_fp32_seed0_c1.pt")):
    rec = torch.load(GOLDEN / fn)
    model = DPTDepthModel(backbone=backbone)
    model.load_state_dict(synthetic.make_state_dict(0, 1, spec=state_dict_spec(1, backbone=backbone)), strict=True)
    model = model.to(dev).eval(); model.keep_taps = True
    with torch.no_grad():
        y = model(make_golden.golden_input(1, seed=0).to(dev)).float().cpu()
    got = {k: nchw(k, v) for k, v in model.taps.items()}
    g = {"output_sub8": rel(y[..., ::8, ::8], rec["output_sub8"])}
    make_golden.N_SAMPLES = 4096
    for name, gv in rec["taps"].items():
        if name in got:
            t = got[name].reshape(-1)
            g[name] = rel(t[make_golden.sample_indices(t.numel(), name)], gv["samples"])
    make_golden.N_SAMPLES = 256
    out[backbone] = {"golden": g, "reference_bf16_drift": {k: float(v) for k, v in rec["bf16_drift"].items()},
                     "reference_bf16_output_drift": float(rec["bf16_output_drift"])}
    print(backbone, "golden", {k: f"{v:.3e}" for k, v in g.items()})
    print(backbone, "reference's own bf16 drift", {k: f"{float(v):.3e}" for k, v in rec["bf16_drift"].items()})
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "diag_taps.json").write_text(json.dumps(out, indent=1))
