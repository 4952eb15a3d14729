"""Turns the measurements of tests/diag_taps_gpu.py (gpurun_out/diag_taps.json, produced on a B200) into the
committed absolute parity ceilings tests/golden/bf16_ceilings.json = measured rel-L2 x 1.2, per model and tap.
The GPU tests assert against these fixed numbers, so a regression cannot hide inside a floating yardstick.

    gpurun -- python tests/diag_taps_gpu.py ; python tests/make_ceilings.py
"""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
MARGIN = 1.2


def main():
    m = json.loads((ROOT / "gpurun_out" / "diag_taps.json").read_text())
    out = {"_how": "rel-L2 measured on B200 by tests/diag_taps_gpu.py, times 1.2 (tests/make_ceilings.py)", "margin": MARGIN}
    for key, rec in m.items():
        e = {}
        if key in ("c1", "c3"):
            e["vs_fp32"] = {k: MARGIN * v["vs_fp32"] for k, v in rec.items() if isinstance(v, dict) and v.get("vs_fp32")}
            e["vs_bf16_oracle"] = {k: MARGIN * v["vs_bf16_oracle"] for k, v in rec.items()
                                   if isinstance(v, dict) and v.get("vs_bf16_oracle")}
            e["torch_autocast_vs_fp32"] = {k: v["autocast_vs_fp32"] for k, v in rec.items()
                                           if isinstance(v, dict) and v.get("autocast_vs_fp32")}
        e["golden"] = {k: MARGIN * v for k, v in rec["golden"].items()}
        out["hybrid_" + key if key in ("c1", "c3") else key] = e
    (ROOT / "tests" / "golden" / "bf16_ceilings.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print("wrote tests/golden/bf16_ceilings.json")


if __name__ == "__main__":
    main()
