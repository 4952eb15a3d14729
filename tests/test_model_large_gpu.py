"""DPT-Large (`backbone='vitl16_384'`) and the plain ViT-B DPT (`'vitb16_384'`) — SURVEY.md 8(f) rank 3: plain
ViT/16 encoders with the ConvTranspose reassemble — through the same kernels, against golden vectors produced by the UNMODIFIED reference class in the build
container (tests/golden/dpt_large_fp32_seed0_c1.pt, oracle/make_golden.py::make_large_golden).

Criterion: at every tap the mismatch against the reference's fp32 result is at most the COMMITTED absolute ceiling
of tests/golden/bf16_ceilings.json (measured on B200 x 1.2, tests/make_ceilings.py) — and, as a sanity anchor, below
the drift the reference module itself shows when it is run entirely in bf16 (recorded in the golden file);
run-to-run, eager-vs-graph and batch-1-vs-batch-2 results are bit-identical.  fp32 mode: 1e-5."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
CEIL = json.loads((GOLDEN / "bf16_ceilings.json").read_text())


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


CASES = {"vitl16_384": "dpt_large_fp32_seed0_c1.pt", "vitb16_384": "dpt_vitb16_fp32_seed0_c1.pt"}


@pytest.fixture(scope="module", params=sorted(CASES))
def setup(lib_built, request):
    from omnidata_b200 import synthetic
    from omnidata_b200.model import DPTDepthModel, state_dict_spec
    from oracle import make_golden
    backbone = request.param
    rec = torch.load(GOLDEN / CASES[backbone])
    spec = state_dict_spec(1, backbone=backbone)
    assert [[k, list(s)] for k, s in spec] == rec["spec"]          # the reference's own key / shape table
    model = DPTDepthModel(backbone=backbone)
    model.load_state_dict(synthetic.make_state_dict(0, 1, spec=spec), strict=True)
    model = model.cuda().eval()
    x = torch.cat([make_golden.golden_input(1, seed=0), make_golden.golden_input(1, seed=7)]).cuda()
    model.keep_taps = True
    with torch.no_grad():
        y = model(x)
    torch.cuda.synchronize()
    taps = {k: v.float().cpu() for k, v in model.taps.items()}
    model.keep_taps = False
    return model, x, y.float().cpu(), taps, rec, backbone


def test_large_against_reference_golden_vectors(setup):
    from oracle import make_golden
    model, x, y, taps, rec, backbone = setup
    ceil = CEIL[backbone]["golden"]
    assert y.shape == (2, 384, 384)
    d_out = rec["bf16_output_drift"]
    err = rel(y[0:1, ::8, ::8], rec["output_sub8"])
    assert err <= ceil["output_sub8"] and err <= d_out, f"output: {err:.3e} (ceiling {ceil['output_sub8']:.3e})"
    make_golden.N_SAMPLES = 4096
    try:
        checked = 0
        for name, g in rec["taps"].items():
            if name not in taps:
                continue                                      # head_pre_relu is fused away in the head kernel
            t = taps[name][0:1]
            if t.dim() == 4:
                t = t.permute(0, 3, 1, 2)                      # channels-last -> the reference's NCHW
            assert list(t.shape) == g["shape"], (name, t.shape, g["shape"])
            idx = make_golden.sample_indices(t.numel(), name)
            err = rel(t.reshape(-1)[idx], g["samples"])
            d = rec["bf16_drift"][name]
            assert err <= ceil[name] and err <= d, f"{name}: {err:.3e} (ceiling {ceil[name]:.3e}, reference bf16 drift {d:.3e})"
            checked += 1
        assert checked >= 12
    finally:
        make_golden.N_SAMPLES = 256


def test_large_graph_eager_batch_independence(setup):
    model, x, y, taps, rec, backbone = setup
    with torch.no_grad():
        model.use_cuda_graph = False
        e1 = model(x).clone()
        e2 = model(x).clone()
        model.use_cuda_graph = True
        g1 = model(x).clone()
        g2 = model(x).clone()
        model.use_cuda_graph = False
        single = model(x[0:1]).clone()
    assert torch.equal(e1, e2) and torch.equal(e1, g1) and torch.equal(g1, g2)
    assert torch.equal(single[0], e1[0])


def test_large_fp32_mode_matches_reference_golden(setup):
    """fp32 correctness mode of the plain-ViT DPTs: 1e-5 against the UNMODIFIED reference class's fp32 vectors."""
    from oracle import make_golden
    model, x, y, taps, rec, backbone = setup
    model.precision = "fp32"
    model.keep_taps = True
    try:
        with torch.no_grad():
            y32 = model(x[0:1]).float().cpu()
        t32 = {k: v.float().cpu() for k, v in model.taps.items()}
    finally:
        model.keep_taps = False
        model.precision = "bf16"
    assert rel(y32[:, ::8, ::8], rec["output_sub8"]) <= 1e-5
    make_golden.N_SAMPLES = 4096
    try:
        checked = 0
        for name, g in rec["taps"].items():
            if name not in t32:
                continue
            t = t32[name]
            if t.dim() == 4 and name != "head_pre_relu":
                t = t.permute(0, 3, 1, 2)
            idx = make_golden.sample_indices(t.numel(), name)
            assert rel(t.reshape(-1)[idx], g["samples"]) <= 1e-5, name
            checked += 1
        assert checked >= 12
    finally:
        make_golden.N_SAMPLES = 256
