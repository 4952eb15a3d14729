"""Reference entry points on the GPU box: demo.py CLI (BASELINE configs[0] plumbing), hub constructors,
two models side by side (configs[2]: depth + normal on the same images), non-384 input sizes."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# rel-L2 of the final (post-ReLU) map of the bf16 path against the fp32 oracle: fixed bound, seeds / sizes vary here
# (tests/golden/bf16_ceilings.json holds the per-tap numbers for the golden input: 3.7e-2 depth, 2.1e-2 normal)
BF16_OUTPUT_CEILING = 6e-2
ROOT = Path(__file__).resolve().parents[1]


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("task", ["normal", "depth"])
def test_demo_cli_writes_reference_outputs(lib_built, tmp_path, task):
    from PIL import Image
    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 256, size=(420, 500, 3), dtype=np.uint8))
    src = tmp_path / "test1.png"
    img.save(src)
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, str(ROOT / "demo.py"), "--task", task, "--img_path", str(src),
                        "--output_path", str(out), "--synthetic_weights", "--weights_dir", str(tmp_path / "none")],
                       capture_output=True, text=True, cwd=str(ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = Image.open(out / f"test1_{task}.png")
    rgb = Image.open(out / "test1_rgb.png")
    assert rgb.size == (512, 512)
    assert res.size == ((384, 384) if task == "normal" else (512, 512))   # demo.py:142-150
    assert np.asarray(res).std() > 0


def test_depth_and_normal_models_side_by_side(lib_built):
    """configs[2]: two complete networks on the same images; instances share no state (the reference's
    module-global `activations` dict, vit.py:158-165, would alias them)."""
    import hubconf
    from omnidata_b200 import synthetic
    from oracle import dpt_oracle, make_golden
    x = make_golden.golden_input(2, seed=11)
    models, sds = {}, {}
    for task, c in (("depth", 1), ("normal", 3)):
        m = hubconf.dpt_hybrid_384(pretrained=False, task=task)
        sds[task] = synthetic.make_state_dict(3 + c, c)
        m.load_state_dict(sds[task])
        models[task] = m.cuda().eval()
    with torch.no_grad():
        xd, xn = x.cuda(), ((x + 1) / 2).cuda()                    # depth: [-1,1]; normal: [0,1] (demo.py:74-95)
        d1 = models["depth"](xd)
        n1 = models["normal"](xn)
        d2 = models["depth"](xd)                                    # interleaved calls do not disturb each other
        n2 = models["normal"](xn)
    assert d1.shape == (2, 384, 384) and n1.shape == (2, 3, 384, 384)
    assert torch.equal(d1, d2) and torch.equal(n1, n2)
    with torch.no_grad():
        rd = dpt_oracle.forward_fp32(sds["depth"], x)
        rn = dpt_oracle.forward_fp32(sds["normal"], (x + 1) / 2)
    # absolute bounds: bf16 production path (operand rounding; stock autocast shows 2-4e-2 on this architecture) ...
    assert rel(d1.float().cpu(), rd) <= BF16_OUTPUT_CEILING
    assert rel(n1.float().cpu(), rn) <= BF16_OUTPUT_CEILING
    # ... and the fp32 correctness mode of the same two instances: the north star's 1e-5
    for m in models.values():
        m.precision = "fp32"
    with torch.no_grad():
        d3, n3 = models["depth"](xd), models["normal"](xn)
    assert rel(d3.float().cpu(), rd) <= 1e-5 and rel(n3.float().cpu(), rn) <= 1e-5


@pytest.mark.parametrize("size", [(256, 256), (320, 384)])
def test_other_input_sizes(lib_built, size):
    """forward_flex's pos-embed resize (vit.py:102-124) and the tile heuristics away from 384x384."""
    from omnidata_b200 import synthetic
    from omnidata_b200.model import DPTDepthModel
    from oracle import dpt_oracle
    sd = synthetic.make_state_dict(0, 1)
    m = DPTDepthModel()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, *size, generator=g) * 2 - 1
    with torch.no_grad():
        y = m(x.cuda()).float().cpu()
        m.precision = "fp32"
        y_fp32 = m(x.cuda()).float().cpu()
        r32 = dpt_oracle.forward_fp32(sd, x)
    assert y.shape == (2, *size)
    assert rel(y, r32) <= BF16_OUTPUT_CEILING
    assert rel(y_fp32, r32) <= 1e-5
