"""Throughput evidence for the widened scope rows: CUDA-event timings of the loss mix
forward + backward, the optimizer step, the demo pre/post-processing kernels and the 3-D refocus augmentation, each
with its algorithmic bytes and the fraction of the measured HBM peak where the kernel is HBM-bound.

    python profiles/widened_rows.py > profiles/r02_widened_rows.json      (on a B200)
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from omnidata_b200 import imageproc, losses, optim, refocus  # noqa

dev = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"hbm_peak_gbs": PEAK, "how": "CUDA events on the launching stream, mean of 20 repetitions after 3 warm-up calls"}
g = torch.Generator().manual_seed(0)

# ---- train-step losses, B = 16 (the per-GPU batch of configs[4]); reference modules: ~0.30 s fwd+bwd on the CPU at B = 8
B = 16
pred = torch.rand(B, 1, 384, 384, generator=g).to(dev)
gt = torch.rand(B, 1, 384, 384, generator=g).to(dev)
mf = (torch.rand(B, 1, 384, 384, generator=g) > 0.1).float().to(dev)
fn = losses.DepthStepLoss((384, 384))
np.random.seed(0)
pts = fn.vnl.select_index()
ms = timeit(lambda: fn(pred, gt, mf, full_mix=True, points=pts))
out["depth_step_loss_fwd_bwd_b16"] = {"ms": round(ms, 4), "images_per_s": round(B / ms * 1e3, 1),
                                      "what": "clamp + make_valid_mask + MidasLoss (ssi, reg) + VNL_Loss, forward AND gradient w.r.t. "
                                              "the prediction (DepthStepLoss: the train step's launch sequence, no host sync)"}
# ---- Adam + clip on the flat buffer: 123.1 M parameters, 28 B per parameter (4 arrays read, 3 written)
n = 123_146_988
flat, grad = torch.randn(n, device=dev) * 0.02, torch.randn(n, device=dev) * 1e-3
opt = optim.FlatAdam(flat, lr=1e-5)
ms = timeit(lambda: opt.step(grad, max_norm=10.0))
byt = n * (4 + 28)                       # norm pass reads the gradient once more
out["clip_plus_adam_123M"] = {"ms": round(ms, 4), "gbs": round(byt / ms / 1e6, 1), "frac_of_hbm_peak": round(byt / ms / 1e6 / PEAK, 3),
                              "bytes": byt, "what": "odb_clip_grad_norm (4 B/param) + odb_adam_step (28 B/param)"}
# ---- demo pre-processing: a 4000 x 3000 8-bit RGB photo -> fp32 [3,384,384] (bit-exact Pillow bilinear + crop + ToTensor + Normalize)
img = torch.randint(0, 256, (3000, 4000, 3), generator=g, dtype=torch.uint8).to(dev)
pre = imageproc.DevicePreprocessor("depth", device=dev)
ms = timeit(lambda: pre(img))
out["pil_resize_crop_to_tensor_4000x3000"] = {"ms": round(ms, 4), "gbs": round(img.numel() / ms / 1e6, 1),
                                              "frac_of_hbm_peak": round(img.numel() / ms / 1e6 / PEAK, 3),
                                              "what": "source image read once (36 MB), two-pass 8-bit fixed-point resample"}
d = torch.rand(32, 384, 384, generator=g).to(dev)
ms = timeit(lambda: imageproc.bicubic_resize(d, (512, 512), clamp_in=True, clamp_out=True, invert=True))
byt = d.numel() * 4 + 32 * 512 * 512 * 4
out["depth_post_bicubic_512_b32"] = {"ms": round(ms, 4), "gbs": round(byt / ms / 1e6, 1), "frac_of_hbm_peak": round(byt / ms / 1e6 / PEAK, 3)}
nrm = torch.rand(3, 384, 384, generator=g).to(dev)
ms = timeit(lambda: imageproc.to_uint8_hwc(nrm, clamp01=True))
out["normal_post_to_uint8"] = {"ms": round(ms, 4)}
# ---- 3-D refocus augmentation, B = 8 at 512 x 512, 8 quantiles
rgb = torch.rand(8, 3, 512, 512, generator=g).to(dev)
dep = torch.rand(8, 1, 512, 512, generator=g).to(dev) + 0.1
aug = refocus.RefocusImageAugmentation(8, 0.01, 0.1)
ms = timeit(lambda: aug(rgb, dep))
out["refocus_b8_512_q8"] = {"ms": round(ms, 4), "images_per_s": round(8 / ms * 1e3, 1)}
print(json.dumps(out, indent=1))
