"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (read-only, /root/reference) in the build
container, with oracle/timm_shim standing in for the un-vendored timm 0.4.12.

Not available on the GPU box (/root/reference does not exist there): callers must check
`reference_available()` and fall back to the committed fixtures in tests/golden/.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path

REFERENCE_TORCH_DIR = Path(os.environ.get("OMNIDATA_REFERENCE", "/root/reference")) / "omnidata_tools" / "torch"
SHIM_DIR = Path(__file__).resolve().parent / "timm_shim"


def reference_available() -> bool:
    return (REFERENCE_TORCH_DIR / "modules" / "midas" / "dpt_depth.py").exists()


def _prepare_path():
    for p in (str(SHIM_DIR), str(REFERENCE_TORCH_DIR)):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_reference_dpt(num_channels: int = 1, backbone: str = "vitb_rn50_384"):
    """reference DPTDepthModel(backbone=..., num_channels=...) exactly as demo.py:63,81-82."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _prepare_path()
    mod = importlib.import_module("modules.midas.dpt_depth")
    return mod.DPTDepthModel(backbone=backbone, num_channels=num_channels)


def load_reference_losses():
    """reference MidasLoss / VNL_Loss, unmodified (np.int was removed from NumPy >= 1.24)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # noqa: the reference uses np.int at losses/virtual_normal_loss.py:64,67,70
    _prepare_path()
    midas = importlib.import_module("losses.midas_loss")
    vnl = importlib.import_module("losses.virtual_normal_loss")
    return midas.MidasLoss, vnl.VNL_Loss


def load_reference_masked_losses():
    """reference masked_l1_loss / masked_cosine_angular_loss, unmodified (losses/masked_losses.py)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _prepare_path()
    m = importlib.import_module("losses.masked_losses")
    return m.masked_l1_loss, m.masked_cosine_angular_loss


def load_reference_refocus():
    """data/refocus_augmentation.py, unmodified.  The module imports matplotlib / seaborn at the top (plotting
    helpers it never calls on this path); both are absent here, so empty stand-in modules are registered first."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.lines"].__dict__.setdefault("Line2D", object)
    sys.modules["matplotlib"].__dict__.setdefault("pyplot", sys.modules["matplotlib.pyplot"])
    sys.modules["matplotlib"].__dict__.setdefault("lines", sys.modules["matplotlib.lines"])
    _prepare_path()
    mod = importlib.import_module("data.refocus_augmentation")
    import torch
    if not torch.cuda.is_available():
        # torch >= 2.5 refuses torch.nn.parallel.parallel_apply without an accelerator; it only runs the given
        # callables on threads, so a sequential map is the same computation
        mod.parallel_apply = lambda modules, args: [m(*a) for m, a in zip(modules, args)]
    return mod
