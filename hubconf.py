"""torch.hub entry points with the names the reference documents (README.md:23-29,
omnidata_tools/torch/README.md:47-53).  The reference's own hubconf lives in an external repo
(alexsax/omnidata_models) that is not part of the reference tree; only names and kwargs are kept.

    model = torch.hub.load('<this repo>', 'depth_dpt_hybrid_384')            # depth, 1 channel
    model = torch.hub.load('<this repo>', 'surface_normal_dpt_hybrid_384')   # normals, 3 channels
    model = torch.hub.load('<this repo>', 'dpt_hybrid_384', pretrained=False, task='normal')

`pretrained=True` looks for the reference checkpoint names (tools/download_*_models.sh) under
./pretrained_models/ — there is no network download here.
"""
dependencies = ["torch"]

import os

_CKPT = {"depth": "omnidata_dpt_depth_v2.ckpt", "normal": "omnidata_dpt_normal_v2.ckpt"}


def _load_checkpoint(model, path):
    import torch
    ckpt = torch.load(path, map_location="cpu")
    if "state_dict" in ckpt:                       # PL checkpoint: strip 'model.' (demo.py:65-68)
        ckpt = {k[6:]: v for k, v in ckpt["state_dict"].items()}
    model.load_state_dict(ckpt)
    return model


def dpt_hybrid_384(pretrained: bool = False, task: str = "depth", weights_dir: str = "./pretrained_models", **kwargs):
    from omnidata_b200.model import DPTDepthModel
    if task not in _CKPT:
        raise ValueError("task should be one of the following: normal, depth")
    model = DPTDepthModel(backbone="vitb_rn50_384", num_channels=3 if task == "normal" else 1, **kwargs)
    if pretrained:
        path = os.path.join(weights_dir, _CKPT[task])
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found (no network: place the reference checkpoint there)")
        _load_checkpoint(model, path)
    return model


def depth_dpt_hybrid_384(pretrained: bool = True, **kwargs):
    return dpt_hybrid_384(pretrained=pretrained, task="depth", **kwargs)


def surface_normal_dpt_hybrid_384(pretrained: bool = True, **kwargs):
    return dpt_hybrid_384(pretrained=pretrained, task="normal", **kwargs)
