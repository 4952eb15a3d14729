"""Thin torch-tensor front end of the C-ABI kernels (pointers + sizes only cross the boundary).

Every function enqueues on torch's current CUDA stream and returns immediately.  Activations are
channels-last bf16 tensors ([B, H, W, C] or [rows, C]); strides are taken from the tensors, so
sliced / strided views (parity planes, token windows) are passed without copies.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _capi
from ._capi import ACT_GELU, ACT_NONE, ACT_RELU, DTYPE_BF16, DTYPE_F32, ConvGemmDesc, View, check, lib

__all__ = [
    "ACT_NONE", "ACT_RELU", "ACT_GELU", "conv_gemm", "linear", "conv1x1", "conv3x3", "conv3x3_s2",
    "layernorm", "attention", "groupnorm_stats", "groupnorm_apply", "stem_gn_relu_maxpool",
    "stem_im2col", "patchify", "upsample2x_add", "write_cls_row", "readout_cls_bias", "pack_conv_weight",
    "cast_f32_bf16", "head_tail_f32",
]

_DTYPES = {torch.bfloat16: DTYPE_BF16, torch.float32: DTYPE_F32}


def _dt(t: torch.Tensor, name: str = "tensor") -> int:
    """odb_dtype of an activation tensor (bf16 in production, fp32 for the residual stream / correctness mode)."""
    if not t.is_cuda:
        raise _capi.OdbError(f"{name}: tensor must live on a CUDA device (no CPU path exists)")
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise _capi.OdbError(f"{name}: expected bfloat16 or float32 storage, got {t.dtype}") from None


def _same_device(*ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _capi.OdbError(f"tensors live on different devices ({dev} and {t.device})")
    return dev


class LaunchTimer:
    """Optional per-launch device timing (bench.py roofline pass): every C-ABI call made while the
    timer is active is bracketed by CUDA events on torch's current stream — the stream the kernel is
    launched on.  Not used on the normal path."""

    active: "Optional[LaunchTimer]" = None

    def __init__(self):
        self.records = []          # (name, info dict, start event, end event)

    def __enter__(self):
        LaunchTimer.active = self
        return self

    def __exit__(self, *exc):
        LaunchTimer.active = None

    def results(self):
        torch.cuda.synchronize()
        return [(n, i, s.elapsed_time(e)) for n, i, s, e in self.records]


def _call(name: str, info: dict, fn, dev, *args):
    """Enqueue one C-ABI call on the current stream of `dev` — the device the tensors live on, made current for the
    duration of the call (kernel attributes, SM counts and TMA descriptors are per device).  The stream is appended
    as the last argument."""
    if dev is None or dev.type != "cuda":
        raise _capi.OdbError(f"{name}: tensors must live on a CUDA device (no CPU path exists)")
    if dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _call(name, info, fn, dev, *args)
    stream = torch.cuda.current_stream(dev).cuda_stream
    t = LaunchTimer.active
    if t is None:
        check(fn(*args, stream), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(*args, stream), name)
    e.record()
    t.records.append((name, info, s, e))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _capi.OdbError(f"{name}: tensor must live on a CUDA device (no CPU path exists)")
    if t.dtype != dtype:
        raise _capi.OdbError(f"{name}: expected {dtype}, got {t.dtype}")


def _view4(t: torch.Tensor, name: str, dtype=torch.bfloat16) -> View:
    """[B,H,W,C] (or [rows,C] -> B=H=1) tensor with unit channel stride -> odb_view."""
    _need(t, dtype, name)
    if t.dim() == 2:
        t = t.unsqueeze(0).unsqueeze(0)
    if t.dim() != 4 or t.stride(3) != 1:
        raise _capi.OdbError(f"{name}: need a channels-last [B,H,W,C] view with unit channel stride")
    b, h, w, c = t.shape
    return View(t.data_ptr(), c, w, h, b, t.stride(2), t.stride(1), t.stride(0))


def conv_gemm(views: Sequence[torch.Tensor], taps: Sequence[Tuple[int, int, int]], weight: torch.Tensor,
              out: Optional[torch.Tensor], *, bias: Optional[torch.Tensor] = None,
              bias_per_image: bool = False, residual: Optional[torch.Tensor] = None, act: int = ACT_NONE,
              out2: Optional[torch.Tensor] = None, tile: Optional[Tuple[int, int]] = None, block_n: int = 0,
              cta_pair: int = 0, halo: int = 0, epilogue: int = 0, gn_stats: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              gn_groups: int = 32, gn_eps: float = 1e-5, out2_act: int = ACT_RELU,
              head: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, bool]] = None,
              out_extent: Optional[Tuple[int, int, int]] = None) -> None:
    """taps: (view index, dx, dy).  head = (w[head_c,32] f32, b[head_c] f32, out[B,head_c,H,W] f32, relu).
    out2_act: activation of the out2 copy (ACT_RELU, or ACT_GELU: `out` keeps the pre-activation)."""
    d = ConvGemmDesc()
    dev = _same_device(*views, weight, out, out2, bias, residual)
    in_t = views[0].dtype
    d.in_dtype = _dt(views[0], "view0")
    d.num_views = len(views)
    for i, v in enumerate(views):
        d.views[i] = _view4(v, f"view{i}", in_t)
    d.num_taps = len(taps)
    for i, (vi, dx, dy) in enumerate(taps):
        d.tap_view[i], d.tap_dx[i], d.tap_dy[i] = vi, dx, dy
    _need(weight, in_t, "weight")
    if not weight.is_contiguous():
        raise _capi.OdbError("weight must be contiguous [n][taps*C]")
    d.weight = weight.data_ptr()
    d.n = weight.shape[0]
    out_t = in_t if out is None else out.dtype
    d.out_dtype = _DTYPES.get(out_t, -1)
    if out is not None:
        d.out = _view4(out, "out", out_t)
    if out2 is not None:
        d.out2 = _view4(out2, "out2", out_t)
        d.out2_act = out2_act
    if bias is not None:
        _need(bias, torch.float32, "bias")
        d.bias = bias.data_ptr()
        d.bias_sb = d.n if bias_per_image else 0
    if residual is not None:
        r = residual
        if r.dim() == 3:  # [rows_per_image, C] broadcast over batch is passed as [1,H,W,C] with sb=0
            r = r.unsqueeze(0)
        rv = _view4(r, "residual", out_t)
        if residual.dim() == 4 and residual.shape[0] == 1 and out is not None and out.dim() == 4 and out.shape[0] > 1:
            rv.sb = 0
        d.residual = rv
    d.act = act
    if tile is not None:
        d.tile_w, d.tile_h = tile
    d.block_n = block_n
    d.cta_pair = cta_pair
    d.halo = halo
    d.epilogue = epilogue
    if head is not None:
        hw, hb, hout, hrelu = head
        _need(hw, torch.float32, "head_w"); _need(hb, torch.float32, "head_b"); _need(hout, torch.float32, "head_out")
        d.head_w, d.head_b, d.head_out = hw.data_ptr(), hb.data_ptr(), hout.data_ptr()
        d.head_c = hw.shape[0]
        d.head_relu = 1 if hrelu else 0
        b, _, h, w = hout.shape
        d.out = View(None, 32, w, h, b, 0, 0, 0)
    rows = d.out.w * d.out.h * d.out.b
    info = {"m": rows, "n": d.n, "k": d.num_taps * d.views[0].c, "taps": d.num_taps, "w": d.out.w, "h": d.out.h,
            "f32": d.in_dtype == DTYPE_F32}
    if gn_stats is None:
        _call("odb_conv_gemm", info, lib().odb_conv_gemm, dev, C.byref(d))
        return
    # fused GroupNorm statistics: the epilogue writes per-warp partial sums, a tiny kernel reduces them
    partial, stats = gn_stats
    _need(partial, torch.float32, "gn partial"); _need(stats, torch.float32, "gn stats")
    plan = (C.c_int32 * 4)()
    check(lib().odb_conv_gemm_plan(C.byref(d), plan), "odb_conv_gemm_plan")
    part_rows = plan[0] * plan[1] * 4
    if partial.numel() < d.out.b * part_rows * gn_groups * 2:
        raise _capi.OdbError("conv_gemm: gn partial buffer too small")
    d.gn_partial = partial.data_ptr()
    d.gn_groups = gn_groups
    _call("odb_conv_gemm", info, lib().odb_conv_gemm, dev, C.byref(d))
    count = float(d.out.w) * float(d.out.h) * (d.n // gn_groups)
    _call("odb_groupnorm_finalize", {}, lib().odb_groupnorm_finalize, dev, partial.data_ptr(), stats.data_ptr(),
          d.out.b, part_rows, gn_groups, count, gn_eps)


TAPS_1 = [(0, 0, 0)]
TAPS_3X3 = [(0, kx - 1, ky - 1) for ky in range(3) for kx in range(3)]


def pack_conv_weight(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """[N, Cin, kh, kw] (any float dtype) -> `dtype` [N, kh*kw*Cin], tap-major / channel-minor."""
    n = w.shape[0]
    return w.permute(0, 2, 3, 1).reshape(n, -1).to(dtype).contiguous()


def linear(x, weight, out, **kw):
    conv_gemm([x], TAPS_1, weight, out, **kw)


def conv1x1(x, weight, out, **kw):
    conv_gemm([x], TAPS_1, weight, out, **kw)


def conv3x3(x, weight, out, **kw):
    conv_gemm([x], TAPS_3X3, weight, out, **kw)


def _parity_taps(mode: str):
    # input index = 2*o + k - pad_before ; parity plane p, plane coordinate o + d
    taps = []
    for ky in range(3):
        for kx in range(3):
            if mode == "same":      # TF-SAME for even sizes: pad (0,1)
                py, dy = (ky & 1), (1 if ky == 2 else 0)
                px, dx = (kx & 1), (1 if kx == 2 else 0)
            elif mode == "sym1":    # padding=1 both sides
                py, dy = ((ky + 1) & 1), (-1 if ky == 0 else 0)
                px, dx = ((kx + 1) & 1), (-1 if kx == 0 else 0)
            else:
                raise ValueError(mode)
            taps.append((py * 2 + px, dx, dy))
    return taps


def conv3x3_s2(x, weight, out, mode: str, **kw):
    """Stride-2 3x3 conv through four parity-plane views of x (no im2col, no copies)."""
    planes = [x[:, py::2, px::2, :] for py in range(2) for px in range(2)]
    conv_gemm(planes, _parity_taps(mode), weight, out, **kw)


def layernorm(x, gamma, beta, out, eps: float = 1e-6):
    """x bf16 or fp32 (the fp32 residual stream), out bf16 (fp32 only together with an fp32 x)."""
    _need(gamma, torch.float32, "gamma"); _need(beta, torch.float32, "beta")
    if not (x.is_contiguous() and out.is_contiguous()):
        raise _capi.OdbError("layernorm: contiguous tensors required")
    rows = x.numel() // x.shape[-1]
    _call("odb_layernorm", {"bytes": x.element_size() * x.numel() + out.element_size() * out.numel()},
          lib().odb_layernorm, _same_device(x, gamma, beta, out), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
          out.data_ptr(), rows, x.shape[-1], eps, _dt(x, "x"), _dt(out, "out"))


def attention(qkv, out, heads: int = 12, scale: float = 0.125, lse=None):
    b, n, c3 = qkv.shape
    if not (qkv.is_contiguous() and out.is_contiguous()) or c3 != 3 * heads * 64:
        raise _capi.OdbError("attention: qkv must be contiguous [B, tokens, 3*heads*64]")
    info = {"flops": 4.0 * b * heads * n * n * 64, "bytes": qkv.element_size() * (qkv.numel() + out.numel())}
    if qkv.dtype == torch.float32:
        _need(out, torch.float32, "out")
        _call("odb_attention_f32", info, lib().odb_attention_f32, _same_device(qkv, out), qkv.data_ptr(), out.data_ptr(),
              b, n, heads, scale)
        return
    _need(qkv, torch.bfloat16, "qkv"); _need(out, torch.bfloat16, "out")
    if lse is not None:
        _need(lse, torch.float32, "lse")
    _call("odb_attention", info, lib().odb_attention, _same_device(qkv, out, lse), qkv.data_ptr(), out.data_ptr(),
          _ptr(lse), b, n, heads, scale)


_GN_SCRATCH = {}


def groupnorm_scratch(device, nbytes: int) -> torch.Tensor:
    """Zero-initialised scratch shared by all GroupNorm statistics launches of one device/stream."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _GN_SCRATCH.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _GN_SCRATCH[key] = t
    return t


def groupnorm_stats(x, stats, groups: int = 32, eps: float = 1e-5, scratch: Optional[torch.Tensor] = None):
    """stats[b, g] = (mean, rstd).  Deterministic; `scratch` must stay zeroed between calls (it does)."""
    _need(stats, torch.float32, "stats")
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    need = int(lib().odb_groupnorm_scratch_bytes(b, hw, c, groups))
    if need < 0:
        raise _capi.OdbError("groupnorm_stats: unsupported shape")
    if scratch is None:
        scratch = groupnorm_scratch(x.device, need)
    _call("odb_groupnorm_stats", {"bytes": x.element_size() * x.numel()}, lib().odb_groupnorm_stats,
          _same_device(x, stats, scratch), x.data_ptr(), stats.data_ptr(), scratch.data_ptr(), scratch.numel(), b, hw, c,
          groups, eps, _dt(x, "x"))


def groupnorm_apply(x, stats, gamma, beta, out, *, relu: bool, res=None, res_stats=None, res_gamma=None,
                    res_beta=None, groups: int = 32):
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    if out.dtype != x.dtype or (res is not None and res.dtype != x.dtype):
        raise _capi.OdbError("groupnorm_apply: x, res and out must share one storage type")
    _call("odb_groupnorm_apply", {"bytes": x.element_size() * x.numel() * (2 + (res is not None))},
          lib().odb_groupnorm_apply, _same_device(x, stats, gamma, beta, res, out),
          x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(res), _ptr(res_stats),
          _ptr(res_gamma), _ptr(res_beta), out.data_ptr(), b, hw, c, groups, 1 if relu else 0, _dt(x, "x"))


def stem_gn_relu_maxpool(x, stats, gamma, beta, out, groups: int = 32):
    b, h, w, c = x.shape
    if out.dtype != x.dtype:
        raise _capi.OdbError("stem_gn_relu_maxpool: x and out must share one storage type")
    _call("odb_stem_gn_relu_maxpool", {"bytes": int(1.25 * x.element_size() * x.numel())}, lib().odb_stem_gn_relu_maxpool,
          _same_device(x, stats, gamma, beta, out), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
          out.data_ptr(), b, h, w, c, groups, _dt(x, "x"))


def stem_im2col(x, cols):
    _need(x, torch.float32, "x")
    b, ch, h, w = x.shape
    if ch != 3 or not x.is_contiguous():
        raise _capi.OdbError("stem_im2col: contiguous [B,3,H,W] fp32 input required")
    _call("odb_stem_im2col", {}, lib().odb_stem_im2col, _same_device(x, cols), x.data_ptr(), cols.data_ptr(), b, h, w,
          cols.shape[-1], _dt(cols, "cols"))


def patchify(x, cols, patch: int = 16):
    """x fp32 [B,3,H,W] -> cols [B*(H/p)*(W/p), 3*p*p] (the im2col of a stride-p, kernel-p convolution)."""
    _need(x, torch.float32, "x")
    b, c, h, w = x.shape
    if c != 3 or not x.is_contiguous() or not cols.is_contiguous() or cols.numel() != b * (h // patch) * (w // patch) * 3 * patch * patch:
        raise _capi.OdbError("patchify: x [B,3,H,W] contiguous, cols [B*gh*gw, 3*p*p]")
    _call("odb_patchify", {"bytes": x.numel() * 4 + cols.numel() * cols.element_size()}, lib().odb_patchify,
          _same_device(x, cols), x.data_ptr(), cols.data_ptr(), b, h, w, patch, _dt(cols, "cols"))


def upsample2x_add(z, out, res=None, out_relu=None):
    b, h, w, c = z.shape
    for t in (out, res, out_relu):
        if t is not None and t.dtype != z.dtype:
            raise _capi.OdbError("upsample2x_add: all tensors must share one storage type")
    n_in = b * h * w * c * z.element_size()
    info = {"bytes": n_in + 4 * n_in * (1 + (res is not None) + (out_relu is not None)), "h": h, "c": c}
    _call("odb_upsample2x_add", info, lib().odb_upsample2x_add, _same_device(z, out, res, out_relu), z.data_ptr(),
          _ptr(res), out.data_ptr(), _ptr(out_relu), b, h, w, c, _dt(z, "z"))


def write_cls_row(tokens, cls, pos0):
    b, n, c = tokens.shape
    _call("odb_write_cls_row", {}, lib().odb_write_cls_row, _same_device(tokens, cls, pos0), tokens.data_ptr(),
          cls.data_ptr(), pos0.data_ptr(), b, n, c, _dt(tokens, "tokens"))


def readout_cls_bias(w, bias, tokens, out):
    b, n, c = tokens.shape
    if w.dtype != tokens.dtype:
        raise _capi.OdbError("readout_cls_bias: w and tokens must share one storage type")
    _call("odb_readout_cls_bias", {}, lib().odb_readout_cls_bias, _same_device(w, bias, tokens, out), w.data_ptr(),
          bias.data_ptr(), tokens.data_ptr(), out.data_ptr(), b, n, c, _dt(tokens, "tokens"))


def cast_f32_bf16(src, dst):
    """dst (bf16) = round(src (fp32)); contiguous, numel a multiple of 8."""
    _need(src, torch.float32, "src"); _need(dst, torch.bfloat16, "dst")
    if not (src.is_contiguous() and dst.is_contiguous()) or src.numel() != dst.numel():
        raise _capi.OdbError("cast_f32_bf16: contiguous tensors of equal size required")
    _call("odb_cast_f32_bf16", {"bytes": 6 * src.numel()}, lib().odb_cast_f32_bf16, _same_device(src, dst),
          src.data_ptr(), dst.data_ptr(), src.numel())


def head_tail_f32(x, w, bias, out, relu: bool, pre=None):
    """fp32 correctness mode: out[b,k,y,x] = relu?(bias[k] + sum_j w[k,j] x[b,y,x,j]), x fp32 [B,H,W,32]."""
    for t_, n_ in ((x, "x"), (w, "w"), (bias, "bias"), (out, "out")):
        _need(t_, torch.float32, n_)
    b, h, wd, c = x.shape
    if c != 32 or not x.is_contiguous() or not out.is_contiguous():
        raise _capi.OdbError("head_tail_f32: x must be contiguous [B,H,W,32]")
    _call("odb_head_tail_f32", {}, lib().odb_head_tail_f32, _same_device(x, w, bias, out, pre), x.data_ptr(),
          w.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(pre), b, h, wd, w.shape[0], 1 if relu else 0)
