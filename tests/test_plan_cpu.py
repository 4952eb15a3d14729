"""Host-side tiling logic of the tensor kernel (odb_conv_gemm_plan: no GPU needed): invariants over a sweep of layer
geometries, and the argument errors of the descriptor."""
import ctypes as C

import pytest

from omnidata_b200 import _capi
from omnidata_b200._capi import ConvGemmDesc, View


def _desc(ow, oh, ob, c, n, taps3x3, head=False, **kw):
    d = ConvGemmDesc()
    d.num_views = 1
    d.views[0] = View(0x1000, c, ow, oh, ob, c, ow * c, oh * ow * c)
    if taps3x3:
        d.num_taps = 9
        for t in range(9):
            d.tap_view[t], d.tap_dx[t], d.tap_dy[t] = 0, t % 3 - 1, t // 3 - 1
    else:
        d.num_taps = 1
    d.weight = 0x2000
    d.n = n
    d.out = View(None if head else 0x3000, n, ow, oh, ob, n, ow * n, oh * ow * n)
    if head:
        d.head_w, d.head_b, d.head_out, d.head_c = 0x4000, 0x5000, 0x6000, 1
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def _plan(d):
    out = (C.c_int32 * 4)()
    rc = _capi.lib().odb_conv_gemm_plan(C.byref(d), out)
    return rc, list(out)


@pytest.mark.parametrize("ow,oh", [(96, 96), (48, 48), (24, 24), (12, 12), (192, 192), (384, 384), (100, 36), (18464, 1),
                                   (577, 1), (130, 7), (20, 36)])
@pytest.mark.parametrize("c,n,taps", [(64, 64, True), (256, 256, True), (256, 128, True), (768, 768, False),
                                      (768, 2304, False), (64, 256, False), (1024, 4096, False)])
def test_plan_invariants(ow, oh, c, n, taps):
    for ob in (1, 32):
        rc, (tx, ty, bn, flags) = _plan(_desc(ow, oh, ob, c, n, taps))
        assert rc == 0, _capi.lib().odb_last_error()
        assert tx >= 1 and ty >= 1 and bn in (64, 128, 256) and n % bn == 0
        pair, halo = flags & 1, flags & 2
        assert not halo                                        # halo tiles are opt-in except for the head tail
        if pair:
            assert bn in (128, 256)
            m_tiles = tx * ty * ob
            assert m_tiles * (n // bn) >= 2 * 148              # pairs only when the problem fills the chip
        # one 128-row MMA tile covers at most 128 output pixels
        assert tx * ty * 128 >= ow * oh


def test_plan_head_tail_uses_the_resident_weights_halo_kernel():
    rc, (tx, ty, bn, flags) = _plan(_desc(384, 384, 32, 128, 32, True, head=True))
    assert rc == 0 and bn == 32 and flags & 2 and not flags & 1
    assert tx == 3 and ty == 384                               # 128 x 1 pixel tiles (130 x 3 halos): every MMA row filled
    rc, (tx, ty, bn, flags) = _plan(_desc(384, 384, 32, 128, 32, True, head=True, halo=-1))
    assert rc == 0 and not flags & 2                           # explicit opt-out: per-tap boxes
    rc, (tx, ty, bn, flags) = _plan(_desc(96, 64, 2, 128, 32, True, head=True))
    assert rc == 0 and flags & 2 and tx == 1                   # narrow maps: whole rows


def test_plan_argument_errors():
    lib = _capi.lib()
    # (channel-count / alignment errors are raised when the tensor maps are encoded, at launch)
    for bad in (_desc(96, 96, 1, 64, 96, False),               # n not a multiple of 64
                _desc(96, 96, 1, 64, 64, False, block_n=48),
                _desc(96, 96, 1, 64, 64, True, cta_pair=1),    # pairs need block_n 256 / 128
                _desc(96, 96, 1, 64, 64, False, halo=1),       # halo needs a 3x3 conv
                _desc(96, 96, 1, 128, 64, True, head=True)):   # the head tail needs n == 32
        rc, _ = _plan(bad)
        assert rc != 0 and len(lib.odb_last_error()) > 0
