"""The ctypes mirror of the C-ABI descriptor structs must have the layout the C compiler gives the header: a tiny
C program (gcc, no CUDA) prints sizeof / offsetof and the test compares them with omnidata_b200/_capi.py."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

PROBE = r'''
#include <stddef.h>
#include <stdio.h>
#include "omnidata_b200.h"
#define F(s, f) printf(#s "." #f " %zu\n", offsetof(s, f));
int main(void) {
  printf("odb_view.size %zu\n", sizeof(odb_view));
  F(odb_view, ptr) F(odb_view, c) F(odb_view, w) F(odb_view, h) F(odb_view, b) F(odb_view, sx) F(odb_view, sy) F(odb_view, sb)
  printf("odb_conv_gemm_desc.size %zu\n", sizeof(odb_conv_gemm_desc));
  F(odb_conv_gemm_desc, num_views) F(odb_conv_gemm_desc, views) F(odb_conv_gemm_desc, num_taps)
  F(odb_conv_gemm_desc, tap_view) F(odb_conv_gemm_desc, tap_dx) F(odb_conv_gemm_desc, tap_dy)
  F(odb_conv_gemm_desc, weight) F(odb_conv_gemm_desc, n) F(odb_conv_gemm_desc, out) F(odb_conv_gemm_desc, out2)
  F(odb_conv_gemm_desc, bias) F(odb_conv_gemm_desc, bias_sb) F(odb_conv_gemm_desc, residual) F(odb_conv_gemm_desc, act)
  F(odb_conv_gemm_desc, tile_w) F(odb_conv_gemm_desc, tile_h) F(odb_conv_gemm_desc, block_n)
  F(odb_conv_gemm_desc, cta_pair) F(odb_conv_gemm_desc, halo) F(odb_conv_gemm_desc, head_w) F(odb_conv_gemm_desc, head_b)
  F(odb_conv_gemm_desc, head_c) F(odb_conv_gemm_desc, head_relu) F(odb_conv_gemm_desc, head_out)
  F(odb_conv_gemm_desc, gn_partial) F(odb_conv_gemm_desc, gn_groups) F(odb_conv_gemm_desc, epilogue)
  F(odb_conv_gemm_desc, in_dtype) F(odb_conv_gemm_desc, out_dtype)
  F(odb_conv_gemm_desc, out2_act)
  printf("abi %d\n", ODB_ABI_VERSION);
  return 0;
}
'''


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_ctypes_structs_match_the_header(tmp_path):
    import ctypes as C
    from omnidata_b200 import _capi
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(line.rsplit(" ", 1) for line in out.strip().splitlines())
    assert int(got.pop("abi")) == _capi.ABI_VERSION
    assert int(got.pop("odb_view.size")) == C.sizeof(_capi.View)
    assert int(got.pop("odb_conv_gemm_desc.size")) == C.sizeof(_capi.ConvGemmDesc)
    mirrors = {"odb_view": _capi.View, "odb_conv_gemm_desc": _capi.ConvGemmDesc}
    checked = 0
    for key, off in got.items():
        struct, field = key.split(".")
        assert getattr(mirrors[struct], field).offset == int(off), key
        checked += 1
    assert checked == len(_capi.View._fields_) + len(_capi.ConvGemmDesc._fields_)


def test_ctypes_signatures_have_the_headers_arity_and_scalar_kinds():
    """Every prototype of include/omnidata_b200.h against its ctypes signature: same number of parameters, and
    pointer / 32-bit / 64-bit / float / double kinds in the same positions."""
    import ctypes as C
    import re
    from omnidata_b200 import _capi
    text = (ROOT / "include" / "omnidata_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = re.findall(r"\b(?:int|int64_t|const char\*)\s+(odb_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 30
    kind_of = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int32: "i32", C.c_int: "i32", C.c_int64: "i64",
               C.c_float: "f32", C.c_double: "f64"}
    seen = set()
    for name, params in protos:
        seen.add(name)
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        res, args = _capi._SIGNATURES[name]
        assert len(args) == len(plist), (name, len(args), len(plist))
        for p, a in zip(plist, args):
            if "*" in p:
                want = "ptr"
            elif p.startswith("int64_t"):
                want = "i64"
            elif p.startswith("int32_t") or p.startswith("int "):
                want = "i32"
            elif p.startswith("float"):
                want = "f32"
            elif p.startswith("double"):
                want = "f64"
            else:
                raise AssertionError(f"{name}: unparsed parameter {p!r}")
            have = "ptr" if isinstance(a, type) and issubclass(a, C._Pointer) else kind_of[a]
            assert have == want, (name, p, a)
    assert seen == set(_capi._SIGNATURES), seen ^ set(_capi._SIGNATURES)
