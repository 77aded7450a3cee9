"""The C-ABI library loads on a GPU-less machine and exports every symbol include/gftorf.h
declares; the ctypes mirror of its structs has the C compiler's layout.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gftorf.h")

from gftorf_b200 import _capi  # noqa: E402


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(gft_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_header_declares_what_capi_lists():
    assert header_functions() == sorted(_capi.EXPORTS)


def test_library_loads_and_exports_every_symbol():
    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_capi.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), name
    assert _capi.lib().gft_abi_version() == _capi.ABI_VERSION


def test_size_queries_do_not_need_a_gpu():
    lib = _capi.lib()
    assert lib.gft_geom_bytes(1000) >= 1000 * (80 + 4 + 4 + 8 + 24 + 4 + 8)
    assert lib.gft_img_bytes(640, 480) >= 640 * 480 * 16 + 1200 * 8
    assert lib.gft_binning_bytes(100000) >= 100000 * 12
    assert lib.gft_geom_bytes_views(1000, 2) >= 1000 * (24 + 2 * (80 + 4 + 4 + 8 + 4 + 8))
    assert lib.gft_backward_scratch_bytes_views(1000, 2) >= 2 * 1000 * 64
    assert lib.gft_backward_scratch_bytes(1000) >= 1000 * 64
    assert lib.gft_dist2_workspace_bytes(1000) > 0
    lay = _capi.GftWorkspaceLayout()
    lib.gft_workspace_layout(1000, 5000, 64, 48, C.byref(lay))
    assert lay.geom_total == lib.gft_geom_bytes(1000)
    assert lay.bin_total == lib.gft_binning_bytes(5000)
    assert lay.img_total == lib.gft_img_bytes(64, 48)


def test_ctypes_structs_match_the_c_layout(tmp_path):
    fields = {
        "GftForwardArgs": [f[0] for f in _capi.GftForwardArgs._fields_],
        "GftBackwardArgs": [f[0] for f in _capi.GftBackwardArgs._fields_],
        "GftWorkspaceLayout": [f[0] for f in _capi.GftWorkspaceLayout._fields_],
        "GftViewArgs": [f[0] for f in _capi.GftViewArgs._fields_],
        "GftForwardViewsArgs": [f[0] for f in _capi.GftForwardViewsArgs._fields_],
        "GftBackwardViewsArgs": [f[0] for f in _capi.GftBackwardViewsArgs._fields_],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gftorf.h"', "int main(void){"]
    for s, fs in fields.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fs:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for s, cls in (("GftForwardArgs", _capi.GftForwardArgs), ("GftBackwardArgs", _capi.GftBackwardArgs),
                   ("GftWorkspaceLayout", _capi.GftWorkspaceLayout), ("GftViewArgs", _capi.GftViewArgs),
                   ("GftForwardViewsArgs", _capi.GftForwardViewsArgs),
                   ("GftBackwardViewsArgs", _capi.GftBackwardViewsArgs)):
        assert int(out[s]) == C.sizeof(cls), s
        for f in fields[s]:
            assert int(out[f"{s}.{f}"]) == getattr(cls, f).offset, (s, f)


def test_product_never_imports_the_oracle():
    """The product package must not reference oracle/ (no CPU fallback, no checker on the path)."""
    pkg = os.path.join(ROOT, "gftorf_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "cpu_oracle" not in txt and "ref_driver" not in txt and \
                    "libgft_oracle" not in txt and "libgftorf_ref" not in txt, os.path.join(dirpath, fn)
