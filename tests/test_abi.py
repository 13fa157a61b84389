"""The C-ABI library loads and exports every symbol include/gpx.h declares (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest


def test_library_built_and_exports_every_declared_symbol(gpx):
    assert os.path.exists(gpx.LIB_PATH), "libgpx.so missing: run __graft_entry__.build()"
    declared = gpx.declared_symbols()
    assert len(declared) >= 30
    L = C.CDLL(gpx.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in gpx.h but not exported: {missing}"


def test_binding_covers_header(gpx):
    L = gpx.lib()
    for s in gpx.declared_symbols():
        assert getattr(L, s).argtypes is not None, f"{s} has no ctypes signature"


def test_struct_layouts_match_header(gpx, orc):
    # sizes follow from the field lists in include/gpx.h (all 4-byte fields + one trailing u64)
    assert C.sizeof(gpx.BodyDesc) == 128
    assert C.sizeof(orc.BodyDesc) == 128
    assert C.sizeof(gpx.WorldConfig) == 44
    assert C.sizeof(gpx.Transform) == 28
    assert gpx.RAY_DTYPE.itemsize == 32 and gpx.HIT_DTYPE.itemsize == 16
    assert gpx.STATS_DTYPE.itemsize == 32


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gpx.h"\nint main(void){gpx_world_config c; (void)c; return sizeof(gpx_body_desc)==128?0:1;}\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_no_cpu_fallback_without_device(gpx):
    """On a box without a GPU the product must refuse to run rather than compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gpx.GpxError):
        gpx.World(worlds=1, max_bodies=8)


def test_product_does_not_reference_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "c-game-engine_b200")
    pat = re.compile(r"oracle/|liborc|\borc_|import orc")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert not pat.search(txt), f"{f} references the oracle"
