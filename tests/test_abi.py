"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/d2r_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from d2r_b200 import build
    return build.build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "d2r_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d2r_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (d2r_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing


def test_ctypes_binding_matches_header(lib_path):
    from d2r_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()
    assert _lib.lib.d2r_abi_version() == 1
    assert _lib.lib.d2r_build_arch() == b"sm_100a"
    assert _lib.lib.d2r_launch_count() == 0


def test_sass_is_blackwell_native(lib_path):
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):   # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass                   # no legacy mma.sync tensor path


def test_struct_sizes_match_c():
    """ctypes mirrors of the argument structs must have the C layout (checked with a tiny C program)."""
    from d2r_b200 import _lib
    import ctypes, tempfile
    prog = r'''
    #include <stdio.h>
    #include "d2r_b200.h"
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(d2r_gemm_args), sizeof(d2r_ptr8), sizeof(d2r_agg_args),
                       sizeof(d2r_agg_bwd_args), sizeof(d2r_saf_args), sizeof(d2r_saf_bwd_args),
                       sizeof(d2r_attn_args), sizeof(d2r_attn_bwd_args)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (_lib.GemmArgs, _lib.Ptr8, _lib.AggArgs, _lib.AggBwdArgs, _lib.SafArgs,
                                       _lib.SafBwdArgs, _lib.AttnArgs, _lib.AttnBwdArgs)]
    assert sizes == mine, (sizes, mine)
