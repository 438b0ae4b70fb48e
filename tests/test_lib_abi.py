"""The C-ABI library loads and exports every symbol include/adaqp_b200.h declares
(no compute calls: runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from adaqp_b200 import build, _lib
    build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "adaqp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adaqp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_bound_and_exported(lib):
    from adaqp_b200 import _lib
    declared = header_functions()
    assert len(declared) >= 20
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (adaqp_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_abi_version_and_sizes(lib):
    from adaqp_b200 import _lib
    assert lib.adaqp_abi_version() == _lib.ADAQP_ABI_VERSION
    # pure host helpers agree with the reference's sizing rule (buffer.py:181-186)
    for N, F, b in [(5, 100, 2), (8, 256, 4), (3, 602, 8), (0, 7, 2), (9, 200, 1)]:
        wpt = 8 // b
        n_round = N + (wpt - N % wpt) % wpt
        assert lib.adaqp_qsize(N, F, b) == int((b * n_round * F + 8) / 8)
        assert lib.adaqp_packed_nbytes(N, F, b) == ((N + wpt - 1) // wpt) * F
    assert lib.adaqp_qsize(4, 4, 3) == -1


def test_argument_errors_without_gpu(lib):
    # invalid bit-width is rejected before any CUDA call
    rc = lib.adaqp_pack_f32(None, None, None, 4, 4, 3, 0, 0, None, None)
    assert rc == -1
    assert b"bits" in lib.adaqp_last_error()


def test_sass_is_sm100a():
    from adaqp_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
