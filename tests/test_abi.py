"""CPU: the C-ABI library loads and exports every symbol include/unfazed_sm100.h declares, and the
ctypes / numpy mirrors have the sizes the header's structs have (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from unfazed_b200 import _lib as L
from unfazed_b200.schema import READ_HDR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "unfazed_sm100.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(L.LIB_PATH)


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"\b(unfz_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
        assert n in L.SYMBOLS, "_lib.py does not bind %s" % n
    assert sorted(L.SYMBOLS) == names


def test_abi_version_without_gpu(lib):
    lib.unfz_abi_version.restype = ctypes.c_int
    assert lib.unfz_abi_version() == L.ABI_VERSION == 2
    lib.unfz_scan_work_bytes.restype = ctypes.c_int64
    lib.unfz_scan_work_bytes.argtypes = [ctypes.c_int64]
    assert lib.unfz_scan_work_bytes(10_000_000) > 0


def test_struct_sizes_match_header(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                    'sizeof(UnfzSiteCols),sizeof(UnfzRead),sizeof(UnfzReadCols),sizeof(UnfzReadSum),sizeof(UnfzParams),'
                    'sizeof(UnfzSegIn),sizeof(UnfzDnm),sizeof(UnfzTally),sizeof(UnfzCall));return 0;}\n' % HEADER)
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [ctypes.sizeof(L.SiteCols), READ_HDR.itemsize, ctypes.sizeof(L.ReadCols), L.RSUM_DTYPE.itemsize,
            ctypes.sizeof(L.Params), L.SEG_DTYPE.itemsize, L.DNM_DTYPE.itemsize, L.TALLY_DTYPE.itemsize, L.CALL_DTYPE.itemsize]
    assert got == want


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from unfazed_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_does_not_import_the_oracle():
    """The product must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "unfazed_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
