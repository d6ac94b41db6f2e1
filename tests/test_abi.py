"""CPU-only: the C-ABI library loads and exports every symbol include/nfdpf.h declares (no compute calls)."""
import os
import re

import normalizing_flows_dpfs_b200._lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nfdpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nfdpf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "libnfdpf.so does not export " + s
        assert s in L.SIGNATURES, "no ctypes signature for " + s
    assert set(L.SIGNATURES) == set(syms)
    assert lib.nfdpf_version() == 100


def test_error_reporting_without_gpu():
    lib = L.load()
    # argument validation happens before any CUDA call
    rc = lib.nfdpf_soft_resample_fwd(None, None, None, None, 0.5, 1, 1, 2, None, None, None, None, None, None, None)
    assert rc == -1 and b"null pointer" in lib.nfdpf_last_error()
