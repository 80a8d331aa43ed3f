"""CPU: the C-ABI library loads and exports every symbol include/pst_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "pst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pstb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    so = os.path.join(ROOT, "pysnptools_b200", "libpst_b200.so")
    assert os.path.exists(so), "build first: bash pysnptools_b200/csrc/build.sh"
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    from pysnptools_b200 import _lib
    assert set(_declared()) == set(_lib.EXPORTS)
    assert _lib.lib.pstb_version() >= 100
    assert _lib.lib.pstb_packed_ld(10000) == 2512 and _lib.lib.pstb_packed_ld(1) == 16


def test_argument_errors_need_no_gpu():
    from pysnptools_b200 import _lib
    ax = _lib.Axis(None, 0, 1, 5)
    rc = _lib.lib.pstb_decode(None, 1, 100, 10, ax, ax, 0, 16, 0, 0, None)   # ld too small; fails before touching memory
    assert rc != 0 and "ld" in _lib.last_error()


def test_host_side_helpers_need_no_gpu():
    """pstb_resolve_low_term (the AUTO rule resolved once for a kernel that is split over several calls) and pstb_numa_bind (no GPU
    here: an error code and a message, nothing changed) are pure host code."""
    import os
    from pysnptools_b200 import _lib
    lib = _lib.lib
    FP16, FP8, AUTO = _lib.LOW_TERM_FP16, _lib.LOW_TERM_FP8, _lib.LOW_TERM_AUTO
    assert lib.pstb_resolve_low_term(AUTO, 500_000, 50_000, _lib.STD_UNIT) == FP8
    assert lib.pstb_resolve_low_term(AUTO, 100_000, 500_000, _lib.STD_UNIT) == FP16      # fewer SNPs than individuals
    assert lib.pstb_resolve_low_term(AUTO, 150_000, 50_000, _lib.STD_BETA) == FP16       # Beta: 4 x as many SNPs needed
    assert lib.pstb_resolve_low_term(AUTO, 250_000, 50_000, _lib.STD_BETA) == FP8
    assert lib.pstb_resolve_low_term(AUTO, 100, 10, _lib.STD_UNIT) == FP16               # too few SNPs altogether
    assert lib.pstb_resolve_low_term(FP16, 500_000, 50_000, _lib.STD_UNIT) == FP16       # explicit modes pass through
    assert lib.pstb_resolve_low_term(FP8, 10, 50_000, _lib.STD_UNIT) == FP8
    before = os.sched_getaffinity(0)
    import torch
    if not torch.cuda.is_available():
        assert lib.pstb_numa_bind(0) == -2 and "pstb_numa_bind" in _lib.last_error()
        assert os.sched_getaffinity(0) == before
