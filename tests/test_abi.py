"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/lpp_b200.h declares, fails loudly
without a GPU (no CPU fallback), and its host-side pieces (tridiagonal solver, continued fraction, sharding) are right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported(lpp):
    header = open(os.path.join(ROOT, "include", "lpp_b200.h")).read()
    declared = set(re.findall(r"\b(lpp_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    L = C.CDLL(lpp._lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared <= set(lpp._lib.SYMBOLS), declared - set(lpp._lib.SYMBOLS)
    assert lpp.lib().lpp_version() >= 100


@pytest.mark.skipif(_has_gpu(), reason="this container check is for GPU-less boxes")
def test_no_cpu_fallback(lpp):
    assert lpp.lib().lpp_device_check(0) == -2
    assert b"no CPU fallback" in lpp.lib().lpp_last_error()
    with pytest.raises(lpp.LppError):
        cases.make_engine(lpp, cases.SMALL_CASES["input0"])


def test_tridiag_and_cf_against_numpy_and_oracle(lpp, oracle):
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 60, 200):
        a = rng.normal(size=n)
        b = rng.uniform(0.1, 1.0, size=n)
        T = np.diag(a) + np.diag(b[:n - 1], 1) + np.diag(b[:n - 1], -1)
        w, z = lpp.tridiag_eig(a, b, vectors=True)
        wr = np.linalg.eigvalsh(T)
        assert np.abs(w - wr).max() < 1e-12 * max(1.0, np.abs(wr).max())
        assert np.abs(T @ z - z * w).max() < 1e-11
        assert np.abs(w - oracle.tridiag_eig(a, b)).max() < 1e-12
        omega = np.linspace(-3, 3, 25)
        g0 = oracle.cf_eval(a, b, -0.3, 1.7, -1, omega, 0.05)
        g1 = lpp.ContinuedFraction(a, b, -0.3, 1.7, -1)(omega, 0.05)
        assert np.abs(g0 - g1).max() < 1e-8 * max(1.0, np.abs(g0).max())


def test_shard_range(lpp):
    f, c = C.c_uint64(), C.c_uint64()
    for n, nr in ((12870, 8), (8008, 4), (7, 8), (48620, 3)):
        tot, nxt = 0, 0
        for r in range(nr):
            assert lpp.lib().lpp_shard_range(n, r, nr, C.byref(f), C.byref(c)) == 0
            assert f.value == nxt
            nxt += c.value
            tot += c.value
        assert tot == n
    assert lpp.lib().lpp_shard_range(10, 3, 2, C.byref(f), C.byref(c)) == -1


def test_solver_options_dispatch(lpp):
    assert lpp.kernel_from_solver_options("none,InternalProductCuda") == lpp.KERNEL_AUTO
    assert lpp.kernel_from_solver_options("InternalProductStored") == lpp.KERNEL_STORED
    assert lpp.kernel_from_solver_options("InternalProductCudaGeneric") == lpp.KERNEL_GENERIC


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under lanczosplusplus_b200/, host/ or include/ may import, include, link or
    load it (a product path that routes through the oracle would void every parity claim)."""
    bad = []
    for sub, exts in (("lanczosplusplus_b200", (".py", ".cu", ".cuh", ".h")), ("host", (".cpp", ".h")), ("include", (".h",))):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if not f.endswith(exts):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r"^\s*(from|import)\s+oracle\b", r"#include\s+[\"<].*oracle", r"liblpp_oracle", r"liblpp_ref",
                            r"lanczos_oracle"):
                    if re.search(pat, text, flags=re.M):
                        bad.append((os.path.join(sub, f), pat))
    assert not bad, bad
    # tools/: only the two fixture generators (their outputs are the committed tests/golden files) use the oracle; every other
    # checker script that does lives under tests/
    allowed = {"make_golden.py", "make_fullsize_pins.py"}
    for f in sorted(os.listdir(os.path.join(ROOT, "tools"))):
        if f.endswith(".py") and f not in allowed:
            text = open(os.path.join(ROOT, "tools", f), errors="replace").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
    # and the build line of the product library names only its own sources
    assert all(s.startswith("lpp_") for s in __import__("lanczosplusplus_b200")._lib.SOURCES)
