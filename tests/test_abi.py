"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol of include/circulantpc.h,
its pure-host helpers are right, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import pytest

import circulantpreconditioner_b200 as cpc
from circulantpreconditioner_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "circulantpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    L = cpc.lib()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in circulantpc.h but not exported"
    assert set(names) == set(_lib.ABI), "python binding table and header disagree"


def test_version_and_error_string():
    L = cpc.lib()
    assert L.cpc_version() == 1
    assert isinstance(L.cpc_last_error(), bytes)


def test_slab_helpers_match_definition():
    for n, P in [(512, 8), (10, 4), (7, 3), (5, 5)]:
        tot = 0
        for r in range(P):
            s, c = cpc.slab_range(n, P, r)
            assert s == tot
            tot += c
        assert tot == n
    L = cpc.lib()
    nx, ny, nz, nc, P = 6, 8, 12, 1, 4
    for r in range(P):
        off_expect = 0
        for q in range(P):
            o, c = ctypes.c_int64(), ctypes.c_int64()
            assert L.cpc_slab_send_chunk(nx, ny, nz, nc, P, r, q, ctypes.byref(o), ctypes.byref(c)) == 0
            assert o.value == off_expect and c.value == (nz // P) * (ny // P) * nx
            off_expect += c.value
            assert L.cpc_slab_recv_chunk(nx, ny, nz, nc, P, r, q, ctypes.byref(o), ctypes.byref(c)) == 0
            assert o.value == q * (nz // P) * (ny // P) * nx
    s, c = ctypes.c_int(), ctypes.c_int()
    assert L.cpc_slab_range(8, 2, 5, ctypes.byref(s), ctypes.byref(c)) == 1       # CPC_ERR_ARG


def test_argument_errors_are_reported():
    L = cpc.lib()
    h = ctypes.c_void_p()
    bad = _lib.PlanDesc(0, 4, 4, 1, 0, 1, 0, None, None, -1)
    assert L.cpc_plan_create(ctypes.byref(h), ctypes.byref(bad)) == 1
    assert b"extents" in L.cpc_last_error()
    bad = _lib.PlanDesc(4, 4, 4, 3, 0, 1, 0, None, None, -1)
    assert L.cpc_plan_create(ctypes.byref(h), ctypes.byref(bad)) == 1
    assert L.cpc_apply(None, None, None, 0) == 1
    assert L.cpc_destroy(None) == 0


def test_no_cpu_fallback():
    if cpc.lib().cpc_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cpc.CpcError) as e:
        cpc.CirculantPlan(8, 8, 8)
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "circulantpreconditioner_b200")
    pat = re.compile(r"(^\s*(from|import)\s+oracle)|(#include\s+[\"<][^\n]*oracle)|(liboracle)", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cxx", ".hxx", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not pat.search(txt), f"{os.path.join(dp, f)} reaches into oracle/"
