"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol of include/circulantpc.h,
its pure-host helpers are right, and compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
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


# ---- pure-host logic: when does a separable symbol take the recurrence form of the middle pass? ----
def _tables(n, lam, shift_y=True):
    from oracle import circulant_oracle as O
    t = [lam[a] * np.fft.fft(O.build_transport_col(n[a])) for a in range(3)]
    if shift_y:
        t[1] = t[1] + 1.0                      # the "+1" (VecShift, FftLinearSolver_3D.c:155) rides on the y table
    return [np.ascontiguousarray(v, dtype=np.complex128) for v in t]


def _recurrence(n, tabs):
    L = cpc.lib()
    lz = ctypes.c_double(-1.0)
    p = [v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for v in tabs]
    ok = L.cpc_symbol_recurrence_lambda(n[0], n[1], n[2], p[0], p[1], p[2], ctypes.byref(lz))
    return ok, lz.value


@pytest.mark.parametrize("n", [(8, 4, 16), (5, 3, 7), (16, 16, 512), (4, 4, 2), (6, 1, 1), (1, 1, 9)])
@pytest.mark.parametrize("lam", [(55.5556, 55.5556, 55.5556), (0.6, 0.15, 0.02), (1.0, 1.0, 4096.0), (2.0, 0.0, 0.0)])
def test_recurrence_gate_accepts_the_reference_symbol(n, lam):
    ok, lz = _recurrence(n, _tables(n, lam))
    assert ok == 1
    assert lz == pytest.approx(lam[2] if n[2] > 1 else 0.0, rel=1e-12, abs=1e-12)


def test_recurrence_gate_rejects_everything_else():
    n = (8, 6, 32)
    base = (2.0, 0.5, 3.0)
    assert _recurrence(n, _tables(n, base))[0] == 1
    assert _recurrence(n, _tables(n, (2.0, 0.5, -0.2)))[0] == 0                # lambda_z < 0
    assert _recurrence(n, _tables(n, (2.0, 0.5, 5000.0)))[0] == 0              # beyond the validated range
    assert _recurrence(n, _tables(n, (-0.3, 0.5, 3.0)))[0] == 0                # Re(alpha) can drop below 1/2
    assert _recurrence(n, _tables(n, (-0.2, 0.5, 3.0)))[0] == 1                # ... here it cannot (min 0.6)
    t = _tables(n, base)
    t[2][5] += 1e-9                                                            # one z entry off: not the upwind column
    assert _recurrence(n, t)[0] == 0
    t = _tables(n, base)
    t[2] = np.conj(t[2])                                                       # downwind column [1, 0, ..., -1]
    assert _recurrence(n, t)[0] == 0
    t = _tables(n, base, shift_y=False)                                        # no "+1" anywhere: alpha can vanish
    assert _recurrence(n, t)[0] == 0
    assert cpc.lib().cpc_symbol_recurrence_lambda(0, 1, 1, None, None, None, None) == -1
