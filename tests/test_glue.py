"""The reference-named interface (glue/) over the C ABI.

CPU part: the shared library builds, loads and exports every name of the reference's two headers
(/root/reference/src/FftLinearSolver_3D.h:21-43, PCSHELLFft_3D.hxx:23-41).
GPU part: the reference's direct-solver tests restated with real assertions (glue/test_fft_solver_3d.cxx),
and the by-value-context entry point driven through ctypes on the 32^3 golden fixture.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT, rel_l2

GLUE = os.path.join(ROOT, "circulantpreconditioner_b200", "glue")
LIB = os.path.join(GLUE, "libfftpreconditioner_b200.so")

C_NAMES = ["build_transport_col", "vec_kronecker_product_identity_left", "vec_kronecker_product_identity_right",
           "build_diag_mat_vec_3D", "solve_3D", "Fft3DSolver", "FftTransportSolver",
           "Fft3DTransportSolver", "Fft2DTransportSolver", "Fft1DTransportSolver", "PetscFft3DTransportSolver",
           "PCShellFFT3DAttach", "getFFTPrec3DContextCreate"]
CXX_NAMES = ["applyFFT3DPrecTransport", "setupFFTPrec3D", "destroyFFTPrec3D", "getFFTPrec3DContext"]


def ensure_built():
    # always run make: a no-op when up to date, and it rebuilds the glue when include/circulantpc.h changed
    # (the glue embeds struct layouts of the C ABI, e.g. cpc_plan_info)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "circulantpreconditioner_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", GLUE])
    return LIB


def test_glue_exports_reference_names():
    ensure_built()
    L = ctypes.CDLL(LIB)
    for n in C_NAMES:
        assert hasattr(L, n), n
    syms = subprocess.run(["nm", "-D", "--demangle", LIB], capture_output=True, text=True).stdout
    for n in CXX_NAMES:
        assert any(line.split(" T ")[-1].startswith(n + "(") for line in syms.splitlines() if " T " in line), n


def test_forwarding_headers_carry_reference_file_names():
    for h in ("FftLinearSolver_3D.h", "PCSHELLFft_3D.hxx"):
        assert "circulantpc_petsc.h" in open(os.path.join(GLUE, h)).read()


@pytest.mark.gpu
def test_reference_direct_solver_tests_restated():
    ensure_built()
    exe = os.path.join(GLUE, "test_fft_solver_3d")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout


class _Vec(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("array", ctypes.POINTER(ctypes.c_double)), ("state", ctypes.c_ulong)]


class _Ctx(ctypes.Structure):       # StructuredTransportContext, reference FftLinearSolver_3D.h:7-19
    _fields_ = [("n_x", ctypes.c_int), ("n_y", ctypes.c_int), ("n_z", ctypes.c_int),
                ("a_x", ctypes.c_double * 2), ("a_y", ctypes.c_double * 2), ("a_z", ctypes.c_double * 2),
                ("dt", ctypes.c_double * 2), ("delta_x", ctypes.c_double * 2), ("delta_y", ctypes.c_double * 2),
                ("delta_z", ctypes.c_double * 2), ("FFT_MAT", ctypes.c_void_p)]


@pytest.mark.gpu
def test_petsc_fft3d_transport_solver_by_value_context():
    ensure_built()
    L = ctypes.CDLL(LIB)
    f = np.load(os.path.join(GOLDEN, "ref_py_3d_32cube_phys.npz"))
    n, N = 32, 32 ** 3
    vp = ctypes.POINTER(_Vec)
    L.VecCreateSeq.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    L.MatCreateFFT.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                               ctypes.POINTER(ctypes.c_void_p)]
    L.PetscFft3DTransportSolver.argtypes = [_Ctx, vp, vp]
    L.ShimLastError.restype = ctypes.c_char_p
    B, X = vp(), vp()
    assert L.VecCreateSeq(0, N, ctypes.byref(B)) == 0 and L.VecCreateSeq(0, N, ctypes.byref(X)) == 0
    mat = ctypes.c_void_p()
    dims = (ctypes.c_int * 3)(n, n, n)
    assert L.MatCreateFFT(0, 3, dims, b"fftw", ctypes.byref(mat)) == 0, L.ShimLastError()
    barr = np.ctypeslib.as_array(B.contents.array, shape=(2 * N,)).view(np.complex128)
    xarr = np.ctypeslib.as_array(X.contents.array, shape=(2 * N,)).view(np.complex128)
    barr[:] = f["b"]
    c2 = lambda v: (ctypes.c_double * 2)(v, 0.0)
    # lambda = a dt / delta = (0.6, 0.15, 0.02): the physics of testFftSolver_3D.py:82-91
    ctx = _Ctx(n, n, n, c2(6.0), c2(3.0), c2(1.0), c2(0.01), c2(0.1), c2(0.2), c2(0.5), mat)
    assert L.PetscFft3DTransportSolver(ctx, B, X) == 0, L.ShimLastError()
    assert rel_l2(xarr.real, f["X_real"]) < 1e-12
    assert rel_l2(xarr, f["X_ref"]) < 1e-12
    assert L.PetscFft3DTransportSolver(ctx, B, B) == 0          # Un, Un aliasing, second step on the same Mat
    assert rel_l2(barr.real, f["X_real"]) < 1e-12
    bad = _Ctx(n, n, 16, c2(6.0), c2(3.0), c2(1.0), c2(0.01), c2(0.1), c2(0.2), c2(0.5), mat)
    assert L.PetscFft3DTransportSolver(bad, B, X) == 62         # PETSC_ERR_ARG_WRONG
