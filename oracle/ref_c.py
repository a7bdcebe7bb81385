"""ctypes loader for oracle/_ref/libreference_fftsolver.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own src/FftLinearSolver_3D.c, unmodified, compiled from /root/reference against the CPU
stand-in for PETSc in oracle/petsc_standin/ (see petsc_standin.h for what is the reference's and what is ours), plus the
plain-pointer wrappers of oracle/ref_entry.c.  It exists to pin the oracle against the reference's actual C code path
(column, Kronecker layout of Diag, divide, scale, the wrappers' lambdas and degenerate axes); only tests/ may use it.
Built by `make -C oracle ref` where /root/reference exists; the built file travels to the GPU box.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libreference_fftsolver.so")
_LIB = None


def available():
    if not os.path.exists(PATH) and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-s", "-C", _HERE, "ref"])
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise ImportError(f"{PATH} is missing and /root/reference is not here to build it from")
        L = ctypes.CDLL(PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        i = ctypes.c_int
        L.ref_transport_solve.argtypes = [i, i, i, i, dp, dp, dp]
        L.ref_build_diag.argtypes = [i, i, i, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp]
        L.ref_solve_3D.argtypes = [i, i, i, dp, dp, dp]
        L.ref_transport_solve_in_place.argtypes = [i, i, i, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _solve(kind, nx, ny, nz, params, b):
    b = np.ascontiguousarray(b, dtype=np.complex128)
    assert b.size == nx * ny * nz
    x = np.empty_like(b)
    p = np.ascontiguousarray(params, dtype=np.float64)
    rc = lib().ref_transport_solve(kind, nx, ny, nz, _p(p), _p(b), _p(x))
    if rc:
        raise RuntimeError(f"reference returned PetscErrorCode {rc}")
    return x


def FftTransportSolver(nx, ny, nz, lx, ly, lz, b):
    return _solve(0, nx, ny, nz, [lx, ly, lz, 0, 0, 0, 0], b)


def Fft3DTransportSolver(nx, ny, nz, ax, ay, az, dt, dx, dy, dz, b):
    return _solve(1, nx, ny, nz, [ax, ay, az, dt, dx, dy, dz], b)


def Fft2DTransportSolver(nx, ny, ax, ay, dt, dx, dy, b):
    return _solve(2, nx, ny, 1, [ax, ay, 0, dt, dx, dy, 1], b)


def Fft1DTransportSolver(nx, ax, dt, dx, b):
    return _solve(3, nx, 1, 1, [ax, 0, 0, dt, dx, 1, 1], b)


def PetscFft3DTransportSolver(nx, ny, nz, ax, ay, az, dt, dx, dy, dz, b):
    return _solve(4, nx, ny, nz, [ax, ay, az, dt, dx, dy, dz], b)


def build_diag(nx, ny, nz, lx, ly, lz):
    d = np.empty(nx * ny * nz, dtype=np.complex128)
    rc = lib().ref_build_diag(nx, ny, nz, lx, ly, lz, _p(d))
    if rc:
        raise RuntimeError(f"reference returned PetscErrorCode {rc}")
    return d


def solve_3D(diag, b, nx, ny, nz):
    diag = np.ascontiguousarray(diag, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    rc = lib().ref_solve_3D(nx, ny, nz, _p(diag), _p(b), _p(x))
    if rc:
        raise RuntimeError(f"reference returned PetscErrorCode {rc}")
    return x


def FftTransportSolver_in_place(nx, ny, nz, lx, ly, lz, u):
    u = np.array(u, dtype=np.complex128)
    rc = lib().ref_transport_solve_in_place(nx, ny, nz, lx, ly, lz, _p(u))
    if rc:
        raise RuntimeError(f"reference returned PetscErrorCode {rc}")
    return u
