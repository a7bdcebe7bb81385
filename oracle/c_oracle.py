"""ctypes loader for oracle/liboracle_c.so (TEST INFRASTRUCTURE ONLY; see circulant_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_c.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_dft3.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_transport_diag.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.oracle_solve_3D.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_Fft3DTransportSolver.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 7 + [dp, dp]
        L.oracle_transport_solve_z_recurrence.argtypes = [dp, dp] + [ctypes.c_int] * 3 + [ctypes.c_double] * 3
        L.oracle_transport_solve_z_recurrence.restype = ctypes.c_int
        L.oracle_transport_solve_z_line_form.argtypes = [dp, dp] + [ctypes.c_int] * 3 + [ctypes.c_double] * 3 + [ctypes.c_int,
                                                                                                           ctypes.c_double]
        L.oracle_transport_solve_z_line_form.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def dft3(v, nx, ny, nz, sign):
    a = np.ascontiguousarray(v, dtype=np.complex128).copy()
    lib().oracle_dft3(_p(a), nx, ny, nz, sign)
    return a


def transport_diag(nx, ny, nz, lx, ly, lz):
    d = np.empty(nx * ny * nz, dtype=np.complex128)
    lib().oracle_transport_diag(_p(d), nx, ny, nz, lx, ly, lz)
    return d


def solve_3D(Diag, b, nx, ny, nz):
    Diag = np.ascontiguousarray(Diag, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    lib().oracle_solve_3D(_p(x), _p(Diag), _p(b), nx, ny, nz)
    return x


def Fft3DTransportSolver(nx, ny, nz, ax, ay, az, dt, dx, dy, dz, b):
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    lib().oracle_Fft3DTransportSolver(nx, ny, nz, ax, ay, az, dt, dx, dy, dz, _p(x), _p(b))
    return x


def transport_solve_z_recurrence(nx, ny, nz, lx, ly, lz, b):
    """The transport solve with the z factor as a cyclic recurrence (what csrc/zsolve.cuh computes)."""
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    rc = lib().oracle_transport_solve_z_recurrence(_p(x), _p(b), nx, ny, nz, lx, ly, lz)
    if rc != 0:
        raise ValueError("the recurrence form needs non-negative lambdas")
    return x



def transport_solve_z_line_form(nx, ny, nz, lx, ly, lz, b, slabs=1, weight_floor=1e-17):
    """The line form / multi-rank owner scheme of the recurrence (csrc/zsolve.cuh) in plain C."""
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.empty_like(b)
    rc = lib().oracle_transport_solve_z_line_form(_p(x), _p(b), nx, ny, nz, lx, ly, lz, slabs, weight_floor)
    if rc != 0:
        raise ValueError("bad lambdas or slab count")
    return x
