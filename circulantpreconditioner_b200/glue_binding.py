"""ctypes binding of glue/libfftpreconditioner_b200.so: the reference-named interface (PCShell callbacks, solve_3D,
the direct-solver wrappers) over the C ABI, on the PETSc stand-in of glue/petsc_shim.h.

Used by tests/ and bench.py to drive the path a PETSc KSP would take -- PCSetUp -> setupFFTPrec3D, PCApply ->
applyFFT3DPrecTransport -> solve_3D (reference src/PCSHELLFft_3D.cxx:10-84) -- with host Vecs and with CUDA Vecs that
wrap torch tensors (VecCreateSeqCUDAWithArray / VecCreateMPICUDAWithArray).
"""
import ctypes
import os

import numpy as np

from . import _lib

_GLUE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glue")
_LIBPATH = os.path.join(_GLUE, "libfftpreconditioner_b200.so")
_L = None

_vp, _i = ctypes.c_void_p, ctypes.c_int


class _c2(ctypes.Structure):         # PetscScalar of a complex build, passed by value like std::complex<double>
    _fields_ = [("re", ctypes.c_double), ("im", ctypes.c_double)]


class GlueError(RuntimeError):
    pass


class FFTPrecTransportContext(ctypes.Structure):      # reference src/PCSHELLFft_3D.hxx:8-21
    _fields_ = [("spaceDim", _i), ("n_x", _i), ("n_y", _i), ("n_z", _i),
                ("lambda_x", _c2), ("lambda_y", _c2), ("lambda_z", _c2),
                ("FFT_MAT", _vp), ("intersectionMatrix", _vp), ("Diag", _vp), ("b_hat", _vp), ("b_cartesien", _vp)]


def lib():
    global _L
    if _L is None:
        _lib.lib()                   # libcirculantpc first (RTLD_GLOBAL), the glue links against it
        if not os.path.exists(_LIBPATH):
            raise ImportError(f"{_LIBPATH} is missing: run `make -C circulantpreconditioner_b200/glue`")
        L = ctypes.CDLL(_LIBPATH, mode=ctypes.RTLD_GLOBAL)
        L.ShimLastError.restype = ctypes.c_char_p
        L.VecCreateSeq.argtypes = [_i, _i, ctypes.POINTER(_vp)]
        L.VecCreateMPI.argtypes = [_i, _i, _i, ctypes.POINTER(_vp)]
        L.VecCreateSeqCUDAWithArray.argtypes = [_i, _i, _i, _vp, ctypes.POINTER(_vp)]
        L.VecCreateMPICUDAWithArray.argtypes = [_i, _i, _i, _i, _vp, ctypes.POINTER(_vp)]
        L.VecDestroy.argtypes = [ctypes.POINTER(_vp)]
        L.VecGetArray.argtypes = [_vp, ctypes.POINTER(_vp)]
        L.VecRestoreArray.argtypes = [_vp, ctypes.POINTER(_vp)]
        L.VecGetLocalSize.argtypes = [_vp, ctypes.POINTER(_i)]
        L.PCCreate.argtypes = [_i, ctypes.POINTER(_vp)]
        L.PCSetUp.argtypes = [_vp]
        L.PCApply.argtypes = [_vp, _vp, _vp]
        L.PCDestroy.argtypes = [ctypes.POINTER(_vp)]
        L.PCShellFFT3DAttach.argtypes = [_vp, ctypes.POINTER(FFTPrecTransportContext)]
        L.getFFTPrec3DContextCreate.argtypes = [_i, _c2, _i, _c2, _c2, _c2, _c2, _c2, _c2, _c2, _c2, _c2,
                                                ctypes.POINTER(ctypes.POINTER(FFTPrecTransportContext))]
        L.FFTPrec3DContextFree.argtypes = [ctypes.POINTER(ctypes.POINTER(FFTPrecTransportContext))]
        L.CPCMatGetPlan.argtypes = [_vp, ctypes.POINTER(_vp)]
        L.ShimWorldSet.argtypes = [_i, _i, _vp]
        L.ShimSetDefaultVecCUDA.argtypes = [_i]
        L.MatCreateSeqAIJFromCSR.argtypes = [_i, _i, _vp, _vp, _vp, ctypes.POINTER(_vp)]
        L.MatDestroy.argtypes = [ctypes.POINTER(_vp)]
        _L = L
    return _L


def check(ierr):
    if ierr != 0:
        raise GlueError(f"PetscErrorCode {ierr}: {lib().ShimLastError().decode(errors='replace')}")


def world_set(size, rank, nccl_id=None):
    """What PETSC_COMM_WORLD stands for in this process (one process per GPU)."""
    buf = ctypes.create_string_buffer(nccl_id, len(nccl_id)) if nccl_id else None
    check(lib().ShimWorldSet(size, rank, ctypes.cast(buf, _vp) if buf else None))


def set_default_vec_cuda(on):
    """The stand-in for -vec_type cuda: MatCreateVecs of an FFT Mat (Diag, b_hat, b_cartesien) makes CUDA Vecs."""
    check(lib().ShimSetDefaultVecCUDA(1 if on else 0))


class Vec:
    def __init__(self, handle, keep=None):
        self.h = handle
        self._keep = keep            # the torch tensor a CUDA Vec wraps

    @classmethod
    def create_host(cls, n, N=None):
        h = _vp()
        if N is None:
            check(lib().VecCreateSeq(1, n, ctypes.byref(h)))
        else:
            check(lib().VecCreateMPI(0, n, N, ctypes.byref(h)))
        return cls(h)

    @classmethod
    def from_device_tensor(cls, t, N=None):
        """A CUDA Vec over the memory of a complex128 CUDA tensor (no copy); N = global size for a z-slab of an MPI Vec."""
        import torch
        assert t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()
        h = _vp()
        if N is None:
            check(lib().VecCreateSeqCUDAWithArray(1, 1, t.numel(), _vp(t.data_ptr()), ctypes.byref(h)))
        else:
            check(lib().VecCreateMPICUDAWithArray(0, 1, t.numel(), N, _vp(t.data_ptr()), ctypes.byref(h)))
        return cls(h, keep=t)

    def local_size(self):
        n = _i()
        check(lib().VecGetLocalSize(self.h, ctypes.byref(n)))
        return n.value

    def numpy(self):
        """The host array (VecGetArray: the host mirror of a CUDA Vec, brought up to date) as a complex128 view."""
        p = _vp()
        check(lib().VecGetArray(self.h, ctypes.byref(p)))
        a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_double)), shape=(2 * self.local_size(),))
        check(lib().VecRestoreArray(self.h, ctypes.byref(p)))
        return a.view(np.complex128)

    def destroy(self):
        if self.h:
            lib().VecDestroy(ctypes.byref(self.h))
            self.h = _vp()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PCShellFFT3D:
    """PCSHELL with the reference's three callbacks attached (PCShellFFT3DAttach), set up for an n_x x n_y x n_z grid
    with the given lambdas; `apply` is PCApply."""

    def __init__(self, ndim, nx, ny, nz, lx, ly, lz, projection=None):
        L = lib()
        c = lambda v: _c2(float(v), 0.0)
        self.ctx = ctypes.POINTER(FFTPrecTransportContext)()
        # getFFTPrec3DContext's arithmetic (n = cbrt(nbCells), lambda = a dt n / L) is exercised by the C++ driver;
        # here the fields are then set exactly so that non-cubic grids and exact lambdas can be used
        check(L.getFFTPrec3DContextCreate(ndim, c(1.0), max(1, nx * ny * nz), c(0.0), c(0.0), c(0.0), c(0.0), c(0.0), c(0.0),
                                          c(1.0), c(1.0), c(1.0), ctypes.byref(self.ctx)))
        k = self.ctx.contents
        k.spaceDim, k.n_x, k.n_y, k.n_z = ndim, nx, ny, nz
        k.lambda_x, k.lambda_y, k.lambda_z = c(lx), c(ly), c(lz)
        self._proj = None
        if projection is not None:
            rows, cols, rowptr, colidx, val = projection
            rp = np.ascontiguousarray(rowptr, dtype=np.int32)
            ci = np.ascontiguousarray(colidx, dtype=np.int32)
            va = np.ascontiguousarray(val, dtype=np.complex128)
            m = _vp()
            check(L.MatCreateSeqAIJFromCSR(rows, cols, _vp(rp.ctypes.data), _vp(ci.ctypes.data), _vp(va.ctypes.data),
                                           ctypes.byref(m)))
            self._proj = m
            k.intersectionMatrix = m
        self.pc = _vp()
        check(L.PCCreate(0, ctypes.byref(self.pc)))
        check(L.PCShellFFT3DAttach(self.pc, self.ctx))
        check(L.PCSetUp(self.pc))

    def apply(self, b: Vec, x: Vec):
        check(lib().PCApply(self.pc, b.h, x.h))

    def _plan(self):
        p = _vp()
        check(lib().CPCMatGetPlan(self.ctx.contents.FFT_MAT, ctypes.byref(p)))
        return p

    def info(self):
        inf = _lib.PlanInfo()
        _lib.check(_lib.lib().cpc_get_info(self._plan(), ctypes.byref(inf)))
        return {f[0]: (list(getattr(inf, f[0])) if f[0] == "fast_path" else getattr(inf, f[0])) for f in inf._fields_}

    def symbol_kind(self):
        return int(self.info()["symbol_kind"])

    def fast_path(self):
        return self.info()["fast_path"]

    def diag(self) -> Vec:
        """ctx->Diag (owned by the context; do not destroy)."""
        v = Vec(_vp(self.ctx.contents.Diag))
        v.destroy = lambda: None
        return v

    def time_table_form(self, vb, vx, steps):
        """ms per PCApply with a Diag that is no longer separable (one entry changed): the N-entry table form."""
        import torch
        d = self.diag().numpy()
        old = complex(d[5])
        d[5] = old + 0.25
        self.apply(vb, vx)                       # uploads the changed Diag
        assert self.symbol_kind() == 2, "a non-separable Diag must be held as a table"
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.apply(vb, vx)
        e1.record()
        torch.cuda.synchronize()
        d = self.diag().numpy()
        d[5] = old
        self.apply(vb, vx)
        return e0.elapsed_time(e1) / steps

    def destroy(self):
        L = lib()
        if self.pc:
            check(L.PCDestroy(ctypes.byref(self.pc)))
            self.pc = _vp()
        if self.ctx:
            L.FFTPrec3DContextFree(ctypes.byref(self.ctx))
        if self._proj:
            L.MatDestroy(ctypes.byref(self._proj))
            self._proj = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()
