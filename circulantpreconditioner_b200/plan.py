"""CirculantPlan: thin object wrapper over the cpc_* C ABI (include/circulantpc.h).

Mirrors the life cycle of the reference's ``FFTPrecTransportContext`` (src/PCSHELLFft_3D.hxx:8-21):
create (== ``setupFFTPrec3D``), set the eigenvalues once, ``apply`` per Krylov iteration
(== ``applyFFT3DPrecTransport`` -> ``solve_3D``), ``destroy`` (== ``destroyFFTPrec3D``).
Arrays may be torch tensors (CUDA, or CPU -- ideally pinned) or numpy arrays (host).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import CpcError, MEM_DEVICE, MEM_HOST, check, lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(_lib.CPC_NCCL_UNIQUE_ID_BYTES)
    check(lib().cpc_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return buf.raw


def slab_range(n, nranks, rank):
    s, c = ctypes.c_int(), ctypes.c_int()
    check(lib().cpc_slab_range(n, nranks, rank, ctypes.byref(s), ctypes.byref(c)))
    return s.value, c.value


def pencil_layout(nx, ny, nz, p_rows, p_cols, rank):
    """What rank (r, c) = (rank % p_rows, rank / p_rows) of a pencil grid holds (include/circulantpc.h,
    cpc_pencil_layout_t) as a dict."""
    out = _lib.PencilLayout()
    check(lib().cpc_pencil_layout(nx, ny, nz, p_rows, p_cols, rank, ctypes.byref(out)))
    return {f[0]: getattr(out, f[0]) for f in out._fields_}


def pencil_steps(nx, ny, nz, p_rows, p_cols, rank):
    """The steps of one apply on a pencil grid, the list the GPU plan executes (cpc_pencil_steps) as dicts."""
    n = ctypes.c_int()
    check(lib().cpc_pencil_steps(nx, ny, nz, p_rows, p_cols, rank, None, 0, ctypes.byref(n)))
    arr = (_lib.PencilStep * n.value)()
    check(lib().cpc_pencil_steps(nx, ny, nz, p_rows, p_cols, rank, arr, n.value, ctypes.byref(n)))
    return [{f[0]: getattr(st, f[0]) for f in st._fields_} for st in arr]


def pencil_group(nx, ny, nz, p_rows, p_cols, rank, step_kind):
    """Ranks of the row (step_kind 4) / column (5) group of `rank`, in chunk order."""
    peers = (ctypes.c_int * (p_rows * p_cols))()
    n = ctypes.c_int()
    check(lib().cpc_pencil_group(nx, ny, nz, p_rows, p_cols, rank, step_kind, peers, ctypes.byref(n)))
    return [peers[i] for i in range(n.value)]


def pencil_apply_lockstep(plans, bs, xs):
    """One apply on every rank of a pencil grid whose plans all live in this process (created with pencil=(pr, pc),
    nranks = pr * pc and no nccl_id): plans[i] is rank i, bs[i] / xs[i] its local arrays (cpc_pencil_apply_lockstep)."""
    n = len(plans)
    if len(bs) != n or len(xs) != n:
        raise ValueError("one input and one output per plan")
    pb, px, keep, kinds = (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)(), [], set()
    for i, (p, b, x) in enumerate(zip(plans, bs, xs)):
        a, ka, k1 = _ptr_and_kind(b, np_dtype=p.np_dtype, count=p.local_elems, device=p.device, what="input")
        c, kc, k2 = _ptr_and_kind(x, writable=True, np_dtype=p.np_dtype, count=p.local_elems, device=p.device, what="output")
        pb[i], px[i] = a, c
        keep += [k1, k2]
        kinds |= {ka, kc}
    if len(kinds) != 1:
        raise ValueError("all inputs and outputs must live in the same memory kind")
    handles = (ctypes.c_void_p * n)(*[p._h for p in plans])
    check(lib().cpc_pencil_apply_lockstep(handles, n, pb, px, kinds.pop()))
    return xs


def _ptr_and_kind(a, writable=False, np_dtype=None, count=None, device=None, what="array"):
    """(address, mem_kind, keepalive) of a torch tensor or numpy array.

    With np_dtype / count / device given, the array's element type, element count and CUDA device are checked against
    them first: the C ABI takes raw addresses, so a mismatch would be an out-of-bounds access, not an error."""
    if torch is not None and isinstance(a, torch.Tensor):
        if not a.is_contiguous():
            raise ValueError(f"{what}: tensor must be contiguous")
        if np_dtype is not None and a.dtype != _TORCH_DTYPES[np.dtype(np_dtype).name]:
            raise ValueError(f"{what}: dtype {a.dtype} does not match the plan's {np.dtype(np_dtype).name}")
        if count is not None and a.numel() != count:
            raise ValueError(f"{what}: {a.numel()} elements, the plan expects {count}")
        if a.is_cuda and device is not None and device >= 0 and a.device.index != device:
            raise ValueError(f"{what}: tensor lives on cuda:{a.device.index}, the plan on cuda:{device}")
        return a.data_ptr(), (MEM_DEVICE if a.is_cuda else MEM_HOST), a
    if isinstance(a, np.ndarray):
        if not a.flags.c_contiguous:
            raise ValueError(f"{what}: array must be C-contiguous")
        if writable and not a.flags.writeable:
            raise ValueError(f"{what}: output array is read-only")
        if np_dtype is not None and a.dtype != np.dtype(np_dtype):
            raise ValueError(f"{what}: dtype {a.dtype} does not match the plan's {np.dtype(np_dtype).name}")
        if count is not None and a.size != count:
            raise ValueError(f"{what}: {a.size} elements, the plan expects {count}")
        return a.ctypes.data, MEM_HOST, a
    raise TypeError(f"unsupported array type {type(a)}")


_TORCH_DTYPES = ({"complex128": torch.complex128, "complex64": torch.complex64, "float64": torch.float64,
                  "float32": torch.float32} if torch is not None else {})


class CirculantPlan:
    def __init__(self, nx, ny=1, nz=1, ncomp=1, dtype="c128", stream=None, device=-1, nranks=1, rank=0,
                 nccl_id: bytes | None = None, pencil=None):
        """pencil=(p_rows, p_cols): a pencil grid instead of z-slabs (cpc_plan_create_pencil); nranks = p_rows * p_cols.
        Without nccl_id such a plan is driven together with its peers by pencil_apply_lockstep."""
        self._h = ctypes.c_void_p()
        self.nx, self.ny, self.nz, self.ncomp = int(nx), int(ny), int(nz), int(ncomp)
        self.dtype = dtype
        self.np_dtype = {"c128": np.complex128, "c64": np.complex64, "f64": np.float64, "f32": np.float32}[dtype]
        self._id_buf = ctypes.create_string_buffer(nccl_id, len(nccl_id)) if nccl_id else None
        if stream is None and torch is not None and torch.cuda.is_available():
            # torch's current stream of the device the plan will live on (a stream belongs to one device)
            stream = (torch.cuda.current_stream(int(device)) if int(device) >= 0 else torch.cuda.current_stream()).cuda_stream
        d = _lib.PlanDesc(self.nx, self.ny, self.nz, self.ncomp, _lib.DTYPES[dtype], int(nranks), int(rank),
                          ctypes.cast(self._id_buf, ctypes.c_void_p) if self._id_buf else None,
                          ctypes.c_void_p(stream or 0), int(device))
        self.pencil = tuple(int(v) for v in pencil) if pencil else None
        if self.pencil:
            check(lib().cpc_plan_create_pencil(ctypes.byref(self._h), ctypes.byref(d), *self.pencil))
        else:
            check(lib().cpc_plan_create(ctypes.byref(self._h), ctypes.byref(d)))
        inf = self.info()
        self.local_elems = int(inf["local_elems"])        # elements of b / x held by this rank
        self.local_cells = self.local_elems // self.ncomp
        self.proj_cols = 0
        self.device = int(device)
        if self.device < 0 and torch is not None and torch.cuda.is_available():
            self.device = torch.cuda.current_device()

    # -- life cycle ----------------------------------------------------------------------------
    def destroy(self):
        if self._h:
            lib().cpc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def set_stream(self, stream):
        check(lib().cpc_set_stream(self._h, ctypes.c_void_p(stream or 0)))

    def sync(self):
        check(lib().cpc_sync(self._h))

    # -- eigenvalue set-up ----------------------------------------------------------------------
    def set_symbol_transport(self, lambda_x, lambda_y=0.0, lambda_z=0.0):
        check(lib().cpc_set_symbol_transport(self._h, lambda_x, lambda_y, lambda_z))

    def set_symbol_separable(self, cx_hat, cy_hat, cz_hat, lambda_x, lambda_y, lambda_z):
        tabs = [np.ascontiguousarray(t, dtype=np.complex128) for t in (cx_hat, cy_hat, cz_hat)]
        for t, n in zip(tabs, (self.nx, self.ny, self.nz)):
            if t.size != n:
                raise ValueError("eigenvalue table has the wrong length")
        p = [t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) for t in tabs]
        check(lib().cpc_set_symbol_separable(self._h, p[0], p[1], p[2], lambda_x, lambda_y, lambda_z))

    def set_symbol_diag(self, diag):
        """Explicit eigenvalues: this rank's z-slab of Diag, complex128 whatever the plan dtype."""
        ptr, kind, keep = _ptr_and_kind(diag, np_dtype=np.complex128, count=self.local_cells, device=self.device,
                                        what="diag")
        check(lib().cpc_set_symbol_diag(self._h, ctypes.c_void_p(ptr), kind))

    def set_symbol_first_column(self, col):
        ptr, kind, keep = _ptr_and_kind(col, np_dtype=self.np_dtype, count=self.local_elems, device=self.device,
                                        what="column")
        check(lib().cpc_set_symbol_first_column(self._h, ctypes.c_void_p(ptr), kind))

    def set_option(self, name, value):
        """Schedule switches (include/circulantpc.h, enum cpc_option): 'z_recurrence', 'l2_chunk_bytes',
        'chain_streams'."""
        check(lib().cpc_set_option(self._h, _lib.OPTIONS[name], int(value)))

    def set_symbol_wave(self, c0, mu_x, mu_y, mu_z):
        check(lib().cpc_set_symbol_wave(self._h, c0, mu_x, mu_y, mu_z))

    def get_diag(self):
        out = np.empty(self.nx * self.ny * self.nz, dtype=np.complex128)
        check(lib().cpc_get_diag(self._h, ctypes.c_void_p(out.ctypes.data), MEM_HOST))
        return out

    # -- hot path -----------------------------------------------------------------------------------
    def _call(self, fn, src, dst, count=None):
        count = self.local_elems if count is None else count
        ps, ks, _k1 = _ptr_and_kind(src, np_dtype=self.np_dtype, count=count, device=self.device, what="input")
        pd, kd, _k2 = _ptr_and_kind(dst, writable=True, np_dtype=self.np_dtype, count=count, device=self.device,
                                    what="output")
        if ks != kd:
            raise ValueError("input and output must live in the same memory kind")
        check(fn(self._h, ctypes.c_void_p(ps), ctypes.c_void_p(pd), ks))
        return dst

    def apply(self, b, x=None):
        """x = (1/N) F^H( F(b) / Lambda )  (reference solve_3D, FftLinearSolver_3D.c:166-190)."""
        if x is None:
            x = b.clone() if (torch is not None and isinstance(b, torch.Tensor)) else np.empty_like(b)
        return self._call(lib().cpc_apply, b, x)

    def forward(self, v, out=None):
        if out is None:
            out = torch.empty_like(v) if (torch is not None and isinstance(v, torch.Tensor)) else np.empty_like(v)
        return self._call(lib().cpc_forward, v, out)

    def inverse(self, v, out=None):
        if out is None:
            out = torch.empty_like(v) if (torch is not None and isinstance(v, torch.Tensor)) else np.empty_like(v)
        return self._call(lib().cpc_inverse, v, out)

    def set_projection(self, cols, rowptr, colidx, val):
        """CSR projection P (N Cartesian rows x `cols` mesh cells, real weights): ctx->intersectionMatrix."""
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(colidx, dtype=np.int32)
        v = np.ascontiguousarray(val, dtype=np.float64)
        if rp.size != self.nx * self.ny * self.nz + 1 or ci.size != v.size or ci.size < rp[-1]:
            raise ValueError("malformed CSR arrays")
        check(lib().cpc_set_projection(self._h, int(cols), rp.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                       ci.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                       v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        self.proj_cols = int(cols)

    def apply_projected(self, b, x=None):
        """x = P^T solve_3D(P b)  (reference applyFFT3DPrecTransport, PCSHELLFft_3D.cxx:10-24, plus back-projection)."""
        if x is None:
            x = torch.empty_like(b) if (torch is not None and isinstance(b, torch.Tensor)) else np.empty_like(b)
        # without a projection the library reports the call-order error itself (CPC_ERR_STATE)
        return self._call(lib().cpc_apply_projected, b, x, count=self.proj_cols or (b.numel() if hasattr(b, "numel") else b.size))

    def apply_profiled(self, b, x):
        """Device-pointer apply that also returns the per-pass durations in ms (CUDA events on the plan stream)."""
        pb, kb, _ = _ptr_and_kind(b, np_dtype=self.np_dtype, count=self.local_elems, device=self.device, what="input")
        px, kx, _ = _ptr_and_kind(x, writable=True, np_dtype=self.np_dtype, count=self.local_elems, device=self.device,
                                  what="output")
        if kb != MEM_DEVICE or kx != MEM_DEVICE:
            raise ValueError("apply_profiled needs CUDA tensors")
        ms = (ctypes.c_float * _lib.CPC_MAX_PASSES)()
        n = ctypes.c_int()
        check(lib().cpc_apply_profiled(self._h, ctypes.c_void_p(pb), ctypes.c_void_p(px), ms, ctypes.byref(n)))
        return [ms[i] for i in range(n.value)]

    def info(self):
        inf = _lib.PlanInfo()
        check(lib().cpc_get_info(self._h, ctypes.byref(inf)))
        return {f[0]: (list(getattr(inf, f[0])) if f[0] == "fast_path" else getattr(inf, f[0])) for f in inf._fields_}


__all__ = ["CirculantPlan", "CpcError", "nccl_unique_id", "slab_range", "pencil_layout", "pencil_steps", "pencil_group",
           "pencil_apply_lockstep"]
