"""ctypes loader for csrc/libcirculantpc.so (the C ABI declared in include/circulantpc.h)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIBPATH = os.path.join(_CSRC, "libcirculantpc.so")
_LIB = None


class CpcError(RuntimeError):
    """A libcirculantpc call returned a non-zero status."""

    def __init__(self, status, message):
        super().__init__(f"libcirculantpc status {status}: {message}")
        self.status = status


class PlanDesc(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int), ("ncomp", ctypes.c_int),
                ("dtype", ctypes.c_int), ("nranks", ctypes.c_int), ("rank", ctypes.c_int),
                ("nccl_unique_id", ctypes.c_void_p), ("stream", ctypes.c_void_p), ("device", ctypes.c_int)]


class PlanInfo(ctypes.Structure):
    _fields_ = [("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int), ("ncomp", ctypes.c_int),
                ("dtype", ctypes.c_int), ("nranks", ctypes.c_int), ("rank", ctypes.c_int),
                ("symbol_kind", ctypes.c_int), ("passes_per_apply", ctypes.c_int), ("dist_mode", ctypes.c_int),
                ("fast_path", ctypes.c_int * 3),
                ("local_elems", ctypes.c_int64), ("bytes_per_apply_alg", ctypes.c_int64),
                ("kernel_launches", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64)]


class PencilLayout(ctypes.Structure):
    _fields_ = [("r", ctypes.c_int), ("c", ctypes.c_int), ("nxl", ctypes.c_int), ("x0", ctypes.c_int),
                ("nyl", ctypes.c_int), ("y0", ctypes.c_int), ("nyl2", ctypes.c_int), ("y02", ctypes.c_int),
                ("nzl", ctypes.c_int), ("z0", ctypes.c_int), ("local_elems", ctypes.c_int64)]


class PencilStep(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("dir", ctypes.c_int), ("a", ctypes.c_int64), ("b", ctypes.c_int64),
                ("inner", ctypes.c_int64), ("scale", ctypes.c_double), ("src_buf", ctypes.c_int), ("dst_buf", ctypes.c_int),
                ("src_buf_staged", ctypes.c_int), ("dst_buf_staged", ctypes.c_int)]


# every symbol include/circulantpc.h declares: name -> (restype, argtypes)
_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_dp = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
ABI = {
    "cpc_plan_create": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(PlanDesc)]),
    "cpc_plan_create_pencil": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(PlanDesc), _i, _i]),
    "cpc_pencil_apply_lockstep": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i]),
    "cpc_destroy": (_i, [_vp]),
    "cpc_set_stream": (_i, [_vp, _vp]),
    "cpc_sync": (_i, [_vp]),
    "cpc_set_symbol_transport": (_i, [_vp, _d, _d, _d]),
    "cpc_set_symbol_separable": (_i, [_vp, _dp, _dp, _dp, _d, _d, _d]),
    "cpc_set_symbol_diag": (_i, [_vp, _vp, _i]),
    "cpc_set_symbol_first_column": (_i, [_vp, _vp, _i]),
    "cpc_set_symbol_wave": (_i, [_vp, _d, _d, _d, _d]),
    "cpc_set_option": (_i, [_vp, _i, ctypes.c_longlong]),
    "cpc_get_diag": (_i, [_vp, _vp, _i]),
    "cpc_build_diag_separable": (_i, [_i, _i, _i, _dp, _dp, _dp, _d, _d, _d, _i, _i, _vp, _i]),
    "cpc_apply": (_i, [_vp, _vp, _vp, _i]),
    "cpc_forward": (_i, [_vp, _vp, _vp, _i]),
    "cpc_inverse": (_i, [_vp, _vp, _vp, _i]),
    "cpc_set_projection": (_i, [_vp, ctypes.c_int64, _i64p, ctypes.POINTER(ctypes.c_int32), _dp]),
    "cpc_apply_projected": (_i, [_vp, _vp, _vp, _i]),
    "cpc_apply_profiled": (_i, [_vp, _vp, _vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i)]),
    "cpc_get_info": (_i, [_vp, ctypes.POINTER(PlanInfo)]),
    "cpc_last_error": (ctypes.c_char_p, []),
    "cpc_version": (_i, []),
    "cpc_device_count": (_i, []),
    "cpc_slab_range": (_i, [_i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "cpc_slab_send_chunk": (_i, [_i, _i, _i, _i, _i, _i, _i, _i64p, _i64p]),
    "cpc_slab_recv_chunk": (_i, [_i, _i, _i, _i, _i, _i, _i, _i64p, _i64p]),
    "cpc_pencil_layout": (_i, [_i, _i, _i, _i, _i, _i, ctypes.POINTER(PencilLayout)]),
    "cpc_pencil_steps": (_i, [_i, _i, _i, _i, _i, _i, ctypes.POINTER(PencilStep), _i, ctypes.POINTER(_i)]),
    "cpc_pencil_group": (_i, [_i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "cpc_pencil_swap_source": (ctypes.c_int64, [ctypes.c_int64] * 4),
    "cpc_symbol_recurrence_lambda": (_i, [_i, _i, _i, _dp, _dp, _dp, _dp]),
    "cpc_nccl_unique_id": (_i, [_vp]),
}

CPC_MAX_PASSES = 16
CPC_NCCL_UNIQUE_ID_BYTES = 128
DTYPES = {"c128": 0, "c64": 1, "f64": 2, "f32": 3}
MEM_DEVICE, MEM_HOST = 0, 1
PSTEP_PASS_X, PSTEP_PASS_Y, PSTEP_MIDDLE, PSTEP_SWAP, PSTEP_A2A_ROW, PSTEP_A2A_COL = range(6)
OPTIONS = {"z_recurrence": 1, "l2_chunk_bytes": 2, "chain_streams": 3, "z_line_form": 4}


def library_path():
    return _LIBPATH


def build_library(jobs=8, verbose=False):
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", _CSRC, f"-j{jobs}"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return _LIBPATH


def lib():
    """The loaded library.  Raises (loudly) if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise ImportError(
                f"{_LIBPATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C circulantpreconditioner_b200/csrc` (there is no CPU fallback)")
        L = ctypes.CDLL(_LIBPATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)     # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(status):
    if status != 0:
        raise CpcError(status, lib().cpc_last_error().decode(errors="replace"))
