"""circulantpreconditioner_b200 -- B200-native (sm_100a) circulant / block-circulant preconditioner apply.

Host-side Python binding of the C ABI in ``include/circulantpc.h`` (``csrc/libcirculantpc.so``).  The product
is the CUDA library and the C++ glue under ``glue/`` that carries the reference's own function names
(``solve_3D``, ``setupFFTPrec3D``, ...); this package exists so that tests, ``bench.py`` and Python users can
drive the same entry points with torch tensors (device memory, streams, ``torch.distributed``).

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every compute call
raises ``CpcError`` when no CUDA device is present, and importing raises if the shared library was not built.
"""
from ._lib import CpcError, lib, library_path, build_library  # noqa: F401
from .plan import (CirculantPlan, nccl_unique_id, slab_range, pencil_layout, pencil_steps, pencil_group,  # noqa: F401
                   pencil_apply_lockstep)

__all__ = ["CirculantPlan", "CpcError", "lib", "library_path", "build_library", "nccl_unique_id", "slab_range",
           "pencil_layout", "pencil_steps", "pencil_group", "pencil_apply_lockstep"]
