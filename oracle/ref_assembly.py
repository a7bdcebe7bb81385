"""ctypes loader for oracle/_ref/libreference_assembly.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own src/TransportEquation.cxx and src/WaveSystem.cxx, unmodified, compiled from
/root/reference against the stand-in for the SOLVERLAB mesh classes in oracle/solverlab_standin/ (see its header for what
is the reference's and what is ours) plus the wrappers of oracle/ref_assembly.cxx.  It assembles the reference's
implicit matrices -- computeDivergenceMatrix (+ the drivers' MatShift(A, 1)) -- on a mesh handed over as plain
finite-volume connectivity, so that the harness's matrix-free restatements (circulantpreconditioner_b200/krylov.py,
meshes.py) and the oracle's wave operator can be pinned against the reference's actual assembly code; only tests/ may
use it.  Built by `make -C oracle ref` where /root/reference exists; the built file travels to the GPU box.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libreference_assembly.so")
_LIB = None
NEUMANN, PERIODIC, WALL = 1, 2, 3


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
        ip, dp, i, d = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double
        L.ref_assemble.argtypes = [i, i, i, i, ip, ip, dp, dp, dp, dp, ip, ip, ip, d, dp, d, i, dp]
        L.ref_jacobian_minus.argtypes = [i, dp, d, d, dp]
        L.ref_initial_conditions.argtypes = [i, i, i, dp, dp, dp]
        _LIB = L
    return _LIB


def cartesian_mesh(shape, spacing, border=NEUMANN):
    """Finite-volume connectivity of an nx x ny x nz box grid, cells numbered i + nx (j + ny k) as everywhere in this
    repository, every cell with its six faces in the order -x, +x, -y, +y, -z, +z; border faces carry the group `border`
    and, for PERIODIC, the opposite border face as their twin."""
    nx, ny, nz = shape
    n = (nx, ny, nz)
    h = tuple(float(v) for v in spacing)
    nc = nx * ny * nz
    cid = np.arange(nc).reshape(nz, ny, nx)
    face_cells, face_measure, face_group, face_twin = [], [], [], []
    cell_faces = [[None] * 6 for _ in range(nc)]
    for d, ax in enumerate((2, 1, 0)):                   # d = 0, 1, 2 <-> x, y, z; ax = numpy axis of cid
        area = h[(d + 1) % 3] * h[(d + 2) % 3]
        lo = np.take(cid, 0, axis=ax).ravel()
        hi = np.take(cid, n[d] - 1, axis=ax).ravel()
        for a, b in zip(np.take(cid, range(0, n[d] - 1), axis=ax).ravel(), np.take(cid, range(1, n[d]), axis=ax).ravel()):
            f = len(face_cells)                          # interior face between a (its +d face) and b (its -d face)
            face_cells.append((a, b)); face_measure.append(area); face_group.append(0); face_twin.append(-1)
            cell_faces[a][2 * d + 1] = f
            cell_faces[b][2 * d] = f
        for a, b in zip(lo, hi):                         # the two border faces of every line along d
            f0 = len(face_cells)
            face_cells.append((a, -1)); face_measure.append(area); face_group.append(border); face_twin.append(f0 + 1)
            face_cells.append((b, -1)); face_measure.append(area); face_group.append(border); face_twin.append(f0)
            cell_faces[a][2 * d] = f0
            cell_faces[b][2 * d + 1] = f0 + 1
    if border != PERIODIC:
        face_twin = [-1] * len(face_twin)
    normals = np.zeros((nc, 6, 3))
    for d in range(3):
        normals[:, 2 * d, d] = -1.0
        normals[:, 2 * d + 1, d] = +1.0
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    centre = np.stack([(i.ravel() + 0.5) * h[0], (j.ravel() + 0.5) * h[1], (k.ravel() + 0.5) * h[2]], axis=1)
    return {"dim": 3, "cell_face_ptr": np.arange(0, 6 * nc + 1, 6), "cell_face_idx": np.asarray(cell_faces).ravel(),
            "cell_face_normal": normals.reshape(-1, 3), "cell_measure": np.full(nc, h[0] * h[1] * h[2]), "cell_centre": centre,
            "face_measure": np.asarray(face_measure), "face_cells": np.asarray(face_cells), "face_group": np.asarray(face_group),
            "face_twin": np.asarray(face_twin)}


def fixture_mesh(mesh):
    """The same from a tests/golden/mesh_*.npz dictionary (interior faces only: the border faces of the transport problem
    are Neumann faces, on which the reference's assembly does nothing)."""
    nc = len(mesh["volume"])
    fc, fa = mesh["face_cells"], mesh["face_area"]
    meas = np.linalg.norm(fa, axis=1)
    unit = fa / meas[:, None]
    per_cell = [[] for _ in range(nc)]
    for f, (a, b) in enumerate(fc):
        per_cell[a].append((f, unit[f]))                 # the area vector points from the first to the second cell
        per_cell[b].append((f, -unit[f]))
    ptr = np.cumsum([0] + [len(p) for p in per_cell])
    idx = np.asarray([f for p in per_cell for f, _ in p])
    nrm = np.asarray([v for p in per_cell for _, v in p])
    return {"dim": 3, "cell_face_ptr": ptr, "cell_face_idx": idx, "cell_face_normal": nrm, "cell_measure": mesh["volume"],
            "cell_centre": mesh["centre"], "face_measure": meas, "face_cells": fc, "face_group": np.zeros(len(fc), dtype=int),
            "face_twin": np.full(len(fc), -1)}


def _ip(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def assemble(kind, mesh, dt, a=(0.0, 0.0, 0.0), c0=700.0, shift=True):
    """Dense matrix of the reference's computeDivergenceMatrix ('transport' or 'wave') on `mesh`, plus the identity of
    the drivers' MatShift(A, 1) when shift."""
    k = {"transport": 0, "wave": 1}[kind]
    nc = len(mesh["cell_measure"])
    n = nc * (1 if k == 0 else mesh["dim"] + 1)
    keep = [_ip(mesh["cell_face_ptr"]), _ip(mesh["cell_face_idx"]), _dp(mesh["cell_face_normal"]), _dp(mesh["cell_measure"]),
            _dp(mesh["cell_centre"]), _dp(mesh["face_measure"]), _ip(mesh["face_cells"]), _ip(mesh["face_group"]),
            _ip(mesh["face_twin"]), _dp(a)]
    out = np.zeros((n, n), dtype=np.float64)
    rc = lib().ref_assemble(k, mesh["dim"], nc, len(mesh["face_measure"]), keep[0][1], keep[1][1], keep[2][1], keep[3][1],
                            keep[4][1], keep[5][1], keep[6][1], keep[7][1], keep[8][1], float(dt), keep[9][1], float(c0),
                            1 if shift else 0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if rc:
        raise RuntimeError(f"reference assembly failed (code {rc})")
    return out


def jacobian_minus(normal, coeff, c0):
    """jacobianMatrices(normal, coeff) of src/WaveSystem.cxx:92-107."""
    nrm = np.ascontiguousarray(normal, dtype=np.float64)
    dim = nrm.size
    out = np.zeros((dim + 1, dim + 1))
    lib().ref_jacobian_minus(dim, nrm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(coeff), float(c0),
                             out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def initial_conditions(kind, centres, lo, hi):
    """initial_conditions_shock of src/TransportEquation.cxx ('transport': temperature per cell) or src/WaveSystem.cxx
    ('wave': [pressure, velocity components] per cell) for cells with the given centres in the box lo .. hi."""
    k = {"transport": 0, "wave": 1}[kind]
    c = np.ascontiguousarray(centres, dtype=np.float64)
    nc = len(c)
    bbox = np.ascontiguousarray(list(lo) + list(hi), dtype=np.float64)
    out = np.zeros(nc * (1 if k == 0 else 4), dtype=np.float64)
    rc = lib().ref_initial_conditions(k, 3, nc, c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      bbox.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if rc:
        raise RuntimeError("reference initial condition failed")
    return out
