"""Unstructured-mesh harness for BASELINE config 5 (TransportEquation on the Kershaw family with the structured circulant
approximation as GMRES preconditioner): harness code around the hot path, like krylov.py.

  fixtures      tests/golden/mesh_*.npz -- finite-volume connectivity of the reference's cube meshes, generated from
                their Gmsh text files by tests/golden/make_mesh_fixtures.py (meshes/3DKershaw/*.med is HDF5: blocked)
  operator      A = I + upwind divergence, reference src/TransportEquation.cxx:75-133 (+ MatShift(A, 1),
                tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:117); border faces do nothing ("Neumann")
  dt            cfl * minRatioVolSurf / |a|, cfl = 1e3 / 3 (same driver :55, :261)
  context       n = floor(cbrt(nbCells)), lambda_d = a_d dt n / (max_d - min_d)  (getFFTPrec3DContext,
                src/PCSHELLFft_3D.cxx:101-151 with the F6 fix)
  projection    ctx->intersectionMatrix (src/PCSHELLFft_3D.hxx:17; never built by the reference, ToDo.md last item):
                here the cell-centre variant -- Cartesian cell k averages (volume-weighted) the mesh cells whose
                centre falls into it; MEDCoupling's exact intersection volumes are not available in this image.
The preconditioner is then x = P^T solve_3D(P b): `cpc_apply_projected` on the GPU, the oracle on the CPU.
"""
from __future__ import annotations

import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_fixture(name):
    f = np.load(os.path.join(GOLDEN, f"mesh_{name}.npz"))
    return {k: f[k] for k in f.files}


def reference_dt(mesh, a, cfl=1e3 / 3.0):
    dx_min = float(np.min(mesh["volume"] / mesh["surface"]))          # SOLVERLAB Mesh::minRatioVolSurf
    return cfl * dx_min / float(np.linalg.norm(a))


def transport_matrix(mesh, a, dt, ref_sign_quirk=False):
    """A = I + dt * upwind divergence as scipy CSR.  Face f between cells i and j with area vector S (from i to j):
    seen from i, un |S| = a.S; un > 0 adds dt a.S / V_i to (i, i), un < 0 puts the inflow on (i, j) -- with the
    reference's sign (src/TransportEquation.cxx:112 writes -dt S/V un, positive) when ref_sign_quirk,
    else with the consistent upwind sign (dt a.S / V_i, negative)."""
    import scipy.sparse as sp
    nc = len(mesh["volume"])
    V = mesh["volume"]
    fc, S = mesh["face_cells"], mesh["face_area"]
    flux = S @ np.asarray(a, dtype=np.float64)             # a.S seen from the first cell
    rows, cols, vals = [np.arange(nc)], [np.arange(nc)], [np.ones(nc)]
    for (me, other, f) in ((fc[:, 0], fc[:, 1], flux), (fc[:, 1], fc[:, 0], -flux)):
        out = f > 0
        rows.append(me[out]); cols.append(me[out]); vals.append(dt * f[out] / V[me[out]])
        inn = f < 0
        coef = dt * f[inn] / V[me[inn]]
        rows.append(me[inn]); cols.append(other[inn]); vals.append(-coef if ref_sign_quirk else coef)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nc, nc)).tocsr()
    A.sum_duplicates()
    return A


def prec_context(mesh, a, dt):
    """(n, lambdas) as getFFTPrec3DContext computes them for a 3-D mesh."""
    nc = len(mesh["volume"])
    n = int(math.floor(nc ** (1.0 / 3.0)))
    while (n + 1) ** 3 <= nc:
        n += 1
    lo, hi = mesh["bbox"]
    return n, tuple(float(a[d]) * dt * n / float(hi[d] - lo[d]) for d in range(3))


def cell_centre_projection(mesh, n, orthonormal=False):
    """P (n^3 x nbCells, scipy CSR) from cell centres: Cartesian cell k (x fastest, as everywhere) gathers the mesh cells
    whose centre lies in it.  Default weights: volume-weighted average (rows sum to 1).  orthonormal=True: weights
    1 / sqrt(cells in k), so that P P^T = I on the non-empty rows and P^T P is an orthogonal projector."""
    import scipy.sparse as sp
    lo, hi = mesh["bbox"]
    ijk = np.clip(np.floor((mesh["centre"] - lo) / (hi - lo) * n).astype(np.int64), 0, n - 1)
    k = ijk[:, 0] + n * (ijk[:, 1] + n * ijk[:, 2])
    nc = len(k)
    if orthonormal:
        cnt = np.bincount(k, minlength=n ** 3)
        return sp.coo_matrix((1.0 / np.sqrt(cnt[k]), (k, np.arange(nc))), shape=(n ** 3, nc)).tocsr()
    P = sp.coo_matrix((mesh["volume"], (k, np.arange(nc))), shape=(n ** 3, nc)).tocsr()
    tot = np.asarray(P.sum(axis=1)).ravel()
    scale = np.where(tot > 0, 1.0 / np.where(tot > 0, tot, 1.0), 0.0)
    return sp.diags(scale) @ P


def two_level_pc(P_apply, projected_solve):
    """M^-1 v = P^T C^-1 P v + (v - P^T P v).

    The reference's form alone, P^T solve_3D(P v) (src/PCSHELLFft_3D.cxx:17-21 plus the back-projection), has rank
    <= n^3 < nbCells (n = floor(cbrt(nbCells)), and distorted meshes leave Cartesian cells empty), so as a left
    preconditioner it makes the system singular; adding the identity on the complement of range(P^T) -- P with
    orthonormal rows -- completes it.  projected_solve(v) = P^T C^-1 P v (one cpc_apply_projected call on the GPU),
    P_apply(v) = P^T P v."""
    return lambda v: projected_solve(v) + v - P_apply(v)


def spherical_step(mesh, inside=650.0, outside=600.0, rmax=0.3):
    """src/TransportEquation.cxx:25-73: 650 inside the sphere of radius 0.3 around the domain centre, 600 outside."""
    lo, hi = mesh["bbox"]
    r = np.linalg.norm(mesh["centre"] - 0.5 * (lo + hi), axis=1)
    return np.where(r < rmax, inside, outside).astype(np.float64)


def torch_operator(A, device):
    """The scipy CSR matrix as a callable on complex torch vectors (real sparse matrix times [re, im] columns)."""
    import torch
    At = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)), torch.from_numpy(A.indices.astype(np.int64)),
                                 torch.from_numpy(A.data), size=A.shape, dtype=torch.float64).to(device)

    def apply(u):
        return torch.view_as_complex((At @ torch.view_as_real(u).contiguous()).contiguous())

    return apply
