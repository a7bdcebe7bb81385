"""PETSc-free Krylov harness for the iteration-count configs: left-preconditioned GMRES(30) with
PETSc's default convergence test, and matrix-free restatements of the reference's upwind operators on Cartesian grids.

This is harness code around the hot path (torch tensors on any device); the preconditioner it calls is either the
CUDA plan (`CirculantPlan.apply`) or, in tests, the CPU oracle.  Reference behaviour mirrored:
  KSP set-up   tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:120-126 (GMRES, rtol = atol = 1e-5, maxits 1000,
               PETSc defaults otherwise: restart 30, left preconditioning, zero initial guess, classical Gram-Schmidt)
  operators    src/TransportEquation.cxx:75-133 (+ MatShift(A,1), tests/TransportEquation_...:117)
               src/WaveSystem.cxx:92-176        (+ MatShift(A,1), tests/WaveSystem_..._impl_mpi.cxx:127)
               (pinned: tests/test_reference_assembly.py compiles those two files, unmodified, against a stand-in for the
               SOLVERLAB mesh classes and compares the assembled matrices with these operators entry by entry)
"""
from __future__ import annotations

import math

import torch


# ---------------------------------------------------------------------------------------------------------------
# z-slab distribution (one process per GPU, torch.distributed): rank r holds planes [r nz/P, (r+1) nz/P), the layout of
# the multi-rank plans.  The operators need one halo plane from each z neighbour, GMRES needs global dots and norms.
# ---------------------------------------------------------------------------------------------------------------
class Slab:
    """Halo exchange and reductions over the z-slabs.  `Slab(None)` (or world size 1) is the single-process case."""

    def __init__(self, group=None, enabled=None):
        import torch.distributed as dist
        self.dist = dist
        self.on = dist.is_available() and dist.is_initialized() if enabled is None else enabled
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.group = group
        self.on = self.on and self.world > 1

    def sum(self, t):
        """In-place global sum of a tensor of partial sums (identical on every rank afterwards)."""
        if self.on:
            if t.is_complex():
                self.dist.all_reduce(torch.view_as_real(t), group=self.group)
            else:
                self.dist.all_reduce(t, group=self.group)
        return t

    def halo(self, U):
        """(plane below this slab's first plane, plane above its last plane) from the z neighbours, cyclic in rank;
        U is [nzl, ...].  Single process: the array's own last / first plane."""
        if not self.on:
            return U[-1], U[0]
        edges = torch.stack([U[0], U[-1]]).contiguous()
        allv = [torch.empty_like(edges) for _ in range(self.world)]
        self.dist.all_gather(allv, edges, group=self.group)
        return allv[(self.rank - 1) % self.world][1], allv[(self.rank + 1) % self.world][0]


# ---------------------------------------------------------------------------------------------------------------
# Operators
# ---------------------------------------------------------------------------------------------------------------
def transport_operator(shape, lam, periodic=False, ref_sign_quirk=False, slab=None):
    """A = I + upwind divergence on an nx x ny x nz Cartesian grid, velocity components >= 0.

    Interior face with outward normal n: un = n.a; un > 0 adds lambda to the diagonal, un < 0 adds -lambda*un/|un| ...
    the reference writes the off-diagonal as ``-dt*S/V*un`` with un < 0, i.e. POSITIVE (src/TransportEquation.cxx:109-112,
    SURVEY.md F11).  ``ref_sign_quirk=True`` reproduces that; the default is the mathematically consistent sign
    (off-diagonal -lambda), which is the matrix the circulant model [1, -1] approximates.  Border faces do nothing
    ("Neumann", :114-131) unless ``periodic``.
    """
    nx, ny, nz = shape
    sgn = +1.0 if ref_sign_quirk else -1.0
    slab = slab or Slab(enabled=False)
    nzl = nz // slab.world                                  # u is this rank's z-slab (all of it for one process)
    first, last = slab.rank == 0, slab.rank == slab.world - 1

    def apply(u):
        U = u.reshape(nzl, ny, nx)
        out = U.clone()
        for ax, n, l in ((2, nx, lam[0]), (1, ny, lam[1]), (0, nz, lam[2])):
            if n == 1 or l == 0:
                continue
            nloc = nzl if ax == 0 else n
            below = slab.halo(U)[0] if (ax == 0 and slab.on) else None      # plane z0 - 1 from the rank below
            sh = torch.zeros_like(U)                          # sh[i] = U[i - 1]
            idx_dst = [slice(None)] * 3
            idx_src = [slice(None)] * 3
            idx_dst[ax] = slice(1, None)
            idx_src[ax] = slice(0, -1)
            sh[tuple(idx_dst)] = U[tuple(idx_src)]
            if periodic:
                if below is not None:
                    sh[0] = below
                else:
                    idx0 = [slice(None)] * 3
                    idxl = [slice(None)] * 3
                    idx0[ax], idxl[ax] = 0, -1
                    sh[tuple(idx0)] = U[tuple(idxl)]
                out = out + l * U + sgn * l * sh
                continue
            diag = torch.ones(nloc, dtype=U.real.dtype, device=U.device)
            if ax != 0 or last:
                diag[-1] = 0.0                              # last cell: its +d face is a border -> nothing
            shp = [1, 1, 1]
            shp[ax] = nloc
            out = out + l * U * diag.reshape(shp)
            if below is not None and not first:
                sh[0] = below                               # (global first cell: its -d face is a border -> nothing)
            out = out + sgn * l * sh
        return out.reshape(-1)

    return apply


def wave_operator(shape, c0, mu, periodic=False, slab=None):
    """A = I + divMat for the wave system, unknowns [p, q_x, q_y, q_z] per cell (interleaved).

    Am(n) = (A(n) - |A|(n))/2 * mu_d  (jacobianMatrices, src/WaveSystem.cxx:92-107); interior faces add Am to (j, nb)
    and -Am to (j, j) (:145-146); wall borders add -Am (2 v v^T), v = (0, n) (:150-158); periodic borders behave as
    interior faces (:159-167).
    """
    nx, ny, nz = shape
    slab = slab or Slab(enabled=False)
    nzl = nz // slab.world
    first, last = slab.rank == 0, slab.rank == slab.world - 1

    def apply(u):
        U = u.reshape(nzl, ny, nx, 4)
        out = U.clone()
        halo = slab.halo(U) if (slab.on and nz > 1) else None
        for d, (ax, n) in enumerate(((2, nx), (1, ny), (0, nz))):
            if n == 1:
                continue
            m = mu[d]
            for s in (+1.0, -1.0):
                Am = torch.zeros(4, 4, dtype=U.dtype, device=U.device)
                Am[0, 0] = -0.5 * c0 * m
                Am[0, d + 1] = 0.5 * c0 * c0 * s * m
                Am[d + 1, 0] = 0.5 * s * m
                Am[d + 1, d + 1] = -0.5 * c0 * m
                nb = torch.roll(U, int(-s), dims=ax)        # neighbour across the face with outward normal s*e_d
                if ax == 0 and halo is not None:            # the neighbour of the slab's edge plane lives on the next rank
                    nb = nb.clone()
                    nb[-1 if s > 0 else 0] = halo[1] if s > 0 else halo[0]
                contrib = (nb - U) @ Am.T
                if not periodic and (ax != 0 or (last if s > 0 else first)):
                    # border cells: replace the interior-face term by the wall term -Am (2 v v^T) U
                    W = torch.zeros(4, 4, dtype=U.dtype, device=U.device)
                    W[0, d + 1] = -c0 * c0 * s * m
                    W[d + 1, d + 1] = c0 * m
                    wall = U @ W.T
                    sel = [slice(None)] * 4
                    sel[ax] = -1 if s > 0 else 0
                    contrib[tuple(sel)] = wall[tuple(sel)]
                out = out + contrib
        return out.reshape(-1)

    return apply


# ---------------------------------------------------------------------------------------------------------------
# GMRES(m), left preconditioned, PETSc's default test on the preconditioned residual
# ---------------------------------------------------------------------------------------------------------------
def gmres(A, b, M=None, rtol=1e-5, atol=1e-5, maxits=1000, restart=30, slab=None):
    """Solve A x = b from x0 = 0.  Returns (x, iterations, reason, residual_history).

    reason follows KSPConvergedReason: 2 = rtol, 3 = atol, -3 = its.  The test is PETSc's KSPConvergedDefault for a
    left-preconditioned method: ||M^-1 r_k|| <= max(rtol * ||M^-1 b||, atol), evaluated from the Givens recurrence.
    """
    ident = M is None
    M = (lambda v: v) if ident else M
    slab = slab or Slab(enabled=False)

    def norm(v):                                             # global 2-norm (b, x, the Krylov vectors are z-slabs)
        return math.sqrt(slab.sum(torch.sum(v.real * v.real + v.imag * v.imag if v.is_complex() else v * v)).item())

    x = torch.zeros_like(b)
    r = M(b)
    beta = norm(r)
    rnorm0 = beta
    hist = [beta]
    its = 0
    if beta <= atol:
        return x, 0, 3, hist
    ttol = max(rtol * rnorm0, atol)
    cdtype = b.dtype
    while True:
        m = restart
        V = [r / beta]
        H = torch.zeros(m + 1, m, dtype=cdtype, device="cpu")
        cs = [None] * m
        sn = [None] * m
        g = torch.zeros(m + 1, dtype=cdtype)
        g[0] = beta
        k_used = 0
        reason = 0
        for k in range(m):
            w = M(A(V[k]))
            # classical Gram-Schmidt with one refinement pass (PETSc: KSPGMRESClassicalGramSchmidtOrthogonalization,
            # refine "if needed"; always refining is the conservative choice and keeps iteration counts stable)
            Vk = torch.stack(V, dim=0)
            h = slab.sum(torch.mv(Vk.conj(), w))
            w = w - torch.mv(Vk.T, h)
            h2 = slab.sum(torch.mv(Vk.conj(), w))
            w = w - torch.mv(Vk.T, h2)
            h = (h + h2).cpu()
            hn = norm(w)
            H[: k + 1, k] = h
            H[k + 1, k] = hn
            for i in range(k):                               # previous rotations
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i].conj() * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            a_, b_ = H[k, k], H[k + 1, k]
            den = math.sqrt(abs(a_) ** 2 + abs(b_) ** 2)
            cs[k] = torch.tensor(float(abs(a_)) / den if den else 1.0, dtype=cdtype)
            ph = a_ / abs(a_) if abs(a_) > 0 else torch.tensor(1.0, dtype=cdtype)
            sn[k] = ph * b_.conj() / den if den else torch.tensor(0.0, dtype=cdtype)
            H[k, k] = cs[k] * a_ + sn[k] * b_
            H[k + 1, k] = 0
            g[k + 1] = -sn[k].conj() * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_used = k + 1
            res = abs(g[k + 1].item())
            hist.append(res)
            if res <= ttol:
                reason = 2 if res <= rtol * rnorm0 else 3
                break
            if its >= maxits:
                reason = -3
                break
            if hn == 0.0:
                reason = 2
                break
            V.append(w / hn)
        y = torch.linalg.solve_triangular(H[:k_used, :k_used], g[:k_used].reshape(-1, 1), upper=True).reshape(-1)
        Vk = torch.stack(V[:k_used], dim=0)
        x = x + torch.mv(Vk.T, y.to(b.device))
        if reason != 0:
            return x, its, reason, hist
        r = M(b - A(x))
        beta = norm(r)
        if beta <= ttol:
            return x, its, 2 if beta <= rtol * rnorm0 else 3, hist


def spherical_step(shape, inside, outside, rmax=0.3, lo=-0.5, hi=0.5, device="cpu", dtype=torch.float64, slab=None):
    """Cell-centred spherical step initial condition (src/TransportEquation.cxx:25-73, src/WaveSystem.cxx:26-76);
    with a slab, this rank's z planes only."""
    nx, ny, nz = shape

    def centres(n):
        d = (hi - lo) / n
        return lo + d * (torch.arange(n, dtype=torch.float64, device=device) + 0.5)
    zc = centres(nz)
    if slab is not None and slab.world > 1:
        nzl = nz // slab.world
        zc = zc[slab.rank * nzl:(slab.rank + 1) * nzl]
    z, y, x = torch.meshgrid(zc, centres(ny), centres(nx), indexing="ij")
    c = 0.5 * (lo + hi)
    r = torch.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    return torch.where(r < rmax, torch.tensor(inside, dtype=dtype, device=device),
                       torch.tensor(outside, dtype=dtype, device=device)).reshape(-1)
