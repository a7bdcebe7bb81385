"""CPU oracle for the circulant / block-circulant preconditioner apply.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path
(``circulantpreconditioner_b200/``) may import this module; it is used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the *checker* (and the timed CPU baseline).

It is a plain numpy restatement of the reference's algorithm
(``/root/reference/src/FftLinearSolver_3D.c`` and the numpy tests under
``/root/reference/tests/FFTDirectSolver/``).  Every function cites the
reference lines it follows.  The arithmetic that the reference delegates to
PETSc ``MATFFTW`` -> FFTW 3.3.x (un-vendored, not installable here) is
restated with (a) an O(n^2) textbook DFT for the tiny known-answer cases and
(b) scipy's pocketfft for everything larger; both are the same mathematical
DFT (forward sign ``exp(-2 pi i ...)``, unnormalised, ``dims = {nz, ny, nx}``
row-major with x fastest) so they differ from FFTW only by rounding
(~1e-16 * sqrt(log N)).

Parity pinning: ``tests/golden/make_golden.py`` runs the reference's own
Python tests (imported from ``/root/reference``) and stores their inputs and
outputs; ``tests/test_oracle.py`` checks this module against those fixtures
and against the integer known-answer vectors of the reference's C tests
(SURVEY.md Appendix B).  Second pin: the reference's own
``src/FftLinearSolver_3D.c``, compiled unmodified against a CPU stand-in for
the PETSc calls it makes (``oracle/petsc_standin/``, ``oracle/_ref/``), must
agree with this module on Diag, on every solver entry point and on the
fixtures (``tests/test_reference_c.py``).
"""
from __future__ import annotations

import os

import numpy as np

try:  # scipy is present in the image; the naive DFT covers its absence for tiny cases
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None


# ---------------------------------------------------------------------------
# DFT back ends (the part the reference delegates to PETSc MATFFTW / FFTW3)
# ---------------------------------------------------------------------------
def dft_naive_1d(v: np.ndarray, sign: int = -1) -> np.ndarray:
    """Textbook O(n^2) unnormalised DFT along the last axis.

    ``sign=-1`` is FFTW_FORWARD (what ``MatMult(FFT_MAT, ...)`` does,
    FftLinearSolver_3D.c:170), ``sign=+1`` is FFTW_BACKWARD
    (``MatMultTranspose``, FftLinearSolver_3D.c:180).  Twiddle exponents are
    reduced mod n in integers so the matrix entries are correctly rounded.
    """
    v = np.asarray(v, dtype=np.complex128)
    n = v.shape[-1]
    k = np.arange(n)
    e = (np.outer(k, k) % n).astype(np.float64)
    w = np.exp(sign * 2j * np.pi * e / n)
    return v @ w.T


def dft_naive_3d(v: np.ndarray, sign: int = -1) -> np.ndarray:
    """Separable naive 3-D DFT of an array shaped (nz, ny, nx)."""
    out = np.asarray(v, dtype=np.complex128)
    for ax in range(out.ndim):
        out = np.moveaxis(dft_naive_1d(np.moveaxis(out, ax, -1), sign), -1, ax)
    return out


def fft3_forward(v3: np.ndarray, workers: int = 1, naive: bool = False) -> np.ndarray:
    """Unnormalised forward 3-D DFT, ``MatMult`` on MATFFTW (FftLinearSolver_3D.c:170)."""
    if naive or _sfft is None:
        return dft_naive_3d(v3, -1)
    return _sfft.fftn(v3, workers=workers)


def fft3_backward(v3: np.ndarray, workers: int = 1, naive: bool = False) -> np.ndarray:
    """Unnormalised backward 3-D DFT, ``MatMultTranspose`` (FftLinearSolver_3D.c:180)."""
    if naive or _sfft is None:
        return dft_naive_3d(v3, +1)
    return _sfft.ifftn(v3, workers=workers, norm="forward")


# ---------------------------------------------------------------------------
# Eigenvalue set-up  (FftLinearSolver_3D.c:80-164)
# ---------------------------------------------------------------------------
def build_transport_col(size: int) -> np.ndarray:
    """First column ``[1, -1, 0, ...]``; all zero when ``size == 1``.

    Follows FftLinearSolver_3D.c:80-90 (``VecSet(c,0)`` then two
    ``VecSetValue`` guarded by ``size>1``).
    """
    c = np.zeros(size, dtype=np.complex128)
    if size > 1:
        c[0] = 1.0
        c[1] = -1.0
    return c


def column_hat(size: int, naive: bool = False) -> np.ndarray:
    """1-D forward DFT of the transport column (FftLinearSolver_3D.c:235-249,
    PCSHELLFft_3D.cxx:60-66): ``c_hat[q] = 1 - exp(-2 pi i q / n)``."""
    c = build_transport_col(size)
    if naive or _sfft is None:
        return dft_naive_1d(c, -1)
    return _sfft.fft(c)


def vec_kronecker_product_identity_left(c, c_size, id_size, lam):
    """``res[j*c_size + i] = lam * c[i]`` (FftLinearSolver_3D.c:92-112) == np.tile."""
    c = np.asarray(c)
    res = np.empty(c_size * id_size, dtype=np.complex128)
    for i in range(c_size):
        res[i::c_size] = c[i] * lam
    return res


def vec_kronecker_product_identity_right(c, c_size, id_size, lam):
    """``res[i*id_size + j] = lam * c[i]`` (FftLinearSolver_3D.c:114-134) == np.repeat."""
    c = np.asarray(c)
    res = np.empty(c_size * id_size, dtype=np.complex128)
    for i in range(c_size):
        res[i * id_size:(i + 1) * id_size] = c[i] * lam
    return res


def build_diag_mat_vec_3D(c_x_hat, c_y_hat, c_z_hat, n_x, n_y, n_z, lambda_x, lambda_y, lambda_z):
    """``Diag = 1 + kpi_x + kpi_y + kpi_z`` with x the fastest index.

    Follows FftLinearSolver_3D.c:136-164 call by call (the same composition
    as ``testFftSolver_3D.py:26-36``: tile / repeat(tile) / repeat).
    """
    kpi_x = vec_kronecker_product_identity_left(c_x_hat, n_x, n_y * n_z, lambda_x)       # :146
    kpi_y_int = vec_kronecker_product_identity_left(c_y_hat, n_y, n_z, lambda_y)         # :147
    kpi_y = vec_kronecker_product_identity_right(kpi_y_int, n_y * n_z, n_x, 1.0)         # :148
    kpi_z = vec_kronecker_product_identity_right(c_z_hat, n_z, n_x * n_y, lambda_z)      # :149
    s = np.zeros(n_x * n_y * n_z, dtype=np.complex128)                                   # :151 (intended zero)
    s += kpi_x                                                                           # :152
    s += kpi_y                                                                           # :153
    s += kpi_z                                                                           # :154
    s += 1.0                                                                             # :155 VecShift
    return s


def transport_diag(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z, naive=False):
    """Set-up path of ``FftTransportSolver`` / ``setupFFTPrec3D``:
    three column DFTs then ``build_diag_mat_vec_3D``
    (FftLinearSolver_3D.c:218-249, PCSHELLFft_3D.cxx:51-69)."""
    return build_diag_mat_vec_3D(column_hat(n_x, naive), column_hat(n_y, naive), column_hat(n_z, naive),
                                 n_x, n_y, n_z, lambda_x, lambda_y, lambda_z)


# ---------------------------------------------------------------------------
# The hot path  (FftLinearSolver_3D.c:166-190, complex-scalar branch)
# ---------------------------------------------------------------------------
def solve_3D(Diag, b, n_x, n_y, n_z, workers: int = 1, naive: bool = False):
    """``X = (1/size) * F^H( F(b) / Diag )``.

    FftLinearSolver_3D.c:166-190: MatMult (:170), VecPointwiseDivide (:174),
    MatMultTranspose (:180), VecScale(1/size) (:184).  ``dims = {n_z,n_y,n_x}``
    (PCSHELLFft_3D.cxx:34), i.e. ``b.reshape(n_z, n_y, n_x)`` as in
    testFftSolver_3D.py:38-52.
    """
    size = n_x * n_y * n_z
    b3 = np.asarray(b, dtype=np.complex128).reshape(n_z, n_y, n_x)
    b_hat = fft3_forward(b3, workers, naive)                       # :170
    b_hat = b_hat / np.asarray(Diag).reshape(n_z, n_y, n_x)        # :174
    X = fft3_backward(b_hat, workers, naive)                       # :180
    X = X * (1.0 / size)                                           # :184
    return X.reshape(-1)


def FftTransportSolver(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z, b, workers=1, naive=False):
    """FftLinearSolver_3D.c:218-264 + Fft3DSolver :192-216 (set-up then one apply)."""
    Diag = transport_diag(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z, naive)
    return solve_3D(Diag, b, n_x, n_y, n_z, workers, naive)


def FftTransportSolver_z_recurrence(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z, b):
    """The same solve with the z part done WITHOUT FFTs -- the form the CUDA middle pass takes for this symbol
    (circulantpreconditioner_b200/csrc/zsolve.cuh); kept here so that the equivalence is pinned on the CPU.

    After the x and y transforms (FftLinearSolver_3D.c:170, two of the three 1-D factors), every z line is a circulant
    system with first column ``[alpha + lz, -lz, 0, ...]``, ``alpha = 1 + lx c_x_hat[kx] + ly c_y_hat[ky]``
    (build_diag_mat_vec_3D :146-157 says its spectrum is ``Diag[k]``), i.e.
    ``(alpha + lz) x_k - lz x_{k-1} = b_k`` cyclically: a first-order recurrence ``y_k = c y_{k-1} + b_k``,
    ``c = lz / (alpha + lz)``, closed by ``y_{-1} = y_{n-1}^{(0)} / (1 - c^n)``, then ``x = y / (alpha + lz)``.
    Requires ``lambda_z >= 0`` and ``Re alpha >= 1`` (then ``|c| < 1``)."""
    b3 = np.asarray(b, dtype=np.complex128).reshape(n_z, n_y, n_x)
    cx = np.fft.fft(build_transport_col(n_x))
    cy = np.fft.fft(build_transport_col(n_y))
    alpha = 1.0 + lambda_x * cx[None, :] + lambda_y * cy[:, None]
    if lambda_z < 0 or alpha.real.min() < 1.0 - 1e-12:
        raise ValueError("the recurrence form needs lambda_z >= 0 and Re(alpha) >= 1")
    bh = np.fft.fft(np.fft.fft(b3, axis=2), axis=1)
    r = 1.0 / (alpha + lambda_z)
    c = lambda_z * r
    y = np.empty_like(bh)
    acc = np.zeros((n_y, n_x), dtype=np.complex128)
    for k in range(n_z):                      # zero carry-in
        acc = c * acc + bh[k]
        y[k] = acc
    carry = acc / (1.0 - c ** n_z)            # cyclic closure
    cp = c.copy()
    for k in range(n_z):
        y[k] = (y[k] + cp * carry) * r
        cp = cp * c
    X = np.fft.ifft(np.fft.ifft(y, axis=1), axis=2)
    return X.reshape(-1)


def FftTransportSolver_z_line_form(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z, b, weight_floor=1e-17, slabs=1):
    """The recurrence form as the CUDA line kernels evaluate it (csrc/zsolve.cuh: zs_end_accum_kernel,
    zs_carry_owner_kernel, zs_dist_line_kernel), restated so that its one approximation is pinned on the CPU:

    * the carry into the first plane of a slab is summed only over the planes whose weight ``|c|^m`` can still reach
      ``weight_floor`` (in whole groups of 16 planes, like the kernel), the rest being below the rounding of the sum;
    * with ``slabs = P`` the line is cut into P z-slabs: every slab's end value from a zero carry-in, the cycle closed
      over the slabs by the line's owner (``Zin_0`` by Horner, ``Zin_{r+1} = e_r + c^(nz/P) Zin_r``), then every slab
      solved from its carry-in -- the multi-rank schedule; ``slabs = 1`` is the single-GPU line form.
    Same operator as solve_3D (FftLinearSolver_3D.c:166-190) to rounding."""
    b3 = np.asarray(b, dtype=np.complex128).reshape(n_z, n_y, n_x)
    cx = np.fft.fft(build_transport_col(n_x))
    cy = np.fft.fft(build_transport_col(n_y))
    alpha = (1.0 + lambda_x * cx[None, :] + lambda_y * cy[:, None]).reshape(-1)
    if lambda_z < 0 or alpha.real.min() < 0.5:
        raise ValueError("the recurrence form needs lambda_z >= 0 and Re(alpha) >= 1/2")
    if n_z % slabs:
        raise ValueError("n_z must be divisible by the number of slabs")
    bh = np.fft.fft(np.fft.fft(b3, axis=2), axis=1).reshape(n_z, -1)
    r = 1.0 / (alpha + lambda_z)
    c = lambda_z * r
    nzl = n_z // slabs
    # planes of a slab that still matter, per line
    c2 = np.abs(c) ** 2
    with np.errstate(divide="ignore"):
        m = np.where(c2 > 0, np.log(weight_floor ** 2) / np.log(np.where(c2 > 0, c2, 0.5)) + 1.0, 1.0)
    start = np.where(m < nzl, nzl - np.minimum(m, nzl).astype(np.int64), 0)
    start = np.maximum(nzl - ((nzl - start + 15) // 16) * 16, 0)
    ends = np.zeros((slabs, alpha.size), dtype=np.complex128)
    for s_ in range(slabs):                   # end value of every slab from a zero carry-in, truncated sum
        acc = np.zeros(alpha.size, dtype=np.complex128)
        for k in range(nzl):
            live = k >= start
            acc = np.where(live, c * acc + bh[s_ * nzl + k], 0.0)
        ends[s_] = acc
    cL = c ** nzl
    acc = np.zeros(alpha.size, dtype=np.complex128)
    for s_ in range(slabs):                   # owner: Zin_0 by Horner over e_0 .. e_{P-1}, closed cyclically
        acc = cL * acc + ends[s_]
    Z = acc / (1.0 - cL ** slabs)
    y = np.empty_like(bh)
    for s_ in range(slabs):                   # second sweep of every slab from its carry-in
        acc = Z.copy()
        for k in range(nzl):
            acc = c * acc + bh[s_ * nzl + k]
            y[s_ * nzl + k] = acc * r
        Z = ends[s_] + cL * Z                 # Zin_{r+1} = e_r + cL Zin_r
    X = np.fft.ifft(np.fft.ifft(y.reshape(n_z, n_y, n_x), axis=1), axis=2)
    return X.reshape(-1)


def Fft3DTransportSolver(n_x, n_y, n_z, a_x, a_y, a_z, dt, delta_x, delta_y, delta_z, b, workers=1, naive=False):
    """``lambda_d = a_d * dt / delta_d`` (FftLinearSolver_3D.c:266-281)."""
    return FftTransportSolver(n_x, n_y, n_z, a_x * dt / delta_x, a_y * dt / delta_y, a_z * dt / delta_z,
                              b, workers, naive)


def Fft2DTransportSolver(n_x, n_y, a_x, a_y, dt, delta_x, delta_y, b, workers=1, naive=False):
    """n_z=1, a_z=0, delta_z=1 (FftLinearSolver_3D.c:283-293)."""
    return Fft3DTransportSolver(n_x, n_y, 1, a_x, a_y, 0.0, dt, delta_x, delta_y, 1.0, b, workers, naive)


def Fft1DTransportSolver(n_x, a_x, dt, delta_x, b, workers=1, naive=False):
    """n_y=n_z=1 (FftLinearSolver_3D.c:295-301)."""
    return Fft3DTransportSolver(n_x, 1, 1, a_x, 0.0, 0.0, dt, delta_x, 1.0, 1.0, b, workers, naive)


def solve_first_column(col3, b, n_x, n_y, n_z, workers=1, naive=False):
    """General circulant: eigenvalues are the 3-D DFT of the first column
    (the 1-D case is testFftSolver_1D.py:11-17 / testFftSolver_1D.c:144-177)."""
    Diag = fft3_forward(np.asarray(col3, dtype=np.complex128).reshape(n_z, n_y, n_x), workers, naive).reshape(-1)
    return solve_3D(Diag, b, n_x, n_y, n_z, workers, naive)


# ---------------------------------------------------------------------------
# Matrix-free / dense operators used to *check* a solve (b := C x_ref  =>  x == x_ref)
# ---------------------------------------------------------------------------
def apply_transport_matrix(x, n_x, n_y, n_z, lambda_x, lambda_y, lambda_z):
    """``C x`` with ``C = I + sum_d lambda_d (I - S_d)``, ``(S_x u)_i = u_{i-1}`` periodic.

    Same matrix as ``build_C_3D`` in testFftSolver_3D.py:12-24 (circulant of
    column [1,-1] Kronecker identities), applied without assembling it.
    Degenerate axes (n=1) contribute nothing (column is zero, :80-90).
    """
    u = np.asarray(x).reshape(n_z, n_y, n_x)
    out = u.copy()
    for ax, (n, lam) in zip((2, 1, 0), ((n_x, lambda_x), (n_y, lambda_y), (n_z, lambda_z))):
        if n > 1:
            out = out + lam * (u - np.roll(u, 1, axis=ax))
    return out.reshape(-1)


def dense_transport_matrix(n_x, n_y, n_z, lambda_x, lambda_y, lambda_z):
    """Explicit Kronecker assembly (testFftSolver_3D.py:12-24); tiny sizes only."""
    def circ(n):
        c = build_transport_col(n).real
        m = np.zeros((n, n))
        for j in range(n):
            m[:, j] = np.roll(c, j)
        return m
    Cx = np.kron(np.eye(n_y * n_z), circ(n_x))
    Cy = np.kron(np.eye(n_z), np.kron(circ(n_y), np.eye(n_x)))
    Cz = np.kron(circ(n_z), np.eye(n_x * n_y))
    return np.eye(n_x * n_y * n_z) + lambda_x * Cx + lambda_y * Cy + lambda_z * Cz


# ---------------------------------------------------------------------------
# Wave system 4x4 block symbol (derived from WaveSystem.cxx:92-107)
# ---------------------------------------------------------------------------
def wave_jacobian_minus(normal, coeff, c0):
    """``Am = (A - |A|)/2 * coeff`` of ``jacobianMatrices`` (WaveSystem.cxx:92-107)."""
    dim = len(normal)
    A = np.zeros((dim + 1, dim + 1))
    absA = np.zeros((dim + 1, dim + 1))
    absA[0, 0] = c0 * coeff
    for i in range(dim):
        A[i + 1, 0] = normal[i] * coeff
        A[0, i + 1] = c0 * c0 * normal[i] * coeff
        for j in range(dim):
            absA[i + 1, j + 1] = c0 * normal[i] * normal[j] * coeff
    return (A - absA) * 0.5


def apply_wave_matrix(u, n_x, n_y, n_z, c0, mu_x, mu_y, mu_z):
    """``(I + divMat) u`` on a fully periodic Cartesian grid.

    Interior/periodic faces of ``computeDivergenceMatrix`` (WaveSystem.cxx:145-146,
    165-166: ``+Am`` at (j, neighbour), ``-Am`` at (j, j)) followed by
    ``MatShift(A, 1)`` (tests/WaveSystem_SphericalExplosion_impl_mpi.cxx:127).
    Unknown ordering ``u[4*cell + comp]`` (:104-115), ``mu_d = dt/delta_d``.
    """
    U = np.asarray(u).reshape(n_z, n_y, n_x, 4)
    out = U.copy()
    for ax, n, mu, d in ((2, n_x, mu_x, 0), (1, n_y, mu_y, 1), (0, n_z, mu_z, 2)):
        if n == 1:
            continue
        for sgn in (+1, -1):
            normal = [0.0, 0.0, 0.0]
            normal[d] = float(sgn)
            Am = wave_jacobian_minus(normal, mu, c0)
            nb = np.roll(U, -sgn, axis=ax)          # neighbour across the face with outward normal sgn*e_d
            out = out + (nb - U) @ Am.T
    return out.reshape(-1)


def wave_symbol(n_x, n_y, n_z, c0, mu_x, mu_y, mu_z):
    """Per-frequency arrow matrix ``M_hat[k,j,i]`` (shape (nz,ny,nx,4,4)), SURVEY.md A.2."""
    th = [2 * np.pi * np.arange(n) / n for n in (n_x, n_y, n_z)]
    one_m_cos = [(1 - np.cos(t)) if n > 1 else np.zeros(1) for t, n in zip(th, (n_x, n_y, n_z))]
    sin = [np.sin(t) if n > 1 else np.zeros(1) for t, n in zip(th, (n_x, n_y, n_z))]
    mu = (mu_x, mu_y, mu_z)
    M = np.zeros((n_z, n_y, n_x, 4, 4), dtype=np.complex128)
    shape = [(1, 1, n_x), (1, n_y, 1), (n_z, 1, 1)]
    M[..., 0, 0] = 1.0
    for d in range(3):
        omc = one_m_cos[d].reshape(shape[d])
        s = sin[d].reshape(shape[d])
        M[..., 0, 0] += c0 * mu[d] * omc
        M[..., 0, d + 1] = 1j * c0 * c0 * mu[d] * s
        M[..., d + 1, 0] = 1j * mu[d] * s
        M[..., d + 1, d + 1] = 1.0 + c0 * mu[d] * omc
    return M


def solve_wave_block(b, n_x, n_y, n_z, c0, mu_x, mu_y, mu_z, workers=1, naive=False, dense=False):
    """Block-circulant apply: 4 forward FFTs, per-frequency 4x4 solve, 4 inverse FFTs.

    ``dense=True`` uses ``np.linalg.solve`` on the 4x4 blocks, otherwise the
    closed-form Schur complement of the arrow matrix.
    """
    N = n_x * n_y * n_z
    B = np.asarray(b, dtype=np.complex128).reshape(n_z, n_y, n_x, 4)
    Bh = np.stack([fft3_forward(B[..., c], workers, naive) for c in range(4)], axis=-1)
    M = wave_symbol(n_x, n_y, n_z, c0, mu_x, mu_y, mu_z)
    if dense:
        Yh = np.linalg.solve(M, Bh[..., None])[..., 0]
    else:
        D = np.stack([M[..., d + 1, d + 1] for d in range(3)], axis=-1)           # D_d
        up = np.stack([M[..., 0, d + 1] for d in range(3)], axis=-1)              # i c0^2 s_d
        lo = np.stack([M[..., d + 1, 0] for d in range(3)], axis=-1)              # i s_d
        num = Bh[..., 0] - np.sum(up * Bh[..., 1:] / D, axis=-1)
        den = M[..., 0, 0] - np.sum(up * lo / D, axis=-1)
        p = num / den
        Yh = np.empty_like(Bh)
        Yh[..., 0] = p
        Yh[..., 1:] = (Bh[..., 1:] - lo * p[..., None]) / D
    Y = np.stack([fft3_backward(Yh[..., c], workers, naive) for c in range(4)], axis=-1) * (1.0 / N)
    return Y.reshape(-1)


# ---------------------------------------------------------------------------
# Physics inputs of the BASELINE configs (structured restatements)
# ---------------------------------------------------------------------------
def spherical_step(n_x, n_y, n_z, inside, outside, rmax=0.3, lo=-0.5, hi=0.5):
    """Cell-centred spherical step (TransportEquation.cxx:25-73 / WaveSystem.cxx:26-76)."""
    def centres(n):
        d = (hi - lo) / n
        return lo + d * (np.arange(n) + 0.5)
    z, y, x = np.meshgrid(centres(n_z), centres(n_y), centres(n_x), indexing="ij")
    c = 0.5 * (lo + hi)
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    return np.where(r < rmax, inside, outside).reshape(-1)


def default_workers() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
