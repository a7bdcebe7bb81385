"""The oracle against the reference's OWN C code path.

oracle/_ref/libreference_fftsolver.so is /root/reference/src/FftLinearSolver_3D.c, unmodified, compiled from where it lies
against a CPU stand-in for the ~30 PETSc calls it makes (oracle/petsc_standin/: sequential complex Vecs, MATFFTW as a
direct O(n^2) DFT with FFTW's conventions).  Everything that file does itself -- build_transport_col (:80-90), the
Kronecker layout of Diag (:92-164), solve_3D's forward / divide / backward / VecScale(1/size) (:166-190), the wrappers
with their lambdas and degenerate axes (:192-312) -- therefore runs exactly as the reference wrote it, and the oracle
(oracle/circulant_oracle.py, circulant_oracle.c) must reproduce it, as must the fixtures generated from the reference's
Python tests.  The reference's own C test programs (tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.c) are built the same
way and must run clean.  Needs the built files (made where /root/reference exists; they travel to the GPU box).
"""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import circulant_oracle as O
from oracle import ref_c as R
from tests.conftest import rel_l2

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libreference_fftsolver.so is not built")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cplx(rng, n):
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


SHAPES = [((4, 3, 2), (1.0, 1.0, 1.0)), ((10, 25, 40), (0.6, 0.15, 0.02)), ((16, 16, 16), (55.5556, 0.0, 0.0)),
          ((32, 32, 32), (0.6, 0.15, 0.02)), ((7, 5, 3), (2.0, 0.5, 4.0)), ((8, 1, 1), (1.0, 0.0, 0.0)),
          ((5, 7, 1), (3.0, 0.3, 0.0)), ((1, 1, 6), (0.0, 0.0, 2.0)), ((2, 2, 2), (1.0, 2.0, 3.0)), ((1, 1, 1), (0.0, 0.0, 0.0))]


@pytest.mark.parametrize("shape,lam", SHAPES)
def test_diag_of_the_reference_setup(shape, lam):
    """build_transport_col + 1-D MatMult x 3 + build_diag_mat_vec_3D, as FftTransportSolver chains them."""
    nx, ny, nz = shape
    want = R.build_diag(nx, ny, nz, *lam)
    assert np.allclose(O.transport_diag(nx, ny, nz, *lam), want, rtol=0, atol=1e-13)
    d = np.empty(nx * ny * nz, dtype=np.complex128)
    CO.lib().oracle_transport_diag(CO._p(d), nx, ny, nz, *lam)
    assert np.allclose(d, want, rtol=0, atol=1e-13)
    # the layout, spelled out: Diag[i + nx (j + ny k)] = 1 + lx cx[i] + ly cy[j] + lz cz[k]
    c = [1.0 - np.exp(-2j * np.pi * np.arange(n) / n) if n > 1 else np.zeros(1) for n in shape]
    full = 1.0 + lam[0] * c[0][None, None, :] + lam[1] * c[1][None, :, None] + lam[2] * c[2][:, None, None]
    assert np.allclose(full.ravel(), want, rtol=0, atol=1e-13)


@pytest.mark.parametrize("shape,lam", SHAPES)
def test_transport_solve_equals_the_reference(shape, lam):
    nx, ny, nz = shape
    b = _cplx(np.random.default_rng(11), nx * ny * nz)
    want = R.FftTransportSolver(nx, ny, nz, *lam, b)
    assert rel_l2(O.FftTransportSolver(nx, ny, nz, *lam, b), want) < 1e-13
    x = np.empty_like(b)
    CO.lib().oracle_Fft3DTransportSolver(nx, ny, nz, *lam, 1.0, 1.0, 1.0, 1.0, CO._p(x), CO._p(b))      # lambda = a dt / delta
    assert rel_l2(x, want) < 1e-13
    # solve_3D with the eigenvalues handed over, and b == x aliasing as the reference's driver uses it
    Diag = R.build_diag(nx, ny, nz, *lam)
    assert rel_l2(R.solve_3D(Diag, b, nx, ny, nz), want) < 1e-14
    assert rel_l2(O.solve_3D(Diag, b, nx, ny, nz), want) < 1e-13
    assert rel_l2(R.FftTransportSolver_in_place(nx, ny, nz, *lam, b), want) < 1e-14
    # the defining identity on the reference itself: b := C x_ref  =>  x = x_ref
    x_ref = np.random.default_rng(12).random(nx * ny * nz)
    bb = O.apply_transport_matrix(x_ref, nx, ny, nz, *lam).astype(np.complex128)
    got = R.FftTransportSolver(nx, ny, nz, *lam, bb)
    assert rel_l2(got.real, x_ref) < 1e-12 and np.abs(got.imag).max() < 1e-12 * max(1.0, np.abs(x_ref).max())


def test_wrappers_lambdas_and_degenerate_axes():
    """Fft3DTransportSolver (lambda = a dt / delta, :274-276), Fft2D / Fft1D (n = 1, a = 0, delta = 1, :283-301) and the
    by-value context of PetscFft3DTransportSolver (:303-312)."""
    rng = np.random.default_rng(13)
    nx, ny, nz = 6, 5, 4
    a, dt, dl = (1.0, 0.5, 0.25), 0.3, (0.1, 0.2, 0.4)
    lam = tuple(a[d] * dt / dl[d] for d in range(3))
    b = _cplx(rng, nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    assert rel_l2(R.Fft3DTransportSolver(nx, ny, nz, *a, dt, *dl, b), want) < 1e-13
    assert rel_l2(R.PetscFft3DTransportSolver(nx, ny, nz, *a, dt, *dl, b), want) < 1e-13
    assert rel_l2(O.Fft3DTransportSolver(nx, ny, nz, *a, dt, *dl, b), want) < 1e-13
    b2 = _cplx(rng, nx * ny)
    assert rel_l2(R.Fft2DTransportSolver(nx, ny, a[0], a[1], dt, dl[0], dl[1], b2),
                  O.FftTransportSolver(nx, ny, 1, lam[0], lam[1], 0.0, b2)) < 1e-13
    b1 = _cplx(rng, nx)
    assert rel_l2(R.Fft1DTransportSolver(nx, a[0], dt, dl[0], b1), O.FftTransportSolver(nx, 1, 1, lam[0], 0.0, 0.0, b1)) < 1e-13


@pytest.mark.parametrize("name", ["ref_c_kat2_3x2", "ref_c_kat3_4x3x2", "ref_py_2d_50x200", "ref_py_3d_10x25x40"])
def test_golden_fixtures_through_the_reference_c_code(name):
    """The fixtures made from the reference's Python tests (tests/golden/make_golden.py) and the integer vectors of its C
    tests, replayed through the reference's C functions: the two halves of the reference agree with each other."""
    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    nx, ny, nz = (int(v) for v in f["n"])
    lam = tuple(float(v) for v in f["lam"])
    assert np.allclose(R.build_diag(nx, ny, nz, *lam), f["Diag"], rtol=0, atol=1e-13)
    got = R.FftTransportSolver(nx, ny, nz, *lam, f["b"].astype(np.complex128))
    assert rel_l2(got, f["X"]) < 1e-13
    assert rel_l2(got.real, f["X_ref"]) < 1e-12


@pytest.mark.parametrize("tag", ["phys", "unit"])
def test_config1_32cube_through_the_reference_c_code(tag):
    """BASELINE config 1 (FFTDirectSolver, 32^3) on the reference's C path."""
    f = np.load(os.path.join(GOLDEN, f"ref_py_3d_32cube_{tag}.npz"))
    nx, ny, nz = (int(v) for v in f["n"])
    lam = tuple(float(v) for v in f["lam"])
    got = R.FftTransportSolver(nx, ny, nz, *lam, f["b"].astype(np.complex128))
    assert rel_l2(got.real, f["X_real"]) < 1e-13 and np.abs(got.imag).max() < 1e-12
    assert rel_l2(got.real, f["X_ref"]) < 1e-12
    assert np.allclose(R.build_diag(nx, ny, nz, *lam)[:64], f["Diag_head"], rtol=0, atol=1e-13)


def test_kat1_first_column_form_is_outside_this_file():
    """KAT-1 (testFftSolver_1D.c: column [1.5, -0.5, 0, 0]) is a general first column; FftLinearSolver_3D.c only builds
    the transport column, so the reference's C path reproduces KAT-1 through solve_3D with Diag = FFT(column)."""
    f = np.load(os.path.join(GOLDEN, "ref_c_kat1_n4.npz"))
    Diag = np.fft.fft(f["col"])
    got = R.solve_3D(Diag, f["b"].astype(np.complex128), 4, 1, 1)
    assert np.allclose(got.real, f["x"], rtol=0, atol=1e-13) and np.allclose(got.real, [6.7, 2.9, 6.3, 20.1], atol=1e-13)


def _run_reference_test_program(name):
    import subprocess
    exe = os.path.join(os.path.dirname(R.PATH), name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} is not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr          # their assert()s held and every PetscCall returned 0
    return r.stdout


def test_the_references_own_c_test_programs_run():
    """tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.c of the reference, compiled unmodified against the stand-in
    (oracle/Makefile): their own assertions hold, and what they print is the known-answer data of SURVEY.md appendix B."""
    import re
    out = _run_reference_test_program("testFftSolver_1D")
    # KAT-1: column [1.5, -0.5, 0, 0], b = [0, 1, 8, 27] -> eigenvalues [1, 1.5+0.5i, 2, 1.5-0.5i], x = [6.7, 2.9, 6.3, 20.1]
    sol = out.split("Vecteur solution x:")[1].split("Matrice circulante")[0]
    vals = [float(t) for t in re.findall(r"^(-?\d+(?:\.\d+)?(?:e[-+]?\d+)?)\s*$", sol, flags=re.M)]
    assert np.allclose(vals, [6.7, 2.9, 6.3, 20.1], rtol=0, atol=1e-13)
    eig = out.split("Valeurs propres lambdas")[1].split("Vecteur second membre b:")[0]
    assert "1.5 + 0.5 i" in eig and "1.5 - 0.5 i" in eig
    res = float(re.search(r"norme du r.sidu = (\S+)", out).group(1))
    assert res < 1e-13
    for name in ("testFftSolver_2D", "testFftSolver_3D"):       # KAT-2 (3 x 2) and KAT-3 (4 x 3 x 2): X_ref[m] = m^3
        out = _run_reference_test_program(name)
        rr = float(re.search(r"Relative residual = (\S+)", out).group(1))
        re_ = float(re.search(r"Relative error = (\S+)", out).group(1))
        assert rr < 1e-14 and re_ < 1e-14, (name, rr, re_)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,lam", [((32, 32, 32), (0.6, 0.15, 0.02)), ((10, 25, 40), (0.6, 0.15, 0.02)),
                                        ((16, 16, 16), (55.5556, 0.0, 0.0)), ((5, 7, 1), (3.0, 0.3, 0.0))])
def test_cuda_path_equals_the_reference_c_code(shape, lam):
    """cpc_apply (and the explicit-Diag route the PCShell set-up takes) against the reference's own C functions."""
    import torch

    import circulantpreconditioner_b200 as cpc
    nx, ny, nz = shape
    b = _cplx(np.random.default_rng(14), nx * ny * nz)
    want = R.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        assert rel_l2(p.apply(torch.from_numpy(b).cuda()).cpu().numpy(), want) < 1e-12
        Diag = R.build_diag(nx, ny, nz, *lam)
        assert np.abs(p.get_diag() - Diag).max() < 1e-13 * max(1.0, np.abs(Diag).max())
        p.set_symbol_diag(Diag)
        assert rel_l2(p.apply(torch.from_numpy(b).cuda()).cpu().numpy(), want) < 1e-12
