"""CPU tests (-m "not gpu"): pin the oracle against the reference's own golden vectors.

Fixtures under tests/golden/ were produced by running the reference's Python tests
(/root/reference/tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.py) through tests/golden/make_golden.py;
the integer known-answer vectors come from the reference's C tests.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import circulant_oracle as O
from tests.conftest import GOLDEN, rel_l2

TOL = 1e-13


def g(name):
    return np.load(os.path.join(GOLDEN, name))


def test_kat1_1d_integer_vector():
    # testFftSolver_1D.c:144-177: N=4, column [1.5,-0.5,0,0], b = i^3  ->  x = [6.7, 2.9, 6.3, 20.1]
    f = g("ref_c_kat1_n4.npz")
    eig = O.fft3_forward(f["col"].astype(complex).reshape(1, 1, 4), naive=True).ravel()
    np.testing.assert_allclose(eig, [1, 1.5 + 0.5j, 2, 1.5 - 0.5j], atol=1e-15)
    for naive in (True, False):
        x = O.solve_first_column(f["col"], f["b"], 4, 1, 1, naive=naive)
        np.testing.assert_allclose(x.real, [6.7, 2.9, 6.3, 20.1], rtol=1e-14)
        np.testing.assert_allclose(x.imag, 0, atol=1e-14)
        assert rel_l2(x, f["x"]) < TOL


def test_kat3_3d_integer_vector():
    # testFftSolver_3D.c:95-141: 4x3x2, lambda=1, X_ref = m^3
    f = g("ref_c_kat3_4x3x2.npz")
    nx, ny, nz = (int(v) for v in f["n"])
    np.testing.assert_allclose(f["b"][:6], [-2267, -2922, -3713, -4606, -4183, -4478])
    D = O.transport_diag(nx, ny, nz, 1.0, 1.0, 1.0, naive=True)
    np.testing.assert_allclose(D[:6], [1, 2 + 1j, 3, 2 - 1j, 2.5 + 0.8660254037844386j, 3.5 + 1.8660254037844386j],
                               atol=1e-14)
    assert abs(np.abs(D).min() - 1.0) < 1e-14 and abs(np.abs(D).max() - 6.557438524302) < 1e-11
    assert np.abs(D - f["Diag"]).max() < 1e-14
    for naive in (True, False):
        X = O.solve_3D(D, f["b"], nx, ny, nz, naive=naive)
        assert rel_l2(X, f["X_ref"]) < TOL
        assert rel_l2(X, f["X"]) < TOL
    Xc = CO.solve_3D(CO.transport_diag(nx, ny, nz, 1.0, 1.0, 1.0), f["b"], nx, ny, nz)
    assert rel_l2(Xc, f["X_ref"]) < TOL


def test_kat2_2d_integer_vector():
    f = g("ref_c_kat2_3x2.npz")
    nx, ny, nz = (int(v) for v in f["n"])
    X = O.Fft2DTransportSolver(nx, ny, 1.0, 1.0, 1.0, 1.0, 1.0, f["b"], naive=True)
    assert rel_l2(X, f["X_ref"]) < TOL
    assert rel_l2(X, f["X"]) < TOL


@pytest.mark.parametrize("name", ["ref_py_2d_50x200.npz", "ref_py_3d_10x25x40.npz"])
def test_python_reference_fixtures(name):
    f = g(name)
    nx, ny, nz = (int(v) for v in f["n"])
    lam = [float(v) for v in f["lam"]]
    D = O.transport_diag(nx, ny, nz, *lam)
    assert np.abs(D - f["Diag"]).max() < 1e-14
    X = O.solve_3D(D, f["b"], nx, ny, nz)
    assert rel_l2(X, f["X"]) < TOL
    assert rel_l2(X, f["X_ref"]) < 1e-12
    # the plain-C restatement agrees with both
    Dc = CO.transport_diag(nx, ny, nz, *lam)
    assert np.abs(Dc - f["Diag"]).max() < 1e-14
    assert rel_l2(CO.solve_3D(Dc, f["b"], nx, ny, nz), f["X"]) < TOL
    # self-consistency identity of the reference tests: b = C X_ref
    assert rel_l2(O.apply_transport_matrix(f["X_ref"], nx, ny, nz, *lam), f["b"]) < TOL


def test_python_reference_1d_fixture():
    f = g("ref_py_1d_n8.npz")
    x = O.solve_first_column(f["col"], f["b"], 8, 1, 1)
    assert rel_l2(x, f["x"]) < TOL
    # [1+l, -l] == 1 + l*[1,-1]: the transport path gives the same answer (FftLinearSolver_3D.c:295-301)
    x2 = O.Fft1DTransportSolver(8, float(f["lam"]), 1.0, 1.0, f["b"])
    assert rel_l2(x2, f["x"]) < TOL


@pytest.mark.parametrize("tag", ["phys", "unit"])
def test_config0_32cube(tag):
    f = g(f"ref_py_3d_32cube_{tag}.npz")
    lam = [float(v) for v in f["lam"]]
    D = O.transport_diag(32, 32, 32, *lam)
    assert np.abs(D[:64] - f["Diag_head"]).max() < 1e-14
    X = O.solve_3D(D, f["b"], 32, 32, 32)
    assert rel_l2(X.real, f["X_real"]) < TOL
    assert np.abs(X.imag).max() < 1e-12
    assert rel_l2(X, f["X_ref"]) < 1e-12
    Xc = CO.Fft3DTransportSolver(32, 32, 32, lam[0], lam[1], lam[2], 1.0, 1.0, 1.0, 1.0, f["b"])
    assert rel_l2(Xc.real, f["X_real"]) < TOL


def test_edge_cases_columns_and_degenerate_axes():
    assert np.all(O.build_transport_col(1) == 0)                       # FftLinearSolver_3D.c:83 (size>1 guard)
    np.testing.assert_allclose(O.column_hat(2), [0, 2], atol=1e-16)
    np.testing.assert_allclose(O.column_hat(1), [0], atol=1e-16)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(12)
    x1 = O.Fft1DTransportSolver(12, 2.0, 0.5, 0.25, b)
    x3 = O.Fft3DTransportSolver(12, 1, 1, 2.0, 0.0, 0.0, 0.5, 0.25, 1.0, 1.0, b)
    assert rel_l2(x1, x3) == 0.0
    C = O.dense_transport_matrix(12, 1, 1, 4.0, 0, 0)
    assert rel_l2(C @ x1.real, b) < TOL


def test_naive_dft_matches_pocketfft_and_c():
    rng = np.random.default_rng(7)
    v = rng.standard_normal((5, 6, 7)) + 1j * rng.standard_normal((5, 6, 7))
    a = O.fft3_forward(v, naive=True)
    assert rel_l2(a, O.fft3_forward(v)) < TOL
    assert rel_l2(CO.dft3(v, 7, 6, 5, -1).reshape(5, 6, 7), a) < TOL
    bk = O.fft3_backward(a, naive=True) / v.size
    assert rel_l2(bk, v) < TOL
    assert rel_l2(CO.dft3(a, 7, 6, 5, +1).reshape(5, 6, 7) / v.size, v) < TOL


def test_wave_block_symbol_against_dense_assembly():
    # FFT + per-frequency 4x4 solve == inverse of the assembled periodic operator
    rng = np.random.default_rng(11)
    nx, ny, nz = 4, 3, 5
    c0, mu = 700.0, (0.0793651, 0.05, 0.03)
    y = rng.standard_normal(nx * ny * nz * 4)
    b = O.apply_wave_matrix(y, nx, ny, nz, c0, *mu)
    # dense operator from unit vectors
    n = y.size
    A = np.stack([O.apply_wave_matrix(e, nx, ny, nz, c0, *mu) for e in np.eye(n)], axis=1)
    y_dense = np.linalg.solve(A, b)
    for dense in (False, True):
        ys = O.solve_wave_block(b, nx, ny, nz, c0, *mu, dense=dense)
        assert np.abs(ys.imag).max() < 1e-9
        assert rel_l2(ys.real, y_dense) < 1e-9       # conditioning ~ c0^2
    assert rel_l2(O.solve_wave_block(b, nx, ny, nz, c0, *mu), O.solve_wave_block(b, nx, ny, nz, c0, *mu, dense=True)) < 1e-12
    # a mild sound speed keeps the system well conditioned: tight check of the closed form
    c0 = 2.0
    b = O.apply_wave_matrix(y, nx, ny, nz, c0, *mu)
    assert rel_l2(O.solve_wave_block(b, nx, ny, nz, c0, *mu).real, y) < 1e-13


def test_spherical_step_counts():
    u = O.spherical_step(16, 16, 16, 650.0, 600.0)
    assert set(np.unique(u)) == {600.0, 650.0}
    frac = (u == 650.0).mean()
    assert abs(frac - 4 / 3 * np.pi * 0.3 ** 3) < 0.02


@pytest.mark.parametrize("shape,lam", [((8, 6, 16), (0.6, 0.15, 0.02)), ((10, 25, 40), (55.5556, 55.5556, 55.5556)),
                                       ((4, 3, 512), (55.5556, 0.0, 55.5556)), ((5, 1, 64), (1.0, 1.0, 4096.0)),
                                       ((3, 2, 1), (2.0, 0.5, 3.0)), ((6, 4, 100), (0.0, 0.0, 1000.0)),
                                       ((7, 5, 33), (3.0, 2.0, 0.0))])
def test_z_recurrence_form_equals_fft_form(shape, lam):
    """The middle pass as a cyclic first-order recurrence (what csrc/zsolve.cuh computes) is the same operator as
    forward-z FFT, division by Diag, backward-z FFT (solve_3D, FftLinearSolver_3D.c:170-184)."""
    nx, ny, nz = shape
    rng = np.random.default_rng(nx * 100 + nz)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    a = O.FftTransportSolver(nx, ny, nz, *lam, b)
    r = O.FftTransportSolver_z_recurrence(nx, ny, nz, *lam, b)
    assert rel_l2(r, a) < 1e-12
    # and both invert the circulant matrix itself
    assert rel_l2(O.apply_transport_matrix(r, nx, ny, nz, *lam), b) < 1e-11
    with pytest.raises(ValueError):
        O.FftTransportSolver_z_recurrence(nx, ny, nz, lam[0], lam[1], -1.0, b)
    # the independent plain-C restatement of the same form
    rc = CO.transport_solve_z_recurrence(nx, ny, nz, *lam, b)
    assert rel_l2(rc, a) < 1e-12
    with pytest.raises(ValueError):
        CO.transport_solve_z_recurrence(nx, ny, nz, lam[0], lam[1], -1.0, b)


def test_oracle_properties_random_shapes():
    """Property checks over random small grids (hypothesis): the FFT form inverts the circulant matrix, is linear, and
    equals the recurrence form of the middle pass -- in both oracle implementations."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=40, deadline=None, derandomize=True)
    @hyp.given(nx=st.integers(1, 12), ny=st.integers(1, 9), nz=st.integers(1, 17),
               lx=st.floats(0.0, 60.0), ly=st.floats(0.0, 60.0), lz=st.floats(0.0, 4096.0), seed=st.integers(0, 2 ** 16))
    def check(nx, ny, nz, lx, ly, lz, seed):
        rng = np.random.default_rng(seed)
        n = nx * ny * nz
        x1 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        x2 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        b1 = O.apply_transport_matrix(x1, nx, ny, nz, lx, ly, lz)
        b2 = O.apply_transport_matrix(x2, nx, ny, nz, lx, ly, lz)
        s1 = O.FftTransportSolver(nx, ny, nz, lx, ly, lz, b1)
        tol = 1e-12 * (1.0 + lx + ly + lz)               # the forward error scales with the conditioning ~ 1 + 2 sum lambda
        assert rel_l2(s1, x1) < tol
        s12 = O.FftTransportSolver(nx, ny, nz, lx, ly, lz, 2.0 * b1 - 3.0j * b2)
        assert rel_l2(s12, 2.0 * x1 - 3.0j * x2) < tol
        assert rel_l2(O.FftTransportSolver_z_recurrence(nx, ny, nz, lx, ly, lz, b1), s1) < tol
        assert rel_l2(CO.transport_solve_z_recurrence(nx, ny, nz, lx, ly, lz, b1), s1) < tol

    check()


# the line form of the middle pass (one thread per z line, carry-in summed over the planes that can still matter) and the
# multi-rank owner scheme, restated in numpy: the truncation is below rounding, the slab closure is exact
@pytest.mark.parametrize("shape", [(6, 5, 64), (4, 4, 200), (8, 3, 96)])
@pytest.mark.parametrize("lam", [(55.5556, 55.5556, 55.5556), (0.6, 0.15, 3000.0), (2.0, 0.5, 0.0)])
@pytest.mark.parametrize("slabs", [1, 2, 4])
def test_z_line_form_equals_fft_form(shape, lam, slabs):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + ny + nz + slabs)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    got = O.FftTransportSolver_z_line_form(nx, ny, nz, *lam, b, slabs=slabs)
    assert rel_l2(got, want) < 1e-12
    got_c = CO.transport_solve_z_line_form(nx, ny, nz, *lam, b, slabs=slabs)       # the plain-C restatement
    # (at lambda_z = 3000 the closure 1/(1 - c^nz) amplifies rounding to ~3e-13 in either restatement)
    assert rel_l2(got_c, want) < 1e-12 and rel_l2(got_c, got) < 1e-12
    # the truncation itself: against the untruncated sum (weight_floor = 0) the difference is at rounding level
    full = O.FftTransportSolver_z_line_form(nx, ny, nz, *lam, b, weight_floor=1e-300, slabs=slabs)
    assert rel_l2(got, full) < 1e-15
    # a careless floor would be visible: the check has teeth
    sloppy = O.FftTransportSolver_z_line_form(nx, ny, nz, *lam, b, weight_floor=1e-3, slabs=slabs)
    if lam[0] > 50.0 and nz >= 96 and slabs == 1:
        assert rel_l2(sloppy, full) > 1e-9
