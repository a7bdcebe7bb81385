"""The harness's operators against the reference's OWN assembly code (BASELINE configs 2, 3 and 5).

oracle/_ref/libreference_assembly.so is /root/reference/src/TransportEquation.cxx and src/WaveSystem.cxx, unmodified, compiled
from where they lie against a stand-in for the SOLVERLAB mesh classes they use (oracle/solverlab_standin/: the mesh is plain
finite-volume connectivity handed in by the test).  computeDivergenceMatrix of both files -- the upwind choice and its
signs, jacobianMatrices, the interior / wall / periodic / Neumann cases, addValue's block placement -- therefore runs
exactly as the reference wrote it, and what the iteration-count tests iterate on must be that matrix:

  circulantpreconditioner_b200/krylov.py   transport_operator(ref_sign_quirk=True), wave_operator(walls / periodic)
  circulantpreconditioner_b200/meshes.py   transport_matrix(ref_sign_quirk=True) on the unstructured fixtures
  oracle/circulant_oracle.py               apply_wave_matrix, wave_jacobian_minus; solve_wave_block inverts the periodic matrix

Needs the built library (made where /root/reference exists; it travels to the GPU box).
"""
import numpy as np
import pytest
import torch

from circulantpreconditioner_b200 import krylov as K
from circulantpreconditioner_b200 import meshes as MS
from oracle import circulant_oracle as O
from oracle import ref_assembly as RA

pytestmark = pytest.mark.skipif(not RA.available(), reason="oracle/_ref/libreference_assembly.so is not built")


def _dense(op, n):
    eye = torch.eye(n, dtype=torch.complex128)
    cols = [op(eye[:, c].contiguous()) for c in range(n)]
    m = torch.stack(cols, dim=1).numpy()
    assert np.abs(m.imag).max() == 0.0
    return m.real


GRIDS = [((5, 4, 3), (0.2, 0.25, 0.5)), ((4, 4, 4), (0.25, 0.25, 0.25)), ((6, 1, 1), (0.1, 1.0, 1.0)), ((3, 5, 1), (0.5, 0.2, 1.0)),
         ((2, 2, 2), (0.5, 0.5, 0.5))]


@pytest.mark.parametrize("shape,h", GRIDS)
def test_transport_operator_is_the_references_matrix(shape, h):
    """src/TransportEquation.cxx:75-133 + MatShift(A, 1): interior faces only ("Neumann" borders do nothing).  The
    reference writes the inflow entry as -dt S/V un with un < 0, i.e. positive: ref_sign_quirk=True is
    the reference's matrix to rounding, the consistent upwind sign differs from it by 2 lambda on every inflow entry."""
    a, dt = (1.0, 0.5, 0.25), 0.3
    lam = tuple(a[d] * dt / h[d] for d in range(3))
    n = int(np.prod(shape))
    A = RA.assemble("transport", RA.cartesian_mesh(shape, h, RA.NEUMANN), dt, a)
    quirk = _dense(K.transport_operator(shape, lam, ref_sign_quirk=True), n)
    assert np.abs(A - quirk).max() < 1e-14
    consistent = _dense(K.transport_operator(shape, lam), n)
    off = ~np.eye(n, dtype=bool)
    assert np.allclose(consistent[off], -A[off], rtol=0, atol=1e-14) and np.allclose(np.diag(consistent), np.diag(A), atol=1e-14)
    assert np.all(A[off] >= 0) and A[off].max() > 0            # the quirk: positive off-diagonal entries


def test_spherical_explosion_config2_parameters():
    """Config 2's own numbers (tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:258-261): a = (1, 0, 0),
    dt = cfl dx / 6 with cfl = 1e3 / 3 on the unit cube -> lambda_x = 55.5556; at 8^3 here."""
    n = 8
    h = 1.0 / n
    dt = (1e3 / 3.0) * (h / 6.0)
    A = RA.assemble("transport", RA.cartesian_mesh((n, n, n), (h, h, h), RA.NEUMANN), dt, (1.0, 0.0, 0.0))
    lam = dt / h
    assert abs(lam - 55.5555555) < 1e-6
    got = _dense(K.transport_operator((n, n, n), (lam, 0.0, 0.0), ref_sign_quirk=True), n ** 3)
    assert np.abs(A - got).max() < 1e-13 * lam


@pytest.mark.parametrize("name", ["hexa_3", "kershaw1"])
def test_unstructured_transport_matrix_is_the_references(name):
    """meshes.transport_matrix on the finite-volume fixtures of the reference's own meshes (config 5)."""
    mesh = MS.load_fixture(name)
    a = (1.0, 0.0, 0.0)
    dt = MS.reference_dt(mesh, a)
    A = RA.assemble("transport", RA.fixture_mesh(mesh), dt, a)
    got = MS.transport_matrix(mesh, a, dt, ref_sign_quirk=True).toarray()
    assert np.abs(A - got).max() < 1e-12 * np.abs(A).max()
    a2 = (0.3, -0.7, 0.5)                                          # a general direction: every face is an in- or outflow face
    A2 = RA.assemble("transport", RA.fixture_mesh(mesh), dt, a2)
    got2 = MS.transport_matrix(mesh, a2, dt, ref_sign_quirk=True).toarray()
    assert np.abs(A2 - got2).max() < 1e-12 * np.abs(A2).max()


FULL_3D = [g for g in GRIDS if min(g[0]) > 1]


@pytest.mark.parametrize("shape,h", FULL_3D)
@pytest.mark.parametrize("c0", [700.0, 3.0])
def test_wave_operator_is_the_references_matrix(shape, h, c0):
    """src/WaveSystem.cxx:109-176 + MatShift(A, 1), unknowns [p, q_x, q_y, q_z] per cell: wall borders (the default of the
    reference's driver) and periodic borders."""
    dt = 0.3 if c0 < 10 else 55.5556 * min(h) / c0
    mu = tuple(dt / h[d] for d in range(3))
    n = 4 * int(np.prod(shape))
    for border, periodic in ((RA.WALL, False), (RA.PERIODIC, True)):
        A = RA.assemble("wave", RA.cartesian_mesh(shape, h, border), dt, c0=c0)
        got = _dense(K.wave_operator(shape, c0, mu, periodic=periodic), n)
        assert np.abs(A - got).max() <= 1e-14 * np.abs(A).max()
    # the oracle's periodic operator, and its block-circulant solve: the symbol inverts the reference's own matrix
    rng = np.random.default_rng(5)
    y = rng.standard_normal(n)
    assert np.allclose(O.apply_wave_matrix(y, *shape, c0, *mu), A @ y, rtol=1e-13, atol=1e-13 * np.abs(A).max())
    b = (A @ y).astype(np.complex128)
    x = O.solve_wave_block(b, *shape, c0, *mu)
    assert np.linalg.norm(x - y) / np.linalg.norm(y) < 1e-10


@pytest.mark.parametrize("shape,h", [g for g in GRIDS if min(g[0]) == 1])
def test_wave_operator_degenerate_axes_periodic(shape, h):
    """An axis of one cell: the harness, the oracle and the GPU plan treat it as absent (4 unknowns per cell, nothing from
    that direction).  On a periodic one-layer mesh the reference's assembly agrees -- the neighbour across the layer is the
    cell itself, +Am and -Am cancel.  (With walls a one-layer 3-D mesh would add wall terms on the layer's faces; the
    reference runs lower-dimensional problems on lower-dimensional meshes instead, dim + 1 unknowns per cell.)"""
    c0, dt = 3.0, 0.3
    mu = tuple(dt / h[d] for d in range(3))
    n = 4 * int(np.prod(shape))
    A = RA.assemble("wave", RA.cartesian_mesh(shape, h, RA.PERIODIC), dt, c0=c0)
    got = _dense(K.wave_operator(shape, c0, mu, periodic=True), n)
    assert np.abs(A - got).max() <= 1e-14 * np.abs(A).max()
    y = np.random.default_rng(6).standard_normal(n)
    assert np.allclose(O.apply_wave_matrix(y, *shape, c0, *mu), A @ y, rtol=1e-13, atol=1e-13 * np.abs(A).max())


def test_neumann_borders_of_the_wave_system_do_nothing():
    shape, h, c0, dt = (3, 3, 2), (0.5, 0.5, 0.5), 3.0, 0.2
    A_neu = RA.assemble("wave", RA.cartesian_mesh(shape, h, RA.NEUMANN), dt, c0=c0)
    A_wall = RA.assemble("wave", RA.cartesian_mesh(shape, h, RA.WALL), dt, c0=c0)
    assert np.abs(A_neu - A_wall).max() > 0.1                      # the wall term exists ...
    d = A_wall - A_neu                                             # ... and sits in the diagonal blocks of border cells only
    nc = int(np.prod(shape))
    for j in range(nc):
        for k in range(nc):
            if j != k:
                assert np.all(d[4 * j:4 * j + 4, 4 * k:4 * k + 4] == 0)


@pytest.mark.parametrize("normal", [(1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (0.6, 0.0, 0.8), (1.0,), (0.0, 1.0)])
def test_jacobian_matrices(normal):
    """jacobianMatrices (src/WaveSystem.cxx:92-107) against the oracle's restatement, any dimension and direction."""
    got = RA.jacobian_minus(normal, 0.37, 700.0)
    assert np.allclose(got, O.wave_jacobian_minus(list(normal), 0.37, 700.0), rtol=1e-15, atol=0)


def test_transport_circulant_model_vs_the_references_matrix():
    """What the preconditioner is to the reference's matrix: C = I + sum_d lambda_d (I - S_d) (the symbol of
    build_diag_mat_vec_3D, src/FftLinearSolver_3D.c:136-164) has the CONSISTENT upwind sign and periodic closure; it
    agrees with the reference's assembled matrix on the diagonal of every cell that has an outflow neighbour and is minus
    its off-diagonal part (F11) -- which is why the quirk runs need hundreds of iterations and the consistent ones two."""
    shape, h, a, dt = (4, 3, 2), (0.25, 0.5, 0.5), (1.0, 0.5, 0.25), 0.3
    lam = tuple(a[d] * dt / h[d] for d in range(3))
    n = int(np.prod(shape))
    A = RA.assemble("transport", RA.cartesian_mesh(shape, h, RA.NEUMANN), dt, a)
    C = O.dense_transport_matrix(*shape, *lam)
    interior = np.zeros(shape[::-1], dtype=bool)
    interior[:-1, :-1, :-1] = True                                 # cells whose three outflow faces are interior faces
    idx = np.flatnonzero(interior.ravel())
    assert np.allclose(np.diag(C)[idx], np.diag(A)[idx], atol=1e-14)
    mask = (A != 0) & ~np.eye(n, dtype=bool)
    assert np.allclose(C[mask], -A[mask], atol=1e-14)


def test_spherical_explosion_initial_data():
    """initial_conditions_shock of both reference files against the three restatements the tests and bench.py use for b
    (oracle.spherical_step, krylov.spherical_step, meshes.spherical_step)."""
    n = 16
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    cen = np.stack([(i.ravel() + 0.5) / n - 0.5, (j.ravel() + 0.5) / n - 0.5, (k.ravel() + 0.5) / n - 0.5], axis=1)
    lo, hi = (-0.5,) * 3, (0.5,) * 3
    T = RA.initial_conditions("transport", cen, lo, hi)
    assert set(np.unique(T)) == {600.0, 650.0}
    assert np.array_equal(T, O.spherical_step(n, n, n, 650.0, 600.0))
    assert np.array_equal(T, K.spherical_step((n, n, n), 650.0, 600.0).numpy())
    W = RA.initial_conditions("wave", cen, lo, hi).reshape(-1, 4)
    assert np.array_equal(W[:, 0], O.spherical_step(n, n, n, 155e5, 70e5)) and np.all(W[:, 1:] == 0.0)
    for name in ("kershaw1", "hexa_3"):
        m = MS.load_fixture(name)
        assert np.array_equal(RA.initial_conditions("transport", m["centre"], m["bbox"][0], m["bbox"][1]), MS.spherical_step(m))


@pytest.mark.gpu
@pytest.mark.parametrize("c0,tol", [(3.0, 1e-11), (700.0, 1e-9)])
def test_cuda_wave_block_inverts_the_references_periodic_matrix(c0, tol):
    """The GPU's fused 4 x 4 block solve against the reference's own assembly: b := A y with A = the matrix
    src/WaveSystem.cxx assembles on a periodic grid (+ MatShift), then cpc_apply(b) must return y.  (The tolerance is the
    block conditioning ~ c0^2 mu times rounding, as in tests/test_gpu_parity.py::test_wave_block_matches_oracle.)"""
    import circulantpreconditioner_b200 as cpc
    shape, h = (6, 5, 4), (0.25, 0.2, 0.5)
    dt = 0.3 if c0 < 10 else 55.5556 * min(h) / c0
    mu = tuple(dt / h[d] for d in range(3))
    A = RA.assemble("wave", RA.cartesian_mesh(shape, h, RA.PERIODIC), dt, c0=c0)
    y = np.random.default_rng(21).standard_normal(A.shape[0])
    b = (A @ y).astype(np.complex128)
    with cpc.CirculantPlan(*shape, ncomp=4) as p:
        p.set_symbol_wave(c0, *mu)
        got = p.apply(torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.linalg.norm(got - O.solve_wave_block(b, *shape, c0, *mu)) / np.linalg.norm(y) < 1e-12 * max(1.0, c0)
    assert np.linalg.norm(got.real - y) / np.linalg.norm(y) < tol
    assert np.abs(got.imag).max() < tol * np.abs(y).max()
