"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerance (BASELINE.json north_star): relative L2 error <= 1e-12 in fp64, <= 1e-5 in fp32.
Nothing here reads /root/reference; golden inputs/outputs come from tests/golden/*.npz.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import circulantpreconditioner_b200 as cpc
from oracle import circulant_oracle as O
from tests.conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5


def dev(a, dtype=torch.complex128):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def host(t):
    return t.detach().cpu().numpy()


def g(name):
    return np.load(os.path.join(GOLDEN, name))


def rand_c(rng, n):
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


# ----------------------------------------------------------------------------------------------------
# golden vectors of the reference (generic any-n kernel path: 4, 8, 3x2, 4x3x2, 50x200, 10x25x40)
# ----------------------------------------------------------------------------------------------------
def test_kat1_first_column_n4():
    f = g("ref_c_kat1_n4.npz")
    with cpc.CirculantPlan(4, 1, 1) as p:
        p.set_symbol_first_column(dev(f["col"]))
        x = host(p.apply(dev(f["b"])))
    np.testing.assert_allclose(x.real, [6.7, 2.9, 6.3, 20.1], rtol=1e-13)
    np.testing.assert_allclose(x.imag, 0, atol=1e-13)
    assert rel_l2(x, f["x"]) < TOL64


@pytest.mark.parametrize("name", ["ref_c_kat2_3x2.npz", "ref_c_kat3_4x3x2.npz", "ref_py_2d_50x200.npz",
                                  "ref_py_3d_10x25x40.npz"])
def test_reference_fixtures_transport(name):
    f = g(name)
    nx, ny, nz = (int(v) for v in f["n"])
    lam = [float(v) for v in f["lam"]]
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        assert np.abs(p.get_diag() - f["Diag"]).max() < 1e-13
        b = dev(f["b"])
        x = host(p.apply(b))
        assert rel_l2(x, f["X"]) < TOL64
        assert rel_l2(x, f["X_ref"]) < TOL64
        # same answer when the caller hands over Diag itself (solve_3D's signature) ...
        p.set_symbol_diag(dev(f["Diag"]))
        assert rel_l2(host(p.apply(b)), f["X"]) < TOL64
        # ... or the three 1-D eigenvalue tables (build_diag_mat_vec_3D's signature)
        p.set_symbol_separable(O.column_hat(nx), O.column_hat(ny), O.column_hat(nz), *lam)
        assert rel_l2(host(p.apply(b)), f["X"]) < TOL64


def test_reference_fixture_1d_n8():
    f = g("ref_py_1d_n8.npz")
    with cpc.CirculantPlan(8) as p:
        p.set_symbol_first_column(dev(f["col"]))
        assert rel_l2(host(p.apply(dev(f["b"]))), f["x"]) < TOL64
        p.set_symbol_transport(float(f["lam"]))
        assert rel_l2(host(p.apply(dev(f["b"]))), f["x"]) < TOL64


@pytest.mark.parametrize("tag", ["phys", "unit"])
def test_config0_32cube_fast_path(tag):
    f = g(f"ref_py_3d_32cube_{tag}.npz")
    lam = [float(v) for v in f["lam"]]
    with cpc.CirculantPlan(32, 32, 32) as p:
        assert p.info()["fast_path"] == [1, 1, 1]
        p.set_symbol_transport(*lam)
        assert np.abs(p.get_diag()[:64] - f["Diag_head"]).max() < 1e-13
        x = host(p.apply(dev(f["b"])))
    assert rel_l2(x.real, f["X_real"]) < TOL64
    assert np.abs(x.imag).max() < 1e-11
    assert rel_l2(x, f["X_ref"]) < TOL64


# ----------------------------------------------------------------------------------------------------
# seeded comparisons with the oracle
# ----------------------------------------------------------------------------------------------------
SHAPES = [(16, 16, 16), (32, 64, 16), (64, 16, 128), (128, 32, 16), (256, 16, 32), (16, 256, 16), (512, 16, 8),
          (16, 8, 512), (1024, 4, 4), (8, 1024, 2), (2048, 2, 2), (2, 4, 2048), (48, 20, 36), (7, 11, 13),
          (64, 64, 1), (128, 1, 1), (1, 1, 64), (1, 1, 1), (30, 1, 17),
          # fast kernels with partial tiles (lines not a multiple of the tile width) and mixed fast / generic axes
          (64, 3, 5), (128, 5, 3), (512, 3, 1), (12, 64, 5), (3, 5, 256), (256, 7, 32),
          # 512-point y / z lines take the 2 x (16 x 16) kernel (its root table must be the transformed axis' own)
          (16, 512, 32), (8, 512, 512), (24, 512, 6),
          (16, 1024, 8), (8, 4, 1024),
          # 2^a * 3 line lengths: radix-6 / radix-12 last stage (prime-factor butterflies), fast kernels on every axis
          (48, 96, 192), (384, 6, 48), (10, 768, 3), (5, 3, 384), (96, 768, 2), (192, 5, 96), (768, 2, 3),
          # generic kernel: radix-4 grouping, odd primes (the reference's own 10 x 25 x 40 grid, testFftSolver_3D.py:82)
          (10, 25, 40), (100, 10, 25), (250, 6, 7), (36, 45, 14),
          # prime factors >= 11 (O(R) sum stages, first / last / only stage) and long generic lines (4 lanes per tile)
          (22, 39, 34), (101, 2, 3), (3, 640, 2), (1536, 2, 2),
          # 2^a * 5^b line lengths: radix 5 / 10 / 20 butterflies, non-power-of-two stage periods (k = jb mod P)
          (100, 200, 160), (250, 8, 400), (320, 3, 500), (500, 100, 2), (800, 2, 250), (2, 3, 1000), (1000, 2, 800),
          (6, 800, 5), (400, 320, 3)]


@pytest.mark.parametrize("shape", SHAPES)
def test_forward_inverse_match_oracle(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx * 7919 + ny * 31 + nz)
    v = rand_c(rng, nx * ny * nz)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        d = dev(v)
        f = host(p.forward(d))
        ref = O.fft3_forward(v.reshape(nz, ny, nx)).ravel()
        assert rel_l2(f, ref) < TOL64
        bk = host(p.inverse(dev(ref)))
        assert rel_l2(bk, v * v.size) < TOL64
        assert np.array_equal(host(d), v)           # input untouched


@pytest.mark.parametrize("shape", SHAPES)
def test_apply_transport_matches_oracle(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + 13 * ny + 101 * nz)
    lam = (55.5556, 0.3, 2.5)
    x_ref = rand_c(rng, nx * ny * nz)
    b = O.apply_transport_matrix(x_ref, nx, ny, nz, *lam)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        d = dev(b)
        out = torch.empty_like(d)
        p.apply(d, out)
        assert rel_l2(host(out), want) < TOL64
        assert rel_l2(host(out), x_ref) < 1e-11
        assert np.array_equal(host(d), b)           # b is read-only
        p.apply(d, d)                               # b == x aliasing (TransportEquationFFT...:111 passes Un, Un)
        assert rel_l2(host(d), want) < TOL64


# the middle pass of a transport symbol: cyclic first-order recurrence along z (csrc/zsolve.cuh) against the
# forward-FFT / divide / backward-FFT form of the reference (solve_3D, src/FftLinearSolver_3D.c:170-184)
@pytest.mark.parametrize("shape", [(64, 32, 512), (24, 16, 384), (8, 8, 96), (8, 8, 100), (4, 4, 1000), (16, 16, 16),
                                   (40, 3, 1024), (8, 5, 200), (12, 7, 64)])
def test_middle_pass_recurrence_matches_fft_form(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + 3 * ny + 7 * nz)
    lam = (55.5556, 55.5556, 55.5556)
    b = rand_c(rng, nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        assert p.info()["fast_path"][2] == 2
        rec = host(p.apply(dev(b)))
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_option("z_recurrence", 0)
        p.set_symbol_transport(*lam)
        assert p.info()["fast_path"][2] != 2
        fft = host(p.apply(dev(b)))
    assert rel_l2(rec, want) < TOL64
    assert rel_l2(fft, want) < TOL64
    assert rel_l2(rec, fft) < TOL64


@pytest.mark.parametrize("shape", [(64, 32, 512), (24, 16, 384), (8, 8, 100), (40, 3, 1024), (8, 5, 200), (12, 7, 63), (16, 16, 2)])
@pytest.mark.parametrize("lam", [(55.5556, 55.5556, 55.5556), (0.6, 0.15, 3000.0)])
def test_middle_pass_line_form(shape, lam):
    """The thread-per-line form of the recurrence (carry-in from the planes that can still matter, then one thread per
    line along z; any nz) against the oracle and the tile-kernel / FFT forms; lambda_z = 3000 makes every plane matter."""
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + 3 * ny + 7 * nz)
    b = rand_c(rng, nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        p.set_option("z_line_form", 1)
        assert p.info()["fast_path"][2] == 2
        line = host(p.apply(dev(b)))
        x = dev(b)
        p.apply(x, x)                                    # in place
        p.set_option("z_line_form", 0)
        other = host(p.apply(dev(b)))
    assert rel_l2(line, want) < TOL64
    assert rel_l2(host(x), want) < TOL64
    assert rel_l2(line, other) < TOL64


def test_middle_pass_recurrence_gating():
    nx, ny, nz = 32, 16, 64
    rng = np.random.default_rng(5)
    b = rand_c(rng, nx * ny * nz)
    col_hat = [np.fft.fft(O.build_transport_col(n)) for n in (nx, ny, nz)]
    with cpc.CirculantPlan(nx, ny, nz) as p:
        # the same tables through the separable entry point (what Fft3DSolver does): recognised
        p.set_symbol_separable(*col_hat, 2.0, 0.5, 3.0)
        assert p.info()["fast_path"][2] == 2
        assert rel_l2(host(p.apply(dev(b))), O.FftTransportSolver(nx, ny, nz, 2.0, 0.5, 3.0, b)) < TOL64
        # large lambda_z, still inside the gate
        p.set_symbol_transport(1.0, 1.0, 4000.0)
        assert p.info()["fast_path"][2] == 2
        assert rel_l2(host(p.apply(dev(b))), O.FftTransportSolver(nx, ny, nz, 1.0, 1.0, 4000.0, b)) < TOL64
        # outside: negative or huge lambda_z, a negative lambda_x, a z table that is not the upwind column, other symbols
        # (lambda_x = -0.3 lets Re(alpha) drop to 0.4, below the 1/2 the recurrence asks for; -0.2 would still qualify)
        for lam in ((1.0, 1.0, -0.2), (1.0, 1.0, 1e5), (-0.3, 1.0, 1.0)):
            p.set_symbol_transport(*lam)
            assert p.info()["fast_path"][2] != 2
            assert rel_l2(host(p.apply(dev(b))), O.FftTransportSolver(nx, ny, nz, *lam, b)) < 1e-11
        cz = col_hat[2].copy()
        cz[3] += 0.25
        p.set_symbol_separable(col_hat[0], col_hat[1], cz, 2.0, 0.5, 3.0)
        assert p.info()["fast_path"][2] != 2
        diag = O.build_diag_mat_vec_3D(col_hat[0], col_hat[1], cz, nx, ny, nz, 2.0, 0.5, 3.0)
        assert rel_l2(host(p.apply(dev(b))), O.solve_3D(diag, b, nx, ny, nz)) < TOL64
        p.set_symbol_diag(diag)
        assert p.info()["fast_path"][2] != 2
        p.set_symbol_transport(2.0, 0.5, 3.0)
        assert p.info()["fast_path"][2] == 2


def test_general_first_column_and_diag_table():
    nx, ny, nz = 32, 16, 64
    rng = np.random.default_rng(3)
    col = np.zeros((nz, ny, nx), dtype=np.complex128)
    col[0, 0, 0] = 7.0
    col[0, 0, 1] = -1.0; col[0, 1, 0] = -1.5; col[1, 0, 0] = -0.5
    col[0, 0, -1] = -0.25; col[-1, 0, 0] = -0.75; col[0, -1, 0] = 0.3j
    b = rand_c(rng, nx * ny * nz)
    want = O.solve_first_column(col.ravel(), b, nx, ny, nz)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_first_column(dev(col.ravel()))
        assert rel_l2(host(p.apply(dev(b))), want) < TOL64
        lam = O.fft3_forward(col).ravel()
        assert rel_l2(p.get_diag(), lam) < TOL64
        p.set_symbol_diag(lam)                       # host pointer
        assert rel_l2(host(p.apply(dev(b))), want) < TOL64


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 16, 64), (64, 32, 16), (6, 5, 4), (16, 16, 1), (128, 16, 32),
                                   (32, 3, 16), (5, 16, 64), (256, 2, 3), (96, 48, 6), (192, 3, 48), (384, 2, 3),
                                   (10, 25, 12)])
def test_wave_block_matches_oracle(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(17)
    c0, mu = 700.0, (0.0793651, 0.0793651, 0.0793651)
    y = rng.standard_normal(nx * ny * nz * 4)
    b = O.apply_wave_matrix(y, nx, ny, nz, c0, *mu).astype(np.complex128)
    want = O.solve_wave_block(b, nx, ny, nz, c0, *mu)
    with cpc.CirculantPlan(nx, ny, nz, ncomp=4) as p:
        p.set_symbol_wave(c0, *mu)
        got = host(p.apply(dev(b)))
    # the two evaluations of the same closed form agree to rounding times the block conditioning (~c0^2 mu)
    assert rel_l2(got, want) < TOL64            # c0 = 700 (BASELINE config 3): measured 7e-14
    c0 = 3.0
    b = O.apply_wave_matrix(y, nx, ny, nz, c0, *mu).astype(np.complex128)
    with cpc.CirculantPlan(nx, ny, nz, ncomp=4) as p:
        p.set_symbol_wave(c0, *mu)
        got = host(p.apply(dev(b)))
    assert rel_l2(got, O.solve_wave_block(b, nx, ny, nz, c0, *mu)) < TOL64
    assert rel_l2(got.real, y) < 1e-11


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 128, 16), (512, 8, 16), (20, 12, 9), (96, 48, 192),
                                   (384, 12, 768), (768, 3, 384), (100, 10, 25), (100, 200, 160), (250, 400, 3),
                                   (320, 3, 500), (1000, 2, 800), (800, 1000, 2)])
def test_fp32_option(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(23)
    lam = (5.0, 0.3, 2.5)
    x_ref = rand_c(rng, nx * ny * nz)
    b = O.apply_transport_matrix(x_ref, nx, ny, nz, *lam)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz, dtype="c64") as p:
        p.set_symbol_transport(*lam)
        got = host(p.apply(dev(b, torch.complex64)))
        assert rel_l2(got, want) < TOL32
        f = host(p.forward(dev(b, torch.complex64)))
        assert rel_l2(f, O.fft3_forward(b.reshape(nz, ny, nx)).ravel()) < TOL32


def test_host_pointer_path_and_counters():
    nx, ny, nz = 64, 32, 48
    rng = np.random.default_rng(29)
    lam = (1.0, 2.0, 3.0)
    b = rand_c(rng, nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        x = np.empty_like(b)
        p.apply(b, x)                                # numpy (pageable host) pointers
        assert rel_l2(x, want) < TOL64
        bt = torch.from_numpy(b).pin_memory()
        xt = torch.empty_like(bt).pin_memory()
        p.apply(bt, xt)                              # pinned host pointers
        assert rel_l2(xt.numpy(), want) < TOL64
        inf = p.info()
        assert inf["h2d_bytes"] == 2 * b.nbytes and inf["d2h_bytes"] == 2 * b.nbytes
        assert inf["kernel_launches"] >= 10 and inf["passes_per_apply"] == 5
        assert inf["bytes_per_apply_alg"] == 160 * b.size
        f = np.empty_like(b)
        p.forward(b, f)
        assert rel_l2(f, O.fft3_forward(b.reshape(nz, ny, nx)).ravel()) < TOL64


def test_error_behaviour_on_device():
    with cpc.CirculantPlan(16, 16, 16) as p:
        b = torch.zeros(16 ** 3, dtype=torch.complex128, device="cuda")
        with pytest.raises(cpc.CpcError) as e:
            p.apply(b)                               # no symbol yet
        assert e.value.status == 4
        with pytest.raises(cpc.CpcError):
            p.set_symbol_wave(700.0, 1, 1, 1)        # ncomp == 1
    with cpc.CirculantPlan(16, 16, 16, ncomp=4) as p:
        with pytest.raises(cpc.CpcError):
            p.set_symbol_transport(1, 1, 1)
    with cpc.CirculantPlan(16, 16, 16, dtype="c64") as p:
        p.set_symbol_transport(1, 1, 1)
        buf = torch.zeros(16 ** 3 + 1, dtype=torch.complex64, device="cuda")
        with pytest.raises(cpc.CpcError) as e:
            p.apply(buf[1:], buf[1:])                # 8-byte aligned only: rejected, not a misaligned-address fault
        assert e.value.status == 1


# ----------------------------------------------------------------------------------------------------
# size-independent properties at the BASELINE sizes (the oracle would take too long here)
# ----------------------------------------------------------------------------------------------------
def _transport_matrix_torch(u, lam):
    out = u.clone()
    for ax, l in zip((2, 1, 0), lam):
        if u.shape[ax] > 1:
            out += l * (u - torch.roll(u, 1, dims=ax))
    return out


@pytest.mark.parametrize("n", [128, 256, 512])
def test_roundtrip_and_linearity_at_baseline_sizes(n):
    lam = (55.5556, 55.5556, 55.5556)
    gen = torch.Generator(device="cuda").manual_seed(n)
    xr = torch.randn(n, n, n, dtype=torch.float64, device="cuda", generator=gen).to(torch.complex128)
    b = _transport_matrix_torch(xr, lam).reshape(-1)
    with cpc.CirculantPlan(n, n, n) as p:
        assert p.info()["fast_path"] == [1, 1, 1]
        p.set_symbol_transport(*lam)
        x = p.apply(b)
        err = (torch.linalg.vector_norm(x - xr.reshape(-1)) / torch.linalg.vector_norm(xr)).item()
        assert err < TOL64, err                      # b := C x_ref  =>  apply(b) == x_ref
        # forward / inverse round trip and Parseval
        f = p.forward(b)
        bb = p.inverse(f)
        N = float(n) ** 3
        assert (torch.linalg.vector_norm(bb / N - b) / torch.linalg.vector_norm(b)).item() < TOL64
        pars = (torch.linalg.vector_norm(f) ** 2 / N / torch.linalg.vector_norm(b) ** 2).item()
        assert abs(pars - 1.0) < 1e-12
        del f, bb
        # linearity: apply(2 b + i b) == (2 + i) apply(b)
        y = p.apply((2.0 + 1.0j) * b)
        assert (torch.linalg.vector_norm(y - (2.0 + 1.0j) * x) / torch.linalg.vector_norm(x)).item() < TOL64


# ----------------------------------------------------------------------------------------------------
# real-scalar plans (CPC_F64 / CPC_F32): r2c / c2r inside, half the bytes
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 32, 16), (512, 8, 16), (128, 64, 32), (256, 16, 1), (1024, 4, 2), (64, 1, 1),
                                   (20, 12, 9), (30, 1, 17), (16, 16, 16), (768, 6, 48), (384, 12, 5), (96, 4, 4),
                                   (192, 4, 4), (100, 6, 4), (200, 100, 3), (400, 3, 160), (1000, 4, 3), (500, 2, 2),
                                   (320, 5, 2), (800, 2, 2)])
def test_real_scalar_plan_matches_oracle(shape):
    nx, ny, nz = shape
    rng = np.random.default_rng(nx + ny + nz)
    lam = (55.5556, 0.3, 2.5)
    x_ref = rng.standard_normal(nx * ny * nz)
    b = O.apply_transport_matrix(x_ref, nx, ny, nz, *lam)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b.astype(np.complex128))
    assert np.abs(want.imag).max() < 1e-9
    with cpc.CirculantPlan(nx, ny, nz, dtype="f64") as p:
        p.set_symbol_transport(*lam)
        d = torch.from_numpy(b).cuda()
        out = torch.empty_like(d)
        p.apply(d, out)
        assert rel_l2(host(out), want.real) < TOL64
        assert np.array_equal(host(d), b)
        p.apply(d, d)                                  # in place
        assert rel_l2(host(d), want.real) < TOL64
        xh = np.empty_like(b)
        p.apply(b, xh)                                 # host pointers
        assert rel_l2(xh, want.real) < TOL64
        assert p.info()["bytes_per_apply_alg"] == 80 * b.size
        with pytest.raises(cpc.CpcError):
            p.forward(d)
    with cpc.CirculantPlan(nx, ny, nz, dtype="f32") as p:
        p.set_symbol_transport(*lam)
        out = p.apply(torch.from_numpy(b.astype(np.float32)).cuda())
        assert rel_l2(host(out), want.real) < TOL32


def test_real_scalar_plan_512cube_roundtrip():
    n = 512
    lam = (55.5556, 55.5556, 55.5556)
    gen = torch.Generator(device="cuda").manual_seed(7)
    xr = torch.randn(n, n, n, dtype=torch.float64, device="cuda", generator=gen)
    b = _transport_matrix_torch(xr, lam).reshape(-1)
    with cpc.CirculantPlan(n, n, n, dtype="f64") as p:
        p.set_symbol_transport(*lam)
        x = p.apply(b)
    err = (torch.linalg.vector_norm(x - xr.reshape(-1)) / torch.linalg.vector_norm(xr)).item()
    assert err < TOL64, err


def test_projected_apply_unstructured_to_cartesian():
    # applyFFT3DPrecTransport with an intersection matrix (reference PCSHELLFft_3D.cxx:17-21): x = P^T solve_3D(P b)
    import scipy.sparse as sp
    nx, ny, nz = 16, 8, 32
    N, M = nx * ny * nz, 3000                      # Cartesian cells, cells of the "unstructured" mesh
    rng = np.random.default_rng(41)
    P = sp.random(N, M, density=4.0 / M, random_state=np.random.RandomState(3), format="csr", dtype=np.float64)
    P.data[:] = rng.random(P.nnz)
    lam = (2.0, 0.5, 1.5)
    b = rand_c(rng, M)
    want = P.T @ O.FftTransportSolver(nx, ny, nz, *lam, P @ b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        with pytest.raises(cpc.CpcError):
            p.apply_projected(dev(b))              # no projection yet
        p.set_projection(M, P.indptr, P.indices, P.data)
        got = host(p.apply_projected(dev(b)))
        assert rel_l2(got, want) < TOL64
        xh = np.empty_like(b)
        p.apply_projected(b, xh)                   # host pointers
        assert rel_l2(xh, want) < TOL64


def test_wave_block_fp32():
    nx, ny, nz = 32, 16, 64
    rng = np.random.default_rng(19)
    c0, mu = 3.0, (0.0793651, 0.05, 0.03)
    y = rng.standard_normal(nx * ny * nz * 4)
    b = O.apply_wave_matrix(y, nx, ny, nz, c0, *mu).astype(np.complex128)
    want = O.solve_wave_block(b, nx, ny, nz, c0, *mu)
    with cpc.CirculantPlan(nx, ny, nz, ncomp=4, dtype="c64") as p:
        p.set_symbol_wave(c0, *mu)
        got = host(p.apply(dev(b, torch.complex64)))
    assert rel_l2(got, want) < TOL32


def test_two_plans_and_streams_do_not_interfere():
    # two plans on two torch streams, interleaved applies: per-plan state only
    n = 64
    rng = np.random.default_rng(31)
    lam1, lam2 = (1.0, 2.0, 3.0), (55.5556, 0.0, 0.0)
    b = rand_c(rng, n ** 3)
    w1 = O.FftTransportSolver(n, n, n, *lam1, b)
    w2 = O.FftTransportSolver(n, n, n, *lam2, b)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    d = dev(b)
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        p1 = cpc.CirculantPlan(n, n, n)
        p1.set_symbol_transport(*lam1)
    with torch.cuda.stream(s2):
        p2 = cpc.CirculantPlan(n, n, n)
        p2.set_symbol_transport(*lam2)
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            o1 = p1.apply(d)
        with torch.cuda.stream(s2):
            o2 = p2.apply(d)
        outs.append((o1, o2))
    torch.cuda.synchronize()
    for o1, o2 in outs:
        assert rel_l2(host(o1), w1) < TOL64 and rel_l2(host(o2), w2) < TOL64
    p1.destroy(); p2.destroy()


def test_largest_sweep_size_1024cube_roundtrip():
    # BASELINE config 4's largest grid: 2^30 points, 17 GB per array -- exercises the 64-bit addressing.
    n = 1024
    free, _total = torch.cuda.mem_get_info()
    if free < 100 * 2 ** 30:
        pytest.skip("needs ~100 GB of free HBM")
    lam = (55.5556, 55.5556, 55.5556)
    gen = torch.Generator(device="cuda").manual_seed(11)
    xr = torch.randn(n, n, n, dtype=torch.float64, device="cuda", generator=gen).to(torch.complex128)
    b = xr.clone()
    for ax, l in zip((2, 1, 0), lam):                 # b = C x_ref, one axis at a time to bound temporaries
        b += l * xr
        b -= l * torch.roll(xr, 1, dims=ax)
    b = b.reshape(-1)
    with cpc.CirculantPlan(n, n, n) as p:
        assert p.info()["fast_path"] == [1, 1, 1]
        p.set_symbol_transport(*lam)
        p.apply(b, b)                                 # in place
    b -= xr.reshape(-1)
    err = (torch.linalg.vector_norm(b) / torch.linalg.vector_norm(xr)).item()
    assert err < TOL64, err


# ---- ADVICE (round 1): the recurrence middle pass on fp32 storage.  Its arithmetic is fp64 whatever the storage type
# (zsolve.cuh), so complex64 plans must match the oracle as well as the FFT form does, also at large lambda_z ----
@pytest.mark.parametrize("lam", [(55.5556, 55.5556, 55.5556), (1.0, 1.0, 1000.0), (1.0, 1.0, 4000.0)])
def test_fp32_recurrence_keeps_fp32_accuracy(lam):
    nx, ny, nz = 64, 32, 512
    rng = np.random.default_rng(77)
    b = rand_c(rng, nx * ny * nz).astype(np.complex64)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b.astype(np.complex128))
    with cpc.CirculantPlan(nx, ny, nz, dtype="c64") as p:
        p.set_symbol_transport(*lam)
        assert p.info()["fast_path"][2] == 2
        rec = host(p.apply(dev(b, torch.complex64)))
        p.set_option("z_recurrence", 0)
        fft = host(p.apply(dev(b, torch.complex64)))
    e_rec, e_fft = rel_l2(rec, want), rel_l2(fft, want)
    assert e_rec < 2e-6 and e_fft < TOL32, (e_rec, e_fft)
    assert e_rec < 4 * e_fft + 2e-7, (e_rec, e_fft)       # no digits lost against the FFT form


# ---- ADVICE (round 1): the Python binding checks dtype, element count and device before handing raw addresses over ----
def test_python_binding_rejects_mismatched_arrays():
    with cpc.CirculantPlan(16, 8, 4) as p:
        p.set_symbol_transport(1.0, 1.0, 1.0)
        good = torch.zeros(16 * 8 * 4, dtype=torch.complex128, device="cuda")
        with pytest.raises(ValueError):
            p.apply(good.to(torch.complex64), good.clone())
        with pytest.raises(ValueError):
            p.apply(good[:100].contiguous(), good.clone())
        with pytest.raises(ValueError):
            p.apply(torch.zeros(16 * 8 * 4, dtype=torch.float64, device="cuda"), good.clone())
        with pytest.raises(ValueError):
            p.set_symbol_diag(np.zeros(10, dtype=np.complex128))
        with pytest.raises(ValueError):
            p.set_symbol_first_column(np.zeros(16 * 8 * 4, dtype=np.float64))
        p.apply(good, good.clone())


# ---- BASELINE config 3 at its own size: 256^3 cells x 4 unknowns, c0 = 700, mu = 55.5556 / 700 ----
def test_wave_block_256cube_config3():
    """b := M y with the periodic wave operator (src/WaveSystem.cxx:92-176 on a periodic Cartesian grid, applied with
    torch ops on the GPU); the block-circulant apply must return y (rel-L2 <= 1e-12, north_star's fp64 tolerance)."""
    from circulantpreconditioner_b200 import krylov as K
    n = 256
    c0, mu = 700.0, (0.0793651,) * 3
    g = torch.Generator(device="cuda").manual_seed(5)
    y = torch.randn(4 * n ** 3, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
    b = K.wave_operator((n, n, n), c0, mu, periodic=True)(y)
    with cpc.CirculantPlan(n, n, n, ncomp=4) as p:
        p.set_symbol_wave(c0, *mu)
        x = p.apply(b)
    err = (torch.linalg.vector_norm(x - y) / torch.linalg.vector_norm(y)).item()
    assert err < TOL64, err


def test_pageable_host_arrays_go_through_the_bounce_buffers():
    """numpy arrays are pageable memory (what a CPU-only PETSc Vec hands over): above 4 MB they are copied by several host
    threads through pinned bounce buffers; the result must equal the device-pointer result bit for bit, in place too."""
    nx, ny, nz = 256, 256, 96                    # 100 MB per array: several 64 MB pieces per z-chunk boundary case
    rng = np.random.default_rng(8)
    b = rand_c(rng, nx * ny * nz)
    lam = (55.5556, 0.3, 2.5)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        want = host(p.apply(dev(b)))
        xh = np.empty_like(b)
        i0 = p.info()
        p.apply(b, xh)                             # pageable in, pageable out
        i1 = p.info()
        assert np.array_equal(xh, want)
        assert i1["h2d_bytes"] - i0["h2d_bytes"] == 16 * b.size and i1["d2h_bytes"] - i0["d2h_bytes"] == 16 * b.size
        inplace = b.copy()
        p.apply(inplace, inplace)
        assert np.array_equal(inplace, want)
        pb = torch.from_numpy(b).pin_memory()      # pinned in, pageable out
        xh2 = np.empty_like(b)
        with pytest.raises(ValueError):
            p.apply(pb, xh2.astype(np.complex64))
    assert rel_l2(want, O.FftTransportSolver(nx, ny, nz, *lam, b)) < TOL64
