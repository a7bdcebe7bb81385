"""The reference-named interface (glue/) over the C ABI.

CPU part: the shared library builds, loads and exports every name of the reference's two headers
(/root/reference/src/FftLinearSolver_3D.h:21-43, PCSHELLFft_3D.hxx:23-41).
GPU part: the reference's direct-solver tests restated with real assertions (glue/test_fft_solver_3d.cxx),
and the by-value-context entry point driven through ctypes on the 32^3 golden fixture.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT, rel_l2

GLUE = os.path.join(ROOT, "circulantpreconditioner_b200", "glue")
LIB = os.path.join(GLUE, "libfftpreconditioner_b200.so")

C_NAMES = ["CPCMatAttachPlan", "CPCMatGetPlan", "CPCApplyProjected", "build_transport_col", "vec_kronecker_product_identity_left", "vec_kronecker_product_identity_right",
           "build_diag_mat_vec_3D", "solve_3D", "Fft3DSolver", "FftTransportSolver",
           "Fft3DTransportSolver", "Fft2DTransportSolver", "Fft1DTransportSolver", "PetscFft3DTransportSolver",
           "PCShellFFT3DAttach", "getFFTPrec3DContextCreate"]
CXX_NAMES = ["applyFFT3DPrecTransport", "setupFFTPrec3D", "destroyFFTPrec3D", "getFFTPrec3DContext"]


def ensure_built():
    # always run make: a no-op when up to date, and it rebuilds the glue when include/circulantpc.h changed
    # (the glue embeds struct layouts of the C ABI, e.g. cpc_plan_info)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "circulantpreconditioner_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", GLUE])
    return LIB


def test_glue_exports_reference_names():
    ensure_built()
    L = ctypes.CDLL(LIB)
    for n in C_NAMES:
        assert hasattr(L, n), n
    syms = subprocess.run(["nm", "-D", "--demangle", LIB], capture_output=True, text=True).stdout
    for n in CXX_NAMES:
        assert any(line.split(" T ")[-1].startswith(n + "(") for line in syms.splitlines() if " T " in line), n


def test_glue_fails_loudly_without_a_gpu():
    """No CPU fallback anywhere on the product path: on a box without a CUDA device MatCreateFFT (the plan behind the
    reference's FFT Mat) must return a PETSc error that says so -- not a Mat that computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    ensure_built()
    from circulantpreconditioner_b200 import glue_binding as G
    L = G.lib()
    L.MatCreateFFT.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                               ctypes.POINTER(ctypes.c_void_p)]
    mat = ctypes.c_void_p()
    dims = (ctypes.c_int * 3)(8, 8, 8)
    ierr = L.MatCreateFFT(0, 3, dims, b"fftw", ctypes.byref(mat))
    assert ierr != 0 and not mat.value
    assert b"no CPU fallback" in L.ShimLastError() or b"CUDA" in L.ShimLastError()
    # the host-side object model of the stand-in works without a GPU (Vecs, views)
    v = G.Vec.create_host(6)
    v.numpy()[:] = np.arange(6) * (1 + 1j)
    assert v.local_size() == 6 and v.numpy()[5] == 5 + 5j


def test_forwarding_headers_carry_reference_file_names():
    for h in ("FftLinearSolver_3D.h", "PCSHELLFft_3D.hxx"):
        assert "circulantpc_petsc.h" in open(os.path.join(GLUE, h)).read()


@pytest.mark.gpu
def test_reference_direct_solver_tests_restated():
    ensure_built()
    exe = os.path.join(GLUE, "test_fft_solver_3d")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout


class _Ctx(ctypes.Structure):       # StructuredTransportContext, reference FftLinearSolver_3D.h:7-19
    _fields_ = [("n_x", ctypes.c_int), ("n_y", ctypes.c_int), ("n_z", ctypes.c_int),
                ("a_x", ctypes.c_double * 2), ("a_y", ctypes.c_double * 2), ("a_z", ctypes.c_double * 2),
                ("dt", ctypes.c_double * 2), ("delta_x", ctypes.c_double * 2), ("delta_y", ctypes.c_double * 2),
                ("delta_z", ctypes.c_double * 2), ("FFT_MAT", ctypes.c_void_p)]


@pytest.mark.gpu
def test_petsc_fft3d_transport_solver_by_value_context():
    ensure_built()
    from circulantpreconditioner_b200 import glue_binding as G
    L = G.lib()
    f = np.load(os.path.join(GOLDEN, "ref_py_3d_32cube_phys.npz"))
    n, N = 32, 32 ** 3
    L.MatCreateFFT.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p,
                               ctypes.POINTER(ctypes.c_void_p)]
    L.PetscFft3DTransportSolver.argtypes = [_Ctx, ctypes.c_void_p, ctypes.c_void_p]
    B, X = G.Vec.create_host(N), G.Vec.create_host(N)
    mat = ctypes.c_void_p()
    dims = (ctypes.c_int * 3)(n, n, n)
    assert L.MatCreateFFT(0, 3, dims, b"fftw", ctypes.byref(mat)) == 0, L.ShimLastError()
    B.numpy()[:] = f["b"]
    c2 = lambda v: (ctypes.c_double * 2)(v, 0.0)
    # lambda = a dt / delta = (0.6, 0.15, 0.02): the physics of testFftSolver_3D.py:82-91
    ctx = _Ctx(n, n, n, c2(6.0), c2(3.0), c2(1.0), c2(0.01), c2(0.1), c2(0.2), c2(0.5), mat)
    assert L.PetscFft3DTransportSolver(ctx, B.h, X.h) == 0, L.ShimLastError()
    assert rel_l2(X.numpy().real, f["X_real"]) < 1e-12
    assert rel_l2(X.numpy(), f["X_ref"]) < 1e-12
    assert L.PetscFft3DTransportSolver(ctx, B.h, B.h) == 0          # Un, Un aliasing, second step on the same Mat
    assert rel_l2(B.numpy().real, f["X_real"]) < 1e-12
    bad = _Ctx(n, n, 16, c2(6.0), c2(3.0), c2(1.0), c2(0.01), c2(0.1), c2(0.2), c2(0.5), mat)
    assert L.PetscFft3DTransportSolver(bad, B.h, X.h) == 62         # PETSC_ERR_ARG_WRONG
    L.MatDestroy(ctypes.byref(mat))


def test_glue_sources_are_petsc_clean():
    """circulantpc_petsc.cxx / circulantpc_pcshell.cxx (and the example of the reference-side driver changes,
    example_petsc_driver.cxx) compile with -DCPC_WITH_PETSC against a header in which Vec, Mat and PC are opaque pointers
    (glue/petsc_opaque_stub.h): they use public PETSc functions only."""
    subprocess.check_call(["make", "-s", "-B", "-C", GLUE, "check-petsc-clean"])
    for src in ("circulantpc_petsc.cxx", "circulantpc_pcshell.cxx"):
        text = open(os.path.join(GLUE, src)).read()
        for inside in ("->array", "->darray", "->hdr", "->rowptr", "->colidx", "petsc_shim.cxx"):
            assert inside not in text, (src, inside)


@pytest.mark.gpu
@pytest.mark.parametrize("cuda_vecs", [True, False])
def test_pcshell_path_takes_the_recurrence_kernels(cuda_vecs):
    """PCSetUp -> setupFFTPrec3D builds Diag (build_diag_mat_vec_3D), PCApply -> applyFFT3DPrecTransport -> solve_3D hands
    it to the plan, which recognises the separable table: symbol kind SEPARABLE, middle pass = recurrence, and with
    CUDA Vecs nothing is staged through the host."""
    import torch
    ensure_built()
    from circulantpreconditioner_b200 import glue_binding as G
    from oracle import circulant_oracle as O
    nx, ny, nz = 64, 32, 128
    lam = (55.5556, 0.3, 2.5)
    rng = np.random.default_rng(11)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    G.set_default_vec_cuda(cuda_vecs)
    try:
        with G.PCShellFFT3D(3, nx, ny, nz, *lam) as pc:
            if cuda_vecs:
                tb = torch.from_numpy(b).cuda()
                tx = torch.empty_like(tb)
                vb, vx = G.Vec.from_device_tensor(tb), G.Vec.from_device_tensor(tx)
            else:
                vb, vx = G.Vec.create_host(b.size), G.Vec.create_host(b.size)
                vb.numpy()[:] = b
            pc.apply(vb, vx)
            i0 = pc.info()
            pc.apply(vb, vx)
            i1 = pc.info()
            got = tx.cpu().numpy() if cuda_vecs else vx.numpy().copy()
            assert rel_l2(got, want) < 1e-12
            assert i1["symbol_kind"] == 1 and i1["fast_path"][2] == 2, i1
            if cuda_vecs:
                assert i1["h2d_bytes"] == i0["h2d_bytes"] and i1["d2h_bytes"] == i0["d2h_bytes"]
            else:
                assert i1["h2d_bytes"] - i0["h2d_bytes"] == 16 * b.size
            # a Diag that is not separable any more is held as a table, and still gives the oracle's answer
            d = pc.diag().numpy()
            d[7] += 0.5
            pc.apply(vb, vx)
            assert pc.symbol_kind() == 2
            Diag = O.transport_diag(nx, ny, nz, *lam)
            Diag[7] += 0.5
            got = tx.cpu().numpy() if cuda_vecs else vx.numpy().copy()
            assert rel_l2(got, O.solve_3D(Diag, b, nx, ny, nz)) < 1e-12
    finally:
        G.set_default_vec_cuda(False)


@pytest.mark.gpu
def test_pcshell_projection_from_python():
    """applyFFT3DPrecTransport with ctx->intersectionMatrix set: x = P^T solve_3D(P b) (reference PCSHELLFft_3D.cxx:17-21)."""
    ensure_built()
    import scipy.sparse as sp
    from circulantpreconditioner_b200 import glue_binding as G
    from oracle import circulant_oracle as O
    n, M = 8, 300
    N = n ** 3
    rng = np.random.default_rng(4)
    P = sp.random(N, M, density=0.02, random_state=3, format="csr")
    lam = (1.2, 0.4, 0.9)
    b = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    want = P.T @ O.FftTransportSolver(n, n, n, *lam, P @ b)
    with G.PCShellFFT3D(3, n, n, n, *lam, projection=(N, M, P.indptr, P.indices, P.data)) as pc:
        vb, vx = G.Vec.create_host(M), G.Vec.create_host(M)
        vb.numpy()[:] = b
        pc.apply(vb, vx)
        assert rel_l2(vx.numpy(), want) < 1e-12
