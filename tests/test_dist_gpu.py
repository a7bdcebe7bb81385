"""Multi-GPU parity (-m gpu, needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`).
One process per GPU, NCCL all-to-all inside libcirculantpc; every rank's slab must match the single-process oracle."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, P, port, shape, lam, b_full, want, ncomp, wave, opts, errs):
    import circulantpreconditioner_b200 as cpc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=P, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nx, ny, nz = shape
    z0, nzl = cpc.slab_range(nz, P, rank)
    plane = nx * ny * ncomp
    loc = torch.from_numpy(b_full[z0 * plane:(z0 + nzl) * plane].copy()).cuda()
    with cpc.CirculantPlan(nx, ny, nz, ncomp=ncomp, nranks=P, rank=rank, nccl_id=idt.cpu().numpy().tobytes()) as p:
        for k, v in opts.items():
            if k not in ("diag", "diag_perturb", "expect_dist_mode", "skip_transforms"):
                p.set_option(k, v)
        if wave:
            p.set_symbol_wave(*lam)
        elif opts.get("diag"):
            # the reference's own set-up product: this rank's z-slab of Diag (solve_3D's argument)
            from oracle import circulant_oracle as O
            Diag = O.transport_diag(nx, ny, nz, *lam)
            if opts.get("diag_perturb"):
                Diag[5] += 0.25                 # no longer a[i] + b[j] + c[k]: must be held as a full table
            p.set_symbol_diag(torch.from_numpy(Diag[z0 * plane:(z0 + nzl) * plane].copy()).cuda())
        else:
            p.set_symbol_transport(*lam)
        if "expect_dist_mode" in opts:
            assert p.info()["dist_mode"] == opts["expect_dist_mode"], p.info()
        out = torch.empty_like(loc)
        p.apply(loc, out)
        e1 = np.linalg.norm(out.cpu().numpy() - want[z0 * plane:(z0 + nzl) * plane]) / np.linalg.norm(want)
        p.apply(loc, loc)                      # in place
        e2 = np.linalg.norm(loc.cpu().numpy() - want[z0 * plane:(z0 + nzl) * plane]) / np.linalg.norm(want)
        # forward then inverse returns N * input
        e3 = 0.0
        if not opts.get("skip_transforms"):
            src = torch.from_numpy(b_full[z0 * plane:(z0 + nzl) * plane].copy()).cuda()
            f = p.forward(src)
            bk = p.inverse(f)
            e3 = (torch.linalg.vector_norm(bk / (nx * ny * nz) - src) / torch.linalg.vector_norm(src)).item()
    errs[rank] = (float(e1), float(e2), float(e3))
    dist.barrier()
    dist.destroy_process_group()


def _run(shape, lam, ncomp=1, wave=False, P=2, opts=None):
    from oracle import circulant_oracle as O
    if torch.cuda.device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    nx, ny, nz = shape
    rng = np.random.default_rng(5)
    n = nx * ny * nz * ncomp
    b = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    if wave:
        want = O.solve_wave_block(b, nx, ny, nz, *lam)
    elif opts and opts.get("diag_perturb"):
        Diag = O.transport_diag(nx, ny, nz, *lam)
        Diag[5] += 0.25
        want = O.solve_3D(Diag, b, nx, ny, nz)
    else:
        want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    mgr = mp.Manager()
    errs = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(P, port, shape, lam, b, want, ncomp, wave, dict(opts or {}), errs), nprocs=P, join=True)
    return dict(errs)


# (256, 256, 32): 65536 lines, 16 planes per rank
@pytest.mark.parametrize("shape", [(64, 64, 64), (128, 32, 16), (32, 16, 256), (20, 12, 10), (96, 48, 24), (16, 32, 1024),
                                   (256, 256, 32)])
def test_two_rank_transport(shape):
    errs = _run(shape, (55.5556, 0.3, 2.5))
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


@pytest.mark.parametrize("shape", [(64, 64, 64), (32, 16, 256)])
def test_two_rank_transport_transposing_schedule(shape):
    """z_recurrence = 0: the FFT form of the middle pass and the transposing schedule that every non-transport symbol
    uses; test_two_rank_transport runs the same shapes through the transpose-free z-slab recurrence
    (csrc/zsolve.cuh).  Both must match the oracle."""
    errs = _run(shape, (55.5556, 0.3, 2.5), opts={"z_recurrence": 0, "expect_dist_mode": 2})
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def test_two_rank_wave_block():
    errs = _run((32, 32, 32), (3.0, 0.0793651, 0.0793651, 0.0793651), ncomp=4, wave=True)
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def _worker_table(rank, P, port, shape, col, b_full, want, errs):
    import circulantpreconditioner_b200 as cpc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=P, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nx, ny, nz = shape
    z0, nzl = cpc.slab_range(nz, P, rank)
    plane = nx * ny
    sl = slice(z0 * plane, (z0 + nzl) * plane)
    with cpc.CirculantPlan(nx, ny, nz, nranks=P, rank=rank, nccl_id=idt.cpu().numpy().tobytes()) as p:
        p.set_symbol_first_column(torch.from_numpy(col[sl].copy()).cuda())      # each rank passes its z-slab of the column
        out = p.apply(torch.from_numpy(b_full[sl].copy()).cuda())
        errs[rank] = float(np.linalg.norm(out.cpu().numpy() - want[sl]) / np.linalg.norm(want))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_general_first_column():
    from oracle import circulant_oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    nx, ny, nz = 32, 64, 16
    rng = np.random.default_rng(9)
    col = np.zeros((nz, ny, nx), dtype=np.complex128)
    col[0, 0, 0] = 7.0
    col[0, 0, 1] = -1.0; col[0, 1, 0] = -1.5; col[1, 0, 0] = -0.5; col[-1, 0, 0] = -0.75; col[0, -1, 0] = 0.3j
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.solve_first_column(col.ravel(), b, nx, ny, nz)
    mgr = mp.Manager()
    errs = mgr.dict()
    mp.spawn(_worker_table, args=(2, 29700 + (os.getpid() % 2000), (nx, ny, nz), col.ravel(), b, want, errs), nprocs=2, join=True)
    for r, e in dict(errs).items():
        assert e < 1e-12, (r, e)


def test_four_rank_transport():
    errs = _run((64, 64, 64), (1.0, 2.0, 3.0), P=4)
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


# ---- 8 ranks (one 8 x B200 node: `gpurun --gpus 8 -- python -m pytest tests/test_dist_gpu.py -m gpu -k eight`) ----
@pytest.mark.parametrize("shape", [(64, 64, 64), (256, 256, 64), (40, 24, 128)])
def test_eight_rank_transport(shape):
    """The transpose-free z-slab schedule at 8 ranks (8 / 16 planes per rank; the (40, 24, 128) grid has generic x / y
    lengths): every rank's slab against the single-process oracle."""
    errs = _run(shape, (55.5556, 55.5556, 55.5556), P=8, opts={"expect_dist_mode": 3})
    assert len(errs) == 8
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def test_eight_rank_transport_transposing_schedule():
    errs = _run((64, 64, 64), (55.5556, 0.3, 2.5), P=8, opts={"z_recurrence": 0, "expect_dist_mode": 2})
    assert len(errs) == 8
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def test_eight_rank_wave_block():
    errs = _run((32, 32, 32), (700.0, 0.0793651, 0.0793651, 0.0793651), ncomp=4, wave=True, P=8)
    assert len(errs) == 8
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


# ---- explicit Diag on z-slabs (the Diag Vec of solve_3D, src/FftLinearSolver_3D.c:166, distributed like b and X) ----
@pytest.mark.parametrize("P", [2, 4])
def test_multi_rank_explicit_diag_is_recognised_as_separable(P):
    errs = _run((32, 16, 64), (2.0, 0.5, 3.0), P=P, opts={"diag": 1, "expect_dist_mode": 3})
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def test_two_rank_explicit_diag_without_recurrence():
    """A Diag whose lambda_z is negative is still separable but not recurrence-capable: transposing schedule."""
    errs = _run((32, 16, 64), (2.0, 0.5, -0.2), P=2, opts={"diag": 1, "expect_dist_mode": 2})
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-11 and e2 < 1e-11 and e3 < 1e-12, (r, e1, e2, e3)


def test_two_rank_explicit_diag_full_table():
    """A Diag that is not separable goes to HBM as N reciprocals, transposed to the layout the fused z pass runs in."""
    errs = _run((32, 16, 64), (2.0, 0.5, 3.0), P=2, opts={"diag": 1, "diag_perturb": 1, "expect_dist_mode": 2})
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (r, e1, e2, e3)


def test_two_rank_ny_not_divisible_takes_the_zslab_schedule():
    """ny = 15 cannot be split over 2 ranks for a transpose; the transport symbol needs none."""
    errs = _run((32, 15, 64), (55.5556, 0.3, 2.5), P=2, opts={"expect_dist_mode": 3, "skip_transforms": 1})
    for r, (e1, e2, e3) in errs.items():
        assert e1 < 1e-12 and e2 < 1e-12, (r, e1, e2, e3)


# ---- the reference-named boundary on several ranks: MatCreateFFT(PETSC_COMM_WORLD, ...) of setupFFTPrec3D
# (src/PCSHELLFft_3D.cxx:35) makes a z-slab plan, Diag / b / x are z-slabs of MPI CUDA Vecs -------------------------
def _worker_pcshell(rank, P, port, shape, lam, b_full, want, errs):
    import circulantpreconditioner_b200 as cpc
    from circulantpreconditioner_b200 import glue_binding as G
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=P, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nx, ny, nz = shape
    N = nx * ny * nz
    z0, nzl = cpc.slab_range(nz, P, rank)
    plane = nx * ny
    G.world_set(P, rank, idt.cpu().numpy().tobytes())
    G.set_default_vec_cuda(True)
    tb = torch.from_numpy(b_full[z0 * plane:(z0 + nzl) * plane].copy()).cuda()
    tx = torch.empty_like(tb)
    with G.PCShellFFT3D(3, nx, ny, nz, *lam) as pc:
        vb, vx = G.Vec.from_device_tensor(tb, N=N), G.Vec.from_device_tensor(tx, N=N)
        pc.apply(vb, vx)
        pc.apply(vb, vx)
        info = pc.info()
        e = np.linalg.norm(tx.cpu().numpy() - want[z0 * plane:(z0 + nzl) * plane]) / np.linalg.norm(want)
        errs[rank] = (float(e), int(info["nranks"]), int(info["dist_mode"]), int(info["symbol_kind"]))
    G.world_set(1, 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [2, 4])
def test_multi_rank_pcshell_boundary(P):
    from oracle import circulant_oracle as O
    if torch.cuda.device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    shape, lam = (32, 16, 64), (55.5556, 0.3, 2.5)
    nx, ny, nz = shape
    rng = np.random.default_rng(21)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    mgr = mp.Manager()
    errs = mgr.dict()
    mp.spawn(_worker_pcshell, args=(P, 29800 + (os.getpid() % 2000), shape, lam, b, want, errs), nprocs=P, join=True)
    assert len(errs) == P
    for r, (e, nranks, mode, kind) in dict(errs).items():
        assert e < 1e-12 and nranks == P and mode == 3 and kind == 1, (r, e, nranks, mode, kind)


# ---- real-scalar plans on z-slabs (PetscScalar of a real PETSc build; r2c / c2r inside the x pass, the carry exchange on
# the half spectrum) ----
def _worker_real(rank, P, port, shape, lam, b_full, want, dtype, errs):
    import circulantpreconditioner_b200 as cpc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=P, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nx, ny, nz = shape
    z0, nzl = cpc.slab_range(nz, P, rank)
    plane = nx * ny
    tdt = torch.float64 if dtype == "f64" else torch.float32
    loc = torch.from_numpy(b_full[z0 * plane:(z0 + nzl) * plane].copy()).to(tdt).cuda()
    with cpc.CirculantPlan(nx, ny, nz, dtype=dtype, nranks=P, rank=rank, nccl_id=idt.cpu().numpy().tobytes()) as p:
        p.set_symbol_transport(*lam)
        assert p.info()["dist_mode"] == 3
        out = p.apply(loc)
        e1 = np.linalg.norm(out.cpu().numpy() - want[z0 * plane:(z0 + nzl) * plane]) / np.linalg.norm(want)
        hb = loc.cpu().numpy()
        hx = np.empty_like(hb)
        p.apply(hb, hx)                                   # host pointers
        e2 = np.linalg.norm(hx - want[z0 * plane:(z0 + nzl) * plane]) / np.linalg.norm(want)
    errs[rank] = (float(e1), float(e2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 1e-5)])
def test_two_rank_real_scalar_plan(dtype, tol):
    from oracle import circulant_oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    shape, lam = (64, 32, 128), (55.5556, 0.3, 2.5)
    nx, ny, nz = shape
    rng = np.random.default_rng(31)
    b = rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b.astype(np.complex128)).real
    mgr = mp.Manager()
    errs = mgr.dict()
    mp.spawn(_worker_real, args=(2, 29900 + (os.getpid() % 2000), shape, lam, b, want, dtype, errs), nprocs=2, join=True)
    assert len(errs) == 2
    for r, (e1, e2) in dict(errs).items():
        assert e1 < tol and e2 < tol, (r, e1, e2)
