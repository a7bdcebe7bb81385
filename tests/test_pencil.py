"""The pencil (p_rows x p_cols) decomposition (csrc/pencil.h, pencil_impl.cuh; BASELINE config 4 "pencil-sharded").

CPU tests: the schedule the GPU plan executes -- the very step list, group lists and reordering index map of the library,
through its pure-host ABI helpers (cpc_pencil_layout / cpc_pencil_steps / cpc_pencil_group / cpc_pencil_swap_source) --
is replayed with numpy 1-D FFTs (a) on virtual ranks in one process and (b) over gloo with one process per rank, and must
return the single-process oracle's result (reference solve_3D, src/FftLinearSolver_3D.c:166-190).

GPU tests: every rank's plan of a grid in ONE process on one GPU (cpc_pencil_apply_lockstep: all kernels and the whole
schedule, exchanges as device copies), and one process per GPU over NCCL (needs p_rows * p_cols GPUs).
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import circulantpreconditioner_b200 as cpc
from circulantpreconditioner_b200 import _lib
from oracle import circulant_oracle as O
from tests.conftest import rel_l2

LAM = (55.5556, 0.3, 2.5)
# (shape, p_rows, p_cols): nx, ny divisible by p_rows; ny, nz by p_cols
GRIDS = [((8, 8, 8), 2, 2), ((12, 6, 4), 3, 2), ((8, 4, 6), 1, 2), ((8, 4, 6), 2, 1), ((4, 8, 8), 2, 4), ((16, 8, 4), 4, 2),
         ((6, 6, 6), 1, 1), ((1, 4, 8), 1, 2), ((10, 15, 9), 5, 3)]


def _tables(shape, lam):
    """The three 1-D tables of build_diag_mat_vec_3D for the transport column (the "+1" on y)."""
    tabs = []
    for a, (n, l) in enumerate(zip(shape, lam)):
        c = 1.0 - np.exp(-2j * np.pi * np.arange(n) / n) if n > 1 else np.zeros(1, dtype=np.complex128)
        tabs.append(l * c + (1.0 if a == 1 else 0.0))
    return tabs


def _x_pencil(full, shape, lay):
    nx, ny, nz = shape
    return full.reshape(nz, ny, nx)[lay["z0"]:lay["z0"] + lay["nzl"], lay["y0"]:lay["y0"] + lay["nyl"], :].copy()


def _local_step(a, step, shape, lay, tabs):
    """One pass or SWAP of the schedule on the flat local array a."""
    nx, ny, nz = shape
    k = step["kind"]
    if k == _lib.PSTEP_SWAP:
        A, B, inner = step["a"], step["b"], step["inner"]
        return (a.reshape(A, B, inner).transpose(1, 0, 2) * step["scale"]).ravel()
    if k == _lib.PSTEP_PASS_X:
        v = a.reshape(lay["nzl"], lay["nyl"], nx)
        v = np.fft.fft(v, axis=2) if step["dir"] < 0 else np.fft.ifft(v, axis=2) * nx
    elif k == _lib.PSTEP_PASS_Y:
        v = a.reshape(lay["nzl"], ny, lay["nxl"])
        v = np.fft.fft(v, axis=1) if step["dir"] < 0 else np.fft.ifft(v, axis=1) * ny
    else:
        assert k == _lib.PSTEP_MIDDLE
        v = a.reshape(nz, lay["nyl2"], lay["nxl"])
        ax = tabs[0][lay["x0"]:lay["x0"] + lay["nxl"]]
        ay = tabs[1][lay["y02"]:lay["y02"] + lay["nyl2"]]
        lam_loc = ax[None, None, :] + ay[None, :, None] + tabs[2][:, None, None]
        # what the middle pass of the nxl x nyl2 x nz sub-plan computes: unnormalised backward z, times 1 / its own N
        v = np.fft.ifft(np.fft.fft(v, axis=0) / lam_loc, axis=0) * nz / (lay["nxl"] * lay["nyl2"] * nz)
    return np.ascontiguousarray(v).ravel()


@pytest.mark.parametrize("shape,pr,pc", GRIDS)
def test_schedule_on_virtual_ranks(shape, pr, pc):
    nx, ny, nz = shape
    P = pr * pc
    rng = np.random.default_rng(3)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *LAM, b).reshape(nz, ny, nx)
    tabs = _tables(shape, LAM)
    lays = [cpc.pencil_layout(nx, ny, nz, pr, pc, r) for r in range(P)]
    steps = [cpc.pencil_steps(nx, ny, nz, pr, pc, r) for r in range(P)]
    assert all(len(s) == len(steps[0]) and [t["kind"] for t in s] == [t["kind"] for t in steps[0]] for s in steps)
    nloc = nx * ny * nz // P
    assert all(lays[r]["local_elems"] == nloc for r in range(P))
    # The plan's arrays, as the library numbers them: 0 = the caller's b, 1 / 2 = the two work buffers, 3 = the caller's x.
    # Three runs: device arrays, device arrays with b == x (in place), host arrays (staged through buffer 1).
    for mode in ("device", "aliased", "staged"):
        sk, dk = ("src_buf_staged", "dst_buf_staged") if mode == "staged" else ("src_buf", "dst_buf")
        bufs = []
        for r in range(P):
            b_loc = _x_pencil(b, shape, lays[r]).ravel()
            x_loc = b_loc if mode == "aliased" else np.full(nloc, np.nan + 0j)
            w0 = b_loc.copy() if mode == "staged" else np.full(nloc, np.nan + 0j)
            bufs.append({0: None if mode == "staged" else b_loc, 1: w0, 2: np.full(nloc, np.nan + 0j),
                         3: None if mode == "staged" else x_loc})
        b_before = [None if mode != "device" else bufs[r][0].copy() for r in range(P)]
        for k, st in enumerate(steps[0]):
            src, dst = st[sk], st[dk]
            assert all(steps[r][k][sk] == src and steps[r][k][dk] == dst for r in range(P))
            assert dst != 0                                            # the caller's b is never written
            if st["kind"] in (_lib.PSTEP_A2A_ROW, _lib.PSTEP_A2A_COL):
                assert src != dst
                for r in range(P):
                    peers = cpc.pencil_group(nx, ny, nz, pr, pc, r, st["kind"])
                    assert r in peers and len(peers) == (pr if st["kind"] == _lib.PSTEP_A2A_ROW else pc)
                    chunk = nloc // len(peers)
                    me = peers.index(r)
                    for q, peer in enumerate(peers):
                        assert cpc.pencil_group(nx, ny, nz, pr, pc, peer, st["kind"]) == peers     # same list on every member
                        bufs[peer][dst][me * chunk:(me + 1) * chunk] = bufs[r][src][q * chunk:(q + 1) * chunk]
            else:
                if st["kind"] == _lib.PSTEP_SWAP:
                    assert src != dst                                  # the reordering kernel cannot run in place
                for r in range(P):
                    bufs[r][dst][:] = _local_step(bufs[r][src], steps[r][k], shape, lays[r], tabs)
        final = steps[0][-1][dk]
        assert mode == "staged" or final == 3                          # device arrays: the last step writes the caller's x
        for r in range(P):
            lay = lays[r]
            got = bufs[r][final].reshape(lay["nzl"], lay["nyl"], nx)
            ref = want[lay["z0"]:lay["z0"] + lay["nzl"], lay["y0"]:lay["y0"] + lay["nyl"], :]
            assert rel_l2(got, ref) < 1e-12, (mode, r, rel_l2(got, ref))
            if mode == "device":
                assert np.array_equal(bufs[r][0], b_before[r])         # b untouched when x is another array
    # the plan's bookkeeping: 5 transform passes (fewer on degenerate axes) + 6 reorderings + 4 exchanges
    kinds = [t["kind"] for t in steps[0]]
    assert kinds.count(_lib.PSTEP_SWAP) == 6 and kinds.count(_lib.PSTEP_A2A_ROW) == 2 and kinds.count(_lib.PSTEP_A2A_COL) == 2
    assert kinds.count(_lib.PSTEP_MIDDLE) == 1 and kinds.count(_lib.PSTEP_PASS_X) == (2 if nx > 1 else 0)


def test_swap_index_map_of_the_kernel():
    """cpc_pencil_swap_source is the function the reordering kernel evaluates per output element."""
    L = cpc.lib()
    for A, B, inner in ((3, 4, 5), (1, 7, 2), (6, 1, 3), (4, 4, 1)):
        src = np.arange(A * B * inner)
        want = src.reshape(A, B, inner).transpose(1, 0, 2).ravel()
        got = np.array([L.cpc_pencil_swap_source(o, A, B, inner) for o in range(A * B * inner)])
        assert np.array_equal(got, want)
        assert L.cpc_pencil_swap_source(A * B * inner, A, B, inner) == -1 and L.cpc_pencil_swap_source(-1, A, B, inner) == -1


def test_layout_covers_the_grid_in_all_three_distributions():
    nx, ny, nz, pr, pc = 12, 6, 8, 3, 2
    seen = [np.zeros((nz, ny, nx), dtype=int) for _ in range(3)]
    for r in range(pr * pc):
        l = cpc.pencil_layout(nx, ny, nz, pr, pc, r)
        assert (l["r"], l["c"]) == (r % pr, r // pr)
        seen[0][l["z0"]:l["z0"] + l["nzl"], l["y0"]:l["y0"] + l["nyl"], :] += 1
        seen[1][l["z0"]:l["z0"] + l["nzl"], :, l["x0"]:l["x0"] + l["nxl"]] += 1
        seen[2][:, l["y02"]:l["y02"] + l["nyl2"], l["x0"]:l["x0"] + l["nxl"]] += 1
    assert all(np.all(s == 1) for s in seen)


def test_bad_grids_are_refused():
    out = _lib.PencilLayout()
    L = cpc.lib()
    for args in ((8, 8, 8, 3, 2, 0), (8, 8, 8, 2, 3, 0), (8, 8, 8, 2, 2, 4), (8, 8, 8, 0, 2, 0), (8, 6, 8, 2, 4, 0)):
        assert L.cpc_pencil_layout(*args, out) != 0
    with pytest.raises(cpc.CpcError):
        cpc.pencil_steps(8, 8, 8, 3, 2, 0)
    with pytest.raises(cpc.CpcError):
        cpc.pencil_group(8, 8, 8, 2, 2, 0, 3)          # not an exchange step


def test_pencil_plan_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(cpc.CpcError, match="no CUDA device"):
        cpc.CirculantPlan(8, 8, 8, pencil=(1, 1))


# ---- one process per rank over gloo ----------------------------------------------------------------------------------
def _group_alltoall(send, peers, rank, world):
    """all-to-all within a group out of all_gather (gloo has no all_to_all): chunk q of my buffer goes to peers[q]."""
    bufs = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(bufs, send)
    chunk = send.numel() // len(peers)
    me = peers.index(rank)
    out = torch.empty_like(send)
    for q, peer in enumerate(peers):
        out[q * chunk:(q + 1) * chunk] = bufs[peer][me * chunk:(me + 1) * chunk]
    return out


def _gloo_worker(rank, P, port, shape, pr, pc, b_full, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    nx, ny, nz = shape
    lay = cpc.pencil_layout(nx, ny, nz, pr, pc, rank)
    tabs = _tables(shape, LAM)
    nloc = lay["local_elems"]
    bufs = {0: _x_pencil(b_full, shape, lay).ravel(), 1: np.zeros(nloc, complex), 2: np.zeros(nloc, complex),
            3: np.zeros(nloc, complex)}
    for st in cpc.pencil_steps(nx, ny, nz, pr, pc, rank):
        src, dst = st["src_buf"], st["dst_buf"]
        if st["kind"] in (_lib.PSTEP_A2A_ROW, _lib.PSTEP_A2A_COL):
            peers = cpc.pencil_group(nx, ny, nz, pr, pc, rank, st["kind"])
            bufs[dst][:] = _group_alltoall(torch.from_numpy(np.ascontiguousarray(bufs[src])), peers, rank, P).numpy()
        else:
            bufs[dst][:] = _local_step(bufs[src], st, shape, lay, tabs)
    a = bufs[3]
    parts = [None] * P
    dist.all_gather_object(parts, (rank, a))
    if rank == 0:
        ret["parts"] = parts
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,pr,pc", [((8, 8, 8), 2, 2), ((12, 6, 4), 3, 1), ((4, 8, 8), 1, 4)])
def test_schedule_over_gloo(shape, pr, pc):
    nx, ny, nz = shape
    P = pr * pc
    rng = np.random.default_rng(4)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *LAM, b).reshape(nz, ny, nx)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29300 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(P, port, shape, pr, pc, b, ret), nprocs=P, join=True)
    for r, a in ret["parts"]:
        lay = cpc.pencil_layout(nx, ny, nz, pr, pc, r)
        ref = want[lay["z0"]:lay["z0"] + lay["nzl"], lay["y0"]:lay["y0"] + lay["nyl"], :]
        assert rel_l2(a.reshape(ref.shape), ref) < 1e-12


# ---- GPU ---------------------------------------------------------------------------------------------------------
# These were written after the GPU budget of the round was spent: their first run on hardware is the driver's own.
# Not strict: a pass is reported as XPASS, a failure does not stop the suite; every body runs in a process of its own.
_first_run = pytest.mark.xfail(reason="pencil plans have not run on hardware yet (written after the round's GPU budget was spent)",
                               strict=False)


def _child(_, fn_name, args, ret):
    ret["value"] = globals()[fn_name](*args)


def _spawn(fn, args, nprocs, seconds=240):
    """mp.spawn with a deadline: a hang (an exchange that never completes) ends as a failed test, not as a stuck suite."""
    import time
    ctx = mp.spawn(fn, args=args, nprocs=nprocs, join=False)
    deadline = time.time() + seconds
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"timed out after {seconds} s")


def _isolated(fn_name, *args):
    """Run a GPU test body in a fresh process: code that has never run on hardware must not be able to take the CUDA
    context of the pytest process (and with it every test that follows) down with it."""
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_child, (fn_name, args, ret), 1, seconds=120)
    return ret["value"]


def _lockstep(shape, pr, pc, dtype="c128", lam=LAM, opts=None, host=False):
    nx, ny, nz = shape
    P = pr * pc
    rng = np.random.default_rng(7)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b).reshape(nz, ny, nx)
    npdt = np.complex128 if dtype == "c128" else np.complex64
    plans = [cpc.CirculantPlan(nx, ny, nz, dtype=dtype, nranks=P, rank=r, pencil=(pr, pc)) for r in range(P)]
    try:
        lays = [cpc.pencil_layout(nx, ny, nz, pr, pc, r) for r in range(P)]
        for p in plans:
            p.set_symbol_transport(*lam)
            for k, v in (opts or {}).items():
                p.set_option(k, v)
            assert p.info()["dist_mode"] == 4 and p.info()["local_elems"] == nx * ny * nz // P
        loc = [_x_pencil(b, shape, l).ravel().astype(npdt) for l in lays]
        if host:
            bs = loc
            xs = [np.empty_like(a) for a in loc]
        else:
            bs = [torch.from_numpy(a).cuda() for a in loc]
            xs = [torch.empty_like(a) for a in bs]
        cpc.pencil_apply_lockstep(plans, bs, xs)
        errs = []
        for r, l in enumerate(lays):
            ref = want[l["z0"]:l["z0"] + l["nzl"], l["y0"]:l["y0"] + l["nyl"], :]
            got = xs[r] if host else xs[r].cpu().numpy()
            errs.append(rel_l2(got.reshape(ref.shape), ref))
        if not host:                                   # in place: b == x on every rank
            cpc.pencil_apply_lockstep(plans, bs, bs)
            for r, l in enumerate(lays):
                ref = want[l["z0"]:l["z0"] + l["nzl"], l["y0"]:l["y0"] + l["nyl"], :]
                errs.append(rel_l2(bs[r].cpu().numpy().reshape(ref.shape), ref))
        return max(errs)
    finally:
        for p in plans:
            p.destroy()


@pytest.mark.gpu
@_first_run
@pytest.mark.parametrize("shape,pr,pc", [((64, 64, 64), 2, 2), ((64, 32, 128), 2, 4), ((128, 64, 32), 4, 2), ((32, 32, 32), 1, 1),
                                         ((24, 12, 20), 3, 2), ((64, 64, 64), 1, 4), ((64, 64, 64), 4, 1), ((256, 256, 256), 2, 2)])
def test_pencil_grid_in_one_process(shape, pr, pc):
    assert _isolated("_lockstep", shape, pr, pc) < 1e-12


@pytest.mark.gpu
@_first_run
def test_pencil_grid_fft_form_fp32_and_host_arrays():
    assert _isolated("_lockstep", (64, 64, 64), 2, 2, "c128", LAM, {"z_recurrence": 0}) < 1e-12    # FFT form of the middle pass
    assert _isolated("_lockstep", (64, 64, 64), 2, 2, "c64", (2.5, 0.3, 2.5)) < 1e-5
    assert _isolated("_lockstep", (32, 64, 32), 2, 2, "c128", LAM, None, True) < 1e-12


@pytest.mark.gpu
@_first_run
def test_pencil_plan_single_rank_apply_and_errors():
    """A 1 x 1 grid goes through cpc_apply itself (no exchange partner needed); unsupported calls say so."""
    assert _isolated("_single_rank_body") == "ok"


def _single_rank_body():
    n = 32
    rng = np.random.default_rng(8)
    b = rng.standard_normal(n ** 3) + 1j * rng.standard_normal(n ** 3)
    want = O.FftTransportSolver(n, n, n, *LAM, b)
    with cpc.CirculantPlan(n, n, n, pencil=(1, 1)) as p:
        with pytest.raises(cpc.CpcError):
            p.apply(torch.from_numpy(b).cuda())                  # no symbol yet
        p.set_symbol_transport(*LAM)
        got = p.apply(torch.from_numpy(b).cuda()).cpu().numpy()
        assert rel_l2(got, want) < 1e-12
        with pytest.raises(cpc.CpcError, match="pencil"):
            p.set_symbol_wave(700.0, 0.1, 0.1, 0.1)
        with pytest.raises(cpc.CpcError, match="pencil"):
            p.forward(torch.from_numpy(b).cuda())
    with pytest.raises(cpc.CpcError):
        cpc.CirculantPlan(n, n, n, nranks=4, rank=0, pencil=(3, 2))          # 3 x 2 != 4
    with pytest.raises(cpc.CpcError):
        cpc.CirculantPlan(n, n, n, ncomp=4, pencil=(1, 1))
    with cpc.CirculantPlan(n, n, n, nranks=4, rank=1, pencil=(2, 2)) as p:    # no communicator: lockstep only
        p.set_symbol_transport(*LAM)
        with pytest.raises(cpc.CpcError, match="lockstep"):
            p.apply(torch.zeros(n ** 3 // 4, dtype=torch.complex128, device="cuda"))
    return "ok"


def _nccl_worker(rank, P, port, shape, pr, pc, b_full, want, errs):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=P, device_id=torch.device("cuda", rank))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nx, ny, nz = shape
    lay = cpc.pencil_layout(nx, ny, nz, pr, pc, rank)
    loc = torch.from_numpy(_x_pencil(b_full, shape, lay).ravel()).cuda()
    ref = want.reshape(nz, ny, nx)[lay["z0"]:lay["z0"] + lay["nzl"], lay["y0"]:lay["y0"] + lay["nyl"], :]
    with cpc.CirculantPlan(nx, ny, nz, nranks=P, rank=rank, nccl_id=idt.cpu().numpy().tobytes(), pencil=(pr, pc)) as p:
        p.set_symbol_transport(*LAM)
        out = p.apply(loc, torch.empty_like(loc))
        e1 = rel_l2(out.cpu().numpy().reshape(ref.shape), ref)
        p.apply(loc, loc)
        e2 = rel_l2(loc.cpu().numpy().reshape(ref.shape), ref)
    errs[rank] = (float(e1), float(e2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@_first_run
@pytest.mark.parametrize("shape,pr,pc", [((64, 64, 64), 2, 1), ((64, 64, 64), 1, 2), ((64, 64, 64), 2, 2), ((128, 64, 64), 2, 4),
                                         ((64, 128, 64), 4, 2)])
def test_pencil_grid_over_nccl(shape, pr, pc):
    P = pr * pc
    if torch.cuda.device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    nx, ny, nz = shape
    rng = np.random.default_rng(9)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *LAM, b)
    mgr = mp.Manager()
    errs = mgr.dict()
    port = 29100 + (os.getpid() % 2000)
    _spawn(_nccl_worker, (P, port, shape, pr, pc, b, want, errs), P)
    for r, (e1, e2) in dict(errs).items():
        assert e1 < 1e-12 and e2 < 1e-12, (r, e1, e2)
