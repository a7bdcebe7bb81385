"""CPU test of the multi-rank (z-slab) schedule: world_size-2 gloo processes replay the distributed apply with numpy
1-D FFTs, using the library's own layout helpers (cpc_slab_range / cpc_slab_send_chunk / cpc_slab_recv_chunk) for the
"zero-pack" chunk layout that the CUDA y-pass writes and NCCL ships.  Result must equal the single-process oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import rel_l2


def _chunk(fn, nx, ny, nz, nc, P, r, q):
    o, c = ctypes.c_int64(), ctypes.c_int64()
    assert fn(nx, ny, nz, nc, P, r, q, ctypes.byref(o), ctypes.byref(c)) == 0
    return o.value, c.value


def _exchange(send, send_chunks, recv_chunks, rank, P):
    """all-to-all built from all_gather (gloo has no all_to_all): chunk q of every rank's send buffer goes to rank q."""
    bufs = [torch.empty_like(send) for _ in range(P)]
    dist.all_gather(bufs, send)
    out = torch.empty_like(send)
    for s in range(P):
        so, sc = send_chunks[s][rank]          # what rank s sends to me
        ro, rc = recv_chunks[s]                # where it lands here
        assert sc == rc
        out[ro:ro + rc] = bufs[s][so:so + sc]
    return out


def _worker(rank, P, port, shape, lam, b_full, ret):
    import circulantpreconditioner_b200 as cpc
    from oracle import circulant_oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    L = cpc.lib()
    nx, ny, nz = shape
    z0, nzl = cpc.slab_range(nz, P, rank)
    y0, nyl = cpc.slab_range(ny, P, rank)
    send_chunks = [[_chunk(L.cpc_slab_send_chunk, nx, ny, nz, 1, P, s, q) for q in range(P)] for s in range(P)]
    recv_chunks = [_chunk(L.cpc_slab_recv_chunk, nx, ny, nz, 1, P, rank, s) for s in range(P)]
    back_recv = [send_chunks[rank][q] for q in range(P)]

    slab = b_full.reshape(nz, ny, nx)[z0:z0 + nzl].copy()
    slab = np.fft.fft(np.fft.fft(slab, axis=2), axis=1)                      # Fx, Fy on the z-slab
    send = np.empty(slab.size, dtype=np.complex128)
    for q in range(P):                                                        # [q][z_loc][y_loc][x]
        yq0, nyq = cpc.slab_range(ny, P, q)
        o, c = send_chunks[rank][q]
        send[o:o + c] = slab[:, yq0:yq0 + nyq, :].ravel()
    tr = _exchange(torch.from_numpy(send), send_chunks, recv_chunks, rank, P).numpy().reshape(nz, nyl, nx)
    Diag = O.transport_diag(nx, ny, nz, *lam).reshape(nz, ny, nx)[:, y0:y0 + nyl, :]
    tr = np.fft.ifft(np.fft.fft(tr, axis=0) / Diag, axis=0) * nz             # Fz . 1/Lambda . Bz (unnormalised Bz)
    # reverse exchange: chunk s of the transposed buffer goes back to rank s
    rsend_chunks = [[_chunk(L.cpc_slab_recv_chunk, nx, ny, nz, 1, P, s, q) for q in range(P)] for s in range(P)]
    back = _exchange(torch.from_numpy(tr.ravel().copy()), rsend_chunks, back_recv, rank, P).numpy()
    slab2 = np.empty((nzl, ny, nx), dtype=np.complex128)
    for q in range(P):
        yq0, nyq = cpc.slab_range(ny, P, q)
        o, c = send_chunks[rank][q]
        slab2[:, yq0:yq0 + nyq, :] = back[o:o + c].reshape(nzl, nyq, nx)
    slab2 = np.fft.ifft(np.fft.ifft(slab2, axis=1), axis=2) * (ny * nx) / (nx * ny * nz)
    parts = [None] * P
    dist.all_gather_object(parts, (z0, slab2))
    if rank == 0:
        full = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0).ravel()
        ret["x"] = full
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,P", [((8, 6, 4), 2), ((5, 10, 6), 2)])   # ny, nz divisible by P, as the CUDA path requires
def test_slab_schedule_world_size_2(shape, P):
    from oracle import circulant_oracle as O
    nx, ny, nz = shape
    rng = np.random.default_rng(1)
    lam = (0.6, 0.15, 2.0)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(P, port, shape, lam, b, ret), nprocs=P, join=True)
    assert rel_l2(ret["x"], want) < 1e-13


# ---------------------------------------------------------------------------------------------------------------
# The transpose-free schedule for the transport symbol (csrc/zsolve.cuh, PlanT::apply_device_zslab): Fx, Fy on the
# local z-slab with the slab's end values accumulated z-chunk by z-chunk (zs_end_accum_kernel), the carry exchange
# through line owners (zs_carry_push_kernel / zs_carry_owner_kernel: rank q owns lines [q lsub, (q+1) lsub), gathers
# their P end values, closes the cycle over the ranks and hands every rank its carry-in), the second sweep from that
# carry-in (ZS_DIST / zs_dist_line_kernel), By, Bx.
# ---------------------------------------------------------------------------------------------------------------
def _worker_zslab(rank, P, port, shape, lam, b_full, ret):
    import circulantpreconditioner_b200 as cpc
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    nx, ny, nz = shape
    z0, nzl = cpc.slab_range(nz, P, rank)
    lx, ly, lz = lam
    slab = b_full.reshape(nz, ny, nx)[z0:z0 + nzl].copy()
    cx = 1.0 - np.exp(-2j * np.pi * np.arange(nx) / nx) if nx > 1 else np.zeros(1)
    cy = 1.0 - np.exp(-2j * np.pi * np.arange(ny) / ny) if ny > 1 else np.zeros(1)
    alpha = (1.0 + lx * cx[None, :] + ly * cy[:, None]).ravel()               # [ny nx] lines
    r = 1.0 / (alpha + lz)
    c = lz * r
    L = nx * ny
    # [Fx, Fy, end-value accumulation] z-chunk by z-chunk: e = c^len e + (end value of the chunk from a zero carry)
    chunk = 3
    e = np.zeros(L, dtype=np.complex128)
    for zb in range(0, nzl, chunk):
        ze = min(nzl, zb + chunk)
        slab[zb:ze] = np.fft.fft(np.fft.fft(slab[zb:ze], axis=2), axis=1)
        for k in range(zb, ze):
            e = c * e + slab[k].ravel()
    # push: the end values of the lines owned by q go to q (all_gather stands in for the peer stores)
    lsub = (L + P - 1) // P
    allv = [torch.empty(L, dtype=torch.complex128) for _ in range(P)]
    dist.all_gather(allv, torch.from_numpy(e.copy()))
    lo, hi = min(L, rank * lsub), min(L, (rank + 1) * lsub)
    G = np.stack([v.numpy()[lo:hi] for v in allv])                            # gbuf[P][lsub] of this owner
    # owner: Zin_0 by Horner over e_0 .. e_{P-1}, closed cyclically, then Zin_{q+1} = e_q + cL Zin_q
    cL = c[lo:hi] ** nzl
    acc = np.zeros(hi - lo, dtype=np.complex128)
    for s_ in range(P):
        acc = cL * acc + G[s_]
    Z = acc / (1.0 - cL ** P)
    zin_owned = np.zeros((P, lsub), dtype=np.complex128)                     # what this owner pushes to every rank
    for q in range(P):
        zin_owned[q, :hi - lo] = Z
        Z = G[q] + cL * Z
    # push back: rank q receives its row from every owner
    back = [torch.empty(P, lsub, dtype=torch.complex128) for _ in range(P)]
    dist.all_gather(back, torch.from_numpy(zin_owned))
    zin = np.concatenate([bk.numpy()[rank] for bk in back])[:L]
    # second sweep from the carry-in (thread-per-line form): y_k = c y_{k-1} + b_k, x_k = r y_k
    acc = zin.copy()
    out = np.empty_like(slab)
    for k in range(nzl):
        acc = c * acc + slab[k].ravel()
        out[k] = (acc * r).reshape(ny, nx)
    slab2 = np.fft.ifft(np.fft.ifft(out, axis=1), axis=2)
    parts = [None] * P
    dist.all_gather_object(parts, (z0, slab2))
    if rank == 0:
        ret["x"] = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0).ravel()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,P", [((8, 6, 8), 2), ((5, 3, 12), 2), ((4, 4, 12), 3), ((5, 3, 8), 4)])   # only nz % P == 0 is needed
def test_zslab_recurrence_schedule(shape, P):
    from oracle import circulant_oracle as O
    nx, ny, nz = shape
    rng = np.random.default_rng(2)
    lam = (0.6, 0.15, 55.5556)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_zslab, args=(P, port, shape, lam, b, ret), nprocs=P, join=True)
    assert rel_l2(ret["x"], want) < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# Distributed GMRES harness (circulantpreconditioner_b200/krylov.py with a Slab): world_size-2 gloo ranks must give
# the single-process operators, residual histories and iteration counts.
# ---------------------------------------------------------------------------------------------------------------
def _worker_krylov(rank, P, port, ret):
    from circulantpreconditioner_b200 import krylov as K
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=P)
    slab = K.Slab()
    out = {}
    for name, shape, lam, periodic, quirk in (("t", (6, 5, 8), (3.0, 0.5, 1.5), False, False),
                                              ("tq", (6, 5, 8), (3.0, 0.5, 1.5), False, True),
                                              ("tp", (6, 5, 8), (3.0, 0.5, 1.5), True, False)):
        nx, ny, nz = shape
        nzl = nz // P
        b_full = K.spherical_step(shape, 650.0, 600.0).to(torch.complex128)
        b = K.spherical_step(shape, 650.0, 600.0, slab=slab).to(torch.complex128)
        assert torch.equal(b, b_full.reshape(nz, -1)[rank * nzl:(rank + 1) * nzl].reshape(-1))
        A = K.transport_operator(shape, lam, periodic=periodic, ref_sign_quirk=quirk, slab=slab)
        x, its, reason, hist = K.gmres(A, b, slab=slab)
        out[name] = (A(b).numpy(), x.numpy(), its, reason, hist)
    shape = (4, 3, 6)
    g = torch.Generator().manual_seed(3)
    u_full = torch.randn(4 * 4 * 3 * 6, dtype=torch.float64, generator=g).to(torch.complex128)
    nloc = u_full.numel() // P
    u = u_full[rank * nloc:(rank + 1) * nloc].clone()
    for name, periodic in (("w", False), ("wp", True)):
        Aw = K.wave_operator(shape, 3.0, (0.08, 0.05, 0.03), periodic=periodic, slab=slab)
        x, its, reason, hist = K.gmres(Aw, u, slab=slab)
        out[name] = (Aw(u).numpy(), x.numpy(), its, reason, hist)
    ret[rank] = out
    dist.destroy_process_group()


def test_distributed_krylov_harness_world_size_2():
    from circulantpreconditioner_b200 import krylov as K
    P = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_krylov, args=(P, 33500 + (os.getpid() % 2000), ret), nprocs=P, join=True)
    got = {r: ret[r] for r in range(P)}

    def cat(name, i):
        return np.concatenate([got[r][name][i] for r in range(P)])

    for name, shape, lam, periodic, quirk in (("t", (6, 5, 8), (3.0, 0.5, 1.5), False, False),
                                              ("tq", (6, 5, 8), (3.0, 0.5, 1.5), False, True),
                                              ("tp", (6, 5, 8), (3.0, 0.5, 1.5), True, False)):
        b = K.spherical_step(shape, 650.0, 600.0).to(torch.complex128)
        A = K.transport_operator(shape, lam, periodic=periodic, ref_sign_quirk=quirk)
        x, its, reason, hist = K.gmres(A, b)
        assert np.allclose(cat(name, 0), A(b).numpy(), rtol=1e-13, atol=1e-10)
        assert got[0][name][2] == its and got[1][name][2] == its and got[0][name][3] == reason
        assert np.allclose(got[0][name][4], hist, rtol=1e-8, atol=1e-12)
        assert np.allclose(cat(name, 1), x.numpy(), rtol=1e-8, atol=1e-8)
    shape = (4, 3, 6)
    g = torch.Generator().manual_seed(3)
    u = torch.randn(4 * 4 * 3 * 6, dtype=torch.float64, generator=g).to(torch.complex128)
    for name, periodic in (("w", False), ("wp", True)):
        Aw = K.wave_operator(shape, 3.0, (0.08, 0.05, 0.03), periodic=periodic)
        x, its, reason, hist = K.gmres(Aw, u)
        assert np.allclose(cat(name, 0), Aw(u).numpy(), rtol=1e-13, atol=1e-12)
        assert got[0][name][2] == its and got[0][name][3] == reason
        assert np.allclose(cat(name, 1), x.numpy(), rtol=1e-8, atol=1e-8)
