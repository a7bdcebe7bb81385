"""Iteration-count parity of GMRES(30) preconditioned by the circulant apply (BASELINE configs 2/3).

The Krylov harness (circulantpreconditioner_b200/krylov.py) is the same code on CPU and GPU; what differs is the
preconditioner: the CPU oracle here, the CUDA plan on the GPU.  KSP parameters are the reference's
(tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:120-126): GMRES, rtol = atol = 1e-5, maxits 1000, restart 30.
"""
import numpy as np
import pytest
import torch

from circulantpreconditioner_b200 import krylov as K
from oracle import circulant_oracle as O

LAM_T = (55.5556, 0.0, 0.0)          # config 2: a = (1,0,0), cfl = 1e3/3, dx_min = dx/6  =>  a dt / dx = 55.5556
C0, MU = 700.0, (0.0793651,) * 3     # config 3: c0 = 700, dt / dx = 55.5556 / 700


def oracle_transport_pc(shape, lam):
    return lambda v: torch.from_numpy(O.FftTransportSolver(*shape, *lam, v.cpu().numpy())).to(v.device)


def oracle_wave_pc(shape):
    return lambda v: torch.from_numpy(O.solve_wave_block(v.cpu().numpy(), *shape, C0, *MU)).to(v.device)


def test_indicative_counts_cpu():
    # 16^3 -> no PC 16 / circulant PC 2 (consistent sign); 15 / 17 with the reference's sign quirk
    shape = (16, 16, 16)
    b = K.spherical_step(shape, 650.0, 600.0).to(torch.complex128)
    M = oracle_transport_pc(shape, LAM_T)
    A = K.transport_operator(shape, LAM_T)
    assert K.gmres(A, b)[1] == 16
    x, its, reason, _ = K.gmres(A, b, M)
    assert its == 2 and reason in (2, 3)
    assert (torch.linalg.vector_norm(A(x) - b) / torch.linalg.vector_norm(b)).item() < 1e-10
    Aq = K.transport_operator(shape, LAM_T, ref_sign_quirk=True)
    assert K.gmres(Aq, b)[1] == 15
    assert K.gmres(Aq, b, M)[1] == 17


def test_circulant_pc_is_exact_on_periodic_grid_cpu():
    shape = (12, 10, 8)
    lam = (3.0, 0.5, 1.5)
    b = K.spherical_step(shape, 650.0, 600.0).to(torch.complex128)
    A = K.transport_operator(shape, lam, periodic=True)
    assert np.allclose(A(b).numpy(), O.apply_transport_matrix(b.numpy(), *shape, *lam))
    x, its, _, _ = K.gmres(A, b, oracle_transport_pc(shape, lam))
    assert its == 1
    # wave operator restatement == the oracle's periodic operator; PC exact there as well
    shape = (6, 5, 4)
    u = torch.randn(4 * 6 * 5 * 4, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).to(torch.complex128)
    Aw = K.wave_operator(shape, 3.0, MU, periodic=True)
    assert np.allclose(Aw(u).numpy(), O.apply_wave_matrix(u.numpy(), *shape, 3.0, *MU), rtol=1e-12, atol=1e-12)


def _gpu_transport_pc(shape, lam):
    import circulantpreconditioner_b200 as cpc
    plan = cpc.CirculantPlan(*shape)
    plan.set_symbol_transport(*lam)
    return plan, (lambda v: plan.apply(v.contiguous()))


@pytest.mark.gpu
@pytest.mark.parametrize("n,quirk", [(16, False), (16, True), (32, False), (32, True), (64, False)])
def test_transport_iteration_parity_gpu_vs_oracle(n, quirk):
    shape = (n, n, n)
    b = K.spherical_step(shape, 650.0, 600.0).to(torch.complex128)
    A_cpu = K.transport_operator(shape, LAM_T, ref_sign_quirk=quirk)
    x_c, its_c, reason_c, hist_c = K.gmres(A_cpu, b, oracle_transport_pc(shape, LAM_T))
    plan, M = _gpu_transport_pc(shape, LAM_T)
    bg = b.cuda()
    x_g, its_g, reason_g, hist_g = K.gmres(K.transport_operator(shape, LAM_T, ref_sign_quirk=quirk), bg, M)
    plan.destroy()
    assert (its_g, reason_g) == (its_c, reason_c)
    # identical counts are the criterion; the residual histories agree to rounding amplified by the conditioning of
    # the Arnoldi recurrence (the sign-quirk system is far from the circulant model)
    assert np.allclose(hist_g, hist_c, rtol=5e-2, atol=1e-7 * hist_c[0])
    # both solves stop at rtol = 1e-5 on the preconditioned residual; with the consistent sign their iterates agree
    # far below that (the sign-quirk system is so ill-conditioned that only the counts are comparable)
    if not quirk:
        assert (torch.linalg.vector_norm(x_g.cpu() - x_c) / torch.linalg.vector_norm(x_c)).item() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("n", [16, 32])
def test_wave_iteration_parity_gpu_vs_oracle(n):
    import circulantpreconditioner_b200 as cpc
    shape = (n, n, n)
    p = K.spherical_step(shape, 155e5, 70e5)
    b = torch.zeros(n ** 3, 4, dtype=torch.complex128)
    b[:, 0] = p
    b = b.reshape(-1)
    A = K.wave_operator(shape, C0, MU)                      # wall boundaries: the circulant block is a preconditioner
    x_c, its_c, reason_c, hist_c = K.gmres(A, b, oracle_wave_pc(shape))
    with cpc.CirculantPlan(*shape, ncomp=4) as plan:
        plan.set_symbol_wave(C0, *MU)
        x_g, its_g, reason_g, hist_g = K.gmres(K.wave_operator(shape, C0, MU), b.cuda(), lambda v: plan.apply(v.contiguous()))
    assert (its_g, reason_g) == (its_c, reason_c)
    assert np.allclose(hist_g, hist_c, rtol=5e-2, atol=1e-7 * hist_c[0])


@pytest.mark.gpu
def test_config2_transport_128cube_gpu():
    shape = (128, 128, 128)
    b = K.spherical_step(shape, 650.0, 600.0, device="cuda").to(torch.complex128)
    plan, M = _gpu_transport_pc(shape, LAM_T)
    A = K.transport_operator(shape, LAM_T)
    x, its, reason, hist = K.gmres(A, b, M)
    plan.destroy()
    assert its == 2 and reason in (2, 3)                     # same count as the CPU-oracle PC at 16^3..64^3
    assert (torch.linalg.vector_norm(A(x) - b) / torch.linalg.vector_norm(b)).item() < 1e-8
