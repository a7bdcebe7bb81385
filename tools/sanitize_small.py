"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel family once, tiny shapes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import circulantpreconditioner_b200 as cpc
from oracle import circulant_oracle as O

rng = np.random.default_rng(0)
worst = 0.0
# the z extents cover the recurrence kernel's forms (16, 8, 10, 5, 4 points per thread), a length that fits none (9)
# and the 2^a * 3 / 2^a * 5 / generic line kernels
for shape in [(32, 16, 64), (64, 128, 16), (512, 8, 8), (16, 256, 16), (20, 6, 9), (1024, 8, 2), (48, 12, 96),
              (100, 8, 200), (8, 14, 100), (24, 10, 40)]:
    nx, ny, nz = shape
    lam = (2.0, 0.5, 1.5)
    b = rng.standard_normal(nx * ny * nz) + 1j * rng.standard_normal(nx * ny * nz)
    want = O.FftTransportSolver(nx, ny, nz, *lam, b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(*lam)
        got = p.apply(torch.from_numpy(b).cuda()).cpu().numpy()
        worst = max(worst, np.linalg.norm(got - want) / np.linalg.norm(want))
        p.set_symbol_diag(O.transport_diag(nx, ny, nz, *lam))
        got = p.apply(torch.from_numpy(b).cuda()).cpu().numpy()
        worst = max(worst, np.linalg.norm(got - want) / np.linalg.norm(want))
    with cpc.CirculantPlan(nx, ny, nz, dtype="f64") as p:
        p.set_symbol_transport(*lam)
        wr = O.FftTransportSolver(nx, ny, nz, *lam, b.real.astype(np.complex128)).real
        got = p.apply(torch.from_numpy(np.ascontiguousarray(b.real)).cuda()).cpu().numpy()
        worst = max(worst, np.linalg.norm(got - wr) / np.linalg.norm(wr))
    with cpc.CirculantPlan(nx, ny, nz, dtype="c64") as p:
        p.set_symbol_transport(*lam)
        p.apply(torch.from_numpy(b.astype(np.complex64)).cuda())
for shape in [(16, 16, 32), (6, 5, 4), (64, 16, 16)]:
    nx, ny, nz = shape
    c0, mu = 3.0, (0.08, 0.07, 0.06)
    b = (rng.standard_normal(4 * nx * ny * nz)).astype(np.complex128)
    want = O.solve_wave_block(b, nx, ny, nz, c0, *mu)
    with cpc.CirculantPlan(nx, ny, nz, ncomp=4) as p:
        p.set_symbol_wave(c0, *mu)
        got = p.apply(torch.from_numpy(b).cuda()).cpu().numpy()
        worst = max(worst, np.linalg.norm(got - want) / np.linalg.norm(want))
torch.cuda.synchronize()
print("sanitize_small worst rel-L2", worst)
assert worst < 1e-12
