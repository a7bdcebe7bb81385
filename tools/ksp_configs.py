"""BASELINE configs 2 and 3 on one GPU: GMRES(30) solve time and iteration count with the circulant / block-circulant
preconditioner (PETSc-free harness, circulantpreconditioner_b200/krylov.py), plus the wave-block apply rate."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc
from circulantpreconditioner_b200 import krylov as K

out = {}
# config 2: transport 128^3
shape = (128,) * 3
lam = (55.5556, 0.0, 0.0)
b = K.spherical_step(shape, 650.0, 600.0, device="cuda").to(torch.complex128)
with cpc.CirculantPlan(*shape) as plan:
    plan.set_symbol_transport(*lam)
    for quirk in (False, True):
        A = K.transport_operator(shape, lam, ref_sign_quirk=quirk)
        M = lambda v: plan.apply(v.contiguous())
        K.gmres(A, b, M, maxits=3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        x, its, reason, hist = K.gmres(A, b, M)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        res = (torch.linalg.vector_norm(A(x) - b) / torch.linalg.vector_norm(b)).item()
        out[f"config2_transport_128cube_{'ref_sign_quirk' if quirk else 'consistent_sign'}"] = {
            "its": its, "reason": reason, "solve_s": dt, "true_rel_residual": res}
# config 3: wave 256^3, wall boundaries
n = int(os.environ.get("WAVE_N", "256"))
shape = (n,) * 3
c0, mu = 700.0, (0.0793651,) * 3
p = K.spherical_step(shape, 155e5, 70e5, device="cuda")
b = torch.zeros(n ** 3, 4, dtype=torch.complex128, device="cuda")
b[:, 0] = p
b = b.reshape(-1)
with cpc.CirculantPlan(*shape, ncomp=4) as plan:
    plan.set_symbol_wave(c0, *mu)
    x = torch.empty_like(b)
    for _ in range(3):
        plan.apply(b, x)
    torch.cuda.synchronize()
    reps, acc = 10, None
    for _ in range(reps):
        ms = plan.apply_profiled(b, x)
        acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
    ms = [a / reps for a in acc]
    tot = sum(ms)
    bytes_alg = 5 * 2 * b.numel() * 16
    out[f"wave_block_apply_{n}cube"] = {"ms": tot, "applies_per_s": 1e3 / tot, "alg_GBps": bytes_alg / tot / 1e6,
                                        "pass_ms": ms, "info": plan.info()["fast_path"]}
    A = K.wave_operator(shape, c0, mu)
    M = lambda v: plan.apply(v.contiguous())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    xs, its, reason, hist = K.gmres(A, b, M, restart=30)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res = (torch.linalg.vector_norm(A(xs) - b) / torch.linalg.vector_norm(b)).item()
    out[f"config3_wave_{n}cube_wall"] = {"its": its, "reason": reason, "solve_s": dt, "true_rel_residual": res}
print(json.dumps(out, indent=1))
