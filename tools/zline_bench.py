"""Tile-kernel vs thread-per-line form of the recurrence middle pass (CPC_OPT_Z_LINE_FORM) on one GPU.
usage: zline_bench.py N [N ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc

for n in [int(s) for s in sys.argv[1:]] or [512, 1024]:
    b = torch.randn(n ** 3, dtype=torch.float64, device="cuda").to(torch.complex128)
    x = torch.empty_like(b)
    with cpc.CirculantPlan(n, n, n) as p:
        p.set_symbol_transport(55.5556, 55.5556, 55.5556)
        ref = None
        for form in (0, 1):
            p.set_option("z_line_form", form)
            for _ in range(3):
                p.apply(b, x)
            acc = None
            for _ in range(5):
                ms = p.apply_profiled(b, x)
                acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
            ms = [a / 5 for a in acc]
            if ref is None:
                ref = x.clone()
            d = (torch.linalg.vector_norm(x - ref) / torch.linalg.vector_norm(ref)).item()
            print(f"n={n} line_form={form} fast_path={p.info()['fast_path']}: passes " + " ".join(f"{m:.3f}" for m in ms)
                  + f" | sum {sum(ms):.3f} ms | z pass {2 * n ** 3 * 16 / ms[2] / 1e6:.0f} GB/s | rel diff vs tile {d:.1e}", flush=True)
    del b, x
