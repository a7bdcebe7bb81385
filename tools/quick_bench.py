"""Per-pass timing of cpc_apply on one GPU (development aid; bench.py is the contract)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc

sizes = [int(s) for s in sys.argv[1:]] or [128, 256, 512]
ONLY = os.environ.get("DTYPES", "c128,c64,f64").split(",")       # e.g. DTYPES=c128
for dtype, tdt, eb in (("c128", torch.complex128, 16), ("c64", torch.complex64, 8), ("f64", torch.float64, 8)):
    if dtype not in ONLY:
        continue
    for n in sizes:
        b = torch.randn(n ** 3, dtype=torch.float64, device="cuda").to(tdt)
        x = torch.empty_like(b)
        with cpc.CirculantPlan(n, n, n, dtype=dtype) as p:
            p.set_symbol_transport(55.5556, 55.5556, 55.5556)
            for _ in range(3):
                p.apply(b, x)
            torch.cuda.synchronize()
            acc = None
            reps = 10
            for _ in range(reps):
                ms = p.apply_profiled(b, x)
                acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
            ms = [a / reps for a in acc]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                p.apply(b, x)
            e1.record()
            torch.cuda.synchronize()
            tot = e0.elapsed_time(e1) / reps
            bytes_pass = 2 * n ** 3 * eb
            print(f"{dtype} n={n}: apply {tot:.3f} ms ({1e3/tot:.1f}/s) alg {5*bytes_pass/tot/1e6:.0f} GB/s | passes ms "
                  + " ".join(f"{m:.3f}" for m in ms) + " | GB/s " + " ".join(f"{bytes_pass/m/1e6:.0f}" for m in ms), flush=True)
        del b, x
