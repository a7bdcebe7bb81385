"""Per-pass timing of cpc_apply for arbitrary nx,ny,nz shapes (development aid): python tools/shape_bench.py 512,256,1024 ..."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc

for arg in sys.argv[1:]:
    nx, ny, nz = (int(v) for v in arg.split(","))
    b = torch.randn(nx * ny * nz, dtype=torch.float64, device="cuda").to(torch.complex128)
    x = torch.empty_like(b)
    with cpc.CirculantPlan(nx, ny, nz) as p:
        p.set_symbol_transport(55.5556, 55.5556, 55.5556)
        for _ in range(3):
            p.apply(b, x)
        torch.cuda.synchronize()
        acc = None
        for _ in range(10):
            ms = p.apply_profiled(b, x)
            acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
        ms = [a / 10 for a in acc]
        bp = 2 * b.numel() * 16
        print(f"{nx}x{ny}x{nz} fast={p.info()['fast_path']}: passes ms " + " ".join(f"{m:.3f}" for m in ms) +
              " | GB/s " + " ".join(f"{bp / m / 1e6:.0f}" for m in ms), flush=True)
    del b, x
