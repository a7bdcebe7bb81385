"""Tuning aid: time each pass of cpc_apply for several kernel variants (CPC_VARIANT_X/Y/Z env hooks).

Variant ids (csrc/plan_impl.cuh enum Variant): 0 wide, 1 narrow, 2 xmap (scalar x pass), 3 wide2, 4 small.
Usage: python tools/sweep_variants.py N  vx,vy,vz[,prefetch_waves] ...   (env NCOMP=4 for the wave block)"""
import os
os.environ["CPC_TUNING"] = "1"      # the library reads its tuning hooks only then
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
combos = [tuple(int(c) for c in a.split(",")) for a in sys.argv[2:]] or [(2, 0, 0), (1, 3, 3), (2, 4, 4), (2, 5, 5), (2, 6, 6), (2, 7, 7)]
combos = [c if len(c) == 4 else c + (0,) for c in combos]
NC = int(os.environ.get("NCOMP", "1"))
DT = os.environ.get("DTYPE", "c128")
b = torch.randn(NC * n ** 3, dtype=torch.float64, device="cuda").to(torch.complex128 if DT == "c128" else torch.complex64)
x = torch.empty_like(b)
ref = None
for vx, vy, vz, pf in combos:
    os.environ["CPC_VARIANT_X"], os.environ["CPC_VARIANT_Y"], os.environ["CPC_VARIANT_Z"] = str(vx), str(vy), str(vz)
    os.environ["CPC_PREFETCH_WAVES"] = str(pf)
    try:
        with cpc.CirculantPlan(n, n, n, ncomp=NC, dtype=DT) as p:
            if NC == 4:
                p.set_symbol_wave(700.0, 0.0793651, 0.0793651, 0.0793651)
            else:
                p.set_symbol_transport(55.5556, 55.5556, 55.5556)
            for _ in range(3):
                p.apply(b, x)
            torch.cuda.synchronize()
            if ref is None:
                ref = x.clone()
            err = (torch.linalg.vector_norm(x - ref) / torch.linalg.vector_norm(ref)).item()
            reps, acc = 10, None
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
            print(f"n={n} var(x,y,z,pf)=({vx},{vy},{vz},{pf}) fast={p.info()['fast_path']} apply {tot:.3f} ms | passes " +
                  " ".join(f"{m:.3f}" for m in ms) + f" | err-vs-first {err:.1e}", flush=True)
    except Exception as e:
        print(f"var ({vx},{vy},{vz}) failed: {e}", flush=True)
