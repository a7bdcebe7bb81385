"""Sweep the z-chunk size (and stream count) of the L2-chained x / y passes (CPC_OPT_L2_CHUNK_BYTES,
CPC_OPT_CHAIN_STREAMS) at N^3 on one GPU.  usage: chain_sweep.py [N] [dtype]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circulantpreconditioner_b200 as cpc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dtype = sys.argv[2] if len(sys.argv) > 2 else "c128"
tdt = {"c128": torch.complex128, "c64": torch.complex64, "f64": torch.float64}[dtype]
b = torch.randn(n ** 3, dtype=torch.float64, device="cuda").to(tdt)
x = torch.empty_like(b)
reps = 20
with cpc.CirculantPlan(n, n, n, dtype=dtype) as p:
    p.set_symbol_transport(55.5556, 55.5556, 55.5556)
    ref = None
    for streams in (1, 2):
        for mb in (0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 48, 64, 96):
            if mb == 0 and streams == 2:
                continue
            p.set_option("l2_chunk_bytes", mb << 20)
            p.set_option("chain_streams", streams)
            for _ in range(3):
                p.apply(b, x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                p.apply(b, x)
            e1.record()
            torch.cuda.synchronize()
            tot = e0.elapsed_time(e1) / reps
            if ref is None:
                ref = x.clone()
            same = bool(torch.equal(ref, x))
            ms = p.apply_profiled(b, x)
            print(f"{dtype} n={n} chunk {mb:3d} MiB streams {streams}: apply {tot:.3f} ms ({1e3/tot:.1f}/s) bitwise_same={same} | passes "
                  + " ".join(f"{m:.3f}" for m in ms), flush=True)
