"""Multi-GPU tuning aid (launch with torch.distributed.run): per-pass CUDA-event times of the z-slab schedule for a few
variants selected through the CPC_TUNING hooks.  usage: dist_variants.py [N] ; one line per variant on rank 0."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import circulantpreconditioner_b200 as cpc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world, rank, lrank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
nzl = n // world
b = torch.randn(n * n * nzl, dtype=torch.float64, device="cuda").to(torch.complex128)
x = torch.empty_like(b)
os.environ["CPC_TUNING"] = "1"
VARIANTS = [{}, {"CPC_FUSED_SYNC": "0"}, {"CPC_XSPLIT": "1"}, {"CPC_FLAG_BARRIER": "0"}]
for var in VARIANTS:
    for k in ("CPC_ZSLAB_LINE", "CPC_END_TRUNC", "CPC_FLAG_BARRIER", "CPC_XSPLIT", "CPC_FUSED_SYNC"):
        os.environ.pop(k, None)
    os.environ.update(var)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    with cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=idt.cpu().numpy().tobytes()) as p:
        p.set_symbol_transport(55.5556, 55.5556, 55.5556)
        for _ in range(5):
            p.apply(b, x)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            p.apply(b, x)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        acc = None
        for _ in range(5):
            ms = p.apply_profiled(b, x)
            acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
        if rank == 0:
            print(f"n={n} P={world} {var or 'default'}: apply {t.item():.4f} ms ({1e3 / t.item():.1f}/s) | passes "
                  + " ".join(f"{a / 5:.4f}" for a in acc), flush=True)
dist.barrier()
dist.destroy_process_group()
