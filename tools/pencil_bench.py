"""z-slabs against pencil grids on the same grid (never run on hardware yet: the pencil plans were written after the round's
GPU budget was spent).  Under torch.distributed.run, one process per GPU:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pencil_bench.py 512 2x4 4x2 1x8

prints rel-L2 against x_ref (b := C x_ref) and ms per apply (CUDA events, max over ranks) for the z-slab plan and for
every P_r x P_c grid given.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import circulantpreconditioner_b200 as cpc

LAM = (55.5556, 55.5556, 55.5556)


def shifted(u, axis):
    return torch.roll(u, 1, dims=axis)


def main():
    n = int(sys.argv[1])
    grids = [tuple(int(v) for v in g.split("x")) for g in sys.argv[2:]]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))

    def nccl_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
        if world > 1:
            dist.broadcast(idt, 0)
        return idt.cpu().numpy().tobytes() if world > 1 else None

    # x_ref and b = C x_ref on the whole grid (fine up to 512^3 per GPU), then cut to what the rank holds
    g = torch.Generator(device="cuda").manual_seed(1)
    x_ref = torch.randn(n, n, n, dtype=torch.float64, device="cuda", generator=g)
    b = x_ref.clone()
    for lam, axis in zip(LAM, (2, 1, 0)):
        b += lam * (x_ref - shifted(x_ref, axis))

    def run(plan, cut):
        plan.set_symbol_transport(*LAM)
        bl = cut(b).contiguous().reshape(-1).to(torch.complex128)
        want = cut(x_ref).contiguous().reshape(-1)
        out = torch.empty_like(bl)
        for _ in range(3):
            plan.apply(bl, out)
        err2 = torch.stack([((out.real - want) ** 2).sum() + (out.imag ** 2).sum(), (want ** 2).sum()])
        if world > 1:
            dist.all_reduce(err2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(20):
            plan.apply(bl, out)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float((err2[0] / err2[1]).sqrt()), float(ms)

    z0, nzl = cpc.slab_range(n, world, rank)
    with cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=nccl_id()) as p:
        e, ms = run(p, lambda a: a[z0:z0 + nzl])
    if rank == 0:
        print(f"{n}^3 on {world} GPUs  z-slabs: {ms:.3f} ms  rel-L2 {e:.1e}", flush=True)
    for pr, pc in grids:
        if pr * pc != world:
            continue
        l = cpc.pencil_layout(n, n, n, pr, pc, rank)
        with cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=nccl_id(), pencil=(pr, pc)) as p:
            e, ms = run(p, lambda a: a[l["z0"]:l["z0"] + l["nzl"], l["y0"]:l["y0"] + l["nyl"], :])
        if rank == 0:
            print(f"{n}^3 on {world} GPUs  pencils {pr} x {pc}: {ms:.3f} ms  rel-L2 {e:.1e}", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
