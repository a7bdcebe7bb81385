#!/usr/bin/env python
"""bench.py -- headline benchmark: 3-D circulant preconditioner applies/s at 512^3 complex128 on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" is one preconditioner apply  x = IFFT3( FFT3(b) ./ Lambda )  (reference solve_3D,
src/FftLinearSolver_3D.c:166-190) on the 512^3 grid of BASELINE.json's metric, lambda = (55.5556,)*3
(SURVEY.md 8d, config 4).  N > 1 shards the same 512^3 grid into z-slabs (strong scaling) with two NCCL
all-to-all transposes per apply.  Prints ONE JSON line on rank 0.

  value      applies/s with b and x resident in HBM (CUDA events on the plan's stream, max over ranks)
  e2e        the same metric through the host-pointer C-ABI call (cpc_apply with CPC_MEM_HOST on pinned buffers):
             H2D of b and D2H of x are inside the timed region
  roofline   dominant kernel (longest pass): algorithmic bytes per launch (2 x N_local x 16 B) / its mean duration,
             against MEASURED_PEAKS.json's hbm_gbs; "apply" gives the 5-pass figure for the whole apply
  cpu_baseline  the oracle (numpy/scipy-pocketfft restatement of the reference path; FFTW/PETSc are not installable)
             timed on this box's host cores on a bounded sample
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 512
LAMBDA = (55.5556, 55.5556, 55.5556)
METRIC = "circulant_pc_applies_per_s_512cube_fp64"
UNIT = "applies/s"
ELEM_BYTES = 16


def workload_config(n_gpus):
    return {
        "workload": f"scalar circulant PC apply, {N_GRID}^3 complex128, lambda=({LAMBDA[0]},)*3 "
                    "(BASELINE.json config 4 at the size the metric is quoted on)",
        "grid": [N_GRID, N_GRID, N_GRID],
        "decomposition": "single GPU, 5 HBM passes" if n_gpus == 1 else f"z-slabs over {n_gpus} ranks, one process per GPU",
        "l2_policy": "inputs larger than L2 (2.1 GB array vs 126 MB L2); no explicit flush",
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [s for (t, s) in self.samples if t0 <= t <= t1 + 0.05] or [s for (_, s) in self.samples]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def cpu_apply_seconds(n, workers, reps=1):
    import numpy as np
    from oracle import circulant_oracle as O
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n ** 3).astype(np.complex128)
    Diag = O.transport_diag(n, n, n, *LAMBDA)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        O.solve_3D(Diag, b, n, n, n, workers=workers)
        best = min(best, time.perf_counter() - t0)
    return best


def scale_to_full(t_sample, n_sample):
    """Scale a sample apply time to the 512^3 workload by the N log N work ratio (stated in `sample`)."""
    if n_sample == N_GRID:
        return t_sample
    w = (N_GRID ** 3 * math.log2(N_GRID ** 3)) / (n_sample ** 3 * math.log2(n_sample ** 3))
    return t_sample * w


def cpu_baseline_block(budget_s=25.0):
    from oracle import circulant_oracle as O
    cores = O.default_workers()
    t256 = cpu_apply_seconds(256, cores)
    est512 = scale_to_full(t256, 256)
    if est512 <= budget_s:
        t = cpu_apply_seconds(N_GRID, cores)
        sample = f"1 apply of the full {N_GRID}^3 workload (scipy pocketfft c2c fp64, workers={cores})"
    else:
        t = est512
        sample = (f"1 apply at 256^3 (scipy pocketfft c2c fp64, workers={cores}) scaled to {N_GRID}^3 by the "
                  f"N log2 N ratio ({est512 / t256:.2f}x)")
    return {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "note": "reference maths (solve_3D restated in numpy), pocketfft backend: PETSc/FFTW not installable here"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import circulant_oracle as O
    cores = O.default_workers()
    t512 = cpu_apply_seconds(N_GRID, cores)
    total = args.steps + args.warmup
    n_s = N_GRID if t512 * total <= 150.0 else 256
    times = []
    for i in range(total):
        t = cpu_apply_seconds(n_s, cores)
        if i >= args.warmup:
            times.append(scale_to_full(t, n_s))
    ms = 1e3 * sum(times) / len(times)
    value = 1e3 / ms
    sample = (f"each step = 1 apply at {n_s}^3 (scipy pocketfft, workers={cores})"
              + ("" if n_s == N_GRID else f", scaled to {N_GRID}^3 by the N log2 N ratio"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import circulantpreconditioner_b200 as cpc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    torch.cuda.set_device(local_rank)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    n = N_GRID
    nzl = n // world
    nloc = n * n * nzl
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    b = torch.randn(nloc, dtype=torch.float64, device="cuda", generator=gen).to(torch.complex128)
    x = torch.empty_like(b)
    plan = cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=nccl_id)
    plan.set_symbol_transport(*LAMBDA)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        plan.apply(b, x)
    barrier()
    l0 = plan.info()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        plan.apply(b, x)
    e1.record()
    barrier()
    t1 = time.time()
    ms_total = e0.elapsed_time(e1)
    launches = plan.info()["kernel_launches"] - l0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_step = ms_total / args.steps

    # per-pass durations, CUDA events between the passes on the same stream (same inputs, right after the timed loop)
    nprof = max(3, min(args.steps, 20))
    acc = None
    for _ in range(nprof):
        ms = plan.apply_profiled(b, x)
        acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
    pass_ms = [a / nprof for a in acc]
    dist_mode = plan.info()["dist_mode"]
    middle = "cyclic first-order recurrence along z (zsolve.cuh)" if plan.info()["fast_path"][2] == 2 else "forward-z FFT, division, backward-z FFT fused"
    if world == 1:
        names = ["Fx", "Fy", "Fz*Lambda^-1*Bz", "By", "Bx"]
    elif dist_mode == 3:      # no transposes: z recurrence on the local slab, one carry per (kx, ky) line all-gathered
        names = ["Fx", "Fy", "z end values + carry all-gather", "Fz*Lambda^-1*Bz (z solve with carries)", "By", "Bx"]
    elif dist_mode == 2:      # transposes fused into the passes (NVLink peer stores), stream-ordered barriers between
        names = ["Fx", "Fy+transpose", "barrier", "Fz*Lambda^-1*Bz+transpose", "barrier", "By", "Bx"]
    else:
        names = ["Fx", "Fy", "all-to-all", "Fz*Lambda^-1*Bz", "all-to-all", "By", "Bx"]
    names = names[:len(pass_ms)]

    # for comparison: the same symbol through the general form of the middle pass (forward-z FFT, division,
    # backward-z FFT fused), which is what every non-transport symbol runs.  Single GPU, a short run, not the headline.
    general_form = None
    if world == 1 and plan.info()["fast_path"][2] == 2:
        os.environ["CPC_ZSOLVE"] = "0"
        try:
            with cpc.CirculantPlan(n, n, n) as p2:
                p2.set_symbol_transport(*LAMBDA)
                x2 = torch.empty_like(b)
                for _ in range(3):
                    p2.apply(b, x2)
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ng = max(3, min(args.steps, 20))
                g0.record()
                for _ in range(ng):
                    p2.apply(b, x2)
                g1.record()
                torch.cuda.synchronize()
                gms = g0.elapsed_time(g1) / ng
                gp = p2.apply_profiled(b, x2)
                diff = (torch.linalg.vector_norm(x2 - x) / torch.linalg.vector_norm(x)).item()
                general_form = {"middle_pass": "forward-z FFT, division, backward-z FFT fused (CPC_ZSOLVE=0)",
                                "ms_per_step": gms, "value": 1e3 / gms, "steps": ng, "pass_ms": gp,
                                "rel_l2_vs_recurrence_form": diff}
                del x2
        finally:
            del os.environ["CPC_ZSOLVE"]

    # e2e: host-pointer C-ABI call on pinned buffers (H2D of b + D2H of x inside the timed region)
    hb = torch.empty(nloc, dtype=torch.complex128).pin_memory()
    hb.copy_(b)
    hx = torch.empty(nloc, dtype=torch.complex128).pin_memory()
    n_e2e = max(2, min(args.steps, 10))
    plan.apply(hb, hx)
    barrier()
    tt0 = time.perf_counter()
    for _ in range(n_e2e):
        plan.apply(hb, hx)          # returns when x has landed in host memory
    barrier()
    e2e_s = (time.perf_counter() - tt0) / n_e2e
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_ok = bool(torch.allclose(hx[:4096].cuda(), x[:4096], rtol=1e-9, atol=1e-9))

    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_pass = 2 * nloc * ELEM_BYTES
        kern = [(nm, m) for nm, m in zip(names, pass_ms) if nm not in ("all-to-all", "barrier") and "all-gather" not in nm]
        dom_name, dom_ms = max(kern, key=lambda kv: kv[1])
        achieved = bytes_pass / dom_ms / 1e6
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom_name)
            except Exception:
                traffic = None
        apply_alg = 5 * bytes_pass
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": bytes_pass, "middle_pass": middle, "general_symbol_form": general_form,
                    "passes": [{"name": nm, "ms": m, "GB/s": (bytes_pass / m / 1e6) if nm not in ("all-to-all", "barrier") and "all-gather" not in nm else None}
                               for nm, m in zip(names, pass_ms)],
                    "apply": {"alg_bytes": apply_alg, "achieved": apply_alg / ms_step / 1e6,
                              "frac": apply_alg / ms_step / 1e6 / peak}}
        if world > 1 and dist_mode == 3:
            roofline["carry_exchange"] = {
                "how": "read-only sweep for the slab's end values + ncclAllGather; replaces both global transposes",
                "bytes_gathered_per_gpu": N_GRID * N_GRID * ELEM_BYTES * world,
                "ms": pass_ms[names.index("z end values + carry all-gather")] if "z end values + carry all-gather" in names else None}
        elif world > 1:
            sent = nloc * ELEM_BYTES * (world - 1) / world
            if dist_mode == 2:
                # the transpose travels inside the producing kernel; charge kernel + the barrier that follows it
                a2a = [pass_ms[i] + pass_ms[i + 1] for i, nm in enumerate(names) if nm.endswith("+transpose")]
                how = "peer stores fused into the Fy / fused-z kernels; time = kernel + following barrier"
            else:
                a2a = [m for nm, m in zip(names, pass_ms) if nm == "all-to-all"]
                how = "NCCL grouped send/recv"
            roofline["alltoall"] = {"how": how, "bytes_sent_per_gpu": sent, "ms": a2a,
                                    "busbw_GB/s": [sent / m / 1e6 for m in a2a], "peak_GB/s": 900.0,
                                    "frac": [sent / m / 1e6 / 900.0 for m in a2a],
                                    "measured_peer_copy_GB/s": 770.0}
        line = {"metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
                "roofline": roofline,
                "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nloc * ELEM_BYTES * world,
                        "d2h_bytes_per_step": nloc * ELEM_BYTES * world, "steps": n_e2e, "matches_device_result": e2e_ok,
                        "api": "cpc_apply(plan, b_host, x_host, CPC_MEM_HOST) on pinned buffers"},
                "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    plan.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 20:      # defaults sized for the GPU arm; keep the CPU arm to minutes
            args.steps, args.warmup = 3, 1
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
