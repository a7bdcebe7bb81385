#!/usr/bin/env python
"""bench.py -- headline benchmark: 3-D circulant preconditioner applies/s at 512^3 complex128 on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" is one preconditioner apply  x = IFFT3( FFT3(b) ./ Lambda )  (reference solve_3D,
src/FftLinearSolver_3D.c:166-190) on the 512^3 grid of BASELINE.json's metric, lambda = (55.5556,)*3
(SURVEY.md 8d, config 4).  N > 1 shards the same 512^3 grid into z-slabs, one process per GPU (strong scaling).
For the transport symbol the headline schedule has NO transposes (the middle pass is a recurrence along z; the slabs
exchange one carry per (kx, ky) line); the transposing schedule every other symbol needs (two all-to-alls fused into
the y / z kernels as NVLink peer stores) is timed beside it as `roofline.general_symbol_form` / `roofline.alltoall`.
Prints ONE JSON line on rank 0.

  value      applies/s with b and x resident in HBM (CUDA events on the plan's stream, max over ranks)
  parity     b := C x_ref for a seeded random x_ref (C = I + sum_d lambda_d (I - S_d), applied with torch ops, every rank
             regenerating its lower neighbour's last plane): rel-L2 of the timed result against x_ref, for the
             headline schedule and for the general (FFT / transposing) form; at N = 1 also against the CPU oracle's
             apply of the same b
  sustained  the same loop run for >= 200 steps with its own clock samples (the headline may be a short burst)
  e2e        the same metric through the host-pointer C-ABI call (cpc_apply with CPC_MEM_HOST on pinned buffers):
             H2D of b and D2H of x are inside the timed region
  pcshell    (N = 1) the same apply through the reference-named PCShell glue (PCApply -> applyFFT3DPrecTransport ->
             solve_3D) on device-resident and host Vecs
  roofline   dominant kernel (longest pass): algorithmic bytes per launch (2 x N_local x 16 B) / its mean duration,
             against MEASURED_PEAKS.json's hbm_gbs; "apply" gives the 5-pass figure for the whole apply
  cpu_baseline  the oracle (numpy/scipy-pocketfft restatement of the reference path; FFTW/PETSc are not installable)
             timed on this box's host cores on one full-size apply
"""
import os

# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank when nproc > 1, which made the CPU legs (rank 0 only:
# --impl reference and cpu_baseline) 3-4x slower than the same code launched plainly.  Rank 0 drops it before numpy /
# scipy / torch load their OpenMP runtimes.
if int(os.environ.get("RANK", "0")) == 0 and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ.pop("OMP_NUM_THREADS", None)

import argparse
import json
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
SEED = 20261018
PARITY_TOL = 1e-12


def workload_config(n_gpus):
    return {
        "workload": f"scalar circulant PC apply, {N_GRID}^3 complex128, lambda=({LAMBDA[0]},)*3 "
                    "(BASELINE.json config 4 at the size the metric is quoted on)",
        "grid": [N_GRID, N_GRID, N_GRID],
        "decomposition": "single GPU, 5 HBM passes" if n_gpus == 1
                         else f"z-slabs over {n_gpus} ranks, one process per GPU",
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def wait_ready(self, timeout=3.0):
        """nvidia-smi needs a moment to start: do not open a (65 ms) timed region before the first sample exists."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.samples and time.time() < t_end:
            time.sleep(0.02)

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)              # let the sample that covers the end of the window arrive
        rows = [s for (t, s) in self.samples if t0 <= t <= t1 + 0.05]
        if not rows and self.samples:  # window shorter than the sampling period: the sample nearest to it
            mid = 0.5 * (t0 + t1)
            rows = [min(self.samples, key=lambda ts: abs(ts[0] - mid))[1]]
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

    def stop(self):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle (cpu_baseline leg and --impl reference): the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------------------------
_DIAG = None


def cpu_apply(b, workers):
    """One reference apply (solve_3D restated) of b on the host; returns (seconds, x)."""
    global _DIAG
    from oracle import circulant_oracle as O
    n = N_GRID
    if _DIAG is None:                                  # set-up, once and not timed (the reference builds Diag once too)
        _DIAG = O.transport_diag(n, n, n, *LAMBDA)
    t0 = time.perf_counter()
    x = O.solve_3D(_DIAG, b, n, n, n, workers=workers)
    return time.perf_counter() - t0, x


def random_b():
    import numpy as np
    rng = np.random.default_rng(0)
    return rng.standard_normal(N_GRID ** 3).astype(np.complex128)


def cpu_baseline_block(b_host=None, x_gpu_host=None):
    import numpy as np
    from oracle import circulant_oracle as O
    cores = O.default_workers()
    b = random_b() if b_host is None else b_host
    t, x = cpu_apply(b, cores)
    out = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"1 apply of the full {N_GRID}^3 workload (scipy pocketfft c2c fp64, workers={cores})",
           "note": "reference maths (solve_3D restated in numpy), pocketfft backend: PETSc/FFTW not installable here"}
    if x_gpu_host is not None:
        out["gpu_vs_oracle_rel_l2"] = float(np.linalg.norm(x_gpu_host - x) / np.linalg.norm(x))
        out["gpu_vs_oracle_tol"] = PARITY_TOL
    return out


def run_reference(args):
    """The reference's CPU path (restated: oracle) on the box's host cores, same 512^3 config at every N; rank 0 only.

    kind stays "port": oracle/_ref does hold the reference's own src/FftLinearSolver_3D.c, but compiled against a stand-in
    for PETSc whose MATFFTW is a direct O(n^2) DFT (oracle/petsc_standin/) -- a checker for parity, not a stand-in for
    FFTW's speed; timing it would misstate the reference by orders of magnitude."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import circulant_oracle as O
    cores = O.default_workers()
    b = random_b()
    times = []
    for i in range(args.steps + args.warmup):
        t, _ = cpu_apply(b, cores)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = 1e3 / ms
    sample = f"each step = 1 apply of the full {N_GRID}^3 workload (scipy pocketfft c2c fp64, workers={cores})"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def make_problem(torch, n, z0, nzl, lam):
    """x_ref (this rank's z-slab, seeded per plane) and b = C x_ref with C = I + sum_d lambda_d (I - S_d),
    (S_d u)_i = u_{i-1} cyclic (reference tests/FFTDirectSolver/testFftSolver_3D.py:12-24)."""
    def plane(z):
        g = torch.Generator(device="cuda").manual_seed(SEED + (z % n))
        return torch.view_as_complex(torch.randn(n, n, 2, dtype=torch.float64, device="cuda", generator=g))
    x_ref = torch.empty(nzl, n, n, dtype=torch.complex128, device="cuda")
    for k in range(nzl):
        x_ref[k] = plane(z0 + k)
    lx, ly, lz = lam
    b = x_ref * (1.0 + lx + ly + lz)
    b.sub_(torch.roll(x_ref, 1, dims=2), alpha=lx)
    b.sub_(torch.roll(x_ref, 1, dims=1), alpha=ly)
    b[1:].sub_(x_ref[:-1], alpha=lz)
    b[0].sub_(plane(z0 - 1), alpha=lz)
    return x_ref.reshape(-1), b.reshape(-1)


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

    def new_nccl_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(cpc.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = N_GRID
    nzl = n // world
    z0 = rank * nzl
    nloc = n * n * nzl
    x_ref, b = make_problem(torch, n, z0, nzl, LAMBDA)
    x = torch.empty_like(b)
    plan = cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=new_nccl_id())
    plan.set_symbol_transport(*LAMBDA)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def parity_of(xv):
        s = torch.stack([torch.sum(torch.abs(xv - x_ref) ** 2), torch.sum(torch.abs(x_ref) ** 2)])
        if world > 1:
            dist.all_reduce(s)
        return float(torch.sqrt(s[0] / s[1]).item())

    def timed_loop(p, steps, xv):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            p.apply(b, xv)
        e1.record()
        barrier()
        return allmax(e0.elapsed_time(e1)) / steps, t0, time.time()

    def profiled(p, xv, reps):
        acc = None
        for _ in range(reps):
            ms = p.apply_profiled(b, xv)
            acc = ms if acc is None else [a + m for a, m in zip(acc, ms)]
        return [a / reps for a in acc]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        plan.apply(b, x)
    if sampler:
        sampler.wait_ready()
    l0 = plan.info()["kernel_launches"]
    ms_step, t0, t1 = timed_loop(plan, args.steps, x)
    launches = plan.info()["kernel_launches"] - l0
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    clocks = sampler.window(t0, t1) if sampler else None
    parity_main = parity_of(x)

    # the same loop run long enough to settle under the power cap
    n_sus = max(200, args.steps)
    sus_ms, s0, s1 = timed_loop(plan, n_sus, x)
    sus_clocks = sampler.window(s0, s1) if sampler else None

    # per-pass durations: CUDA events after every launch on the plan's stream, summed per pass (same inputs)
    nprof = max(3, min(args.steps, 20))
    pass_ms = profiled(plan, x, nprof)
    info = plan.info()
    dist_mode = info["dist_mode"]
    middle = ("cyclic first-order recurrence along z (zsolve.cuh)" if info["fast_path"][2] == 2
              else "forward-z FFT, division, backward-z FFT fused")
    if world == 1:
        names = ["Fx", "Fy", "Fz*Lambda^-1*Bz", "By", "Bx"]
    elif dist_mode == 3:      # no transposes: z recurrence on the local slab, one carry per (kx, ky) line exchanged
        names = ["Fx", "Fy", "z end values (read-only sweep)", "carry exchange", "Fz*Lambda^-1*Bz (z solve with carries)",
                 "By", "Bx"]
    elif dist_mode == 2:      # transposes fused into the passes (NVLink peer stores), stream-ordered barriers between
        names = ["Fx", "Fy+transpose", "barrier", "Fz*Lambda^-1*Bz+transpose", "barrier", "By", "Bx"]
    else:
        names = ["Fx", "Fy", "all-to-all", "Fz*Lambda^-1*Bz", "all-to-all", "By", "Bx"]
    names = names[:len(pass_ms)]

    # the general form of the middle pass (forward-z FFT, division, backward-z FFT fused) on the same symbol and the
    # same b: what every non-transport symbol runs; at N > 1 it needs the two global transposes (all-to-all).
    general_form, alltoall = None, None
    if info["fast_path"][2] == 2:
        with cpc.CirculantPlan(n, n, n, nranks=world, rank=rank, nccl_id=new_nccl_id()) as p2:
            p2.set_option("z_recurrence", 0)
            p2.set_symbol_transport(*LAMBDA)
            x2 = torch.empty_like(b)
            for _ in range(3):
                p2.apply(b, x2)
            ng = max(3, min(args.steps, 20))
            gms, _, _ = timed_loop(p2, ng, x2)
            gpar = parity_of(x2)
            gp = profiled(p2, x2, 3)
            g_mode = p2.info()["dist_mode"]
            general_form = {"middle_pass": "forward-z FFT, division, backward-z FFT fused (CPC_OPT_Z_RECURRENCE = 0)",
                            "dist_mode": g_mode, "ms_per_step": gms, "value": 1e3 / gms, "steps": ng, "pass_ms": gp,
                            "parity": {"rel_l2_vs_x_ref": gpar, "tol": PARITY_TOL, "ok": gpar <= PARITY_TOL}}
            if world > 1:
                sent = nloc * ELEM_BYTES * (world - 1) / world
                if g_mode == 2:     # Fx | Fy+transpose | barrier | fused z+transpose | barrier | By | Bx
                    a2a = [gp[1] + gp[2], gp[3] + gp[4]]
                    how = "peer stores fused into the Fy / fused-z kernels (NVLink); time = kernel + following barrier"
                else:               # Fx | Fy | all-to-all | fused z | all-to-all | By | Bx
                    a2a = [gp[2], gp[4]]
                    how = "NCCL grouped send/recv"
                alltoall = {"how": how, "bytes_sent_per_gpu": sent, "ms": a2a,
                            "busbw_GB/s": [sent / m / 1e6 for m in a2a], "peak_GB/s": 900.0,
                            "frac": [sent / m / 1e6 / 900.0 for m in a2a],
                            "frac_of_measured_peer_copy_770": [sent / m / 1e6 / 770.0 for m in a2a]}
            del x2

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
    e2e_s = allmax((time.perf_counter() - tt0) / n_e2e)
    e2e_par = parity_of(hx.cuda())

    pcshell = None
    if world == 1 and not args.no_pcshell:
        try:
            pcshell = pcshell_block(torch, b, x_ref, max(3, min(args.steps, 20)))
        except Exception as exc:      # the glue is optional for the headline; say why it is missing
            pcshell = {"unavailable": f"{type(exc).__name__}: {exc}"}

    # side reference only (north_star: "cuFFT is timed only as a side reference"): the same three steps with
    # torch.fft (cuFFT Z2Z) + a pointwise division, device-resident; nothing of it is on the product path
    side = None
    if world == 1 and not args.no_side_reference:
        try:
            side = cufft_side_reference(torch, b, x_ref, max(3, min(args.steps, 10)))
        except Exception as exc:
            side = {"unavailable": f"{type(exc).__name__}: {exc}"}

    ksp = None
    if not args.no_ksp:
        del x_ref, b, x
        torch.cuda.empty_cache()
        ksp = ksp_block(torch, cpc, world, rank, new_nccl_id)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        # N = 1: the oracle applies the very b of the timed run and the GPU result is compared with it
        cpu = cpu_baseline_block(hb.numpy(), hx.numpy()) if world == 1 else cpu_baseline_block()
    barrier()

    if rank == 0:
        peak, peak_src = measured_peaks()
        bytes_pass = 2 * nloc * ELEM_BYTES
        not_kernel = ("all-to-all", "barrier")
        kern = [(nm, m) for nm, m in zip(names, pass_ms)
                if nm not in not_kernel and "exchange" not in nm and "end values" not in nm]
        dom_name, dom_ms = max(kern, key=lambda kv: kv[1])
        achieved = bytes_pass / dom_ms / 1e6
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if world == 1 and os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get(dom_name)
                traffic_src = tj.get("_source", "profiles/traffic.json (one ncu --set full capture, committed)")
            except Exception:
                traffic = None
        apply_alg = 5 * bytes_pass
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "alg_bytes_per_launch": bytes_pass, "middle_pass": middle, "general_symbol_form": general_form,
                    "passes": [{"name": nm, "ms": m,
                                "GB/s": (bytes_pass / m / 1e6) if nm not in not_kernel and "exchange" not in nm
                                and "end values" not in nm and m > 0 else None}
                               for nm, m in zip(names, pass_ms)],
                    "apply": {"alg_bytes": apply_alg, "achieved": apply_alg / ms_step / 1e6,
                              "frac": apply_alg / ms_step / 1e6 / peak}}
        if world > 1 and dist_mode == 3:
            roofline["carry_exchange"] = {
                "how": "end values of the local lines (sweep over the planes whose weight can reach 1e-17), pushed to line "
                       "owners over NVLink, cycle closed by the owner, carry-in pushed back; two peer-flag barriers; "
                       "replaces both global transposes",
                "bytes_sent_per_gpu": 2 * N_GRID * N_GRID * ELEM_BYTES * (world - 1) / world,
                "end_values_ms": pass_ms[names.index("z end values (read-only sweep)")],
                "ms": pass_ms[names.index("carry exchange")]}
        if alltoall is not None:
            roofline["alltoall"] = alltoall
        line = {"metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
                "parity": {"rel_l2_vs_x_ref": parity_main, "tol": PARITY_TOL, "ok": parity_main <= PARITY_TOL,
                           "dist_mode": dist_mode,
                           "how": "b := C x_ref (torch ops, seeded per plane); x = apply(b) of the timed loop vs x_ref"},
                "sustained": {"value": 1e3 / sus_ms, "ms_per_step": sus_ms, "steps": n_sus,
                              "sm_mhz": (sus_clocks or {}).get("sm_mhz"), "power_w": (sus_clocks or {}).get("power_w_max"),
                              "reasons": (sus_clocks or {}).get("reasons")},
                "roofline": roofline,
                "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": nloc * ELEM_BYTES * world,
                        "d2h_bytes_per_step": nloc * ELEM_BYTES * world, "steps": n_e2e,
                        "parity_rel_l2_vs_x_ref": e2e_par,
                        "api": "cpc_apply(plan, b_host, x_host, CPC_MEM_HOST) on pinned buffers"},
                "gpu_launches": launches, "clocks": clocks}
        if pcshell is not None:
            line["pcshell"] = pcshell
        if ksp is not None:
            line["ksp"] = ksp
        if side is not None:
            line["side_reference"] = side
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if sampler:
        sampler.stop()
    plan.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# cuFFT side reference (not product code): fftn -> divide by Diag -> ifftn with torch.fft on the same b
# ---------------------------------------------------------------------------------------------------------------
def cufft_side_reference(torch, b, x_ref, steps):
    n = N_GRID
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    chat = 1.0 - torch.exp(-2j * torch.pi * k / n)                     # DFT of the upwind column [1, -1, 0, ...]
    diag = (1.0 + LAMBDA[0] * chat[None, None, :] + LAMBDA[1] * chat[None, :, None] + LAMBDA[2] * chat[:, None, None])
    b3 = b.reshape(n, n, n)

    def apply():
        return torch.fft.ifftn(torch.fft.fftn(b3) / diag)

    for _ in range(2):
        xs = apply()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        xs = apply()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    err = float((torch.linalg.vector_norm(xs.reshape(-1) - x_ref) / torch.linalg.vector_norm(x_ref)).item())
    del xs, diag
    torch.cuda.empty_cache()
    return {"what": "torch.fft.fftn (cuFFT Z2Z) -> / Diag (N-entry table) -> torch.fft.ifftn, device-resident, same b",
            "value": 1e3 / ms, "ms_per_step": ms, "steps": steps, "rel_l2_vs_x_ref": err,
            "note": "side reference only; cuFFT is not linked into libcirculantpc.so"}


# ---------------------------------------------------------------------------------------------------------------
# Full Krylov solves (BASELINE configs 2 and 3) on the same ranks: PETSc-free GMRES(30) harness (krylov.py: the
# reference's KSP settings, tests/TransportEquation_SphericalExplosion_impl_mpi.cxx:120-126), z-slab operators with a
# one-plane halo, global dots through torch.distributed, the preconditioner = the multi-rank plan
# ---------------------------------------------------------------------------------------------------------------
def ksp_block(torch, cpc, world, rank, new_nccl_id):
    from circulantpreconditioner_b200 import krylov as K
    slab = K.Slab()
    out = {"harness": "krylov.gmres: GMRES(30), left PC, rtol = atol = 1e-5, maxits 1000; z-slabs over the ranks"}

    def solve(A, b, plan):
        M = lambda v: plan.apply(v.contiguous())
        K.gmres(A, b, M, maxits=2, slab=slab)                    # warm-up (plan buffers, NCCL)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        x, its, reason, hist = K.gmres(A, b, M, slab=slab)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r = A(x) - b
        num = slab.sum(torch.sum(torch.abs(r) ** 2)).item()
        den = slab.sum(torch.sum(torch.abs(b) ** 2)).item()
        return {"its": its, "reason": reason, "solve_s": dt, "true_rel_residual": (num / den) ** 0.5}

    # config 2: transport 128^3, a = (1, 0, 0): lambda = (55.5556, 0, 0)
    shape, lam = (128,) * 3, (55.5556, 0.0, 0.0)
    b = K.spherical_step(shape, 650.0, 600.0, device="cuda", slab=slab).to(torch.complex128)
    with cpc.CirculantPlan(*shape, nranks=world, rank=rank, nccl_id=new_nccl_id()) as plan:
        plan.set_symbol_transport(*lam)
        # (the consistent upwind sign; with the reference's sign quirk, SURVEY.md F11, the circulant model is the wrong
        # matrix and GMRES does not converge in 1000 iterations at this size -- tools/ksp_configs.py runs both)
        A = K.transport_operator(shape, lam, slab=slab)
        out["config2_transport_128cube"] = solve(A, b, plan)
    # config 3: wave system 256^3 x 4 unknowns, wall boundaries, c0 = 700, dt / dx = 55.5556 / 700
    n = 256
    shape, c0, mu = (n,) * 3, 700.0, (0.0793651,) * 3
    p0 = K.spherical_step(shape, 155e5, 70e5, device="cuda", slab=slab)
    b = torch.zeros(p0.numel(), 4, dtype=torch.complex128, device="cuda")
    b[:, 0] = p0
    b = b.reshape(-1)
    with cpc.CirculantPlan(*shape, ncomp=4, nranks=world, rank=rank, nccl_id=new_nccl_id()) as plan:
        plan.set_symbol_wave(c0, *mu)
        xw = torch.empty_like(b)
        for _ in range(3):
            plan.apply(b, xw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            plan.apply(b, xw)
        e1.record()
        torch.cuda.synchronize()
        out["wave_block_apply_256cube"] = {"applies_per_s": 1e4 / e0.elapsed_time(e1), "ms": e0.elapsed_time(e1) / 10,
                                           "dist_mode": plan.info()["dist_mode"]}
        del xw
        out["config3_wave_256cube_wall"] = solve(K.wave_operator(shape, c0, mu, slab=slab), b, plan)
    return out


# ---------------------------------------------------------------------------------------------------------------
# The reference-named boundary: PCApply -> applyFFT3DPrecTransport -> solve_3D (glue/libfftpreconditioner_b200.so)
# ---------------------------------------------------------------------------------------------------------------
def pcshell_block(torch, b, x_ref, steps):
    from circulantpreconditioner_b200 import glue_binding as G
    n = N_GRID
    out = {"api": "PCShellFFT3DAttach + PCSetUp + PCApply (reference names, src/PCSHELLFft_3D.hxx:23-25) over the C ABI",
           "steps": steps}
    with G.PCShellFFT3D(3, n, n, n, *LAMBDA) as pc:
        # device-resident Vecs (VecCreateSeqCUDA-like; the glue asks VecGetArrayReadAndMemType and gets device pointers)
        vb = G.Vec.from_device_tensor(b)
        xd = torch.empty_like(b)
        vx = G.Vec.from_device_tensor(xd)
        for _ in range(3):
            pc.apply(vb, vx)
        torch.cuda.synchronize()
        # solve_3D handed ctx->Diag to the plan on the first apply; 1 = separable tables (recognised), 2 = N-entry table
        out["symbol_kind"] = pc.symbol_kind()
        out["middle_pass_is_recurrence"] = pc.fast_path()[2] == 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pc.apply(vb, vx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        par = float((torch.linalg.vector_norm(xd - x_ref) / torch.linalg.vector_norm(x_ref)).item())
        out["device_vecs"] = {"value": 1e3 / ms, "ms_per_step": ms, "parity_rel_l2_vs_x_ref": par}
        # host Vecs (what a CPU-only PETSc build hands over): H2D + D2H inside every PCApply
        hb = G.Vec.create_host(n ** 3)
        hb.numpy()[:] = b.cpu().numpy()
        hx = G.Vec.create_host(n ** 3)
        pc.apply(hb, hx)
        nh = max(2, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(nh):
            pc.apply(hb, hx)
        th = (time.perf_counter() - t0) / nh
        xh = torch.from_numpy(hx.numpy()).cuda()
        out["host_vecs"] = {"value": 1.0 / th, "ms_per_step": 1e3 * th, "steps": nh,
                            "parity_rel_l2_vs_x_ref": float((torch.linalg.vector_norm(xh - x_ref) /
                                                             torch.linalg.vector_norm(x_ref)).item())}
        # the same Diag with one entry changed is no longer separable: the N-entry table form of the middle pass
        tab_ms = pc.time_table_form(vb, vx, max(3, min(steps, 10)))
        out["table_symbol"] = {"value": 1e3 / tab_ms, "ms_per_step": tab_ms,
                               "note": "non-separable Diag held as N reciprocals in HBM (176 N bytes per apply)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pcshell", action="store_true")
    ap.add_argument("--no-ksp", action="store_true")
    ap.add_argument("--no-side-reference", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 20:      # defaults sized for the GPU arm; keep the CPU arm to minutes
            args.steps, args.warmup = 3, 1
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
