"""bench.py's reference arm under torch.distributed.run (the way the driver launches it for N > 1): rank 0 alone times the
CPU restatement of the reference path on the FULL 512^3 workload and prints one JSON line with the contract's keys; the
other rank exits 0 without work; the OMP_NUM_THREADS=1 that torchrun exports (the cause of round 1's 4.5x slowdown of
this arm at N > 1) is dropped on rank 0 before numpy loads."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_under_torchrun_world_size_2():
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(34500 + os.getpid() % 1000), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout                      # rank 0 only
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"] == "circulant_pc_applies_per_s_512cube_fp64" and d["unit"] == "applies/s" and d["higher_is_better"]
    assert d["config"]["grid"] == [512, 512, 512]         # the full workload at every N, never a scaled-down sample
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-9 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["omp_num_threads_env"] is None               # torchrun's OMP_NUM_THREADS=1 did not reach numpy / scipy
    assert d["gpu_launches"] == 0
