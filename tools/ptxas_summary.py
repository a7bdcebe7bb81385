"""Compile one .cu with -Xptxas -v and print a terse per-kernel table (registers, spills, smem)."""
import re
import subprocess
import sys

src = sys.argv[1]
extra = sys.argv[2:]
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
       "-c", "-o", "/dev/null", src] + extra
out = subprocess.run(cmd, capture_output=True, text=True).stderr
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1), "spill": "0/0"}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["spill"] = f"{m.group(2)}/{m.group(3)}"
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = m.group(1)


def pretty(name):
    m = re.search(r"fft_pass_kernelI([df])((?:Li\d+E)+)", name)
    if m:
        nums = re.findall(r"Li(\d+)E", m.group(2)); xm = "X" if re.search(r"Lb1E", name) else "-"
        keys = ["N", "R0", "R1", "R2", "E", "TX", "G", "MODE", "MINB"]
        return ("f64 " if m.group(1) == "d" else "f32 ") + " ".join(f"{k}={v}" for k, v in zip(keys, nums)) + " xmap=" + xm
    return name[:70]


for r in rows:
    print(f"{pretty(r['name']):70s} regs={r.get('regs','?'):>4s} spill(st/ld)={r['spill']}")
