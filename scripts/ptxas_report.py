#!/usr/bin/env python
"""Static resource table of every kernel in llm-inference-engine_b200/csrc (no GPU needed): registers, spill bytes, static shared memory as
`nvcc -Xptxas -v` reports them for sm_100a.  usage: python scripts/ptxas_report.py > profiles/<tag>_ptxas_resources.txt"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "llm-inference-engine_b200", "csrc")
rows = []
for cu in sorted(glob.glob(os.path.join(SRC, "*.cu"))):
    out = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xptxas", "-v", "-c", cu, "-o",
                          "/dev/null"], capture_output=True, text=True).stderr
    name = None
    spill = (0, 0, 0)
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = tuple(int(x) for x in m.groups())
            continue
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
        if m and name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(.*\)$", "", dem).replace("b200::", "").replace("void ", "")
            rows.append((os.path.basename(cu), dem, int(m.group(1)), spill[1], spill[2], int(m.group(4) or 0)))
            name, spill = None, (0, 0, 0)
print(f"{'file':22s} {'kernel':78s} {'regs':>4s} {'spill st':>8s} {'spill ld':>8s} {'static smem':>11s}")
for f, k, r, st, ld, sm in rows:
    print(f"{f:22s} {k[:78]:78s} {r:4d} {st:8d} {ld:8d} {sm:11d}")
print(f"\n{len(rows)} kernels; {sum(1 for r in rows if r[3] or r[4])} with register spills")
