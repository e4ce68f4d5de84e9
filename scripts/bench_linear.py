#!/usr/bin/env python
"""Micro-benchmark of b200_linear on Llama-2-7B shapes: GB/s of weight bytes (decode shapes) and TFLOP/s (prefill shapes).
Weights rotate over enough distinct buffers to exceed the 126 MB L2.  CUDA events on the launching stream."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
mod = importlib.import_module("llm-inference-engine_b200")
mod.lib()
mod.ensure_workspace()
dev = torch.device("cuda")
dt = torch.bfloat16
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0); TF = peaks.get("bf16_tflops", 1590.0)

def bench(M, K, N, reps=20, nbuf=None):
    wbytes = N * K * 2
    nbuf = nbuf or max(2, int(400e6 // wbytes) + 1)
    ws = [torch.randn(N, K, device=dev, dtype=dt) * 0.02 for _ in range(nbuf)]
    x = torch.randn(M, K, device=dev, dtype=dt)
    y = torch.empty(M, N, device=dev, dtype=dt)
    for i in range(3):
        mod.linear(x, ws[i % nbuf], mod.LAYOUT_NK, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        mod.linear(x, ws[i % nbuf], mod.LAYOUT_NK, out=y)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ref = (x.float() @ ws[(reps - 1) % nbuf].float().T)
    err = ((y.float() - ref).norm() / ref.norm()).item()
    return us, wbytes / us / 1e3, 2.0 * M * N * K / us / 1e6, err

shapes = [("qkv", 4096, 12288), ("o", 4096, 4096), ("gate_up", 4096, 22016), ("down", 11008, 4096), ("lm_head", 4096, 32000)]
only = [a for a in sys.argv[1:] if not a.isdigit()]
if only:
    shapes = [sh for sh in shapes if sh[0] in only]
Ms = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 4, 8, 16, 32, 64, 128, 512, 2048]
print(f"peaks: HBM {HBM} GB/s, bf16 {TF} TFLOP/s; B200_FORCE_TC={os.environ.get('B200_FORCE_TC')}")
for M in Ms:
    for name, K, N in shapes:
        us, gbs, tfs, err = bench(M, K, N)
        print(f"M={M:5d} {name:8s} K={K:6d} N={N:6d}  {us:9.2f} us  {gbs:8.1f} GB/s ({gbs / HBM * 100:5.1f}% hbm)  {tfs:8.2f} TFLOP/s ({tfs / TF * 100:5.1f}% tc)  relerr {err:.2e}", flush=True)
