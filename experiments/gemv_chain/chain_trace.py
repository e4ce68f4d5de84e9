"""Diagnostics: per-CTA, per-phase timeline of the chained GEMV kernel on the 7B bf16 B=1 workload (b200_decoder_debug_trace)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mod = importlib.import_module("llm-inference-engine_b200")
mod.lib()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
h, H, Hkv, d, I, L, V = 4096, 32, 32, 128, 11008, 32, 32000
B = int(os.environ.get("TRACE_BATCH", "1"))
ctx = 1024
S = 1152
dt = torch.bfloat16
gen = torch.Generator(device=dev)
gen.manual_seed(1)


def randw(n, k):
    return torch.empty((n, k), dtype=dt, device=dev).normal_(0.0, 0.02, generator=gen)


dc = mod.DecoderConfig(h, H, Hkv, d, I, L, S, B, mod.BF16, mod.W_DENSE, 128, 1e-5, d, 10000.0, 1, 0)
dec = mod.Decoder(dc, dev)
for l in range(L):
    g1 = torch.ones(h, dtype=dt, device=dev)
    dec.set_layer(l, dict(g1=g1, qkv=randw((H + 2 * Hkv) * d, h), o=randw(h, H * d), g2=g1.clone(), gate_up=randw(2 * I, h), down=randw(h, I)))
kc = torch.empty((L, B, Hkv, S, d), dtype=dt, device=dev).normal_(0.0, 0.5, generator=gen)
vc = torch.empty((L, B, Hkv, S, d), dtype=dt, device=dev).normal_(0.0, 0.5, generator=gen)
hidden = torch.randn(B, h, device=dev).to(dt)
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    for _ in range(200):
        dec.step(hidden, kc, vc, ctx)
    st.synchronize()
    tr = dec.debug_trace(True)
    for _ in range(3):
        dec.step(hidden, kc, vc, ctx)
    st.synchronize()
t = tr.cpu().numpy().astype(np.int64).reshape(L, -1, 4, 8)
nsm = t.shape[1]
names = ["O", "gate_up", "down", "qkv"]
print(f"SMs {nsm}; all times in ns; layers 4..27, median / p90 / max over CTAs and layers")


def stat(x):
    x = np.asarray(x).ravel()
    return f"{np.median(x):8.0f} {np.percentile(x, 90):8.0f} {x.max():8.0f}"


lay = slice(4, 28)
for p in range(4):
    T = t[lay, :, p, :]
    print(f"--- phase {p} ({names[p]})")
    if p > 0:
        Tp = t[lay, :, p - 1, :]
        print("  compute-done skew over the grid (max t4 - min t4)  ", stat(Tp[..., 4].max(axis=1) - Tp[..., 4].min(axis=1)))
        print("  reducer: last store - compute warp0 done (t5-t4)   ", stat(Tp[..., 5] - Tp[..., 4]))
        print("  staging done - last reducer store of the grid      ", stat(T[..., 3] - Tp[..., 5].max(axis=1, keepdims=True)))
    print("  CTA-local barrier wait (t2-t0)                     ", stat(T[..., 2] - T[..., 0]))
    print("  staging incl. polling (t3-t2)                      ", stat(T[..., 3] - T[..., 2]))
    print("  compute (t4-t3)                                    ", stat(T[..., 4] - T[..., 3]))
    print("  phase span over the grid: max(t4) - min(t3)        ", stat(T[..., 4].max(axis=1) - T[..., 3].min(axis=1)))
    print("  producer: start issuing - compute start (t1-t3)    ", stat(T[..., 1] - T[..., 3]))
    print("  producer: last copy issued - compute end (t6-t4)   ", stat(T[..., 6] - T[..., 4]))
tot = t[lay, :, 3, 4].max(axis=1) - t[lay, :, 0, 0].min(axis=1)
print("chain kernel span (first t0 .. last qkv compute end)   ", stat(tot))
gap = t[5:28, :, 0, 0].min(axis=1) - t[4:27, :, 3, 4].max(axis=1)
print("between chain kernels (attention + 2 boundaries)       ", stat(gap))
# is the per-SM speed systematic?  compute time of the gate_up phase per smid, correlation between layers
smid = t[10, :, 1, 7]
gu = (t[lay, :, 1, 4] - t[lay, :, 1, 3]).astype(np.float64)  # [layers, cta]
order = np.argsort(smid)
print("smid range", smid.min(), smid.max(), "unique", len(np.unique(smid)))
same_map = all((t[l, :, 1, 7] == smid).all() for l in range(4, 28))
print("blockIdx -> smid mapping identical in every layer:", same_map)
m = gu.mean(axis=0)
c = np.corrcoef(gu)
print("gate_up compute time per CTA: mean over layers min/median/max", m.min(), np.median(m), m.max(),
      "; mean pairwise correlation between layers", (c.sum() - len(c)) / (len(c) * (len(c) - 1)))
print("per-smid mean gate_up compute time (us), sorted by smid:")
print(" ".join(f"{int(s)}:{v / 1e3:.1f}" for s, v in zip(smid[order], m[order])))
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "chain_trace.npz"), t=t)
