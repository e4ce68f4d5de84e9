#!/usr/bin/env python
"""bench.py -- decode tokens/s + fraction of HBM roofline for the Llama-2 decoder-layer hot path (BASELINE.json).

A "step" is one decode step over one batch of synthetic input: embedding gather -> 32 fused decoder layers over the KV cache ->
final RMSNorm + LM head -> top-k -> sampling.  Default workload (N=1): BASELINE.json configs[1], Llama-2-7B, 32 layers, bf16,
batch 1, 1024-token context.  With --gpus N > 1 the same model runs tensor-parallel over N ranks (column-sharded QKV / gate-up,
row-sharded O / down, head-sharded KV cache, one all-reduce per attention and per MLP block -- fused into the neighbouring kernels over NVLink
peer memory by default, NCCL with B200_TP_NCCL=1): total work is fixed, so "scaling" is "strong".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 7b|70b] [--batch B] [--ctx C]
                  [--wformat bf16|fp8|int4]

`value`  : device-resident decode tokens/s (inputs in HBM, CUDA-graph replay, timed with CUDA events, max over ranks); the MEDIAN of
           `--regions` (5) timed regions of exactly --steps steps each, all listed in `timed_regions`.
`e2e`    : the same through the C ABI with HOST buffers: token ids copied H2D from pinned memory and sampled ids copied D2H
           inside the timed region, every step (median of 3 regions).
`roofline`: the weight-streaming linears of one step (the dominant kernels: >96 % of a step's bytes at batch 1), issued by ONE C call
           (b200_decoder_linears_only: exactly the engine's launches minus attention; on this rank's shard under tensor parallelism) and
           timed live with CUDA events; algorithmic bytes = the packed weight bytes.  `roofline.kernel` names the kernel
           this configuration dispatches to; `traffic` is a STATIC figure from the committed ncu capture, labelled as such.
`cpu_baseline`: the CPU restatement of the reference's path (oracle/, "port": the reference has no CPU inference path of its own,
           SURVEY.md 8d) with OpenMP on the host cores, bounded sample, extrapolated from one layer; `.single_thread` the same on one thread;
           `.reference_loops` the reference's own unit-test CPU loops (oracle/_ref/libref.so) composed into the layer.  Runs in a CHILD
           process: the process that loads and times libb200llm.so never loads oracle/*.so.
`tp_parity` (--gpus N > 1): before the timed region the sharded engine runs a small Llama-shaped model through the same exchange and every
           rank compares its output with the un-sharded CPU oracle (child process of rank 0); `lm_head`: vocab-sharded under TP.
--impl reference: those reference loops (kind "reference"; written single-threaded, so every linear's weight rows are cut into one block
           per host thread, each through the same unmodified loop) when oracle/_ref/libref.so travelled, with the single-thread figure and
           the all-cores port beside them; the port alone otherwise.  `"extrapolated": true` (one layer + LM head per step, scaled).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "7b": dict(name="Llama-2-7B", hidden=4096, head_num=32, kv_head_num=32, head_size=128, inter=11008, layers=32, vocab=32000),
    "70b": dict(name="Llama-2-70B-shaped", hidden=8192, head_num=64, kv_head_num=8, head_size=128, inter=28672, layers=80, vocab=32000),
    # diagnostics: ONE rank's shard of the 70B-shaped model under TP-8, run on one GPU without the exchange (kernel-level profiling)
    "70b-tp8-rank": dict(name="Llama-2-70B-shaped, one TP-8 shard", hidden=8192, head_num=8, kv_head_num=1, head_size=128, inter=3584, layers=80, vocab=32000),
}
WBYTES = {"bf16": 2.0, "fp8": 1.0, "int4": 0.5}


def algorithmic_bytes(cfg, batch, ctx, wformat, tp=1, group=128, head_sharded=False):
    """Bytes one decode step must read from HBM on ONE rank: packed layer weights (+ quantisation
    scales / zero points) + the KV rows of [0, ctx) + the bf16 LM head.  Activations, gammas, the appended KV row and the
    embedding row (< 0.01 %) are excluded."""
    h, H, Hkv, d, I, L, V = (cfg[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "layers", "vocab"))
    params = h * (H + 2 * Hkv) * d + H * d * h + 3 * h * I
    wb = params * WBYTES[wformat] / tp
    if wformat == "fp8":
        wb += 4 * ((H + 2 * Hkv) * d + h + 2 * I + h) / tp
    if wformat == "int4":
        wb += params / group * 3 / tp
    kv = batch * 2 * (Hkv // tp if Hkv >= tp else 1) * d * ctx * 2
    head = V * h * 2 / (tp if head_sharded else 1)
    return L * (wb + kv) + head, L * wb + head


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled in-process every ~2 ms
    (a timed region can be a few tens of ms); `nvidia-smi -lms` is the fallback when NVML cannot be loaded."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index, uuid=None):
        self.index = index
        self.uuid = uuid
        self.rows = []  # (sm_mhz, sm_max_mhz, reasons bitmask)
        self.proc = None
        self.thread = None
        self.stop = threading.Event()
        self.source = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        if self.uuid:
            for u in (self.uuid, "GPU-" + self.uuid):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(u if isinstance(u, bytes) else u.encode())
                except Exception:
                    try:
                        return pynvml, pynvml.nvmlDeviceGetHandleByUUID(u)
                    except Exception:
                        pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll_nvml(self, nv, h):
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while True:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            if self.stop.wait(0.002):
                break

    def _poll_smi(self):
        names = [n for n, _ in self.REASONS[:4]]
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                mask = 0
                for (n, bit), v in zip(self.REASONS[:4], r[2:6]):
                    if v.lower().startswith("active"):
                        mask |= bit
                self.rows.append((float(r[0]), float(r[1]), mask))
            except Exception:
                pass

    def __enter__(self):
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[2]
        reasons = sorted(n for n, bit in self.REASONS if mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": self.source}


def cpu_reference_tokens_per_s(cfg, batch, ctx, threads, budget_s=20.0):
    """Time the oracle (fp32 restatement of the reference's layer, OpenMP over output columns) on a bounded sample:
    ONE decoder layer + the LM head, extrapolated to layers + head.  Returns (tokens/s, sample description, cores)."""
    import numpy as np

    from oracle import oracle

    oracle.set_threads(threads)
    rng = np.random.default_rng(0)
    h, H, Hkv, d, I, V = (cfg[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "vocab"))
    S = ctx + 8

    def rnd(*shape, scale=1.0):
        return (rng.random(shape, dtype=np.float32) - 0.5) * (2 * scale)

    w = dict(g1=1 + rnd(h, scale=0.1), wqkv=rnd((H + 2 * Hkv) * d, h, scale=0.03), bqkv=None, wo=rnd(h, H * d, scale=0.03), bo=None,
             g2=1 + rnd(h, scale=0.1), wgu=rnd(2 * I, h, scale=0.03), wd=rnd(h, I, scale=0.03))
    kc, vc = rnd(1, batch, Hkv, S, d), rnd(1, batch, Hkv, S, d)
    lm = rnd(V, h, scale=0.03)
    ocfg = dict(head_num=H, kv_head_num=Hkv, head_size=d, inter=I, eps=1e-6, rot_dim=d, base=10000.0)
    x = rnd(batch, h)
    oracle.decoder_layer(x.copy(), w, kc, vc, ocfg, ctx, 0)  # warm caches / page in
    reps, t_layer = 0, 0.0
    t_end = time.perf_counter() + budget_s * 0.7
    while reps < 3 or (time.perf_counter() < t_end and reps < 50):
        xx = x.copy()
        t0 = time.perf_counter()
        oracle.decoder_layer(xx, w, kc, vc, ocfg, ctx, 0)
        t_layer += time.perf_counter() - t0
        reps += 1
    t_layer /= reps
    t0 = time.perf_counter()
    for _ in range(2):
        oracle.linear(x, lm, "nk")
    t_lm = (time.perf_counter() - t0) / 2
    t_step = cfg["layers"] * t_layer + t_lm
    sample = (f"1 of {cfg['layers']} decoder layers ({reps} reps, {t_layer * 1e3:.1f} ms each) + LM head ({t_lm * 1e3:.1f} ms), fp32, "
              f"batch {batch}, ctx {ctx}, extrapolated to {cfg['layers']} layers")
    return batch / t_step, sample, threads


def reference_cpulinear(ref, a, wt, threads=1):
    """y[M,N] = a[M,K] . wt[N,K]^T through the reference's own CPUlinear (test_linear.cu:17-33; it accumulates, so y starts at zero).  The loop
    itself is single-threaded; with threads > 1 the weight ROWS are cut into contiguous blocks and each block is handed to the same unmodified
    loop on its own host thread (ctypes releases the GIL): every output element is still produced by the reference's arithmetic, in its order."""
    import ctypes as C

    import numpy as np

    M, K = a.shape
    N = wt.shape[0]

    def p(t):
        return t.ctypes.data_as(C.c_void_p)

    if threads <= 1 or N < 2 * threads:
        y = np.zeros((M, N), dtype=np.float32)
        ref.refcpu_linear(p(a), p(wt), p(y), M, K, N)
        return y
    from concurrent.futures import ThreadPoolExecutor

    bounds = [N * t // threads for t in range(threads + 1)]
    parts = [np.zeros((M, bounds[t + 1] - bounds[t]), dtype=np.float32) for t in range(threads)]

    def run(t):
        blk = wt[bounds[t]:bounds[t + 1]]  # a view of contiguous rows
        ref.refcpu_linear(p(a), p(blk), p(parts[t]), M, K, blk.shape[0])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(run, range(threads)))
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def reference_loops_decoder_layer(ref, x, w, k_cache, v_cache, cfg, step, threads=1):
    """One decode layer on x[B,h] (fp32, returns the new hidden state) out of the reference's OWN single-threaded unit-test CPU loops,
    compiled unmodified into oracle/_ref/libref.so (`ref`): CPUlinear test_linear.cu:17-33, CPUfusedresidandRMSNorm test_rmsnorm.cu:10-27,
    CPUresidual test_add_residual.cu:10-21, CPUSwiGLU test_silu_and_mul.cu:16-32, in the order of self_decoder.cpp:69-119.  RoPE and the
    decode attention (2 % of the arithmetic at ctx 1024; the reference holds no valid CPU loop for them) come from the C oracle.
    w / cfg as oracle.decoder_layer (no biases); position = step - 1; tests/test_oracle_golden.py checks it against oracle.decoder_layer."""
    import ctypes as C

    import numpy as np

    from oracle import oracle

    H, Hkv, d, I = cfg["head_num"], cfg["kv_head_num"], cfg["head_size"], cfg["inter"]
    batch, h = x.shape
    eps = C.c_float(cfg["eps"])

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    def linear(a, wt):
        return reference_cpulinear(ref, a, wt, threads)

    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    res = x.copy()
    ref.refcpu_rmsnorm(p(x), p(w["g1"]), eps, h, batch)
    qkv = np.ascontiguousarray(linear(x, w["wqkv"]).reshape(batch, H + 2 * Hkv, d))
    oracle.rope_decode(qkv, H, Hkv, step, cfg["rot_dim"], cfg["base"])
    att = np.ascontiguousarray(oracle.decode_mha(qkv, None, k_cache, v_cache, H, Hkv, step, 0).reshape(batch, H * d))
    o = linear(att, w["wo"])
    ref.refcpu_add_residual(p(res), p(o), h, batch)
    res = o.copy()
    ref.refcpu_rmsnorm(p(o), p(w["g2"]), eps, h, batch)
    gu = linear(o, w["wgu"])
    act = np.empty((batch, I), dtype=np.float32)
    ref.refcpu_swiglu(p(gu), p(act), batch, I)
    y = linear(act, w["wd"])
    ref.refcpu_add_residual(p(res), p(y), h, batch)
    return y


class ReferenceLoopsSample:
    """Synthetic one-layer inputs + LM head for the reference's own unit-test CPU loops (reference_loops_decoder_layer).  `available` is False
    where oracle/_ref/libref.so did not travel / does not load.  measure() times ONE decoder layer and the LM head once and returns the
    extrapolated seconds per decode step."""

    def __init__(self, cfg, batch, ctx, threads=1):
        import numpy as np

        from oracle import oracle

        self.ref = oracle.ref_lib()
        self.available = self.ref is not None and hasattr(self.ref, "refcpu_linear")
        self.threads = max(int(threads), 1)
        if not self.available:
            return
        oracle.set_threads(self.threads)
        rng = np.random.default_rng(1)
        h, H, Hkv, d, I, V = (cfg[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "vocab"))
        S = ctx + 8

        def rnd(*shape, scale=1.0):
            return (rng.random(shape, dtype=np.float32) - 0.5) * (2 * scale)

        self.cfg, self.batch, self.ctx, self.V = cfg, batch, ctx, V
        self.w = dict(g1=1 + rnd(h, scale=0.1), wqkv=rnd((H + 2 * Hkv) * d, h, scale=0.03), wo=rnd(h, H * d, scale=0.03),
                      g2=1 + rnd(h, scale=0.1), wgu=rnd(2 * I, h, scale=0.03), wd=rnd(h, I, scale=0.03))
        self.kc, self.vc, self.lm = rnd(1, batch, Hkv, S, d), rnd(1, batch, Hkv, S, d), rnd(V, h, scale=0.03)
        self.ocfg = dict(head_num=H, kv_head_num=Hkv, head_size=d, inter=I, eps=1e-6, rot_dim=d, base=10000.0)
        self.x = rnd(batch, h)
        self.t_layer = self.t_lm = 0.0
        self.reps = 0

    def measure(self):
        import ctypes as C

        import numpy as np

        t0 = time.perf_counter()
        reference_loops_decoder_layer(self.ref, self.x, self.w, self.kc, self.vc, self.ocfg, self.ctx + 1, self.threads)
        t_layer = time.perf_counter() - t0
        t0 = time.perf_counter()
        reference_cpulinear(self.ref, self.x, self.lm, self.threads)
        t_lm = time.perf_counter() - t0
        self.t_layer += t_layer
        self.t_lm += t_lm
        self.reps += 1
        return self.cfg["layers"] * t_layer + t_lm

    def describe(self):
        how = ("single-threaded as written" if self.threads == 1 else
               f"CPUlinear's weight rows cut into {self.threads} blocks, one unmodified loop per host thread")
        return (f"the reference's own unit-test CPU loops (oracle/_ref/libref.so, {how}), 1 of {self.cfg['layers']} decoder "
                f"layers ({self.reps} reps, {self.t_layer / max(self.reps, 1) * 1e3:.1f} ms each) + LM head "
                f"({self.t_lm / max(self.reps, 1) * 1e3:.1f} ms), fp32, batch {self.batch}, ctx {self.ctx}, extrapolated to {self.cfg['layers']} layers; "
                f"RoPE + decode attention from the C oracle")


def reference_loops_tokens_per_s(cfg, batch, ctx, reps=2, threads=1):
    """cpu_baseline.reference_loops of the default arm: a bounded sample (one page-in pass + `reps` timed passes) or None."""
    smp = ReferenceLoopsSample(cfg, batch, ctx, threads)
    if not smp.available:
        return None
    smp.measure()  # page in
    smp.t_layer = smp.t_lm = 0.0
    smp.reps = 0
    t = sum(smp.measure() for _ in range(reps)) / reps
    return {"value": batch / t, "unit": "tokens/s", "cores": smp.threads, "kind": "reference", "sample": smp.describe()}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's own CPU arithmetic for the path on this box's host cores.  The reference has no CPU inference path;
    what it does hold are the single-threaded CPU loops of its unit tests, compiled unmodified into oracle/_ref/libref.so in the build
    container: when that library travelled and loads, the arm times those loops (kind "reference", 1 core: they are written single-threaded)
    and reports the OpenMP port on all cores beside it; otherwise it times the port (kind "port").  Every step is a bounded sample (one
    decoder layer + the LM head, extrapolated to all layers); the whole arm stays under ~2 minutes whatever --steps says."""
    if rank != 0:
        return
    from oracle import oracle

    threads = oracle.max_threads()
    t_all = time.perf_counter()
    n = args.warmup + args.steps
    try:
        smp = ReferenceLoopsSample(cfg, args.batch, args.ctx, threads)
        if smp.available:
            smp.measure()  # page in; also proves the library's loops run here before the arm commits to them
    except Exception as e:
        print(f"[bench] the reference's CPU loops are not usable here ({type(e).__name__}: {e}); timing the port", file=sys.stderr)
        smp = None
    vals = []
    if smp is not None and smp.available:
        for i in range(n):
            if i == args.warmup:
                smp.t_layer = smp.t_lm = 0.0
                smp.reps = 0
            t_step = smp.measure()
            if i >= args.warmup:
                vals.append(args.batch / t_step)
            if time.perf_counter() - t_all > 80 and vals:
                break
        value = sum(vals) / len(vals)
        port_v, port_sample, port_cores = cpu_reference_tokens_per_s(cfg, args.batch, args.ctx, threads, budget_s=8.0)
        single = reference_loops_tokens_per_s(cfg, args.batch, args.ctx, reps=1, threads=1)
        cpu = {"value": value, "unit": "tokens/s", "cores": smp.threads, "kind": "reference", "sample": smp.describe(), "single_thread": single,
               "port_all_cores": {"value": port_v, "unit": "tokens/s", "cores": port_cores, "kind": "port", "sample": port_sample}}
        note = ("the reference has no CPU inference path (its CUDA path is fp32/sm_86 single-GPU, SURVEY.md 8d); this arm times the CPU loops of "
                "its own unit tests (oracle/_ref/libref.so) composed into the decoder layer; cpu_baseline.port_all_cores is the C restatement "
                "(oracle/llama_oracle.c) with OpenMP on all host cores")
    else:
        for i in range(n):
            budget = min(20.0, max(1.5, 80.0 / n))
            v, sample, cores = cpu_reference_tokens_per_s(cfg, args.batch, args.ctx, threads, budget_s=budget)
            if i >= args.warmup:
                vals.append(v)
            if time.perf_counter() - t_all > 90 and vals:
                break
        vals = vals or [v]
        value = sum(vals) / len(vals)
        cpu = {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}
        note = ("the reference has no CPU inference path and its CUDA path is fp32/sm_86 single-GPU (SURVEY.md 8d); oracle/_ref/libref.so is not "
                "loadable here, so this arm times the CPU restatement of its decoder layer (oracle/llama_oracle.c) with OpenMP on all host cores")
    cpu["extrapolated"] = True
    line = {"impl": "reference", "extrapolated": True, "metric": "decode tokens/s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": note + "; EXTRAPOLATED: every timed step is one decoder layer + the LM head, scaled to all layers (a bounded sample, as the "
                    "measurement contract asks); the figure moves with the host's core count and is a stated baseline, not a target"}
    print(json.dumps(line), flush=True)


def cpu_baseline_dict(cfg, batch, ctx):
    """cpu_baseline of the product arm (runs in a child process: `--cpu-leg baseline`)."""
    from oracle import oracle  # checker only: the CPU baseline leg

    v, sample, cores = cpu_reference_tokens_per_s(cfg, batch, ctx, oracle.max_threads(), budget_s=15.0)
    cpu = {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample, "extrapolated": True}
    # the same port on ONE thread: the reference's own CPU loops (tests/unit_tests/*.cu) are single-threaded (SURVEY.md 8d i)
    try:  # extras: never allowed to cost the headline line
        v1, sample1, _ = cpu_reference_tokens_per_s(cfg, batch, ctx, 1, budget_s=3.0)
        cpu["single_thread"] = {"value": v1, "unit": "tokens/s", "cores": 1, "sample": sample1}
        # the reference's own unit-test loops (oracle/_ref) on all host threads, or None where libref.so is not loadable
        cpu["reference_loops"] = reference_loops_tokens_per_s(cfg, batch, ctx, threads=oracle.max_threads())
    except Exception as e:
        cpu["extras_error"] = f"{type(e).__name__}: {e}"
    return cpu


def run_cpu_leg(args, cfg):
    """Child process of the product arm.  `baseline`: prints the cpu_baseline object as one JSON line.  `tp-oracle`: writes the un-sharded
    oracle's outputs for the tensor-parallel parity cases to --cpu-leg-out (npz).  Keeping these in a child keeps oracle/*.so out of the
    process that loads libb200llm.so and times it."""
    if args.cpu_leg == "baseline":
        print(json.dumps(cpu_baseline_dict(cfg, args.batch, args.ctx)), flush=True)
        return
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_decoder_engine import make_model, run_oracle

    out = {}
    for i, (dtype, batch, step) in enumerate(TP_CHECK_CASES):
        ccfg = tp_check_cfg(args.gpus)
        ref, rkc, _ = run_oracle(make_model(ccfg, seed=17), ccfg, dtype, batch, step, storage="f32")
        out[f"ref{i}"], out[f"k{i}"] = ref, rkc[:, :, :, step - 1]
    np.savez(args.cpu_leg_out, **out)


def child(args, leg, extra=()):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", leg, "--gpus", str(args.gpus), "--config", args.config, "--batch", str(args.batch),
           "--ctx", str(args.ctx)] + list(extra)
    if args.layers:
        cmd += ["--layers", str(args.layers)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)


# ---------------------------------------------------------------- tensor-parallel parity leg (runs before the timed region at --gpus > 1)
TP_CHECK_CASES = (("f32", 2, 37), ("bf16", 1, 130), ("bf16", 4, 64), ("bf16", 6, 21))  # (dtype, batch, step): fused GEMVs and the batched path


def tp_check_cfg(world):
    return dict(hidden=1024, head_num=8, kv_head_num=8, head_size=128, inter=2048, layers=3, max_seq=160, eps=1e-6, base=10000.0)


def tp_parity_check(args, mod, tpmod, dist, dev, rank, world, fused):
    """The sharded engine on `world` GPUs, through the same exchange as the timed run, against the UN-SHARDED CPU oracle on the same seeded
    model (a small Llama-shaped model: 8 heads, hidden 1024, 3 layers).  The oracle runs in a child process of rank 0; every rank compares
    its own (replicated) output.  Returns a dict for the JSON line; raises nothing: a failure is reported as ok = false."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    # model / input generators of the parity tests (numpy only); the module-level `from oracle import oracle` there is NOT wanted in this
    # process, so the two generators are restated through a stub module
    import types

    stub = types.ModuleType("oracle")
    stub.oracle = None
    saved = sys.modules.get("oracle")
    sys.modules["oracle"] = stub
    try:
        import test_decoder_engine as tde
    finally:
        if saved is not None:
            sys.modules["oracle"] = saved
        else:
            del sys.modules["oracle"]
    from util import to_dev, to_np

    path = os.path.join("/tmp", f"b200_tp_oracle_{os.environ.get('MASTER_PORT', '0')}_{world}.npz")
    err = ""
    if rank == 0:
        p = child(args, "tp-oracle", ["--cpu-leg-out", path])
        if p.returncode != 0:
            err = "oracle child failed: " + p.stderr[-300:]
    dist.barrier()
    ccfg = tp_check_cfg(world)
    cases, ok = [], not err
    try:
        want = np.load(path) if not err else None
        for i, (dtype, batch, step) in enumerate(TP_CHECK_CASES):
            if want is None:
                break
            model = tde.make_model(ccfg, seed=17)
            lcfg = tpmod.local_cfg(dict(head_num=ccfg["head_num"], kv_head_num=ccfg["kv_head_num"], head_size=ccfg["head_size"], inter=ccfg["inter"]), world)
            dc = mod.DecoderConfig(ccfg["hidden"], lcfg["head_num"], lcfg["kv_head_num"], ccfg["head_size"], lcfg["inter"], ccfg["layers"], ccfg["max_seq"],
                                   batch, {"f32": 0, "f16": 1, "bf16": 2}[dtype], 0, 128, ccfg["eps"], ccfg["head_size"], ccfg["base"], world, rank)
            dec = mod.Decoder(dc, dev)
            for l, w in enumerate(model["layers"]):
                sh = tpmod.shard_layer(w, ccfg, rank, world)
                dec.set_layer(l, dict(g1=to_dev(sh["g1"], dtype), qkv=to_dev(sh["wqkv"], dtype), qkv_bias=to_dev(sh["bqkv"], dtype), o=to_dev(sh["wo"], dtype),
                                      o_bias=to_dev(sh["bo"], dtype), g2=to_dev(sh["g2"], dtype), gate_up=to_dev(sh["wgu"], dtype), down=to_dev(sh["wd"], dtype)))
            x, kc, vc = tde.make_inputs(ccfg, batch, step, model["seed"])
            hidden = to_dev(x, dtype)
            kcd = to_dev(tpmod.shard_kv_cache(kc, ccfg["kv_head_num"], rank, world), dtype)
            vcd = to_dev(tpmod.shard_kv_cache(vc, ccfg["kv_head_num"], rank, world), dtype)
            if fused:
                dec.tp_attach(dist)
                dec.step_tp(hidden, kcd, vcd, step)
            else:
                y_attn, y_ffn = torch.empty_like(hidden), torch.empty_like(hidden)

                def attn_block(l, h, pending):
                    dec.attn_block(l, h, pending, kcd, vcd, y_attn, step)
                    return y_attn

                def ffn_block(l, pending):
                    dec.ffn_block(l, pending, y_ffn)
                    return y_ffn

                tpmod.decode_step_tp(ccfg["layers"], hidden, attn_block, ffn_block, lambda h, pending: dec.fold(h, pending), dist.all_reduce)
            torch.cuda.synchronize()
            got, ref = to_np(hidden).astype(np.float64), want[f"ref{i}"].astype(np.float64)
            e = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
            mine = to_np(kcd)[:, :, :, step - 1].astype(np.float64)
            wk = tpmod.shard_kv_cache(want[f"k{i}"][:, :, :, None], ccfg["kv_head_num"], rank, world)[:, :, :, 0].astype(np.float64)
            ke = float(np.linalg.norm(mine - wk) / np.linalg.norm(wk))
            tol = 1e-5 if dtype == "f32" else 1e-2
            good = bool(np.isfinite(got).all() and e <= tol and ke <= tol and (not fused or dec.tp_error() == 0))
            ok = ok and good
            cases.append({"dtype": dtype, "batch": batch, "step": step, "rel_err": e, "appended_k_rel_err": ke, "tol": tol, "ok": good})
    except Exception as ex:  # a broken check must not take the measurement down with it; it is reported
        ok, err = False, f"{type(ex).__name__}: {ex}"
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    return {"ok": int(flag.item()) == 0, "vs": "un-sharded CPU oracle (child process of rank 0), same seeded model, every rank compares its own output",
            "exchange": "fused NVLink peer-memory exchange" if fused else "nccl all-reduce", "model": ccfg, "cases_rank0": cases, "error": err}


def dominant_linear_kernel(batch, wformat):
    """Name of the kernel the weight-streaming linears of a decode step dispatch to (decoder.cu norm_linear / plain_linear, gemv_f32.cu,
    linear.cu)."""
    if batch == 1 or (wformat != "bf16" and batch <= 4):
        return "gemv_nk_kernel" if wformat == "bf16" else "gemv_q_kernel"
    if batch <= 16:
        return "gemv_mma_kernel"
    if wformat == "bf16":
        return "gemm_tc_kernel (tcgen05, swap-AB + stream-K)" if batch <= 128 else "gemm_tc_kernel (tcgen05)"
    return "gemv_mma_kernel (passes of <= 16 tokens)" if batch <= 64 else "generic SIMT fallback"


def run_prefill(args, cfg, mod, dec, dev, dt, kc, vc, rank):
    """BASELINE configs[2], first half: a T-token prompt per sequence through all layers (b200_decoder_prefill): tensor-core linears +
    tcgen05 context attention.  Prints one JSON line: prefill tokens/s and the fraction of the measured bf16 tensor peak."""
    import torch

    h, H, Hkv, d, I, L = (cfg[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "layers"))
    B, Tq = args.batch, args.prefill_tokens
    T = B * Tq
    assert Tq <= kc.shape[3], "KV cache too short for the prompt: raise --ctx"
    x0 = torch.randn(T, h, device=dev).to(dt)
    x = x0.clone()
    il = torch.full((B,), Tq, dtype=torch.int32, device=dev)
    hl = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            x.copy_(x0)
            dec.prefill(x, kc, vc, il, hl, il, Tq)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(dev.index or 0, gpu_uuid(dev)) as clocks:
            e0.record(stream)
            for _ in range(args.steps):
                dec.prefill(x, kc, vc, il, hl, il, Tq)
            e1.record(stream)
            stream.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
    params = h * (H + 2 * Hkv) * d + H * d * h + 3 * h * I
    flops_lin = 2.0 * T * params * L
    flops_attn = 4.0 * B * H * Tq * Tq * d * L / 2  # causal: only the lower triangle is computed
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved = (flops_lin + flops_attn) / (ms * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps({
            "metric": "prefill tokens/s", "value": T * 1e3 / ms, "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} {L}-layer {args.wformat} prefill, batch {B} x {Tq} tokens", "batch": B, "prompt_tokens": Tq, "weights": args.wformat,
                       "cache_policy": "weights (13 GB) larger than L2"},
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel + context_attn_tc_kernel (whole prefill pass)", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1400",
                         "flops": {"linears": flops_lin, "attention_causal": flops_attn}},
            "clocks": clocks.summary()}), flush=True)


def run_serve(args, cfg, mod, dec, dev, dt, emb, final_gamma, lm_head, kc, vc, L, Hkv, d, V, S):
    """Continuous batching over the paged cache (b200_batcher_*) on a stream of requests of mixed lengths, against static batching of the
    same stream (b200_generate_ragged on groups of --batch requests in arrival order: a group runs until its longest member is done).
    Host ids in, host ids out; eager launches; the scheduler synchronises once per iteration."""
    import numpy as np
    import torch

    B, n_req = args.batch, args.requests
    rng = np.random.default_rng(0)
    plens, news = rng.integers(32, 513, n_req), rng.integers(32, 257, n_req)
    prompts = [rng.integers(3, V, int(n)).astype(np.int32) for n in plens]
    mp = (512 + 256 + 63) // 64
    assert mp * 64 <= S, f"--ctx too small for serve mode: the engine reaches {S} positions, a request up to {mp * 64}"
    num_pages = B * mp + B
    kp = torch.empty((L, num_pages, Hkv, 64, d), dtype=dt, device=dev)
    vp = torch.empty_like(kp)

    def continuous(reqs):
        bat = mod.Batcher(B, num_pages, mp, 2048)
        rids = [bat.submit(prompts[i], int(news[i])) for i in reqs]
        torch.cuda.synchronize()
        t0, it = time.perf_counter(), 0
        while bat.pending():
            bat.step(dec, emb, final_gamma, lm_head, kp, vp, top_k=1, end_id=-1)
            it += 1
        sec = time.perf_counter() - t0
        return sec, it, sum(bat.preemptions(r) for r in rids), [bat.result(r)[0] for r in rids]

    def static(reqs):
        torch.cuda.synchronize()
        t0, out = time.perf_counter(), []
        for g0 in range(0, len(reqs), B):
            grp = list(reqs[g0:g0 + B])
            lens = [int(plens[i]) for i in grp] + [1] * (B - len(grp))
            padded = np.full((B, max(lens)), 3, np.int32)
            for b, i in enumerate(grp):
                padded[b, :lens[b]] = prompts[i]
            n_new = int(max(news[i] for i in grp))
            ids, _ = dec.generate(padded, emb, final_gamma, lm_head, kc, vc, n_new, top_k=1, end_id=-1, prompt_lens=lens)
            out += [ids[b, :int(news[i])] for b, i in enumerate(grp)]
        return time.perf_counter() - t0, out

    warm = list(range(min(2 * B, n_req)))
    continuous(warm)
    static(warm[:B])
    allr = list(range(n_req))
    with ClockSampler(dev.index or 0, gpu_uuid(dev)) as clk:
        sec_c, iters, pre, out_c = continuous(allr)
    sec_s, out_s = static(allr)
    new_tokens = int(news.sum())
    # same greedy ids?  (the batch composition differs between the two schedules, and with it the GEMV kernel: ties inside the bf16
    # tolerance may resolve differently -- reported, not asserted)
    same = sum(int(np.array_equal(a, b_)) for a, b_ in zip(out_c, out_s))
    print(json.dumps({
        "metric": "serving tokens/s (continuous batching, paged KV cache)", "value": new_tokens / sec_c, "unit": "tokens/s", "n_gpus": 1,
        "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{cfg['name']} {L}-layer {args.wformat}, {n_req} requests (prompts 32..512, 32..256 new tokens, all queued at t=0), "
                               f"max_batch {B}, {num_pages} pages of 64 positions", "launch_mode": "eager (PDL-chained), one host sync per iteration"},
        "seconds": sec_c, "iterations": iters, "preemptions": pre, "prompt_tokens": int(plens.sum()), "new_tokens": new_tokens,
        "static_batching": {"what": "b200_generate_ragged on groups of max_batch requests in arrival order (contiguous cache)", "seconds": sec_s,
                            "tokens_per_s": new_tokens / sec_s},
        "speedup_vs_static": sec_s / sec_c, "requests_with_identical_ids": f"{same}/{n_req}", "clocks": clk.summary()}), flush=True)


def gpu_uuid(dev):
    try:
        import torch

        return str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        return None


def workload_config(args, cfg):
    return {"workload": f"{cfg['name']} {cfg['layers']}-layer {args.wformat} decode, batch {args.batch}, {args.ctx}-token context",
            "batch": args.batch, "context": args.ctx, "weights": args.wformat, "kv_cache": "bf16",
            "parallelism": f"tp{args.gpus}" if args.gpus > 1 else "single-gpu",
            "cache_policy": "inputs larger than L2 (13.7 GB of weights + KV per step vs 126 MB L2); no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="7b", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--wformat", default="bf16", choices=list(WBYTES))
    ap.add_argument("--layers", type=int, default=0, help="override the layer count (debug only: makes the number INVALID)")
    ap.add_argument("--requests", type=int, default=64, help="--mode serve: requests in the stream (prompts of 32..512, 32..256 new tokens)")
    ap.add_argument("--mode", default="decode", choices=["decode", "prefill", "generate", "serve"],
                    help="decode (default, the BASELINE metric); prefill: one pass of --prefill-tokens tokens through all layers; generate: "
                         "the whole loop through b200_generate (prompt of --prefill-tokens tokens, --new-tokens sampled tokens, host in / host out)")
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--prefill-tokens", type=int, default=2048)
    ap.add_argument("--preheat", type=float, default=1.5, help="seconds of untimed steps before the warm-up (clock ramp)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tp-check", action="store_true", help="skip the tensor-parallel parity leg that runs before the timed region at --gpus > 1")
    ap.add_argument("--regions", type=int, default=5, help="timed regions of --steps steps each; the median region is reported")
    ap.add_argument("--cpu-leg", default="", choices=["", "baseline", "tp-oracle"], help=argparse.SUPPRESS)  # child processes of this script
    ap.add_argument("--cpu-leg-out", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.mode == "prefill":
        args.ctx = max(args.ctx, args.prefill_tokens)
    if args.mode == "generate":
        args.ctx = max(args.ctx, args.prefill_tokens + args.new_tokens)
    cfg = dict(CONFIGS[args.config])
    if args.layers:
        cfg["layers"] = args.layers
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.cpu_leg:  # child process of the product arm: the only place of that arm where oracle/ is loaded
        run_cpu_leg(args, cfg)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    mod = importlib.import_module("llm-inference-engine_b200")
    mod.lib()  # fails loudly if libb200llm.so is missing
    tpmod = importlib.import_module("llm-inference-engine_b200.tp")
    tp = world
    assert tp == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if tp > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, H, Hkv, d, I, L, V = (cfg[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "layers", "vocab"))
    assert H % tp == 0 and I % tp == 0 and (Hkv % tp == 0), "tensor-parallel degree must divide heads / kv heads / inter"
    Hl, Hkvl, Il = H // tp, Hkv // tp, I // tp
    B, ctx = args.batch, args.ctx
    S = ((ctx + 64 + 127) // 128) * 128
    step = ctx  # positions [0, ctx) are attended: ctx-1 cached rows + the token being appended
    wfmt = {"bf16": mod.W_DENSE, "fp8": mod.W_FP8, "int4": mod.W_INT4}[args.wformat]
    dt = torch.bfloat16

    # ---------------- synthetic model (seeded; every rank generates its own shard from a rank-specific seed)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def randw(n, k):
        w = torch.empty((n, k), dtype=dt, device=dev)
        w.normal_(0.0, 0.02, generator=gen)
        if wfmt == mod.W_FP8:
            return mod.quantize_fp8(w)
        if wfmt == mod.W_INT4:
            return mod.quantize_int4(w, 128)
        return w

    dc = mod.DecoderConfig(h, Hl, Hkvl, d, Il, L, S, B, mod.BF16, wfmt, 128, 1e-5, d, 10000.0, tp, rank)
    dec = mod.Decoder(dc, dev)
    ggen = torch.Generator(device=dev)
    ggen.manual_seed(99)  # replicated tensors: same seed on every rank
    for l in range(L):
        g1 = (1 + 0.1 * torch.randn(h, device=dev, generator=ggen)).to(dt)
        g2 = (1 + 0.1 * torch.randn(h, device=dev, generator=ggen)).to(dt)
        dec.set_layer(l, dict(g1=g1, qkv=randw((Hl + 2 * Hkvl) * d, h), o=randw(h, Hl * d), g2=g2, gate_up=randw(2 * Il, h), down=randw(h, Il)))
    final_gamma = (1 + 0.1 * torch.randn(h, device=dev, generator=ggen)).to(dt)
    lm_head = torch.empty((V, h), dtype=dt, device=dev).normal_(0.0, 0.02, generator=ggen)
    emb = torch.empty((V, h), dtype=dt, device=dev).normal_(0.0, 1.0, generator=ggen)
    kc = torch.empty((L, B, Hkvl, S, d), dtype=dt, device=dev).normal_(0.0, 0.5, generator=gen)
    vc = torch.empty((L, B, Hkvl, S, d), dtype=dt, device=dev).normal_(0.0, 0.5, generator=gen)
    K_TOP, END_ID = 5, 2
    ids_dev = torch.randint(3, V, (B,), dtype=torch.int32, device=dev, generator=ggen)
    hidden = torch.empty((B, h), dtype=dt, device=dev)
    y_attn = torch.empty((B, h), dtype=dt, device=dev)
    y_ffn = torch.empty((B, h), dtype=dt, device=dev)
    bufs = dict(logits=torch.empty((B, V), dtype=torch.float32, device=dev), tmp_ids=torch.empty((B, 8, K_TOP), dtype=torch.int32, device=dev),
                tmp_vals=torch.empty((B, 8, K_TOP), dtype=torch.float32, device=dev), topk_ids=torch.empty((B, K_TOP), dtype=torch.int32, device=dev),
                topk_vals=torch.empty((B, K_TOP), dtype=torch.float32, device=dev), seq_len=torch.full((B,), ctx, dtype=torch.int32, device=dev),
                finished=torch.zeros(B, dtype=torch.uint8, device=dev), output_id=torch.zeros(B, dtype=torch.int32, device=dev))
    # embedding, per layer [norm,] QKV, attention, O, [norm,] gate/up, down (the two norms are fused into the GEMVs at batch 1), fold,
    # [final norm,] LM head (16 tokens per pass), top-k x 2, sampling
    fused_norm = B == 1 or (args.wformat != "bf16" and B <= 4)
    launches_per_step = 1 + L * (5 if fused_norm else 7) + 1 + ((1 if B == 1 else 1 + (B + 15) // 16)) + 2 + 1
    tp_mode = "none"
    if tp > 1:
        tp_mode = "nccl all-reduce"
        if not os.environ.get("B200_TP_NCCL"):
            try:  # one-shot all-reduce over NVLink peer memory, fused into the consuming kernels (no NCCL call on the path)
                dec.tp_attach(dist)
                tp_mode = "fused NVLink peer-memory exchange"
                launches_per_step += 1
            except Exception as e:
                print(f"[bench] fused tensor-parallel exchange unavailable ({e}); using NCCL all-reduce", file=sys.stderr)

    tp_parity = None
    if tp > 1 and not args.no_tp_check and args.mode == "decode":
        tp_parity = tp_parity_check(args, mod, tpmod, dist, dev, rank, tp, tp_mode.startswith("fused"))
        if rank == 0 and not tp_parity["ok"]:
            print(f"[bench] tensor-parallel parity leg FAILED: {json.dumps(tp_parity)[:1500]}", file=sys.stderr)

    sharded_head = None
    if tp > 1 and V % tp == 0 and not os.environ.get("B200_TP_REPLICATED_HEAD"):
        sharded_head = tpmod.VocabShardedHead(mod, dec, tpmod.shard_lm_head_rows(lm_head, rank, tp), V, rank, tp, B, K_TOP, dev)
        launches_per_step += 8  # local top-k x 2, pack x 2, candidate copies x 2, index widen + gather (torch), merge top-k x 2 - replaced tail

    if args.mode == "prefill":
        run_prefill(args, cfg, mod, dec, dev, dt, kc, vc, rank)
        return
    if args.mode == "serve":
        assert tp == 1, "serve mode is single-GPU"
        run_serve(args, cfg, mod, dec, dev, dt, emb, final_gamma, lm_head, kc, vc, L, Hkvl, d, V, S)
        return
    if args.mode == "generate":
        # end to end through the generation loop of the C ABI: host prompt ids in, host token ids out; eager launches (the position
        # changes every step), prefill included in the time
        assert tp == 1, "generate mode is single-GPU"
        prompt = np.random.default_rng(0).integers(3, V, size=(B, args.prefill_tokens)).astype(np.int32)
        times = []
        for i in range(max(args.warmup, 1) + args.steps):
            t0 = time.perf_counter()
            ids, ngen = dec.generate(prompt, emb, final_gamma, lm_head, kc, vc, args.new_tokens, top_k=K_TOP, end_id=-1)
            times.append(time.perf_counter() - t0)
        tt = sorted(times[max(args.warmup, 1):])
        sec = tt[len(tt) // 2]
        print(json.dumps({"metric": "generate tokens/s (end to end, prefill included)", "value": B * args.new_tokens / sec, "unit": "tokens/s",
                          "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_call": sec * 1e3,
                          "ms_per_new_token": sec * 1e3 / args.new_tokens, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": f"{cfg['name']} {L}-layer {args.wformat} b200_generate, batch {B}, prompt {args.prefill_tokens}, "
                                                 f"{args.new_tokens} new tokens", "launch_mode": "eager (PDL-chained)", "sampled": int(ngen.sum())}}), flush=True)
        return

    def decode_step():
        """embedding -> L layers -> fold -> final norm + LM head -> top-k -> sampling; all on the current stream."""
        mod.check(mod.lib().b200_input_embedding(mod.ptr(ids_dev), mod.ptr(emb), mod.ptr(hidden), B, h, mod.BF16, mod.stream()))
        if tp == 1:
            dec.step(hidden, kc, vc, step)
        elif tp_mode.startswith("fused"):
            dec.step_tp(hidden, kc, vc, step)
        else:  # one NCCL all-reduce per attention block and per MLP block (llm-inference-engine_b200/tp.py)
            def attn_block(l, h, pending):
                dec.attn_block(l, h, pending, kc, vc, y_attn, step)
                return y_attn

            def ffn_block(l, pending):
                dec.ffn_block(l, pending, y_ffn)
                return y_ffn

            tpmod.decode_step_tp(L, hidden, attn_block, ffn_block, lambda h, pending: dec.fold(h, pending), dist.all_reduce)
        if sharded_head is not None:  # vocab-sharded LM head: local top-k, one all-gather of k (value, id) pairs, merge, sampling
            sharded_head.run(dist, hidden, final_gamma, bufs["seq_len"], bufs["finished"], bufs["output_id"], step, END_ID)
        else:
            dec.lm_head_topk_sample(hidden, final_gamma, lm_head, bufs, K_TOP, step, END_ID)

    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        for _ in range(2):
            decode_step()
        stream.synchronize()
        if not args.no_graph:
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    decode_step()
                graph.replay()
                stream.synchronize()
            except Exception as e:  # report, and measure the eager path instead
                print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
                graph = None
                torch.cuda.synchronize()

        def run_step():
            if graph is not None:
                graph.replay()
            else:
                decode_step()

        def barrier():
            if tp > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks(ms):
            if tp == 1:
                return ms
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # ---------------- value: device-resident, CUDA events, max over ranks
        # pre-heat: an idle GPU sits at low clocks and a few warm-up steps (tens of ms) do not bring it up; run untimed steps for
        # ~1.5 s first (every rank the same count), then the W warm-up steps the contract asks for
        preheat_steps = 0
        if tp > 1:  # ranks meet inside every step (exchange flags / all-reduce): every rank must run the same number of steps
            preheat_steps = 384 if args.preheat > 0 else 0
            for _ in range(preheat_steps):
                run_step()
            stream.synchronize()
        else:
            t_pre = time.perf_counter()
            while time.perf_counter() - t_pre < args.preheat:
                for _ in range(32):
                    run_step()
                stream.synchronize()
                preheat_steps += 32
        for _ in range(args.warmup):
            run_step()
        barrier()
        # `regions` timed regions of EXACTLY --steps steps, each bracketed by barrier + synchronize, each the max over ranks; the MEDIAN region
        # is the reported one (a single region of a latency-bound tensor-parallel step moved by +-10 % between two looks at the same graph)
        region_ms = []
        with ClockSampler(local_rank, gpu_uuid(dev)) as clocks:
            for _ in range(max(args.regions, 1)):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e0.record(stream)
                for _ in range(args.steps):
                    run_step()
                e1.record(stream)
                barrier()
                region_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        ms_total = sorted(region_ms)[len(region_ms) // 2]
        ms_step = ms_total / args.steps
        value = B * 1e3 / ms_step

        # ---------------- e2e: host buffers through the C ABI, H2D + D2H inside the timed region every step
        ids_host = torch.randint(3, V, (B,), dtype=torch.int32).pin_memory()
        out_host = torch.zeros(B, dtype=torch.int32).pin_memory()
        for _ in range(3):
            ids_dev.copy_(ids_host, non_blocking=True)
            run_step()
            out_host.copy_(bufs["output_id"], non_blocking=True)
            stream.synchronize()
        e2e_regions = []
        for _ in range(3 if args.regions > 1 else 1):
            barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record(stream)
            for _ in range(args.steps):
                ids_dev.copy_(ids_host, non_blocking=True)
                run_step()
                out_host.copy_(bufs["output_id"], non_blocking=True)
                stream.synchronize()  # the next token id is only known once this one is on the host
                ids_host.copy_(out_host.clamp_(min=0))
            e3.record(stream)
            barrier()
            e2e_regions.append(max_over_ranks(e2.elapsed_time(e3)))
        ms_e2e = sorted(e2e_regions)[len(e2e_regions) // 2] / args.steps
        e2e_value = B * 1e3 / ms_e2e

        # ---------------- roofline of the dominant kernel: every GEMV launch of one step, back to back, CUDA events
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        step_bytes, gemv_bytes = algorithmic_bytes(cfg, B, ctx, args.wformat, tp, head_sharded=sharded_head is not None)
        xin = torch.randn(B, h, device=dev).to(dt)
        xin_i = torch.randn(B, Il, device=dev).to(dt)
        xin_a = torch.randn(B, Hl * d, device=dev).to(dt)
        outs = {n: torch.empty((B, n), dtype=dt, device=dev) for n in {(Hl + 2 * Hkvl) * d, h, 2 * Il, V}}

        def lin(x, w, n):
            if isinstance(w, (tuple, list)):
                q, s, z = (list(w) + [None])[:3]
                mod.linear(x, q, mod.LAYOUT_NK, wfmt, s, z, 128, N=n, out=outs[n])
            else:
                mod.linear(x, w, mod.LAYOUT_NK, N=n, out=outs[n])

        lin_launches = [4 * L]

        def gemv_pass():
            # exactly the weight-streaming launches of the step, issued by one C call (no Python between launches); under tensor
            # parallelism the same launches on this rank's shard without the exchange
            lin_launches[0] = dec.linears_only(B)
            lin(xin, lm_head, V)

        for _ in range(2):
            gemv_pass()
        stream.synchronize()
        n_gemv = lin_launches[0] + 1
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        r0.record(stream)
        for _ in range(reps):
            gemv_pass()
        r1.record(stream)
        stream.synchronize()
        gemv_ms = r0.elapsed_time(r1) / reps
        achieved = gemv_bytes / (gemv_ms * 1e-3) / 1e9
        kernel_name = "%s (the %d weight-streaming linears of one step%s, back to back%s) + LM head" % (
            dominant_linear_kernel(B, args.wformat), 4 * L, "" if fused_norm else " with their %d norm kernels" % (2 * L),
            ", this rank's shard, no exchange" if tp > 1 else "")
        # DRAM traffic of the same launches: a STATIC figure from the committed ncu capture of this command (read + write bytes over
        # algorithmic bytes), not measured in this run -- only quoted for the configuration that was captured
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if args.config == "7b" and args.wformat == "bf16" and B == 1 and tp == 1:
                traffic = gemv_bytes * float(tr["traffic_over_algorithmic"])
                traffic_src = ("static: profiles/r2_traffic.json (ncu --set full capture of this command: dram__bytes_read+write of the QKV / O / "
                               "gate_up / down launches over their algorithmic bytes, scaled to the step); not measured in this run")
        except Exception:
            pass

    if rank == 0:
        cpu = None
        if tp == 1 and not args.no_cpu_baseline:
            # the CPU leg runs in a child process: this process (the one that loaded libb200llm.so and was timed) never loads oracle/*.so
            try:
                pr = child(args, "baseline")
                cpu = json.loads(pr.stdout.strip().splitlines()[-1])
            except Exception as e:
                cpu = {"error": f"cpu baseline leg failed: {type(e).__name__}: {e}"}
        line = {
            "metric": "decode tokens/s", "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, cfg),
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4 * B, "ms_per_step": ms_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "launch_mode": "cuda-graph replay" if graph is not None else "eager", "preheat_steps": preheat_steps,
            "tp_exchange": tp_mode, "tp_parity": tp_parity,
            "lm_head": ("vocab-sharded: local top-k + one NCCL all-gather of k (value, id) pairs per rank + merge" if sharded_head is not None else
                        "replicated" if tp > 1 else "single GPU"),
            "timed_regions": {"count": len(region_ms), "steps_each": args.steps, "ms": region_ms, "reported": "median", "e2e_ms": e2e_regions},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "launches": n_gemv, "avg_launch_us": gemv_ms * 1e3 / n_gemv,
                         "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "bytes_per_step_launches": gemv_bytes, "ms": gemv_ms,
                         "whole_step": {"bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                                        "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak_gbs,
                                        "frac_of_8000_nominal": step_bytes / (ms_step * 1e-3) / 1e9 / 8000.0}},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if tp > 1 and tp_mode.startswith("fused") and dec.tp_error():
        print(f"[bench] rank {rank}: the tensor-parallel exchange timed out on a peer: the number above is INVALID", file=sys.stderr)
    if tp > 1:
        # every rank is done once rank 0 has printed; leave without the NCCL teardown (observed to hang after graph-captured
        # collectives on this stack) so the launcher returns immediately
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
