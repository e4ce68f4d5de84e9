"""Generate tests/golden/ref_cpu_golden.npz from the REFERENCE'S OWN unit-test CPU loops.

Runs only in the build container: needs oracle/_ref/libref.so (make -C oracle ref), which is compiled from
/root/reference/tests/unit_tests/test_*.cu where they lie.  Each fixture = inputs + the output of the
reference's CPU function on them; tests/test_oracle_golden.py replays the inputs through oracle/llama_oracle.c.
Input patterns follow the reference tests (cited), at reduced sizes so that the fixture stays small.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    ref = oracle.ref_lib()
    assert ref is not None, "build oracle/_ref/libref.so first: make -C oracle ref"
    rng = np.random.default_rng(20261018)
    g = {}

    # test_rmsnorm.cu:44-72: x = i*i%3+1, gamma = i%3+1, eps = 1e-6, h = 4096 (T reduced 64 -> 4)
    t, h = 4, 4096
    idx = np.arange(t * h, dtype=np.int64)
    x = ((idx * idx) % 3 + 1).astype(np.float32).reshape(t, h)
    gamma = (np.arange(h) % 3 + 1).astype(np.float32)
    y = x.copy()
    ref.refcpu_rmsnorm(p(y), p(gamma), C.c_float(1e-6), h, t)
    g["rmsnorm_pattern_out"] = y
    x = rng.standard_normal((3, 512)).astype(np.float32)
    gamma = (1 + 0.1 * rng.standard_normal(512)).astype(np.float32)
    y = x.copy()
    ref.refcpu_rmsnorm(p(y), p(gamma), C.c_float(1e-5), 512, 3)
    g["rmsnorm_rand_x"], g["rmsnorm_rand_gamma"], g["rmsnorm_rand_out"] = x, gamma, y

    # test_add_residual.cu:10-21
    r = rng.standard_normal((5, 256)).astype(np.float32)
    o = rng.standard_normal((5, 256)).astype(np.float32)
    y = o.copy()
    ref.refcpu_add_residual(p(r), p(y), 256, 5)
    g["addres_res"], g["addres_in"], g["addres_out"] = r, o, y

    # test_linear.cu:17-33,54-82: small integers (rand()%3 there), W in [N,K] order
    m, k, n = 4, 256, 96
    x = rng.integers(0, 3, (m, k)).astype(np.float32)
    w = rng.integers(0, 3, (n, k)).astype(np.float32)
    y = np.zeros((m, n), np.float32)
    ref.refcpu_linear(p(x), p(w), p(y), m, k, n)
    g["linear_x"], g["linear_w_nk"], g["linear_out"] = x, w, y
    x = rng.standard_normal((2, 128)).astype(np.float32)
    w = (0.05 * rng.standard_normal((40, 128))).astype(np.float32)
    y = np.zeros((2, 40), np.float32)
    ref.refcpu_linear(p(x), p(w), p(y), 2, 128, 40)
    g["linear_rand_x"], g["linear_rand_w_nk"], g["linear_rand_out"] = x, w, y

    # test_build_causal_mask.cu:13-31,60-66: random lens
    b, mq, mk = 8, 16, 40
    ql = rng.integers(1, mq + 1, b).astype(np.int32)
    kl = rng.integers(1, mk + 1, b).astype(np.int32)
    mask = np.zeros((b, mq, mk), np.float32)
    ref.refcpu_causal_mask(p(mask), p(ql), p(kl), mq, mk, b)
    g["mask_q_lens"], g["mask_k_lens"], g["mask_out"] = ql, kl, mask.astype(np.uint8)

    # test_silu_and_mul.cu:16-32
    x = rng.standard_normal((3, 2, 384)).astype(np.float32) * 3
    y = np.zeros((3, 384), np.float32)
    ref.refcpu_swiglu(p(x), p(y), 3, 384)
    g["swiglu_in"], g["swiglu_out"] = x, y

    # test_qkv_bias_and_rope.cu:14-72 (CPUfunc; head_size must be 128: it hard-codes +64); no padding
    b, s, hn, hkv, d = 2, 4, 3, 1, 128
    qkv = rng.standard_normal((b * s, hn + 2 * hkv, d)).astype(np.float32)
    hist = np.array([0, 5], np.int32)
    ilen = np.full(b, s, np.int32)
    po = np.zeros(b * s, np.int32)
    q = np.zeros((b, hn, s, d), np.float32)
    kk = np.zeros((b, hkv, s, d), np.float32)
    v = np.zeros((b, hkv, s, d), np.float32)
    ref.refcpu_qkv_rope(p(q), p(kk), p(v), p(qkv), p(po), p(hist), p(ilen), b, s, b * s, hn, hkv, d, d, C.c_float(10000.0))
    g["rope_qkv"], g["rope_hist"], g["rope_q"], g["rope_k"] = qkv, hist, q, kk

    # test_input_embedding.cu:15-23
    ids = rng.integers(0, 100, 8).astype(np.int32)
    table = rng.standard_normal((100, 64)).astype(np.float32)
    out = np.zeros((8, 64), np.float32)
    ref.refcpu_embedding(p(ids), p(out), p(table), 8, 64, 100)
    g["emb_ids"], g["emb_table"], g["emb_out"] = ids, table, out

    path = os.path.join(ROOT, "tests", "golden", "ref_cpu_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
