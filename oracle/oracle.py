"""ctypes binding of oracle/liboracle.so (the CPU restatement of the reference's decoder-layer path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.

All arrays are numpy, fp32 / int32 / uint8, C-contiguous; functions that the reference runs in place
operate in place on the arrays passed in.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref.so")


def build(force=False):
    src = os.path.join(_HERE, "llama_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_xorwow_uniform_subseq0.restype = C.c_float
        _lib.oracle_xorwow_uniform_subseq0.argtypes = [C.c_ulonglong]
        _lib.oracle_e4m3_decode.restype = C.c_float
        _lib.oracle_e4m3_decode.argtypes = [C.c_uint8]
        _lib.oracle_e4m3_encode.restype = C.c_uint8
        _lib.oracle_e4m3_encode.argtypes = [C.c_float]
        _lib.oracle_round_bf16.restype = C.c_float
        _lib.oracle_round_bf16.argtypes = [C.c_float]
    return _lib


def ref_lib():
    """oracle/_ref/libref.so (the unmodified reference kernels + its tests' CPU loops) or None."""
    if not os.path.exists(_REF_PATH):
        return None
    try:
        return C.CDLL(_REF_PATH)
    except OSError:
        return None


def _p(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous numpy"
    return a.ctypes.data_as(C.c_void_p)


def _f(a):
    assert a is None or a.dtype == np.float32
    return _p(a)


def _i(a):
    assert a is None or a.dtype == np.int32
    return _p(a)


def _u8(a):
    assert a is None or a.dtype in (np.uint8, np.bool_)
    return _p(a)


def set_threads(n):
    lib().oracle_set_threads(int(n))


def max_threads():
    return int(lib().oracle_max_threads())


def set_storage(dtype):
    """Storage-type emulation for the composed paths (decoder_layer): 'f32' (default), 'f16' or 'bf16'.  See llama_oracle.c."""
    lib().oracle_set_storage({"f32": 0, "f16": 1, "bf16": 2}[dtype])


def rmsnorm(x, residual, gamma, eps):
    t, h = x.shape
    lib().oracle_rmsnorm(_f(x), _f(residual), _f(gamma), C.c_float(eps), t, h)


def fused_add_bias_residual_rmsnorm(residual, out, bias, gamma, eps):
    t, h = out.shape
    lib().oracle_fused_add_bias_residual_rmsnorm(_f(residual), _f(out), _f(bias), _f(gamma), C.c_float(eps), t, h)


def add_residual(residual, out):
    t, h = out.shape
    lib().oracle_add_residual(_f(residual), _f(out), t, h)


def linear(x, w, layout="nk", wide=False):
    """y[M,N] = x[M,K] @ W.  layout 'kn': w is [K,N]; 'nk': w is [N,K]."""
    m, k = x.shape
    n = w.shape[1] if layout == "kn" else w.shape[0]
    assert (w.shape[0] if layout == "kn" else w.shape[1]) == k
    y = np.zeros((m, n), np.float32)
    lib().oracle_linear(_f(x), _f(w), _f(y), m, k, n, 0 if layout == "kn" else 1, int(wide))
    return y


def batched_gemm(a, b, trans_b):
    batch, m, k = a.shape
    n = b.shape[1] if trans_b else b.shape[2]
    c = np.zeros((batch, m, n), np.float32)
    lib().oracle_batched_gemm(_f(a), _f(b), _f(c), batch, m, n, k, int(trans_b))
    return c


def rope_decode(qkv, head_num, kv_head_num, step, rot_dim, base):
    b, hh, d = qkv.shape
    assert hh == head_num + 2 * kv_head_num
    lib().oracle_rope_decode(_f(qkv), b, head_num, kv_head_num, d, step, rot_dim, C.c_float(base))


def decode_mha(qkv, bias, k_cache, v_cache, head_num, kv_head_num, step, layer):
    b, hh, d = qkv.shape
    s = k_cache.shape[3]
    out = np.zeros((b, head_num * d), np.float32)
    lib().oracle_decode_mha(_f(qkv), _f(bias), _f(k_cache), _f(v_cache), _f(out), b, head_num, kv_head_num, d, s, step, layer)
    return out


def cal_padding_offset(input_lengths, max_q_len, fill=0):
    b = input_lengths.shape[0]
    po = np.full((b, max_q_len), fill, np.int32)
    cum = np.zeros(b + 1, np.int32)
    lib().oracle_cal_padding_offset(_i(po), _i(cum), _i(input_lengths), b, max_q_len)
    return po, cum


def build_causal_masks(q_lens, k_lens, max_q_len, max_k_len):
    b = q_lens.shape[0]
    mask = np.zeros((b, max_q_len, max_k_len), np.float32)
    lib().oracle_build_causal_masks(_f(mask), _i(q_lens), _i(k_lens), b, max_q_len, max_k_len)
    return mask


def qkv_bias_transpose_rope(qkv, padding_offset, history_len, batch, seq_len, head_num, kv_head_num, rot_dim, base):
    t, hh, d = qkv.shape
    q = np.zeros((batch, head_num, seq_len, d), np.float32)
    k = np.zeros((batch, kv_head_num, seq_len, d), np.float32)
    v = np.zeros((batch, kv_head_num, seq_len, d), np.float32)
    lib().oracle_qkv_bias_transpose_rope(_f(q), _f(k), _f(v), _f(qkv), _i(padding_offset), _i(history_len), batch, seq_len, t,
                                         head_num, kv_head_num, d, rot_dim, C.c_float(base))
    return q, k, v


def concat_kv_cache(src, cache, cur_len, history_len, layer):
    b, hkv, mq, d = src.shape
    s = cache.shape[3]
    lib().oracle_concat_kv_cache(_f(src), _f(cache), _i(cur_len), _i(history_len), layer, b, hkv, mq, s, d)


def repeat_kv_cache(cache, context_len, layer, head_num, max_k_len):
    _, b, hkv, s, d = cache.shape
    dst = np.zeros((b, head_num, max_k_len, d), np.float32)
    lib().oracle_repeat_kv_cache(_f(cache), _f(dst), _i(context_len), layer, b, head_num, hkv, max_k_len, s, d)
    return dst


def scale_mask_softmax(qk, mask, scale):
    b, h, ql, kl = qk.shape
    out = np.zeros_like(qk)
    lib().oracle_scale_mask_softmax(_f(qk), _f(mask), _f(out), C.c_float(scale), b, h, ql, kl)
    return out


def transpose_remove_padding(src, padding_offset, num_tokens):
    b, h, s, d = src.shape
    dst = np.zeros((num_tokens, h, d), np.float32)
    lib().oracle_transpose_remove_padding(_f(src), _i(padding_offset), _f(dst), num_tokens, b, s, h, d)
    return dst


def context_attention(q, k_cache, v_cache, padding_offset, input_len, context_len, layer, num_tokens, max_k_len, scale):
    b, h, mq, d = q.shape
    _, _, hkv, s, _ = k_cache.shape
    out = np.zeros((num_tokens, h, d), np.float32)
    lib().oracle_context_attention(_f(q), _f(k_cache), _f(v_cache), _f(out), _i(padding_offset), _i(input_len), _i(context_len),
                                   layer, b, h, hkv, mq, max_k_len, s, d, num_tokens, C.c_float(scale))
    return out


def silu_and_mul(x):
    t, two, inter = x.shape
    assert two == 2
    out = np.zeros((t, inter), np.float32)
    lib().oracle_silu_and_mul(_f(x), _f(out), t, inter)
    return out


def input_embedding(ids, table):
    out = np.zeros((ids.shape[0], table.shape[1]), np.float32)
    lib().oracle_input_embedding(_i(ids), _f(table), _f(out), ids.shape[0], table.shape[1])
    return out


def topk(logits, k):
    rows, vocab = logits.shape
    ids = np.zeros((rows, k), np.int32)
    vals = np.zeros((rows, k), np.float32)
    lib().oracle_topk(_f(logits), _i(ids), _f(vals), rows, vocab, k)
    return ids, vals


def xorwow_uniform_subseq0(seed):
    return float(lib().oracle_xorwow_uniform_subseq0(int(seed)))


def sampling(topk_id, topk_val, seq_len, finished, uniform, end_id, vocab):
    b, k = topk_id.shape
    out = np.zeros(b, np.int32)
    fin = finished.view(np.uint8)
    lib().oracle_sampling(_i(topk_id), _f(topk_val), _i(seq_len), _u8(fin), _i(out), _f(uniform), b, k, end_id, vocab)
    return out


def quantize_fp8(w):
    n, k = w.shape
    q = np.zeros((n, k), np.uint8)
    sc = np.zeros(n, np.float32)
    lib().oracle_quantize_fp8(_f(w), _u8(q), _f(sc), n, k)
    return q, sc


def dequantize_fp8(q, sc):
    n, k = q.shape
    w = np.zeros((n, k), np.float32)
    lib().oracle_dequantize_fp8(_u8(q), _f(sc), _f(w), n, k)
    return w


def quantize_int4(w, group, scale_round=1):
    n, k = w.shape
    q = np.zeros((n, k // 2), np.uint8)
    sc = np.zeros((n, k // group), np.float32)
    z = np.zeros((n, k // group), np.uint8)
    lib().oracle_quantize_int4(_f(w), _u8(q), _f(sc), _u8(z), n, k, group, scale_round)
    return q, sc, z


def dequantize_int4(q, sc, z, group):
    n, kh = q.shape
    w = np.zeros((n, kh * 2), np.float32)
    lib().oracle_dequantize_int4(_u8(q), _f(sc), _u8(z), _f(w), n, kh * 2, group)
    return w


def round_bf16(a):
    """fp32 array -> nearest-even bf16 values, returned as fp32."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32)


def decoder_layer(hidden, w, k_cache, v_cache, cfg, step, layer):
    """One decode layer in place on hidden[B,h].  w: dict of fp32 arrays in [N,K] layout
    (g1, wqkv, bqkv|None, wo, bo|None, g2, wgu, wd); cfg: dict(head_num, kv_head_num, head_size, inter, eps, rot_dim, base)."""
    b, h = hidden.shape
    lib().oracle_decoder_layer(_f(hidden), _f(w["g1"]), _f(w["wqkv"]), _f(w.get("bqkv")), _f(w["wo"]), _f(w.get("bo")),
                               _f(w["g2"]), _f(w["wgu"]), _f(w["wd"]), _f(k_cache), _f(v_cache), b, h, cfg["head_num"],
                               cfg["kv_head_num"], cfg["head_size"], cfg["inter"], k_cache.shape[3], step, layer,
                               C.c_float(cfg["eps"]), cfg["rot_dim"], C.c_float(cfg["base"]))
