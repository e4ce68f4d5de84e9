"""llm-inference-engine_b200 -- Python plumbing over libb200llm.so (the C ABI of include/b200llm.h).

The product is the shared library (hand-written sm_100a CUDA behind a C ABI) and the C++ shim in
`shim/` that re-provides the reference's launch* / layer / weight classes.  This module only binds the
C ABI with ctypes so that tests and bench.py can drive it with torch tensors as device memory.  There is
no CPU or PyTorch fallback: if the library is missing or a call fails, an exception is raised.

The directory name contains a hyphen; import it with
    importlib.import_module("llm-inference-engine_b200")
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200llm.so")

F32, F16, BF16 = 0, 1, 2
W_DENSE, W_FP8, W_INT4 = 0, 1, 2
LAYOUT_KN, LAYOUT_NK = 0, 1
TOPK_BLOCKS, TOPK_MAX_K = 8, 8


class B200Error(RuntimeError):
    pass


class DecoderConfig(C.Structure):
    _fields_ = [(n, t) for n, t in [
        ("hidden", C.c_int), ("head_num", C.c_int), ("kv_head_num", C.c_int), ("head_size", C.c_int),
        ("inter_size", C.c_int), ("num_layers", C.c_int), ("max_seq_len", C.c_int), ("max_batch", C.c_int),
        ("dtype", C.c_int), ("w_format", C.c_int), ("group", C.c_int), ("rmsnorm_eps", C.c_float),
        ("rotary_dim", C.c_int), ("rotary_base", C.c_float), ("tp_world", C.c_int), ("tp_rank", C.c_int)]]


class LinearWeight(C.Structure):
    _fields_ = [("w", C.c_void_p), ("scales", C.c_void_p), ("zeros", C.c_void_p)]


class LayerWeights(C.Structure):
    _fields_ = [("attn_norm_gamma", C.c_void_p), ("qkv", LinearWeight), ("qkv_bias", C.c_void_p), ("o", LinearWeight),
                ("o_bias", C.c_void_p), ("ffn_norm_gamma", C.c_void_p), ("gate_up", LinearWeight), ("down", LinearWeight)]


class GenerateParams(C.Structure):
    """b200_generate_params_t (include/b200llm.h)."""
    _fields_ = [("embedding", C.c_void_p), ("final_gamma", C.c_void_p), ("lm_head", C.c_void_p), ("vocab", C.c_int), ("top_k", C.c_int),
                ("end_id", C.c_int), ("max_new_tokens", C.c_int), ("check_every", C.c_int)]


class BatcherConfig(C.Structure):
    """b200_batcher_config_t (include/b200llm.h)."""
    _fields_ = [("max_batch", C.c_int), ("num_pages", C.c_int), ("max_pages_per_seq", C.c_int), ("max_prefill_tokens", C.c_int)]


class BatchPlan(C.Structure):
    """b200_batch_plan_t (include/b200llm.h)."""
    _fields_ = [("n_prefill", C.c_int), ("prefill_tokens", C.c_int), ("prefill_max_len", C.c_int), ("n_decode", C.c_int),
                ("decode_max_step", C.c_int), ("n_preempted", C.c_int), ("free_pages", C.c_int), ("n_waiting", C.c_int)]


KV_PAGE_SIZE = 64
REQ_WAITING, REQ_RUNNING, REQ_FINISHED, REQ_REJECTED = 0, 1, 2, 3
(PLAN_PREFILL_IDS, PLAN_PREFILL_LENS, PLAN_PREFILL_REQUESTS, PLAN_PREFILL_BLOCK_TABLE, PLAN_PREFILL_LAST_ROWS, PLAN_DECODE_TOKENS,
 PLAN_DECODE_STEPS, PLAN_DECODE_REQUESTS, PLAN_DECODE_BLOCK_TABLE) = range(9)

_P, _I, _F, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_size_t
# name -> argtypes (restype int unless listed in _RESTYPES); must list every symbol of include/b200llm.h
SIGNATURES = {
    "b200_last_error_string": [],
    "b200_abi_version": [],
    "b200_sm_count": [],
    "b200_workspace_default_bytes": [],
    "b200_workspace_set": [_P, _SZ],
    "b200_workspace_ensure": [_SZ],
    "b200_rmsnorm": [_P, _P, _P, _F, _I, _I, _I, _P],
    "b200_fused_add_bias_residual_rmsnorm": [_P, _P, _P, _P, _F, _I, _I, _I, _P],
    "b200_add_residual": [_P, _P, _I, _I, _I, _P],
    "b200_linear": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b200_linear_swiglu": [_P, _P, _P, _I, _I, _I, _I, _P],
    "b200_batched_gemm": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_quantize_fp8": [_P, _P, _P, _I, _I, _I, _P],
    "b200_quantize_int4": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200_dequantize": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "b200_transpose2d": [_P, _P, _I, _I, _I, _P],
    "b200_rope_decode": [_P, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_decode_mha": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_decode_mha_ragged": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_qkv_bias_transpose_rope": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_concat_kv_cache": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b200_repeat_kv_cache": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "b200_scale_mask_softmax": [_P, _P, _P, _F, _I, _I, _I, _I, _I, _P],
    "b200_build_causal_masks": [_P, _P, _P, _I, _I, _I, _I, _P],
    "b200_cal_padding_offset": [_P, _P, _P, _I, _I, _P],
    "b200_transpose_remove_padding": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_context_attention": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_silu_and_mul": [_P, _P, _I, _I, _I, _P],
    "b200_input_embedding": [_P, _P, _P, _I, _I, _I, _P],
    "b200_topk": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200_sampling": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_xorwow_uniform": [_P, _I, C.c_ulonglong, _P],
    "b200_decoder_create": [C.POINTER(DecoderConfig)],
    "b200_decoder_destroy": [_P],
    "b200_decoder_get_config": [_P, C.POINTER(DecoderConfig)],
    "b200_generate_workspace_bytes": [_P, C.POINTER(GenerateParams), _I, _I],
    "b200_generate": [_P, C.POINTER(GenerateParams), _P, _I, _I, _P, _P, _P, _SZ, _P, _P, _P],
    "b200_generate_ragged": [_P, C.POINTER(GenerateParams), _P, _P, _I, _I, _P, _P, _P, _SZ, _P, _P, _P],
    "b200_decoder_set_layer": [_P, _I, C.POINTER(LayerWeights)],
    "b200_decoder_scratch_bytes": [_P],
    "b200_decoder_set_scratch": [_P, _P, _SZ],
    "b200_decoder_step": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200_decoder_step_ragged": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _P],
    "b200_decoder_linears_only": [_P, _I, C.POINTER(C.c_int), _P],
    "b200_decoder_prefill_scratch_bytes": [_P, _I, _I, _I],
    "b200_decoder_prefill": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _SZ, _I, _I, _P],
    "b200_decoder_attn_block": [_P, _I, _P, _P, _P, _P, _P, _I, _I, _P],
    "b200_decoder_ffn_block": [_P, _I, _P, _P, _P, _I, _P],
    "b200_decoder_fold": [_P, _P, _P, _I, _P],
    "b200_decoder_tp_buffer_bytes": [_P],
    "b200_tp_alloc_exported": [_SZ, C.POINTER(C.c_void_p), _P],
    "b200_tp_open": [_P, C.POINTER(C.c_void_p)],
    "b200_decoder_tp_attach": [_P, _I, _I, C.POINTER(C.c_void_p)],
    "b200_decoder_tp_error": [_P],
    "b200_decoder_step_tp": [_P, _P, _P, _P, _I, _I, _P],
    "b200_lm_head_topk_sample": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200_decode_mha_paged": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_context_attention_paged": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_decoder_step_paged": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_decoder_prefill_paged": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _SZ, _I, _I, _P],
    "b200_decoder_prefill_tp": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _SZ, _I, _I, _P, _P, _P],
    "b200_batcher_create": [C.POINTER(BatcherConfig)],
    "b200_batcher_destroy": [_P],
    "b200_batcher_submit": [_P, _P, _I, _I],
    "b200_batcher_plan": [_P, C.POINTER(BatchPlan)],
    "b200_batcher_plan_array": [_P, _I],
    "b200_batcher_commit": [_P, _P, _P, _I],
    "b200_batcher_abort": [_P],
    "b200_batcher_result": [_P, _I, _P, _I, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "b200_batcher_pending": [_P],
    "b200_batcher_free_pages": [_P],
    "b200_batcher_preemptions": [_P, _I],
    "b200_batcher_workspace_bytes": [_P, _P, C.POINTER(GenerateParams)],
    "b200_batcher_step": [_P, _P, C.POINTER(GenerateParams), _P, _P, _P, _SZ, C.POINTER(C.c_int), _P],
}
_RESTYPES = {"b200_last_error_string": C.c_char_p, "b200_workspace_default_bytes": _SZ, "b200_decoder_create": _P,
             "b200_decoder_destroy": None, "b200_decoder_scratch_bytes": _SZ, "b200_decoder_prefill_scratch_bytes": _SZ,
             "b200_decoder_tp_buffer_bytes": _SZ, "b200_generate_workspace_bytes": _SZ, "b200_batcher_create": _P,
             "b200_batcher_destroy": None, "b200_batcher_plan_array": C.POINTER(C.c_int), "b200_batcher_workspace_bytes": _SZ}

_lib = None


def lib():
    """The loaded libb200llm.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (make -C llm-inference-engine_b200/csrc)")
        handle = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise B200Error(f"b200 status {rc}: {lib().b200_last_error_string().decode()}")


# ------------------------------------------------------------------ torch plumbing (device memory + streams only)
def _torch():
    import torch

    return torch


def dtype_code(t):
    torch = _torch()
    return {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}[t.dtype]


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


_ws_holder = {}


def ensure_workspace(nbytes=None):
    """Hand the library a torch-allocated workspace for the current device (once)."""
    torch = _torch()
    dev = torch.cuda.current_device()
    nbytes = nbytes or lib().b200_workspace_default_bytes()
    if dev not in _ws_holder or _ws_holder[dev].numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
        check(lib().b200_workspace_set(ptr(buf), nbytes))
        _ws_holder[dev] = buf


# thin op wrappers (same argument meaning as the reference launchers; see include/b200llm.h)
def rmsnorm(x, residual, gamma, eps):
    check(lib().b200_rmsnorm(ptr(x), ptr(residual), ptr(gamma), eps, x.shape[0], x.shape[1], dtype_code(x), stream()))


def fused_add_bias_residual_rmsnorm(residual, out, bias, gamma, eps):
    check(lib().b200_fused_add_bias_residual_rmsnorm(ptr(residual), ptr(out), ptr(bias), ptr(gamma), eps, out.shape[0], out.shape[1],
                                                     dtype_code(out), stream()))


def add_residual(residual, out):
    check(lib().b200_add_residual(ptr(residual), ptr(out), out.shape[0], out.shape[1], dtype_code(out), stream()))


def linear(x, w, layout=LAYOUT_NK, w_format=W_DENSE, scales=None, zeros=None, group=0, N=None, out=None):
    torch = _torch()
    M, K = x.shape
    if N is None:
        N = w.shape[0] if layout == LAYOUT_NK else w.shape[1]
    y = out if out is not None else torch.empty((M, N), dtype=x.dtype, device=x.device)
    if layout == LAYOUT_KN or M > 4:
        ensure_workspace()  # split-K partials
    check(lib().b200_linear(ptr(x), ptr(w), ptr(scales), ptr(zeros), ptr(y), M, K, N, dtype_code(x), w_format, layout, group, stream()))
    return y


def linear_swiglu(x, w_gate_up):
    """act[M, I] = silu(x . Wgate^T) * (x . Wup^T) in one tensor-core kernel (b200_linear_swiglu): w_gate_up [2I, K] dense, gate rows first."""
    torch = _torch()
    M, K = x.shape
    inter = w_gate_up.shape[0] // 2
    act = torch.empty((M, inter), dtype=x.dtype, device=x.device)
    ensure_workspace()
    check(lib().b200_linear_swiglu(ptr(x), ptr(w_gate_up), ptr(act), M, K, inter, dtype_code(x), stream()))
    return act


def batched_gemm(a, b, trans_b):
    torch = _torch()
    batch, M, K = a.shape
    N = b.shape[1] if trans_b else b.shape[2]
    c = torch.empty((batch, M, N), dtype=a.dtype, device=a.device)
    check(lib().b200_batched_gemm(ptr(a), ptr(b), ptr(c), batch, M, N, K, int(trans_b), dtype_code(a), stream()))
    return c


def quantize_fp8(w):
    torch = _torch()
    N, K = w.shape
    q = torch.empty((N, K), dtype=torch.uint8, device=w.device)
    sc = torch.empty(N, dtype=torch.float32, device=w.device)
    check(lib().b200_quantize_fp8(ptr(w), ptr(q), ptr(sc), N, K, dtype_code(w), stream()))
    return q, sc


def quantize_int4(w, group):
    torch = _torch()
    N, K = w.shape
    q = torch.empty((N, K // 2), dtype=torch.uint8, device=w.device)
    sc = torch.empty((N, K // group), dtype=w.dtype, device=w.device)
    z = torch.empty((N, K // group), dtype=torch.uint8, device=w.device)
    check(lib().b200_quantize_int4(ptr(w), ptr(q), ptr(sc), ptr(z), N, K, group, dtype_code(w), stream()))
    return q, sc, z


def dequantize(q, scales, zeros, w_format, group, dtype, K):
    torch = _torch()
    N = q.shape[0]
    dst = torch.empty((N, K), dtype=dtype, device=q.device)
    check(lib().b200_dequantize(ptr(q), ptr(scales), ptr(zeros), ptr(dst), N, K, w_format, group, dtype_code(dst), stream()))
    return dst


def transpose2d(src):
    torch = _torch()
    rows, cols = src.shape
    dst = torch.empty((cols, rows), dtype=src.dtype, device=src.device)
    check(lib().b200_transpose2d(ptr(src), ptr(dst), rows, cols, dtype_code(src), stream()))
    return dst


def rope_decode(qkv, head_num, kv_head_num, step, rot_dim, base):
    B, _, d = qkv.shape
    check(lib().b200_rope_decode(ptr(qkv), B, head_num, kv_head_num, d, step, rot_dim, base, dtype_code(qkv), stream()))


def decode_mha(qkv, bias, k_cache, v_cache, head_num, kv_head_num, step, layer, apply_rope=False, rot_dim=0, base=10000.0, steps=None):
    """steps (optional): int32 device tensor [B] of per-row positions (ragged batch); `step` must then be >= their maximum."""
    torch = _torch()
    ensure_workspace()
    B, _, d = qkv.shape
    S = k_cache.shape[3]
    out = torch.empty((B, head_num * d), dtype=qkv.dtype, device=qkv.device)
    if steps is not None:
        check(lib().b200_decode_mha_ragged(ptr(qkv), ptr(bias), ptr(k_cache), ptr(v_cache), ptr(out), ptr(steps), B, head_num, kv_head_num, d,
                                           S, step, layer, int(apply_rope), rot_dim, base, dtype_code(qkv), stream()))
        return out
    check(lib().b200_decode_mha(ptr(qkv), ptr(bias), ptr(k_cache), ptr(v_cache), ptr(out), None, B, head_num, kv_head_num, d, S, step,
                                layer, int(apply_rope), rot_dim, base, dtype_code(qkv), stream()))
    return out


def decode_mha_paged(qkv, bias, k_pool, v_pool, block_table, steps, head_num, kv_head_num, layer, apply_rope=False, rot_dim=0, base=10000.0):
    """k_pool / v_pool [L, num_pages, Hkv, 64, d]; block_table int32 [B, max_pages_per_seq] and steps int32 [B] on the device."""
    torch = _torch()
    ensure_workspace()
    B, _, d = qkv.shape
    out = torch.empty((B, head_num * d), dtype=qkv.dtype, device=qkv.device)
    check(lib().b200_decode_mha_paged(ptr(qkv), ptr(bias), ptr(k_pool), ptr(v_pool), ptr(out), ptr(block_table), ptr(steps), B, head_num,
                                      kv_head_num, d, k_pool.shape[1], block_table.shape[1], int(steps.max().item()), layer, int(apply_rope),
                                      rot_dim, base, dtype_code(qkv), stream()))
    return out


def context_attention_paged(q, k_pool, v_pool, block_table, input_len, context_len, layer, num_tokens, scale):
    """q [B, H, max_q_len, d]; pools [L, num_pages, Hkv, 64, d]; returns [num_tokens, H * d] (un-padded)."""
    torch = _torch()
    ensure_workspace()
    B, H, mq, d = q.shape
    out = torch.empty((num_tokens, H * d), dtype=q.dtype, device=q.device)
    check(lib().b200_context_attention_paged(ptr(q), ptr(k_pool), ptr(v_pool), ptr(out), ptr(block_table), ptr(input_len), ptr(context_len), layer,
                                             B, H, k_pool.shape[2], mq, k_pool.shape[1], block_table.shape[1], d, scale, dtype_code(q), stream()))
    return out


def silu_and_mul(x):
    torch = _torch()
    t, _, inter = x.shape
    out = torch.empty((t, inter), dtype=x.dtype, device=x.device)
    check(lib().b200_silu_and_mul(ptr(x), ptr(out), t, inter, dtype_code(x), stream()))
    return out


def input_embedding(ids, table):
    torch = _torch()
    out = torch.empty((ids.shape[0], table.shape[1]), dtype=table.dtype, device=table.device)
    check(lib().b200_input_embedding(ptr(ids), ptr(table), ptr(out), ids.shape[0], table.shape[1], dtype_code(table), stream()))
    return out


def topk(logits, k):
    torch = _torch()
    rows, vocab = logits.shape
    dev = logits.device
    tmp_i = torch.empty((rows, TOPK_BLOCKS, k), dtype=torch.int32, device=dev)
    tmp_v = torch.empty((rows, TOPK_BLOCKS, k), dtype=logits.dtype, device=dev)
    ids = torch.empty((rows, k), dtype=torch.int32, device=dev)
    vals = torch.empty((rows, k), dtype=logits.dtype, device=dev)
    check(lib().b200_topk(ptr(logits), ptr(tmp_i), ptr(tmp_v), ptr(ids), ptr(vals), rows, vocab, k, dtype_code(logits), stream()))
    return ids, vals


def sampling(topk_id, topk_val, seq_len, finished, step, end_id, vocab):
    torch = _torch()
    B, k = topk_id.shape
    out = torch.empty(B, dtype=torch.int32, device=topk_id.device)
    check(lib().b200_sampling(ptr(topk_id), ptr(topk_val), ptr(seq_len), ptr(finished), ptr(out), B, k, step, end_id, vocab,
                              dtype_code(topk_val), stream()))
    return out


def xorwow_uniform(n, seed, device):
    torch = _torch()
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(lib().b200_xorwow_uniform(ptr(out), n, seed, stream()))
    return out


def cal_padding_offset(input_lengths, max_q_len, fill=0):
    torch = _torch()
    B = input_lengths.shape[0]
    po = torch.full((B, max_q_len), fill, dtype=torch.int32, device=input_lengths.device)
    cum = torch.empty(B + 1, dtype=torch.int32, device=input_lengths.device)
    check(lib().b200_cal_padding_offset(ptr(po), ptr(cum), ptr(input_lengths), B, max_q_len, stream()))
    return po, cum


def build_causal_masks(q_lens, k_lens, max_q_len, max_k_len, dtype):
    torch = _torch()
    B = q_lens.shape[0]
    mask = torch.empty((B, max_q_len, max_k_len), dtype=dtype, device=q_lens.device)
    check(lib().b200_build_causal_masks(ptr(mask), ptr(q_lens), ptr(k_lens), B, max_q_len, max_k_len, dtype_code(mask), stream()))
    return mask


def qkv_bias_transpose_rope(qkv, padding_offset, history_len, input_len, batch, seq_len, head_num, kv_head_num, rot_dim, base):
    torch = _torch()
    T, _, d = qkv.shape
    q = torch.zeros((batch, head_num, seq_len, d), dtype=qkv.dtype, device=qkv.device)
    k = torch.zeros((batch, kv_head_num, seq_len, d), dtype=qkv.dtype, device=qkv.device)
    v = torch.zeros((batch, kv_head_num, seq_len, d), dtype=qkv.dtype, device=qkv.device)
    check(lib().b200_qkv_bias_transpose_rope(ptr(q), ptr(k), ptr(v), ptr(qkv), None, ptr(padding_offset), ptr(history_len), ptr(input_len),
                                             batch, seq_len, T, head_num, kv_head_num, d, rot_dim, base, dtype_code(qkv), stream()))
    return q, k, v


def concat_kv_cache(k_src, v_src, k_cache, v_cache, cur_len, history_len, layer):
    B, Hkv, mq, d = k_src.shape
    S = k_cache.shape[3]
    check(lib().b200_concat_kv_cache(ptr(k_src), ptr(v_src), ptr(k_cache), ptr(v_cache), ptr(cur_len), ptr(history_len), layer, B, Hkv, mq,
                                     S, d, dtype_code(k_src), stream()))


def repeat_kv_cache(k_cache, v_cache, context_len, layer, head_num, max_k_len):
    torch = _torch()
    _, B, Hkv, S, d = k_cache.shape
    kd = torch.zeros((B, head_num, max_k_len, d), dtype=k_cache.dtype, device=k_cache.device)
    vd = torch.zeros((B, head_num, max_k_len, d), dtype=k_cache.dtype, device=k_cache.device)
    check(lib().b200_repeat_kv_cache(ptr(k_cache), ptr(v_cache), ptr(kd), ptr(vd), ptr(context_len), layer, B, head_num, Hkv, max_k_len, S,
                                     d, dtype_code(k_cache), stream()))
    return kd, vd


def scale_mask_softmax(qk, mask, scale, out=None):
    torch = _torch()
    B, H, ql, kl = qk.shape
    out = out if out is not None else torch.empty_like(qk)
    check(lib().b200_scale_mask_softmax(ptr(qk), ptr(mask), ptr(out), scale, B, H, ql, kl, dtype_code(qk), stream()))
    return out


def transpose_remove_padding(src, padding_offset, num_tokens):
    torch = _torch()
    B, H, S, d = src.shape
    dst = torch.empty((num_tokens, H, d), dtype=src.dtype, device=src.device)
    check(lib().b200_transpose_remove_padding(ptr(src), ptr(padding_offset), ptr(dst), num_tokens, B, S, H, d, dtype_code(src), stream()))
    return dst


def context_attention(q, k_cache, v_cache, padding_offset, input_len, context_len, layer, num_tokens, scale):
    torch = _torch()
    ensure_workspace()
    B, H, mq, d = q.shape
    _, _, Hkv, S, _ = k_cache.shape
    out = torch.empty((num_tokens, H, d), dtype=q.dtype, device=q.device)
    check(lib().b200_context_attention(ptr(q), ptr(k_cache), ptr(v_cache), ptr(out), ptr(padding_offset), ptr(input_len), ptr(context_len),
                                       layer, B, H, Hkv, mq, S, d, num_tokens, scale, dtype_code(q), stream()))
    return out


class Decoder:
    """The fused decode engine (b200_decoder_*): weights are torch tensors in the packed [N,K] layout."""

    def __init__(self, cfg: DecoderConfig, device):
        torch = _torch()
        self.cfg = cfg
        self.device = device
        ensure_workspace()  # split-K / stream-K partials of the batched (M > 4) linears
        self.handle = lib().b200_decoder_create(C.byref(cfg))
        if not self.handle:
            raise B200Error(lib().b200_last_error_string().decode())
        nbytes = lib().b200_decoder_scratch_bytes(self.handle)
        self.scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        base = (self.scratch.data_ptr() + 255) // 256 * 256
        check(lib().b200_decoder_set_scratch(self.handle, C.c_void_p(base), nbytes))
        self._keep = []

    def set_layer(self, layer, w):
        """w: dict with g1, qkv, o, g2, gate_up, down (each linear = tensor or (q, scales, zeros)), optional qkv_bias, o_bias."""
        def lw(x):
            if isinstance(x, (tuple, list)):
                q, s, z = (list(x) + [None, None])[:3]
                return LinearWeight(q.data_ptr(), s.data_ptr() if s is not None else None, z.data_ptr() if z is not None else None)
            return LinearWeight(x.data_ptr(), None, None)

        def p(x):
            return x.data_ptr() if x is not None else None

        s = LayerWeights(p(w["g1"]), lw(w["qkv"]), p(w.get("qkv_bias")), lw(w["o"]), p(w.get("o_bias")), p(w["g2"]), lw(w["gate_up"]),
                         lw(w["down"]))
        self._keep.append(w)
        check(lib().b200_decoder_set_layer(self.handle, layer, C.byref(s)))

    def step(self, hidden, k_cache, v_cache, step, layer_begin=0, layer_end=None):
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        check(lib().b200_decoder_step(self.handle, ptr(hidden), ptr(k_cache), ptr(v_cache), hidden.shape[0], step, layer_begin, layer_end,
                                      stream()))

    def step_ragged(self, hidden, k_cache, v_cache, steps, max_step, layer_begin=0, layer_end=None):
        """steps: int32 device tensor [batch] of per-row 1-based positions; max_step >= max(steps)."""
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        check(lib().b200_decoder_step_ragged(self.handle, ptr(hidden), ptr(k_cache), ptr(v_cache), hidden.shape[0], ptr(steps), max_step,
                                             layer_begin, layer_end, stream()))

    def step_paged(self, hidden, k_pool, v_pool, block_table, steps, max_step, layer_begin=0, layer_end=None):
        """One decode step over a paged cache: pools [L, num_pages, Hkv, 64, d], block_table int32 [batch, max_pages_per_seq], steps int32 [batch]."""
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        check(lib().b200_decoder_step_paged(self.handle, ptr(hidden), ptr(k_pool), ptr(v_pool), ptr(block_table), ptr(steps), hidden.shape[0],
                                            max_step, k_pool.shape[1], block_table.shape[1], layer_begin, layer_end, stream()))

    def prefill_paged(self, hidden, k_pool, v_pool, block_table, input_len, history_len, context_len, max_q_len, layer_begin=0, layer_end=None):
        torch = _torch()
        ensure_workspace()
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        B, T = input_len.shape[0], hidden.shape[0]
        nbytes = lib().b200_decoder_prefill_scratch_bytes(self.handle, B, max_q_len, T)
        if getattr(self, "_prefill_scratch", None) is None or self._prefill_scratch.numel() < nbytes + 256:
            self._prefill_scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=hidden.device)
        base = (self._prefill_scratch.data_ptr() + 255) // 256 * 256
        check(lib().b200_decoder_prefill_paged(self.handle, ptr(hidden), ptr(k_pool), ptr(v_pool), ptr(block_table), ptr(input_len),
                                               ptr(history_len), ptr(context_len), B, max_q_len, T, k_pool.shape[1], block_table.shape[1],
                                               C.c_void_p(base), nbytes, layer_begin, layer_end, stream()))

    def prefill_tp(self, hidden, k_cache, v_cache, input_len, history_len, context_len, max_q_len, dist_module, layer_begin=0, layer_end=None):
        """Tensor-parallel prefill on this rank's shard (b200_decoder_prefill_tp): torch.distributed.all_reduce is the collective, called
        by the library after the O projection and after the down projection of every layer on the partial [T, hidden] tensor."""
        torch = _torch()
        ensure_workspace()
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        B, T = input_len.shape[0], hidden.shape[0]
        nbytes = lib().b200_decoder_prefill_scratch_bytes(self.handle, B, max_q_len, T)
        if getattr(self, "_prefill_scratch", None) is None or self._prefill_scratch.numel() < nbytes + 256:
            self._prefill_scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=hidden.device)
        scratch = self._prefill_scratch
        base = (scratch.data_ptr() + 255) // 256 * 256
        esize = hidden.element_size()
        failure = []

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p)
        def reduce(buf, count, dtype, user, stream_):
            try:  # the partial lives inside the scratch tensor: all-reduce a view of it (no copy)
                off = buf - scratch.data_ptr()
                view = scratch[off:off + count * esize].view(hidden.dtype)
                dist_module.all_reduce(view)
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                failure.append(e)
                return 1

        rc = lib().b200_decoder_prefill_tp(self.handle, ptr(hidden), ptr(k_cache), ptr(v_cache), ptr(input_len), ptr(history_len),
                                           ptr(context_len), B, max_q_len, T, C.c_void_p(base), nbytes, layer_begin, layer_end, C.cast(reduce, C.c_void_p), None,
                                           stream())
        if failure:
            raise failure[0]
        check(rc)

    def generate(self, prompt_ids, embedding, final_gamma, lm_head, k_cache, v_cache, max_new_tokens, top_k=1, end_id=2, check_every=0,
                 prompt_lens=None):
        """The generation loop (b200_generate / b200_generate_ragged): prompt_ids = int array [batch, prompt_len] on the HOST, prompt_lens
        (optional) = the real length of every row (the rest of a row is padding); returns (ids [batch, max_new_tokens], n_generated [batch])
        as numpy arrays.  embedding / lm_head: [vocab, hidden] tensors of the engine's dtype."""
        import numpy as np

        torch = _torch()
        prompt = np.ascontiguousarray(np.asarray(prompt_ids, dtype=np.int32))
        B, T = prompt.shape
        gp = GenerateParams(C.c_void_p(embedding.data_ptr()), C.c_void_p(final_gamma.data_ptr()), C.c_void_p(lm_head.data_ptr()),
                            int(lm_head.shape[0]), int(top_k), int(end_id), int(max_new_tokens), int(check_every))
        nbytes = lib().b200_generate_workspace_bytes(self.handle, C.byref(gp), B, T)
        if nbytes == 0:
            raise B200Error(lib().b200_last_error_string().decode())
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = (ws.data_ptr() + 255) // 256 * 256
        out = np.empty((B, max_new_tokens), dtype=np.int32)
        ngen = np.empty(B, dtype=np.int32)
        if prompt_lens is not None:
            lens = np.ascontiguousarray(np.asarray(prompt_lens, dtype=np.int32))
            assert lens.shape == (B,)
            check(lib().b200_generate_ragged(self.handle, C.byref(gp), prompt.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), B, T,
                                             ptr(k_cache), ptr(v_cache), C.c_void_p(base), nbytes, out.ctypes.data_as(C.c_void_p),
                                             ngen.ctypes.data_as(C.c_void_p), stream()))
            return out, ngen
        check(lib().b200_generate(self.handle, C.byref(gp), prompt.ctypes.data_as(C.c_void_p), B, T, ptr(k_cache), ptr(v_cache),
                                  C.c_void_p(base), nbytes, out.ctypes.data_as(C.c_void_p), ngen.ctypes.data_as(C.c_void_p), stream()))
        return out, ngen

    def linears_only(self, batch):
        """Diagnostic: the weight-streaming launches of one decode step (no attention, no fold); returns the number of launches."""
        n = C.c_int(0)
        check(lib().b200_decoder_linears_only(self.handle, batch, C.byref(n), stream()))
        return n.value

    def prefill(self, hidden, k_cache, v_cache, input_len, history_len, context_len, max_q_len, layer_begin=0, layer_end=None):
        """hidden [T, h] in/out; input_len / history_len / context_len: int32 device tensors [B]."""
        torch = _torch()
        ensure_workspace()
        layer_end = self.cfg.num_layers if layer_end is None else layer_end
        B, T = input_len.shape[0], hidden.shape[0]
        nbytes = lib().b200_decoder_prefill_scratch_bytes(self.handle, B, max_q_len, T)
        if getattr(self, "_prefill_scratch", None) is None or self._prefill_scratch.numel() < nbytes + 256:
            self._prefill_scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=hidden.device)
        base = (self._prefill_scratch.data_ptr() + 255) // 256 * 256
        check(lib().b200_decoder_prefill(self.handle, ptr(hidden), ptr(k_cache), ptr(v_cache), ptr(input_len), ptr(history_len), ptr(context_len),
                                         B, max_q_len, T, C.c_void_p(base), nbytes, layer_begin, layer_end, stream()))

    def tp_attach(self, dist_module):
        """Fused tensor-parallel exchange: allocate this rank's exchange buffer, swap CUDA-IPC handles with the other ranks of the
        node through torch.distributed (host-side plumbing only) and map theirs."""
        world, rank = self.cfg.tp_world, self.cfg.tp_rank
        nbytes = lib().b200_decoder_tp_buffer_bytes(self.handle)
        mine = C.c_void_p()
        handle = (C.c_char * 64)()
        check(lib().b200_tp_alloc_exported(nbytes, C.byref(mine), handle))
        handles = [None] * world
        dist_module.all_gather_object(handles, bytes(handle.raw))
        bases = (C.c_void_p * world)()
        for r in range(world):
            if r == rank:
                bases[r] = mine.value
            else:
                p = C.c_void_p()
                buf = (C.c_char * 64).from_buffer_copy(handles[r])
                check(lib().b200_tp_open(buf, C.byref(p)))
                bases[r] = p.value
        check(lib().b200_decoder_tp_attach(self.handle, world, rank, bases))
        dist_module.barrier()  # every rank has mapped every buffer before anybody signals into one

    def step_tp(self, hidden, k_cache, v_cache, step):
        check(lib().b200_decoder_step_tp(self.handle, ptr(hidden), ptr(k_cache), ptr(v_cache), hidden.shape[0], step, stream()))

    def tp_error(self):
        return lib().b200_decoder_tp_error(self.handle)

    def attn_block(self, layer, hidden, pending, k_cache, v_cache, partial, step):
        batch = partial.shape[0]
        check(lib().b200_decoder_attn_block(self.handle, layer, ptr(hidden), ptr(pending), ptr(k_cache), ptr(v_cache), ptr(partial), batch,
                                            step, stream()))

    def ffn_block(self, layer, pending, partial):
        check(lib().b200_decoder_ffn_block(self.handle, layer, None, ptr(pending), ptr(partial), partial.shape[0], stream()))

    def fold(self, hidden, pending):
        check(lib().b200_decoder_fold(self.handle, ptr(hidden), ptr(pending), hidden.shape[0], stream()))

    def lm_head_topk_sample(self, hidden, final_gamma, lm_head, bufs, k, step, end_id):
        """bufs: dict(logits[B,V] f32, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len, finished(uint8), output_id)."""
        vocab = lm_head.shape[0]
        g = bufs.get
        check(lib().b200_lm_head_topk_sample(self.handle, ptr(hidden), ptr(final_gamma), ptr(lm_head), vocab, ptr(bufs["logits"]),
                                             ptr(g("tmp_ids")), ptr(g("tmp_vals")), ptr(g("topk_ids")), ptr(g("topk_vals")),
                                             ptr(g("seq_len")), ptr(g("finished")), ptr(g("output_id")), hidden.shape[0], k, step, end_id,
                                             stream()))

    def __del__(self):
        try:
            if self.handle:
                lib().b200_decoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Batcher:
    """Continuous batching over a paged KV cache (b200_batcher_*): the scheduler half (submit / plan / commit) is host only."""

    def __init__(self, max_batch, num_pages, max_pages_per_seq, max_prefill_tokens):
        self.cfg = BatcherConfig(max_batch, num_pages, max_pages_per_seq, max_prefill_tokens)
        self.handle = lib().b200_batcher_create(C.byref(self.cfg))
        if not self.handle:
            raise B200Error(lib().b200_last_error_string().decode())
        self._ws = None

    def __del__(self):
        if getattr(self, "handle", None):
            lib().b200_batcher_destroy(self.handle)
            self.handle = None

    def submit(self, prompt_ids, max_new_tokens):
        import numpy as np

        ids = np.ascontiguousarray(np.asarray(prompt_ids, dtype=np.int32))
        rid = lib().b200_batcher_submit(self.handle, ids.ctypes.data_as(C.c_void_p), int(ids.size), int(max_new_tokens))
        if rid < 0:
            raise B200Error(f"b200 status {rid}: {lib().b200_last_error_string().decode()}")
        return rid

    def plan(self):
        import numpy as np

        p = BatchPlan()
        check(lib().b200_batcher_plan(self.handle, C.byref(p)))
        mp = self.cfg.max_pages_per_seq

        def arr(which, n):
            if n == 0:
                return np.zeros(0, np.int32)
            return np.ctypeslib.as_array(lib().b200_batcher_plan_array(self.handle, which), shape=(n,)).copy()

        views = dict(prefill_ids=arr(PLAN_PREFILL_IDS, p.prefill_tokens), prefill_lens=arr(PLAN_PREFILL_LENS, p.n_prefill),
                     prefill_requests=arr(PLAN_PREFILL_REQUESTS, p.n_prefill),
                     prefill_block_table=arr(PLAN_PREFILL_BLOCK_TABLE, p.n_prefill * mp).reshape(p.n_prefill, mp),
                     prefill_last_rows=arr(PLAN_PREFILL_LAST_ROWS, p.n_prefill), decode_tokens=arr(PLAN_DECODE_TOKENS, p.n_decode),
                     decode_steps=arr(PLAN_DECODE_STEPS, p.n_decode), decode_requests=arr(PLAN_DECODE_REQUESTS, p.n_decode),
                     decode_block_table=arr(PLAN_DECODE_BLOCK_TABLE, p.n_decode * mp).reshape(p.n_decode, mp))
        return p, views

    def commit(self, prefill_sampled, decode_sampled, end_id):
        import numpy as np

        a = np.ascontiguousarray(np.asarray(prefill_sampled, dtype=np.int32))
        b = np.ascontiguousarray(np.asarray(decode_sampled, dtype=np.int32))
        n = lib().b200_batcher_commit(self.handle, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), int(end_id))
        if n < 0:
            raise B200Error(f"b200 status {n}: {lib().b200_last_error_string().decode()}")
        return n

    def abort(self):
        check(lib().b200_batcher_abort(self.handle))

    def result(self, request, capacity=4096):
        import numpy as np

        out = np.zeros(capacity, np.int32)
        n, st = C.c_int(0), C.c_int(0)
        check(lib().b200_batcher_result(self.handle, request, out.ctypes.data_as(C.c_void_p), capacity, C.byref(n), C.byref(st)))
        return out[:n.value].copy(), st.value

    def pending(self):
        return lib().b200_batcher_pending(self.handle)

    def free_pages(self):
        return lib().b200_batcher_free_pages(self.handle)

    def preemptions(self, request):
        return lib().b200_batcher_preemptions(self.handle, request)

    def step(self, dec, embedding, final_gamma, lm_head, k_pool, v_pool, top_k=1, end_id=2):
        """One iteration on the GPU (b200_batcher_step); returns the number of requests it finished."""
        torch = _torch()
        ensure_workspace()
        gp = GenerateParams(C.c_void_p(embedding.data_ptr()), C.c_void_p(final_gamma.data_ptr()), C.c_void_p(lm_head.data_ptr()),
                            int(lm_head.shape[0]), int(top_k), int(end_id), 1, 0)
        if self._ws is None:
            nbytes = lib().b200_batcher_workspace_bytes(self.handle, dec.handle, C.byref(gp))
            if nbytes == 0:
                raise B200Error(lib().b200_last_error_string().decode())
            self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=k_pool.device)
            self._ws_bytes = nbytes
        base = (self._ws.data_ptr() + 255) // 256 * 256
        fin = C.c_int(0)
        check(lib().b200_batcher_step(self.handle, dec.handle, C.byref(gp), ptr(k_pool), ptr(v_pool), C.c_void_p(base), self._ws_bytes,
                                      C.byref(fin), stream()))
        return fin.value
