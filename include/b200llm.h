/*
 * b200llm.h -- C ABI of the B200-native (sm_100a) Llama-2 decoder-layer hot path.
 *
 * This is the drop-in boundary for chongchen1999/llm-inference-engine: one entry point per
 * `launch*` function of the reference's src/kernels/includes/ headers (cited per function below,
 * paths relative to the reference root), plus the fused decode engine that the reference's
 * src/layers classes map onto.  Plain pointers and sizes only: no C++ types, no torch types,
 * no exceptions across the boundary.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - tensors are row-major, innermost dimension last, exactly as the reference lays them out;
 *   - `dtype` is the activation / KV-cache / output element type (b200_dtype_t);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference uses everywhere);
 *   - every function returns B200_OK (0) or a negative b200_status_t; the text of the last
 *     error on the calling thread is returned by b200_last_error_string();
 *   - launchers borrow every pointer, never allocate, never free, never synchronise.
 *
 * There is no CPU fallback: if the CUDA launch cannot be made the call fails with an error.
 */
#ifndef B200LLM_H
#define B200LLM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LLM_ABI_VERSION 1

typedef void *b200_stream_t; /* cudaStream_t */

typedef enum {
    B200_OK = 0,
    B200_ERR_INVALID_ARG = -1,  /* shape / alignment / enum violation (reference: LLM_CHECK throw, src/utils/macro.h:74-95) */
    B200_ERR_UNSUPPORTED = -2,  /* combination not implemented */
    B200_ERR_CUDA = -3,         /* cudaGetLastError() != cudaSuccess after the launch */
    B200_ERR_WORKSPACE = -4,    /* library workspace missing or too small */
    B200_ERR_STATE = -5         /* engine handle used out of order */
} b200_status_t;

typedef enum {
    B200_F32 = 0,  /* float            (reference DataType::FP32, src/utils/tensor.h:24-32) */
    B200_F16 = 1,  /* __half           (reference DataType::FP16) */
    B200_BF16 = 2  /* __nv_bfloat16    (new) */
} b200_dtype_t;

/* Weight storage of a linear (reference WeightType, src/weights/includes/base_weights.h:7-12, extended). */
typedef enum {
    B200_W_DENSE = 0,   /* same element type as the activations */
    B200_W_FP8E4M3 = 1, /* 1 byte / weight, fp32 scale per output channel: w = scale[n] * fp8(n,k)           */
    B200_W_INT4 = 2     /* 4 bit / weight, group along K: w = (q - zero[n,k/g]) * scale[n,k/g]; see DESIGN.md */
} b200_wformat_t;

/* Memory order of a linear's weight. */
typedef enum {
    B200_LAYOUT_KN = 0, /* [K,N] row-major: what launchLinearGemm really reads (src/kernels/linear.cu:47-81)        */
    B200_LAYOUT_NK = 1  /* [N,K] row-major: the HF / CPUlinear order (tests/unit_tests/test_linear.cu:25-32);      */
                        /* the engine's packed streaming format                                                    */
} b200_wlayout_t;

const char *b200_last_error_string(void);
int b200_abi_version(void);
/* Number of SMs of the current device (grid sizing is a multiple of this). */
int b200_sm_count(void);

/* ---- library workspace (split-K partials, split-KV partials, tickets) -------------------------------
 * Launchers never allocate.  The workspace is either handed over by the caller or allocated once by
 * b200_workspace_ensure() (cudaMalloc; call it outside stream capture). */
size_t b200_workspace_default_bytes(void);
int b200_workspace_set(void *ptr, size_t bytes);
int b200_workspace_ensure(size_t bytes);

/* ===================================== normalisation / residual ===================================== */

/* launchRMSNorm, src/kernels/includes/rmsnorm.cuh:9-15 (kernel src/kernels/rmsnorm.cu:35-80):
 * residual[t,:] = x[t,:];  x[t,j] = (x[t,j]*gamma[j]) * rsqrt(mean_j(x^2) + eps).  residual may be NULL. */
int b200_rmsnorm(void *x, void *residual, const void *gamma, float eps, int tokens, int hidden,
                 int dtype, b200_stream_t stream);

/* launchFusedAddBiasResidualAndRMSNorm, src/kernels/includes/add_residual_and_rmsnorm.cuh:11-18
 * (kernel src/kernels/add_residual_and_rmsnorm.cu:43-121):
 * o = out + residual; residual = o (before bias); o += bias; out = gamma * o * rsqrt(mean(o^2)+eps).
 * residual, bias and gamma may each be NULL (skip that step; gamma NULL leaves `out` = o un-normalised
 * exactly as the reference does). */
int b200_fused_add_bias_residual_rmsnorm(void *residual, void *out, const void *bias, const void *gamma,
                                         float eps, int tokens, int hidden, int dtype, b200_stream_t stream);

/* launchAddResidual, src/kernels/includes/add_residual.cuh:9-14: out += residual. */
int b200_add_residual(const void *residual, void *out, int tokens, int hidden, int dtype,
                      b200_stream_t stream);

/* ============================================== linears ============================================= */

/* launchLinearGemm, src/kernels/includes/linear.cuh:13-20 (src/kernels/linear.cu:10-87):
 * y[M,N] = x[M,K] * W, fp32 accumulate, alpha=1, beta=0, no bias.
 *   w_layout = B200_LAYOUT_KN: W memory is [K,N] (the reference's behaviour for both trans_b values);
 *   w_layout = B200_LAYOUT_NK: W memory is [N,K].
 *   w_format = DENSE: `w` has element type `dtype`; scales/zeros ignored.
 *   w_format = FP8E4M3: `w` is uint8[N,K] (NK only), `scales` is float[N].
 *   w_format = INT4: `w` is uint8[N,K/2] (NK only; element k of a row sits in byte k/2, low nibble for
 *              even k), `scales` has element type `dtype` [N,K/group], `zeros` is uint8[N,K/group]. */
int b200_linear(const void *x, const void *w, const void *scales, const void *zeros, void *y,
                int M, int K, int N, int dtype, int w_format, int w_layout, int group,
                b200_stream_t stream);

/* launchLinearGemm (gate_up) + launchSiluAndMul in ONE kernel for prefill-sized token counts (src/layers/ffn.cpp:105-129,
 * src/kernels/silu_and_mul.cu:6-41): act[M, inter] = silu(x . Wgate^T) * (x . Wup^T), w_gate_up dense `dtype` [2*inter, K] (gate rows
 * first, the engine's packed layout).  Bit-identical to b200_linear followed by b200_silu_and_mul (gate and up are rounded to `dtype`
 * before the activation, as the two-launcher path stores them); the [M, 2*inter] intermediate is never written.  16-bit dtypes, M > 128,
 * K % 8 == 0, inter >= 128; B200_ERR_UNSUPPORTED otherwise. */
int b200_linear_swiglu(const void *x, const void *w_gate_up, void *act, int M, int K, int inter_size, int dtype,
                       b200_stream_t stream);

/* launchLinearStridedBatchGemm, src/kernels/includes/linear.cuh:22-29 (src/kernels/linear.cu:89-158):
 * for each of `batch` matrices C[M,N] = A[M,K] * op(B); trans_b=0: B is [K,N]; trans_b=1: B is [N,K]
 * and the product is the true A*B^T (the reference's QK^T defect D4 is NOT reproduced). */
int b200_batched_gemm(const void *a, const void *b, void *c, int batch, int M, int N, int K,
                      int trans_b, int dtype, b200_stream_t stream);

/* Weight-only quantisers (device side, used by the weight classes and the tests):
 * src: dense [N,K] (NK) weights of type `dtype`. */
int b200_quantize_fp8(const void *src, void *w_out, float *scales_out, int N, int K, int dtype,
                      b200_stream_t stream);
int b200_quantize_int4(const void *src, void *w_out, void *scales_out, void *zeros_out, int N, int K,
                       int group, int dtype, b200_stream_t stream);
/* Dequantise back to dense [N,K] `dtype` (test oracle helper: "oracle on the dequantised weights"). */
int b200_dequantize(const void *w, const void *scales, const void *zeros, void *dst, int N, int K,
                    int w_format, int group, int dtype, b200_stream_t stream);
/* [K,N] <-> [N,K] re-layout of a dense weight (load-time packing). rows x cols in, cols x rows out. */
int b200_transpose2d(const void *src, void *dst, int rows, int cols, int dtype, b200_stream_t stream);

/* ======================================== rotary / attention ======================================== */

/* launchRope, src/kernels/includes/rope.cuh:12-17 (src/kernels/rope.cu:4-43): in-place rotate-half RoPE
 * of the q and k heads of qkv[B, H+2Hkv, d] at position step-1.  Correct batch stride (reference D5)
 * and each k head rotated once (reference D6). */
int b200_rope_decode(void *qkv, int batch, int head_num, int kv_head_num, int head_size, int step,
                     int rotary_dim, float rotary_base, int dtype, b200_stream_t stream);

/* launchDecoderMaskedMultiHeadAttention, src/kernels/includes/decoder_self_attention.cuh:11-22
 * (kernel src/kernels/decoder_self_attention.cu:56-188).
 * qkv[B, H+2Hkv, d]; qkv_bias[(H+2Hkv)*d] or NULL (added AFTER RoPE, as the reference does);
 * k_cache/v_cache [L, B, Hkv, S, d] (layer offset applied here from `layer`);
 * writes k,v of the new token at position step-1, attends over positions [0, step), softmax with the
 * reference's +1e-6 denominator, out[B, H*d].
 * apply_rope != 0 additionally rotates q,k at position step-1 first (fusion of launchRope).
 * finished may be NULL (unused by the reference kernel too). */
int b200_decode_mha(const void *qkv, const void *qkv_bias, void *k_cache, void *v_cache, void *out,
                    const uint8_t *finished, int batch, int head_num, int kv_head_num, int head_size,
                    int max_seq_len, int step, int layer, int apply_rope, int rotary_dim,
                    float rotary_base, int dtype, b200_stream_t stream);
/* The same over a RAGGED batch (an extension: the reference shares one step across the batch, self_decoder.cpp:33-39 "step" is a CPU
 * int[1]): row b sits at its own 1-based position steps[b] (DEVICE int[batch], clamped to [1, max_step]); max_step >= max(steps) plans
 * the KV splits.  Row b's result is what b200_decode_mha returns for that row at step = steps[b]. */
int b200_decode_mha_ragged(const void *qkv, const void *qkv_bias, void *k_cache, void *v_cache, void *out,
                           const int *steps, int batch, int head_num, int kv_head_num, int head_size,
                           int max_seq_len, int max_step, int layer, int apply_rope, int rotary_dim,
                           float rotary_base, int dtype, b200_stream_t stream);
/* The same over a PAGED cache (the reference's cache is one static [L,B,Hkv,S,d] tensor, src/models/llama/llama.cpp:47-48):
 * k_pool / v_pool [L, num_pages, Hkv, B200_KV_PAGE_SIZE, d]; position p of batch row b lives in page
 * block_table[b * max_pages_per_seq + p / B200_KV_PAGE_SIZE] (DEVICE int32), row p % B200_KV_PAGE_SIZE.  One page of one kv head is
 * exactly one 16 KiB stage of the kernel's K / V ring (16-bit), so paging costs no extra copies -- only the table look-up, made one tile
 * ahead.  steps as in b200_decode_mha_ragged.  Head size 128 only.  Results are bit-identical to the contiguous kernel on the same rows. */
#define B200_KV_PAGE_SIZE 64
int b200_decode_mha_paged(const void *qkv, const void *qkv_bias, void *k_pool, void *v_pool, void *out,
                          const int *block_table, const int *steps, int batch, int head_num, int kv_head_num,
                          int head_size, int num_pages, int max_pages_per_seq, int max_step, int layer,
                          int apply_rope, int rotary_dim, float rotary_base, int dtype, b200_stream_t stream);

/* launchFusedQKVAddBiasAndTransposeAndRope, src/kernels/includes/qkv_bias_and_rope.cuh:12-23
 * (kernel src/kernels/qkv_bias_and_rope.cu:5-79): QKV[T, H+2Hkv, d] -> q[B,H,Sq,d], k,v[B,Hkv,Sq,d]
 * (re-padded via padding_offset), RoPE at position history_len[b] + local_token.  The bias argument is
 * accepted and ignored, exactly like the reference kernel. */
int b200_qkv_bias_transpose_rope(void *q, void *k, void *v, const void *qkv, const void *qkv_bias,
                                 const int *padding_offset, const int *history_len, const int *input_len,
                                 int batch, int seq_len, int num_tokens, int head_num, int kv_head_num,
                                 int head_size, int rotary_dim, float rotary_base, int dtype,
                                 b200_stream_t stream);

/* launchConcatKVCache, src/kernels/includes/concat_past_kv.cuh:9-18 (src/kernels/concat_past_kv.cu:10-89):
 * cache[layer,b,h,history_len[b]+t,:] = src[b,h,t,:] for t < cur_len[b]; K and V in one launch. */
int b200_concat_kv_cache(const void *k_src, const void *v_src, void *k_cache, void *v_cache,
                         const int *cur_query_len, const int *history_len, int layer, int batch,
                         int kv_head_num, int max_q_len, int max_seq_len, int head_size, int dtype,
                         b200_stream_t stream);

/* launchRepeatKVCache, src/kernels/includes/repeat_kv.cuh:9-17 (src/kernels/repeat_kv.cu:13-106):
 * dst[b, h, s, :] = cache[layer, b, h / (H/Hkv), s, :] for s < context_len[b] (the intended semantics;
 * the reference's source index defect D8 is NOT reproduced). */
int b200_repeat_kv_cache(const void *k_cache, const void *v_cache, void *k_dst, void *v_dst,
                         const int *context_len, int layer, int batch, int head_num, int kv_head_num,
                         int max_k_len, int max_seq_len, int head_size, int dtype, b200_stream_t stream);

/* launchFusedScaleMaskAndSoftmax, src/kernels/includes/scale_and_mask_and_softmax.cuh:10-16
 * (kernel src/kernels/scale_and_mask_and_softmax.cu:64-127): s = scale*qk + (1-mask)*(-10000);
 * p = exp(s - max(max_k s, FLT_MIN)) / (sum + 1e-6).  qk/out [B,H,Sq,Sk] (may alias), mask [B,Sq,Sk].
 * Any Sk (the reference is limited to 4096). */
int b200_scale_mask_softmax(const void *qk, const void *mask, void *out, float scale, int batch,
                            int head_num, int q_len, int k_len, int dtype, b200_stream_t stream);

/* launchBuildCausalMasks, src/kernels/includes/build_causal_mask.cuh:9-14 (src/kernels/build_causal_mask.cu:4-42):
 * mask[b,q,k] = (q<q_len[b]) && (k<k_len[b]) && (k <= q + k_len[b]-q_len[b]) as 1/0 of `dtype`. */
int b200_build_causal_masks(void *mask, const int *q_lens, const int *k_lens, int batch, int max_q_len,
                            int max_k_len, int dtype, b200_stream_t stream);

/* launchCalPaddingOffset, src/kernels/includes/cal_padding_offset.cuh:16-20 (src/kernels/cal_padding_offset.cu:17-70):
 * cum_seqlens[B+1], padding_offset[token] = pad slots before that token; entries past the total
 * token count are left untouched, like the reference. */
int b200_cal_padding_offset(int *padding_offset, int *cum_seqlens, const int *input_lengths, int batch,
                            int max_q_len, b200_stream_t stream);

/* launchFusedTransposeAndRemovePadding, src/kernels/includes/transpose_and_remove_padding.cuh:8-13
 * (src/kernels/transpose_and_remove_padding.cu:15-74): [B,H,Sq,d] -> [T,H,d] dropping pad rows. */
int b200_transpose_remove_padding(const void *src, const int *padding_offset, void *dst, int num_tokens,
                                  int batch, int seq_len, int head_num, int head_size, int dtype,
                                  b200_stream_t stream);

/* Fused prefill ("context") attention: causal flash-style attention over q[B,H,Sq,d], k/v cache
 * [L,B,Hkv,S,d] for k positions < context_len[b], without materialising [B,H,Sq,Sk] and without
 * launchRepeatKVCache.  Semantics = the chain src/layers/context_attention.cpp:221-289 with a true QK^T:
 * mask from (input_len, context_len) as in launchBuildCausalMasks, additive -10000, +1e-6 denominator.
 * out[T, H, d] un-padded (padding_offset applied). */
int b200_context_attention(const void *q, const void *k_cache, const void *v_cache, void *out,
                           const int *padding_offset, const int *input_len, const int *context_len,
                           int layer, int batch, int head_num, int kv_head_num, int max_q_len,
                           int max_seq_len, int head_size, int num_tokens, float scale, int dtype,
                           b200_stream_t stream);
/* The same over a paged cache (see b200_decode_mha_paged): a 128-key tile of the tcgen05 kernel is two pages, fetched as four TMA boxes
 * of 64 rows; a page past the context is never requested.  16-bit dtypes, head size 128 (B200_ERR_UNSUPPORTED otherwise). */
int b200_context_attention_paged(const void *q, const void *k_pool, const void *v_pool, void *out, const int *block_table,
                                 const int *input_len, const int *context_len, int layer, int batch, int head_num,
                                 int kv_head_num, int max_q_len, int num_pages, int max_pages_per_seq, int head_size,
                                 float scale, int dtype, b200_stream_t stream);

/* ========================================== MLP / embedding ========================================= */

/* launchSiluAndMul, src/kernels/includes/silu_and_mul.cuh:9-13 (src/kernels/silu_and_mul.cu:6-82):
 * out[t,i] = silu(in[t,0,i]) * in[t,1,i]. */
int b200_silu_and_mul(const void *in, void *out, int tokens, int inter_size, int dtype,
                      b200_stream_t stream);

/* launchInputEmbedding, src/kernels/includes/input_embedding.cuh:7-12 (src/kernels/input_embedding.cu:4-51):
 * out[t,:] = table[ids[t],:]. */
int b200_input_embedding(const int *ids, const void *table, void *out, int tokens, int hidden, int dtype,
                         b200_stream_t stream);

/* =========================================== sampling tail ========================================== */

/* launchTopKForBeamSearch, src/kernels/includes/topk.cuh:44-51 (src/kernels/topk.cu:24-140).
 * logits[rows, vocab] -> final_ids/final_vals[rows, K], descending, ties -> lower id.
 * tmp_ids/tmp_vals: [rows, B200_TOPK_BLOCKS, K] scratch (the reference's round-1 outputs, same shape).
 * K in 1..B200_TOPK_MAX_K (the reference hard-codes 5). */
#define B200_TOPK_BLOCKS 8
#define B200_TOPK_MAX_K 8
int b200_topk(const void *logits, int *tmp_ids, void *tmp_vals, int *final_ids, void *final_vals,
              int rows, int vocab, int k, int dtype, b200_stream_t stream);

/* launchSampling, src/kernels/includes/sampling.cuh:11-19 (src/kernels/sampling.cu:14-102):
 * w_i = exp(val_i - val_0) written back into topk_val; cuRAND XORWOW curand_init(seed=step,
 * subsequence=batch row, offset 0), thr = curand_uniform*sum; first i with (thr -= w_i) < 0, default
 * candidate 0; id %= vocab; ++seq_len unless finished; finished = (id == end_id). */
int b200_sampling(const int *topk_id, void *topk_val, int *seq_len, uint8_t *finished, int *output_id,
                  int batch, int k, int step, int end_id, int vocab, int dtype, b200_stream_t stream);

/* Test helper: out[b] = the first curand_uniform() of curand_init(seed, subsequence=b, offset=0) (XORWOW),
 * i.e. the random number launchSampling draws for batch row b at step == seed. */
int b200_xorwow_uniform(float *out, int n, unsigned long long seed, b200_stream_t stream);

/* ======================================= fused decode engine ========================================
 * The reference's LlamaSelfDecoder<T>::forward (src/layers/self_decoder.cpp:24-122) + the sampling
 * tail of src/models/llama/llama.cpp:247-311 as one stream-ordered, allocation-free, CUDA-graph-
 * capturable sequence of fused kernels.  See DESIGN.md for the kernel list. */

typedef struct b200_decoder b200_decoder_t;

typedef struct {
    int hidden;        /* model hidden size h (= head_num_total * head_size)                       */
    int head_num;      /* q heads held by THIS rank                                                */
    int kv_head_num;   /* kv heads held by THIS rank                                               */
    int head_size;
    int inter_size;    /* FFN intermediate columns held by THIS rank                               */
    int num_layers;
    int max_seq_len;   /* S of the KV cache [L,B,Hkv,S,d]                                          */
    int max_batch;
    int dtype;         /* activations + KV cache                                                   */
    int w_format;      /* b200_wformat_t of the 4 layer linears                                    */
    int group;         /* INT4 group size                                                          */
    float rmsnorm_eps;
    int rotary_dim;
    float rotary_base;
    int tp_world;      /* 1 = no tensor parallelism                                                */
    int tp_rank;
} b200_decoder_config_t;

/* One linear in the engine's packed format: [N,K] row-major (B200_LAYOUT_NK). */
typedef struct {
    const void *w;
    const void *scales; /* FP8: float[N]; INT4: dtype[N,K/group]; DENSE: NULL */
    const void *zeros;  /* INT4: uint8[N,K/group]; else NULL                  */
} b200_linear_weight_t;

typedef struct {
    const void *attn_norm_gamma;   /* [h]                                                        */
    b200_linear_weight_t qkv;      /* N = (H+2Hkv)*d, K = h                                      */
    const void *qkv_bias;          /* [(H+2Hkv)*d] or NULL                                       */
    b200_linear_weight_t o;        /* N = h, K = H*d                                             */
    const void *o_bias;            /* [h] or NULL (added before the FFN norm only, as reference) */
    const void *ffn_norm_gamma;    /* [h]                                                        */
    b200_linear_weight_t gate_up;  /* N = 2*I (gate rows then up rows), K = h                    */
    b200_linear_weight_t down;     /* N = h, K = I                                               */
} b200_layer_weights_t;

b200_decoder_t *b200_decoder_create(const b200_decoder_config_t *cfg);
void b200_decoder_destroy(b200_decoder_t *dec);
/* Copies the configuration the engine was created with. */
int b200_decoder_get_config(const b200_decoder_t *dec, b200_decoder_config_t *cfg);
int b200_decoder_set_layer(b200_decoder_t *dec, int layer, const b200_layer_weights_t *w);
/* Device scratch the engine needs (activations, split-KV partials); caller-owned. */
size_t b200_decoder_scratch_bytes(const b200_decoder_t *dec);
int b200_decoder_set_scratch(b200_decoder_t *dec, void *ptr, size_t bytes);

/* One decode step over layers [layer_begin, layer_end): hidden[B,h] in/out (the residual stream),
 * k_cache/v_cache [L,B,Hkv,S,d], step = 1-based count of tokens including the current one.
 * LlamaSelfDecoder<T>::forward, src/layers/self_decoder.cpp:24-122.  One token: 5 launches per layer (the norms run inside the
 * GEMVs); 2..16 tokens of a 16-bit model: 7 (norm kernel + tensor-core GEMV, every batch reads the weights once); more: tcgen05 GEMM.
 * tp_world == 1 only. */
int b200_decoder_step(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch,
                      int step, int layer_begin, int layer_end, b200_stream_t stream);

/* b200_decoder_step over a ragged batch: row b decodes at its own position steps[b] (DEVICE int[batch]; max_step >= max(steps)).  Only
 * the attention kernel depends on positions; see b200_decode_mha_ragged. */
int b200_decoder_step_ragged(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch,
                             const int *steps, int max_step, int layer_begin, int layer_end, b200_stream_t stream);
/* b200_decoder_step over a paged cache: pools [L, num_pages, Hkv, B200_KV_PAGE_SIZE, d], block_table DEVICE int[batch, max_pages_per_seq],
 * steps DEVICE int[batch].  batch <= max_batch: the pool has no batch dimension, so sequences join and leave the batch between steps
 * without moving a byte of cache (continuous batching, b200_batcher_* below). */
int b200_decoder_step_paged(b200_decoder_t *dec, void *hidden, void *k_pool, void *v_pool, const int *block_table,
                            const int *steps, int batch, int max_step, int num_pages, int max_pages_per_seq,
                            int layer_begin, int layer_end, b200_stream_t stream);

/* Diagnostic (roofline measurement): exactly the weight-streaming launches of b200_decoder_step -- the QKV / O / gate_up / down linears
 * of every layer (reference src/layers/self_attention.cpp:79-86,131-138, src/layers/ffn.cpp:105-139), as the step
 * runs them (with their norm kernels where the step has them: 2+ tokens) -- without attention and without the final fold; a
 * tensor-parallel engine runs them on its shard, without the exchange.  Call after at least one real step; outputs are meaningless.
 * *n_launches (optional) receives the number of kernels launched. */
int b200_decoder_linears_only(b200_decoder_t *dec, int batch, int *n_launches, b200_stream_t stream);

/* Prefill ("context") pass over layers [layer_begin, layer_end): LlamaContextDecoder<T>::forward
 * (src/layers/context_decoder.cpp:58-199): padding offsets -> per layer RMSNorm -> QKV linear -> split/transpose/RoPE ->
 * KV append at history_len -> causal attention over context_len keys -> O linear -> add-bias-residual-RMSNorm -> gate/up ->
 * SwiGLU -> down -> add-residual.  hidden[num_tokens, h] in/out (un-padded tokens, sequences back to back);
 * input_len / history_len / context_len: device int[batch] (context = history + input); max_q_len >= max(input_len).
 * The cache batch dimension is the engine's max_batch; max_q_len <= max_seq_len (rows past the cache slab are never written).  A qkv bias is
 * added with the decode step's convention (RoPE result stored, then + bias: the rows a prompt leaves in the cache are those the decode
 * kernel would have appended; the reference's own prefill launcher drops the bias, which b200_qkv_bias_transpose_rope keeps).  `scratch`:
 * caller-owned device memory of at least b200_decoder_prefill_scratch_bytes(); linears run on the tensor-core GEMM (library workspace
 * must be set). */
size_t b200_decoder_prefill_scratch_bytes(const b200_decoder_t *dec, int batch, int max_q_len, int num_tokens);
int b200_decoder_prefill(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, const int *input_len,
                         const int *history_len, const int *context_len, int batch, int max_q_len, int num_tokens,
                         void *scratch, size_t scratch_bytes, int layer_begin, int layer_end, b200_stream_t stream);
/* b200_decoder_prefill with the K / V rows written into, and read back from, the page pool (16-bit dtypes, head size 128; batch <=
 * max_batch).  Scratch as b200_decoder_prefill_scratch_bytes. */
int b200_decoder_prefill_paged(b200_decoder_t *dec, void *hidden, void *k_pool, void *v_pool, const int *block_table,
                               const int *input_len, const int *history_len, const int *context_len, int batch,
                               int max_q_len, int num_tokens, int num_pages, int max_pages_per_seq, void *scratch,
                               size_t scratch_bytes, int layer_begin, int layer_end, b200_stream_t stream);
/* Tensor-parallel prefill (the reference is single-GPU; north_star's sharding: column-sharded QKV / gate_up, row-sharded O / down,
 * head-sharded cache, one all-reduce per attention and per MLP block): b200_decoder_prefill on this rank's shard, with `reduce` called
 * after the O projection and after the down projection of every layer to all-reduce (sum, in place, on `stream`) the partial
 * [num_tokens, hidden] tensor of `dtype`.  The library does not link NCCL: the host supplies the collective (ncclAllReduce, or
 * torch.distributed.all_reduce from Python).  The callback returns 0 on success. */
typedef int (*b200_allreduce_fn)(void *buf, size_t count, int dtype, void *user, b200_stream_t stream);
int b200_decoder_prefill_tp(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, const int *input_len,
                            const int *history_len, const int *context_len, int batch, int max_q_len, int num_tokens,
                            void *scratch, size_t scratch_bytes, int layer_begin, int layer_end,
                            b200_allreduce_fn reduce, void *user, b200_stream_t stream);

/* Tensor-parallel halves of one layer.  Each leaves this rank's PARTIAL sum of the row-sharded linear
 * in partial[B,h] (`dtype`); the caller all-reduces it (NCCL) and passes the reduced tensor as
 * `pending` to the next call, which folds it into the residual stream (hidden += pending) before its
 * RMSNorm.  pending == NULL means nothing to fold (first layer). */
int b200_decoder_attn_block(b200_decoder_t *dec, int layer, void *hidden, const void *pending,
                            void *k_cache, void *v_cache, void *partial, int batch, int step,
                            b200_stream_t stream);
int b200_decoder_ffn_block(b200_decoder_t *dec, int layer, void *hidden, const void *pending,
                           void *partial, int batch, b200_stream_t stream);
/* hidden += pending (the fold after the last layer). */
int b200_decoder_fold(b200_decoder_t *dec, void *hidden, const void *pending, int batch,
                      b200_stream_t stream);

/* Fused tensor-parallel decode step: no NCCL on the path.  Every rank owns one EXCHANGE BUFFER (b200_decoder_tp_buffer_bytes) that all
 * other ranks of the node map through CUDA IPC (b200_tp_alloc_exported on the owner, b200_tp_open on the peers; the 64-byte handles
 * travel over any host channel, e.g. torch.distributed.all_gather_object).  The row-sharded O / down linears PUSH their partial sums
 * into every rank's buffer from their epilogue as 8-byte {payload, flag} words (posted NVLink stores; an aligned 8-byte store is never
 * torn, so a word whose flag shows the block's sequence number carries valid data).  The first kernel of the next block POLLS those words
 * in its own memory -- no flag exchange, no fence -- and adds the P partials in rank order: a one-shot all-reduce fused into the
 * producing linear's epilogue and the residual-add + RMSNorm of the consuming one (inside the GEMV at one token, in the norm kernel for
 * more).  A peer that never delivers trips a sticky error word after ~2 s: b200_decoder_tp_error() then reads non-zero and the affected
 * results are NaN, not plausible numbers -- check it at every synchronisation point.  bases[r] = rank r's buffer as mapped in THIS
 * process.  Replaces nothing in the reference (single GPU); semantics = src/layers/self_decoder.cpp:69-119 under Megatron sharding. */
size_t b200_decoder_tp_buffer_bytes(const b200_decoder_t *dec);
int b200_tp_alloc_exported(size_t bytes, void **ptr, void *handle64);
int b200_tp_open(const void *handle64, void **ptr);
int b200_decoder_tp_attach(b200_decoder_t *dec, int world, int rank, void *const *bases);
int b200_decoder_tp_error(const b200_decoder_t *dec);
int b200_decoder_step_tp(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch, int step,
                         b200_stream_t stream);

/* Sampling tail: final RMSNorm (gamma) + LM head (dense `dtype` [V,h], NK) -> logits[B,V] (float) ->
 * top-k -> sampling.  logits/topk buffers caller-owned.  src/models/llama/llama.cpp:247-311. */
int b200_lm_head_topk_sample(b200_decoder_t *dec, const void *hidden, const void *final_gamma,
                             const void *lm_head, int vocab, float *logits, int *tmp_ids, float *tmp_vals,
                             int *topk_ids, float *topk_vals, int *seq_len, uint8_t *finished,
                             int *output_id, int batch, int k, int step, int end_id,
                             b200_stream_t stream);

/* ========================================== generation loop ========================================= */

/* The loop LlamaModel<T>::response / generateFirstToken / generateNextToken / LMHeadAndTopKSample intend
 * (src/models/llama/llama.cpp:165-398; dead code in the reference -- SURVEY.md 8f rank 2):
 * prompt -> launchInputEmbedding -> context decoder (KV cache filled from position 0) -> final RMSNorm + LM head on the last prompt
 * token -> top-k -> sampling; then per new token: embedding of the sampled id -> self decoder step -> the same tail.
 * `step` (sampling seed and 1-based position count) starts at the prompt length and grows by one per token, as in the reference. */
typedef struct {
    const void *embedding;   /* [vocab, hidden] of the engine's dtype (pre_decoder_embedding_weight)          */
    const void *final_gamma; /* [hidden] (out_rmsnorm_weight)                                                */
    const void *lm_head;     /* [vocab, hidden] dense, engine dtype (post_decoder_embedding_weight)          */
    int vocab;
    int top_k;               /* 1..B200_TOPK_MAX_K; 1 = greedy                                               */
    int end_id;
    int max_new_tokens;      /* tokens to produce per sequence, including the one sampled from the prompt    */
    int check_every;         /* every this many steps, stop if every sequence has produced end_id (0: always run max_new_tokens) */
} b200_generate_params_t;

/* Device workspace b200_generate needs for this batch / prompt length (0 on a bad argument: see b200_last_error_string). */
size_t b200_generate_workspace_bytes(const b200_decoder_t *dec, const b200_generate_params_t *p, int batch, int prompt_len);

/* prompt_ids: HOST int[batch, prompt_len] (every sequence the same length -- b200_generate_ragged takes prompts of different lengths;
 * batch must equal the engine's max_batch, which is the cache's batch dimension).  k_cache / v_cache: caller-owned [L, batch, Hkv, max_seq_len, d], overwritten from position 0.
 * out_ids: HOST int[batch, max_new_tokens]: the sampled ids; everything from a sequence's first end_id on is end_id.
 * n_generated (optional): HOST int[batch], tokens before the first end_id.  Synchronises the stream before returning. */
int b200_generate(b200_decoder_t *dec, const b200_generate_params_t *p, const int *prompt_ids, int batch, int prompt_len,
                  void *k_cache, void *v_cache, void *workspace, size_t workspace_bytes, int *out_ids, int *n_generated,
                  b200_stream_t stream);

/* The same loop over prompts of DIFFERENT lengths (the "dynamic prompt length" of SURVEY.md 8f rank 2; the reference hard-codes one
 * 13-token prompt, llama.cpp:327-341).  prompt_ids: HOST int[batch, prompt_len], row b holds prompt_lens[b] ids followed by padding that
 * is never read; prompt_lens: HOST int[batch], 1 <= prompt_lens[b] <= prompt_len (NULL: all rows prompt_len long = b200_generate).
 * The prompts run through the context decoder packed back to back (no padded token is computed); every decode step then runs the whole
 * batch with row b at position prompt_lens[b] + i (b200_decoder_step_ragged).  Row b's greedy ids are those of the same prompt
 * generated alone.  The sampling seed `step` stays shared by the batch as in the reference (sampling.cu:44-52 seeds with (step, row)):
 * it starts at the LONGEST prompt length.  Workspace: b200_generate_workspace_bytes(dec, p, batch, prompt_len). */
int b200_generate_ragged(b200_decoder_t *dec, const b200_generate_params_t *p, const int *prompt_ids, const int *prompt_lens,
                         int batch, int prompt_len, void *k_cache, void *v_cache, void *workspace, size_t workspace_bytes,
                         int *out_ids, int *n_generated, b200_stream_t stream);

/* ===================================== continuous batching over a paged KV cache ===================================== */

/* SURVEY.md 8f rank 4 (absent from the reference: one static [L,B,Hkv,S,d] cache, src/models/llama/llama.cpp:47-48; one prompt at a time,
 * :327-398).  A SCHEDULER (host only: FCFS queue, allocator of B200_KV_PAGE_SIZE-position pages, per-iteration plan, preemption by
 * recomputation) and an ITERATION that runs the plan on the engine: admitted requests are prefilled in one packed pass straight into
 * their pages (b200_decoder_prefill_paged), running ones take one ragged decode step through the block table (b200_decoder_step_paged);
 * both sample one token per sequence.  A row of the batch is nothing but a block-table row: sequences join and leave without moving
 * cache bytes.  Greedy ids of a request equal those of the same prompt generated alone (up to ties inside the 16-bit tolerance: the
 * batch size selects the GEMV kernel).  The sampling seed of an iteration is the longest position in its batch (sampling.cu:44-52 seeds
 * with (step, row): shared by the batch as in the reference). */
typedef struct b200_batcher b200_batcher_t;
typedef struct {
    int max_batch;          /* sequences decoding together; <= the engine's max_batch                                   */
    int num_pages;          /* pages in the pool: k_pool / v_pool are [L, num_pages, Hkv, B200_KV_PAGE_SIZE, d]          */
    int max_pages_per_seq;  /* block-table row length; the longest sequence is max_pages_per_seq * B200_KV_PAGE_SIZE     */
    int max_prefill_tokens; /* prompt tokens of one prefill pass (bounds the workspace); also the longest single request */
} b200_batcher_config_t;
typedef struct {
    int n_prefill, prefill_tokens, prefill_max_len; /* admitted this iteration: sequences, packed tokens, longest        */
    int n_decode, decode_max_step;                  /* running sequences taking a decode step; their largest step        */
    int n_preempted;                                /* running sequences pushed back to the queue (pages freed)          */
    int free_pages, n_waiting;                      /* after planning                                                    */
} b200_batch_plan_t;
enum { B200_REQ_WAITING = 0, B200_REQ_RUNNING = 1, B200_REQ_FINISHED = 2, B200_REQ_REJECTED = 3 /* a prompt id outside the vocabulary */ };
enum {
    B200_PLAN_PREFILL_IDS = 0,         /* int[prefill_tokens]: the admitted sequences' tokens, packed back to back           */
    B200_PLAN_PREFILL_LENS = 1,        /* int[n_prefill]                                                                     */
    B200_PLAN_PREFILL_REQUESTS = 2,    /* int[n_prefill]: request ids                                                        */
    B200_PLAN_PREFILL_BLOCK_TABLE = 3, /* int[n_prefill, max_pages_per_seq], -1 where no page is held                        */
    B200_PLAN_PREFILL_LAST_ROWS = 4,   /* int[n_prefill]: packed row of every sequence's last token                          */
    B200_PLAN_DECODE_TOKENS = 5,       /* int[n_decode]: the token each running sequence feeds                               */
    B200_PLAN_DECODE_STEPS = 6,        /* int[n_decode]: its 1-based position count including that token                     */
    B200_PLAN_DECODE_REQUESTS = 7,
    B200_PLAN_DECODE_BLOCK_TABLE = 8
};

/* ---- scheduler: host only, no CUDA call (tests/test_batcher.py drives it on a CPU-only box) */
b200_batcher_t *b200_batcher_create(const b200_batcher_config_t *cfg); /* NULL on a bad configuration */
void b200_batcher_destroy(b200_batcher_t *b);
/* Queue a request; returns its id (>= 0) or a negative error.  prompt_len + max_new_tokens - 1 positions must fit a block-table row,
 * max_prefill_tokens (a preempted request is recomputed in one prefill) and the pool. */
int b200_batcher_submit(b200_batcher_t *b, const int *prompt_ids, int prompt_len, int max_new_tokens);
/* Plan the next iteration: (1) every running sequence gets the page its next position needs -- none free: the most recently admitted
 * running sequence is preempted (pages freed, re-queued at the FRONT with prompt + generated tokens); (2) unless something was
 * preempted, waiting requests are admitted first come first served while there is a batch slot, their pages plus one spare, the
 * prefill token budget and its padded-query budget (n * longest <= 2 * max_prefill_tokens). */
int b200_batcher_plan(b200_batcher_t *b, b200_batch_plan_t *plan);
const int *b200_batcher_plan_array(const b200_batcher_t *b, int which); /* host arrays of the current plan, valid until commit */
/* Feed the iteration's sampled ids (plan order).  A sequence finishes on end_id or after max_new_tokens: its pages return to the pool.
 * Returns the number of requests that finished in this iteration (or a negative error). */
int b200_batcher_commit(b200_batcher_t *b, const int *prefill_sampled, const int *decode_sampled, int end_id);
/* Drop the current plan without results (a launch failed): admitted requests give their pages back and return to the front of the queue
 * in order; b200_batcher_plan can be called again.  b200_batcher_step does this itself when one of its launches fails. */
int b200_batcher_abort(b200_batcher_t *b);
/* Generated ids so far (end_id included if it was sampled), state = B200_REQ_*. */
int b200_batcher_result(const b200_batcher_t *b, int request, int *out_ids, int capacity, int *n_generated, int *state);
int b200_batcher_pending(const b200_batcher_t *b);     /* waiting + running */
int b200_batcher_free_pages(const b200_batcher_t *b);
int b200_batcher_preemptions(const b200_batcher_t *b, int request);

/* ---- one iteration on the GPU: plan -> prefill pass of the admitted -> decode step of the running -> commit.  Pools caller-owned,
 * [L, num_pages, Hkv, B200_KV_PAGE_SIZE, d] of the engine's (16-bit) dtype; workspace caller-owned device memory, 256-byte aligned.
 * Synchronises the stream (the scheduler needs the sampled ids).  *n_finished (optional): requests finished by this iteration. */
size_t b200_batcher_workspace_bytes(const b200_batcher_t *b, const b200_decoder_t *dec, const b200_generate_params_t *p);
int b200_batcher_step(b200_batcher_t *b, b200_decoder_t *dec, const b200_generate_params_t *p, void *k_pool, void *v_pool,
                      void *workspace, size_t workspace_bytes, int *n_finished, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LLM_H */
