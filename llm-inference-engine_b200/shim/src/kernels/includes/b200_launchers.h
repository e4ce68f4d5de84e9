// b200_launchers.h -- the reference's 18 launch* host functions (src/kernels/includes/*.cuh) re-provided on top of the
// C ABI of libb200llm.so.  Same names, argument order, shape conventions and LLM_CHECK failures as the reference
// launchers cited per function; T may be float, half or __nv_bfloat16.  The per-kernel headers of this directory
// (rmsnorm.cuh, linear.cuh, ...) all include this file, so `#include "src/kernels/includes/rmsnorm.cuh"` keeps working.
//
// Streams: the reference launches everything on the legacy default stream.  The shim does the same unless the caller
// installs a stream with b200SetStream() (thread-local), which is how the layer classes make the whole forward
// stream-ordered and CUDA-graph-capturable.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "b200llm.h"
#include "cublas_utils.cuh"
#include "../../models/llama/llama_params.h"
#include "../../utils/macro.h"
#include "../../utils/params.h"
#include "../../utils/tensor.h"
#include "../../weights/includes/base_weights.h"
#include "../../weights/includes/embedding_weights.h"
#include "../../weights/includes/norm_weights.h"

inline cudaStream_t &b200StreamRef() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}
inline void b200SetStream(cudaStream_t s) { b200StreamRef() = s; }
inline cudaStream_t b200GetStream() { return b200StreamRef(); }

namespace b200shim {
inline void ensure_workspace() {
    static thread_local int done_for_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done_for_device != dev) {
        B200_CALL(b200_workspace_ensure(0));
        done_for_device = dev;
    }
}
inline int flat_cols(const std::vector<int> &shape) {
    int n = 1;
    for (size_t i = 1; i < shape.size(); ++i) n *= shape[i];
    return n;
}
}  // namespace b200shim

// reference: src/kernels/includes/rmsnorm.cuh:9-15, rmsnorm.cu:130-159
template <typename T>
void launchRMSNorm(TensorWrapper<T> *decoder_out, TensorWrapper<T> *decoder_residual, LayerNormWeight<T> *attention_norm_weight, float eps,
                   bool is_last = false) {
    (void)is_last;  // ignored by the reference kernel too
    const int num_tokens = decoder_out->shape[0];
    const int hidden_units = decoder_out->shape[1];
    B200_CALL(b200_rmsnorm(decoder_out->data, decoder_residual ? decoder_residual->data : nullptr, attention_norm_weight->gamma, eps,
                           num_tokens, hidden_units, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/add_residual_and_rmsnorm.cuh:11-18, add_residual_and_rmsnorm.cu:170-201
// (`norm` carries the BIAS of the preceding projection; `scale` is the RMSNorm gamma)
template <typename T>
void launchFusedAddBiasResidualAndRMSNorm(TensorWrapper<T> *residual, TensorWrapper<T> *decoder_out, BaseWeight<T> *norm, T *scale, float eps) {
    const int num_tokens = decoder_out->shape[0];
    const int hidden_units = decoder_out->shape[1];
    B200_CALL(b200_fused_add_bias_residual_rmsnorm(residual ? residual->data : nullptr, decoder_out->data, norm ? norm->bias : nullptr, scale,
                                                   eps, num_tokens, hidden_units, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/add_residual.cuh:9-14, add_residual.cu:51-76
template <typename T> void launchAddResidual(TensorWrapper<T> *residual, TensorWrapper<T> *decoder_out, bool is_print = false) {
    (void)is_print;
    const int num_tokens = decoder_out->shape[0];
    const int hidden_units = decoder_out->shape[1];
    LLM_CHECK_WITH_INFO(residual->shape[0] == num_tokens && residual->shape[1] == hidden_units, "residual shape must match decoder_out");
    B200_CALL(b200_add_residual(residual->data, decoder_out->data, num_tokens, hidden_units, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/linear.cuh:13-20, linear.cu:10-87.  Shape rules identical to the reference
// (3-D inputs/outputs flattened, trans flags swap the declared dims for the CHECKS only); the weight memory is read as
// [K,N] row-major exactly like the reference's cublasGemmEx call does -- unless the weight carries a packed
// B200 copy (BaseWeight::packed, made by packForB200()), in which case the [N,K] streaming kernels are used.
template <typename T>
void launchLinearGemm(TensorWrapper<T> *input, BaseWeight<T> *weight, TensorWrapper<T> *output, CublasWrapper *cublas_wrapper,
                      bool trans_a = false, bool trans_b = false) {
    (void)cublas_wrapper;  // kept for API compatibility; no library GEMM on this path
    int Am = input->shape[0], An = b200shim::flat_cols(input->shape);
    int Bm = weight->shape[0], Bn = weight->shape[1];
    const int Cm = output->shape[0], Cn = b200shim::flat_cols(output->shape);
    if (trans_a) std::swap(Am, An);
    if (trans_b) std::swap(Bm, Bn);
    LLM_CHECK_WITH_INFO(An == Bm, "2nd dim of weight MUST = 1st dim of input");
    LLM_CHECK_WITH_INFO(Am == Cm && Bn == Cn, "output shape should be equal to weight shape");
    LLM_CHECK_WITH_INFO(!trans_a, "trans_a is never used by the reference layers and is not supported");
    const int M = Cm, K = An, N = Cn;
    if (weight->packed && weight->packed_from == weight->data) {
        const int fmt = weight->packed_type == WeightType::FP8_W ? B200_W_FP8E4M3 : (weight->packed_type == WeightType::INT4_W ? B200_W_INT4 : B200_W_DENSE);
        B200_CALL(b200_linear(input->data, weight->packed, weight->packed_scales, weight->packed_zeros, output->data, M, K, N, b200DType<T>(),
                              fmt, B200_LAYOUT_NK, weight->group_size, b200GetStream()));
        return;
    }
    b200shim::ensure_workspace();
    B200_CALL(b200_linear(input->data, weight->data, nullptr, nullptr, output->data, M, K, N, b200DType<T>(), B200_W_DENSE, B200_LAYOUT_KN, 0,
                          b200GetStream()));
}

// reference: src/kernels/includes/linear.cuh:22-29, linear.cu:89-158.  [bs, heads, rows, cols] operands; trans_b = true is the
// TRUE q.k^T (the reference's call computes q.reshape(k) instead -- SURVEY D4, documented divergence).
template <typename T>
void launchLinearStridedBatchGemm(TensorWrapper<T> *input1, TensorWrapper<T> *input2, TensorWrapper<T> *output, CublasWrapper *cublas_wrapper,
                                  bool trans_a = false, bool trans_b = false) {
    (void)cublas_wrapper;
    int Am = input1->shape[2], An = input1->shape[3];
    int Bm = input2->shape[2], Bn = input2->shape[3];
    const int Cm = output->shape[2], Cn = output->shape[3];
    if (trans_a) std::swap(Am, An);
    if (trans_b) std::swap(Bm, Bn);
    LLM_CHECK_WITH_INFO(An == Bm, "2nd dim of weight MUST = 1st dim of input");
    LLM_CHECK_WITH_INFO(Am == Cm && Bn == Cn, "output shape should be equal to weight shape");
    LLM_CHECK_WITH_INFO(!trans_a, "trans_a is never used by the reference layers and is not supported");
    const int batch = input1->shape[0] * input1->shape[1];
    B200_CALL(b200_batched_gemm(input1->data, input2->data, output->data, batch, Cm, Cn, An, trans_b ? 1 : 0, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/rope.cuh:12-17, rope.cu:60-98
template <typename T> void launchRope(TensorWrapper<T> *qkv_buf, TensorWrapper<int> *step, LlamaAttentionStaticParams *static_params) {
    const int batch_size = qkv_buf->shape[0];
    const int qkv_head_num = qkv_buf->shape[1];
    const int head_size = qkv_buf->shape[2];
    const int head_num = static_params->head_num;
    const int kv_head_num = (qkv_head_num - head_num) / 2;
    LLM_CHECK(batch_size == 1 || batch_size > 0);
    LLM_CHECK_WITH_INFO(qkv_head_num == head_num + 2 * kv_head_num, "qkv head count must be head_num + 2 * kv_head_num");
    const int rot = static_params->rotary_embedding_dim < head_size ? static_params->rotary_embedding_dim : head_size;
    B200_CALL(b200_rope_decode(qkv_buf->data, batch_size, head_num, kv_head_num, head_size, step->getVal(), rot,
                               static_params->rotary_embedding_base, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/qkv_bias_and_rope.cuh:12-23, qkv_bias_and_rope.cu:86-138
template <typename T>
void launchFusedQKVAddBiasAndTransposeAndRope(TensorWrapper<T> *q_buf, TensorWrapper<T> *k_buf, TensorWrapper<T> *v_buf, TensorWrapper<T> *QKV,
                                              BaseWeight<T> *qkv, TensorWrapper<int> *padding_offset, TensorWrapper<int> *history_length,
                                              TensorWrapper<int> *input_length, LlamaAttentionStaticParams *static_params) {
    const int token_num = QKV->shape[0];
    const int qkv_head_num = QKV->shape[1];
    const int head_size = QKV->shape[2];
    const int batch_size = q_buf->shape[0];
    const int head_num = q_buf->shape[1];
    const int seq_len = q_buf->shape[2];
    LLM_CHECK_WITH_INFO(k_buf->shape[1] == v_buf->shape[1], "k and v should have same head_num");
    LLM_CHECK_WITH_INFO(k_buf->shape[1] == (qkv_head_num - head_num) / 2, "k and v should have same head_num");
    LLM_CHECK_WITH_INFO(q_buf->shape[3] == head_size, "head_size does not match!");
    const int kv_head_num = k_buf->shape[1];
    const int rot = static_params->rotary_embedding_dim < head_size ? static_params->rotary_embedding_dim : head_size;
    B200_CALL(b200_qkv_bias_transpose_rope(q_buf->data, k_buf->data, v_buf->data, QKV->data, qkv ? qkv->bias : nullptr, padding_offset->data,
                                           history_length->data, input_length->data, batch_size, seq_len, token_num, head_num, kv_head_num,
                                           head_size, rot, static_params->rotary_embedding_base, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/decoder_self_attention.cuh:11-22, decoder_self_attention.cu:211-270
template <typename T>
void launchDecoderMaskedMultiHeadAttention(TensorWrapper<T> *qkv_buf, BaseWeight<T> *qkv, TensorWrapper<int> *layer_id, TensorWrapper<T> *k_cache,
                                           TensorWrapper<T> *v_cache, TensorWrapper<bool> *finished, TensorWrapper<int> *step,
                                           TensorWrapper<T> *mha_output, LlamaAttentionStaticParams *static_params) {
    (void)static_params;  // RoPE parameters are unused by the reference kernel: RoPE is launchRope's job
    const int batch_size = qkv_buf->shape[0];
    const int qkv_head_num = qkv_buf->shape[1];
    const int head_size = qkv_buf->shape[2];
    const int kv_head_num = k_cache->shape[2];
    const int max_seq_len = k_cache->shape[3];
    const int head_num = qkv_head_num - 2 * kv_head_num;
    b200shim::ensure_workspace();
    B200_CALL(b200_decode_mha(qkv_buf->data, qkv ? qkv->bias : nullptr, k_cache->data, v_cache->data, mha_output->data,
                              finished ? reinterpret_cast<const uint8_t *>(finished->data) : nullptr, batch_size, head_num, kv_head_num,
                              head_size, max_seq_len, step->getVal(), layer_id->getVal(), 0, 0, 0.0f, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/concat_past_kv.cuh:9-18, concat_past_kv.cu:44-89
template <typename T>
void launchConcatKVCache(TensorWrapper<T> *k_src, TensorWrapper<T> *v_src, TensorWrapper<int> *layer_id, TensorWrapper<int> *cur_query_length,
                         TensorWrapper<int> *history_length, TensorWrapper<T> *k_dst, TensorWrapper<T> *v_dst) {
    const int batch_size = k_src->shape[0];
    const int kv_head_num = k_src->shape[1];
    const int max_q_len = k_src->shape[2];
    const int head_size = k_src->shape[3];
    const int max_seq_len = k_dst->shape[3];
    B200_CALL(b200_concat_kv_cache(k_src->data, v_src->data, k_dst->data, v_dst->data, cur_query_length->data, history_length->data,
                                   layer_id->getVal(), batch_size, kv_head_num, max_q_len, max_seq_len, head_size, b200DType<T>(),
                                   b200GetStream()));
}

// reference: src/kernels/includes/repeat_kv.cuh:9-17, repeat_kv.cu:51-106 (intended semantics)
template <typename T>
void launchRepeatKVCache(TensorWrapper<T> *k_cache_src, TensorWrapper<T> *v_cache_src, TensorWrapper<int> *context_length,
                         TensorWrapper<int> *layer_id, TensorWrapper<T> *k_cache_dst, TensorWrapper<T> *v_cache_dst) {
    const int batch_size = context_length->shape[0];
    const int kv_head_num = k_cache_src->shape[2];
    const int max_seq_len = k_cache_src->shape[3];
    const int head_num = k_cache_dst->shape[1];
    const int max_k_len = k_cache_dst->shape[2];
    const int head_size = k_cache_dst->shape[3];
    B200_CALL(b200_repeat_kv_cache(k_cache_src->data, v_cache_src->data, k_cache_dst->data, v_cache_dst->data, context_length->data,
                                   layer_id->getVal(), batch_size, head_num, kv_head_num, max_k_len, max_seq_len, head_size, b200DType<T>(),
                                   b200GetStream()));
}

// reference: src/kernels/includes/scale_and_mask_and_softmax.cuh:10-16, scale_and_mask_and_softmax.cu:213-341
template <typename T>
void launchFusedScaleMaskAndSoftmax(TensorWrapper<T> *qk, TensorWrapper<T> *mask, TensorWrapper<T> *attention_weights, float scale) {
    const int batch_size = qk->shape[0];
    const int head_nums = qk->shape[1];
    const int q_length = qk->shape[2];
    const int k_length = qk->shape[3];
    B200_CALL(b200_scale_mask_softmax(qk->data, mask->data, attention_weights->data, scale, batch_size, head_nums, q_length, k_length,
                                      b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/build_causal_mask.cuh:9-14, build_causal_mask.cu:25-42
template <typename T> void launchBuildCausalMasks(TensorWrapper<T> *mask, TensorWrapper<int> *q_lens, TensorWrapper<int> *k_lens) {
    const int batch_size = mask->shape[0];
    const int max_q_len = mask->shape[1];
    const int max_k_len = mask->shape[2];
    B200_CALL(b200_build_causal_masks(mask->data, q_lens->data, k_lens->data, batch_size, max_q_len, max_k_len, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/cal_padding_offset.cuh:16-20, cal_padding_offset.cu:45-70
inline void launchCalPaddingOffset(TensorWrapper<int> *padding_offset, TensorWrapper<int> *cum_seqlens, TensorWrapper<int> *input_lengths) {
    const int batch_size = padding_offset->shape[0];
    const int max_q_len = padding_offset->shape[1];
    LLM_CHECK_WITH_INFO(batch_size == input_lengths->shape[0], "input lengths numbers should equal to padding offset bs dim!");
    LLM_CHECK_WITH_INFO(batch_size == cum_seqlens->shape[0] - 1, "cum seqlen numbers should equal to padding offset bs dim + 1!");
    B200_CALL(b200_cal_padding_offset(padding_offset->data, cum_seqlens->data, input_lengths->data, batch_size, max_q_len, b200GetStream()));
}

// reference: src/kernels/includes/transpose_and_remove_padding.cuh:8-13, transpose_and_remove_padding.cu:45-74
template <typename T>
void launchFusedTransposeAndRemovePadding(TensorWrapper<T> *padded_qkv_buf, TensorWrapper<int> *padding_offset, TensorWrapper<T> *qkv_buf_without_padding) {
    const int batch_size = padded_qkv_buf->shape[0];
    const int head_num = padded_qkv_buf->shape[1];
    const int seq_len = padded_qkv_buf->shape[2];
    const int head_size = padded_qkv_buf->shape[3];
    const int num_tokens = qkv_buf_without_padding->shape[0];
    B200_CALL(b200_transpose_remove_padding(padded_qkv_buf->data, padding_offset->data, qkv_buf_without_padding->data, num_tokens, batch_size,
                                            seq_len, head_num, head_size, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/silu_and_mul.cuh:9-13, silu_and_mul.cu:61-82.  input [tokens, 2, inter]
template <typename T> void launchSiluAndMul(TensorWrapper<T> *input, TensorWrapper<T> *output) {
    const int batch_size = input->shape[0];
    LLM_CHECK(input->shape[1] == 2);
    const int intermedia_size = input->shape[2];
    B200_CALL(b200_silu_and_mul(input->data, output->data, batch_size, intermedia_size, b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/input_embedding.cuh:7-12, input_embedding.cu:24-51
template <typename T> void launchInputEmbedding(TensorWrapper<int> *input_ids, TensorWrapper<T> *output, EmbeddingWeight<T> *embed_table) {
    const int max_context_token_num = output->shape[0];
    const int hidden_size = output->shape[1];
    LLM_CHECK_WITH_INFO(max_context_token_num == input_ids->shape[0], "input ids 1st shape should equal to 1st shape of output");
    B200_CALL(b200_input_embedding(input_ids->data, embed_table->data, output->data, max_context_token_num, hidden_size, b200DType<T>(),
                                   b200GetStream()));
}

// reference: src/kernels/includes/topk.cuh:44-51, topk.cu:104-140.  K is taken from final_topk_ids->shape.back() (the reference
// hard-codes 5, which is also what its callers allocate).
template <typename T>
void launchTopKForBeamSearch(TensorWrapper<T> *probs, TensorWrapper<int> *topk_ids, TensorWrapper<T> *topk_vals, TensorWrapper<int> *final_topk_ids,
                             TensorWrapper<T> *final_topk_vals) {
    const int bsxbm = probs->shape[0];
    const int vocab_size = probs->shape[1];
    const int K = final_topk_ids->shape.empty() ? 5 : final_topk_ids->shape.back();
    B200_CALL(b200_topk(probs->data, topk_ids->data, topk_vals->data, final_topk_ids->data, final_topk_vals->data, bsxbm, vocab_size, K,
                        b200DType<T>(), b200GetStream()));
}

// reference: src/kernels/includes/sampling.cuh:11-19, sampling.cu:73-102
template <typename T>
void launchSampling(TensorWrapper<int> *topk_id, TensorWrapper<T> *topk_val, TensorWrapper<int> *seqlen, TensorWrapper<bool> *is_finished,
                    TensorWrapper<int> *output_id, MapStringToInt *params) {
    const int batch_size = topk_id->shape[0];
    const int K = topk_id->shape[1];
    B200_CALL(b200_sampling(topk_id->data, topk_val->data, seqlen->data, reinterpret_cast<uint8_t *>(is_finished->data), output_id->data, batch_size,
                            K, params->at("step"), params->at("end_id"), params->at("vocab_size"), b200DType<T>(), b200GetStream()));
}
