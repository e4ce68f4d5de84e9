// Drop-in for the reference's src/layers/includes/context_attention.h:15-67 + src/layers/context_attention.cpp:143-303: prefill
// ("context") attention.
//   inputs  {"attention_input" [T, hidden], "padding_offset" [bs, max_q_len] int, "history_length", "input_length", "context_length"
//            [bs] int (GPU), "attention_mask" [bs, max_q_len, max_k_len], "layer_id" (CPU int)}
//   outputs {"attention_output" [T, hidden], "all_k_cache", "all_v_cache" [L, bs, Hkv, S, d]}
// Steps 1-3 and 8-9 are the reference's launcher sequence (QKV linear, bias/transpose/RoPE, KV append, ..., output linear).  Steps
// 4-8 (repeat-KV gather, QK^T, scale+mask+softmax, PV, transpose/un-pad) run as ONE fused causal flash-style kernel
// (b200_context_attention): the [bs, H, Sq, Sk] score tensor and the repeated K/V are never materialised.  The mask it applies is
// the one launchBuildCausalMasks would build from (input_length, context_length); "attention_mask" is accepted and not read.
// setUnfused(true) switches to the reference's literal 9-launcher chain (debugging / parity of the individual launchers).
#pragma once

#include <cmath>
#include "b200_layer_common.h"
#include "../../weights/includes/attention_weights.h"
#include "../../kernels/includes/linear.cuh"
#include "../../kernels/includes/scale_and_mask_and_softmax.cuh"
#include "../../kernels/includes/qkv_bias_and_rope.cuh"
#include "../../kernels/includes/transpose_and_remove_padding.cuh"
#include "../../kernels/includes/concat_past_kv.cuh"
#include "../../kernels/includes/repeat_kv.cuh"

template <typename T> class LlamaContextAttentionLayer {
private:
    int head_num;
    int head_size;
    int hidden_units;
    int repeats_per_kv;
    int kv_head_num;
    float scale;
    LlamaAttentionStaticParams *attention_static_params;
    cudaStream_t stream;
    BaseAllocator *allocator;
    CublasWrapper *cublas_wrapper;
    b200shim::Workspace workspace;
    cudaStream_t active_stream = nullptr;
    bool unfused = false;
    int alloc_key[4] = {-1, -1, -1, -1};

    TensorWrapper<T> *lineared_qkv = nullptr;             // [T, H + 2 Hkv, d]
    TensorWrapper<T> *padded_q = nullptr;                 // [bs, H, max_q_len, d]
    TensorWrapper<T> *padded_k = nullptr;                 // [bs, Hkv, max_q_len, d]
    TensorWrapper<T> *padded_v = nullptr;
    TensorWrapper<T> *k_cache = nullptr;                  // un-fused only: [bs, H, max_k_len, d]
    TensorWrapper<T> *v_cache = nullptr;
    TensorWrapper<T> *qkT = nullptr;                      // un-fused only: [bs, H, max_q_len, max_k_len]
    TensorWrapper<T> *padded_qkTv = nullptr;              // un-fused only: [bs, H, max_q_len, d]
    TensorWrapper<T> *transposed_unpadded_qkv = nullptr;  // [T, H, d]

    void dropViews() {
        delete lineared_qkv;
        delete padded_q;
        delete padded_k;
        delete padded_v;
        delete k_cache;
        delete v_cache;
        delete qkT;
        delete padded_qkTv;
        delete transposed_unpadded_qkv;
        lineared_qkv = padded_q = padded_k = padded_v = k_cache = v_cache = qkT = padded_qkTv = transposed_unpadded_qkv = nullptr;
    }

public:
    LlamaContextAttentionLayer(int head_num, int kv_head_num, int head_size, LlamaAttentionStaticParams *attention_static_params,
                               cudaStream_t stream, CublasWrapper *cublas_wrapper, BaseAllocator *allocator)
        : head_num(head_num), head_size(head_size), hidden_units(head_num * head_size), repeats_per_kv(head_num / kv_head_num),
          kv_head_num(kv_head_num), scale(1.0f / std::sqrt((float)head_size)), attention_static_params(attention_static_params), stream(stream),
          allocator(allocator), cublas_wrapper(cublas_wrapper), workspace(allocator) {}
    ~LlamaContextAttentionLayer() { freeBuf(); }
    LlamaContextAttentionLayer(const LlamaContextAttentionLayer &) = delete;
    LlamaContextAttentionLayer &operator=(const LlamaContextAttentionLayer &) = delete;

    LlamaAttentionStaticParams *getAttentionStaticParams() { return attention_static_params; }
    void setStream(cudaStream_t s) { active_stream = s; }
    void setUnfused(bool on) { unfused = on, alloc_key[0] = -1; }

    void allocateMemory(LlamaAttentionDynamicParams *params) {
        const int bs = params->batch_size, T_ = params->num_tokens, mq = params->max_q_len, mk = params->max_k_len;
        if (alloc_key[0] == bs && alloc_key[1] == T_ && alloc_key[2] == mq && alloc_key[3] == mk && lineared_qkv) return;
        const int qkv_heads = head_num + 2 * kv_head_num;
        using W = b200shim::Workspace;
        const size_t n_lin = (size_t)T_ * qkv_heads * head_size, n_q = (size_t)bs * head_num * mq * head_size,
                     n_kv = (size_t)bs * kv_head_num * mq * head_size, n_out = (size_t)T_ * head_num * head_size,
                     n_rep = (size_t)bs * head_num * mk * head_size, n_qk = (size_t)bs * head_num * mq * mk;
        size_t bytes = W::padded(n_lin, sizeof(T)) + W::padded(n_q, sizeof(T)) + 2 * W::padded(n_kv, sizeof(T)) + W::padded(n_out, sizeof(T));
        if (unfused) bytes += 2 * W::padded(n_rep, sizeof(T)) + W::padded(n_qk, sizeof(T)) + W::padded(n_q, sizeof(T));
        workspace.reserve(bytes);
        dropViews();
        const DataType dt = getTensorType<T>();
        lineared_qkv = new TensorWrapper<T>(Device::GPU, dt, {T_, qkv_heads, head_size}, workspace.template take<T>(n_lin));
        padded_q = new TensorWrapper<T>(Device::GPU, dt, {bs, head_num, mq, head_size}, workspace.template take<T>(n_q));
        padded_k = new TensorWrapper<T>(Device::GPU, dt, {bs, kv_head_num, mq, head_size}, workspace.template take<T>(n_kv));
        padded_v = new TensorWrapper<T>(Device::GPU, dt, {bs, kv_head_num, mq, head_size}, workspace.template take<T>(n_kv));
        transposed_unpadded_qkv = new TensorWrapper<T>(Device::GPU, dt, {T_, head_num, head_size}, workspace.template take<T>(n_out));
        if (unfused) {
            k_cache = new TensorWrapper<T>(Device::GPU, dt, {bs, head_num, mk, head_size}, workspace.template take<T>(n_rep));
            v_cache = new TensorWrapper<T>(Device::GPU, dt, {bs, head_num, mk, head_size}, workspace.template take<T>(n_rep));
            qkT = new TensorWrapper<T>(Device::GPU, dt, {bs, head_num, mq, mk}, workspace.template take<T>(n_qk));
            padded_qkTv = new TensorWrapper<T>(Device::GPU, dt, {bs, head_num, mq, head_size}, workspace.template take<T>(n_q));
        }
        alloc_key[0] = bs, alloc_key[1] = T_, alloc_key[2] = mq, alloc_key[3] = mk;
    }
    void freeBuf() {
        dropViews();
        workspace.release();
        alloc_key[0] = -1;
    }

    void forward(TensorMap *inputs, TensorMap *outputs, LlamaAttentionWeights<T> *weights, LlamaAttentionDynamicParams *params,
                 LlamaAttentionStaticParams *static_params) {
        b200shim::StreamScope scope(active_stream);
        allocateMemory(params);
        LlamaAttentionStaticParams sp = *static_params;
        sp.head_num = head_num, sp.kv_head_num = kv_head_num, sp.head_size = head_size;
        Tensor *attention_input = inputs->at("attention_input");
        Tensor *padding_offset = inputs->at("padding_offset");
        Tensor *history_length = inputs->at("history_length");
        Tensor *input_length = inputs->at("input_length");
        Tensor *context_length = inputs->at("context_length");
        Tensor *layer_id = inputs->at("layer_id");  // CPU
        Tensor *all_k_cache = outputs->at("all_k_cache");
        Tensor *all_v_cache = outputs->at("all_v_cache");
        Tensor *attention_output = outputs->at("attention_output");
        // 1. QKV linear  2. split + transpose + re-pad + RoPE  3. append to the KV cache
        launchLinearGemm(attention_input->wrap<T>(), &weights->qkv, lineared_qkv, cublas_wrapper, false, weights->qkv.is_transposed);
        launchFusedQKVAddBiasAndTransposeAndRope(padded_q, padded_k, padded_v, lineared_qkv, &weights->qkv, padding_offset->wrap<int>(),
                                                 history_length->wrap<int>(), input_length->wrap<int>(), &sp);
        launchConcatKVCache(padded_k, padded_v, layer_id->wrap<int>(), input_length->wrap<int>(), history_length->wrap<int>(),
                            all_k_cache->wrap<T>(), all_v_cache->wrap<T>());
        if (!unfused) {
            // 4-8. fused causal attention straight off the head-sharded cache, output already [T, H, d]
            const std::vector<int> &cs = all_k_cache->shape;  // [L, bs, Hkv, S, d]
            LLM_CHECK_WITH_INFO(cs.size() == 5 && cs[1] == params->batch_size && cs[2] == kv_head_num && cs[4] == head_size,
                                "all_k_cache must be [num_layers, batch, kv_head_num, max_seq_len, head_size]");
            B200_CALL(b200_context_attention(padded_q->data, all_k_cache->wrap<T>()->data, all_v_cache->wrap<T>()->data, transposed_unpadded_qkv->data,
                                             padding_offset->wrap<int>()->data, input_length->wrap<int>()->data, context_length->wrap<int>()->data,
                                             layer_id->wrap<int>()->getVal(), params->batch_size, head_num, kv_head_num, params->max_q_len, cs[3],
                                             head_size, params->num_tokens, scale, b200DType<T>(), b200GetStream()));
        } else {
            Tensor *attention_mask = inputs->at("attention_mask");
            launchRepeatKVCache(all_k_cache->wrap<T>(), all_v_cache->wrap<T>(), context_length->wrap<int>(), layer_id->wrap<int>(), k_cache, v_cache);
            launchLinearStridedBatchGemm(padded_q, k_cache, qkT, cublas_wrapper, false, true);
            launchFusedScaleMaskAndSoftmax(qkT, attention_mask->wrap<T>(), qkT, scale);
            launchLinearStridedBatchGemm(qkT, v_cache, padded_qkTv, cublas_wrapper, false, false);
            launchFusedTransposeAndRemovePadding(padded_qkTv, padding_offset->wrap<int>(), transposed_unpadded_qkv);
        }
        // 9. output linear ([T, H*d] view of the [T, H, d] buffer)
        launchLinearGemm(transposed_unpadded_qkv, &weights->output, attention_output->wrap<T>(), cublas_wrapper, false, weights->output.is_transposed);
    }
};
