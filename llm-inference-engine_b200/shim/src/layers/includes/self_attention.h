// Drop-in for the reference's src/layers/includes/self_attention.h + src/layers/self_attention.cpp:63-151: decode attention.
//   inputs  {"attention_input" [bs, hidden], "layer_id" (CPU int), "step" (CPU int), "finished" [bs] bool}
//   outputs {"attention_output" [bs, hidden], "all_k_cache", "all_v_cache" [L, bs, Hkv, S, d]}
//   qkv = x * Wqkv -> RoPE(step-1) -> masked MHA (+ qkv bias, KV append) -> * Wo
#pragma once

#include <cmath>
#include "b200_layer_common.h"
#include "../../weights/includes/attention_weights.h"

template <typename T> class LlamaSelfAttentionLayer {
private:
    const int head_num;
    const int kv_head_num;
    const int head_size;
    const int hidden_units;
    const int repeats_per_kv;
    float scale;
    LlamaAttentionStaticParams *attention_static_params;
    cudaStream_t stream;
    BaseAllocator *allocator;
    CublasWrapper *cublas_wrapper;
    b200shim::Workspace workspace;
    cudaStream_t active_stream = nullptr;
    TensorWrapper<T> *qkv_buf = nullptr;     // [bs, H + 2 Hkv, d]
    TensorWrapper<T> *mha_output = nullptr;  // [bs, hidden]
    LlamaAttentionStaticParams effective_params;

public:
    LlamaSelfAttentionLayer(int head_num, int kv_head_num, int head_size, LlamaAttentionStaticParams *attention_params, cudaStream_t stream,
                            CublasWrapper *cublas_wrapper, BaseAllocator *allocator)
        : head_num(head_num), kv_head_num(kv_head_num), head_size(head_size), hidden_units(head_num * head_size),
          repeats_per_kv(head_num / kv_head_num), scale(1.0f / std::sqrt((float)head_size)), attention_static_params(attention_params),
          stream(stream), allocator(allocator), cublas_wrapper(cublas_wrapper), workspace(allocator) {}
    ~LlamaSelfAttentionLayer() { freeBuf(); }

    LlamaAttentionStaticParams *getAttentionStaticParams() { return attention_static_params; }
    void setStream(cudaStream_t s) { active_stream = s; }

    void allocateMemory(LlamaAttentionDynamicParams *dynamic_params) {
        const int bs = dynamic_params->batch_size;
        const int qkv_heads = head_num + 2 * kv_head_num;
        const size_t nq = (size_t)bs * qkv_heads * head_size, no = (size_t)bs * hidden_units;
        workspace.reserve(b200shim::Workspace::padded(nq, sizeof(T)) + b200shim::Workspace::padded(no, sizeof(T)));
        delete qkv_buf;
        delete mha_output;
        qkv_buf = new TensorWrapper<T>(Device::GPU, getTensorType<T>(), {bs, qkv_heads, head_size}, workspace.take<T>(nq));
        mha_output = new TensorWrapper<T>(Device::GPU, getTensorType<T>(), {bs, hidden_units}, workspace.take<T>(no));
    }
    void freeBuf() {
        delete qkv_buf;
        delete mha_output;
        qkv_buf = mha_output = nullptr;
        workspace.release();
    }

    void forward(TensorMap *inputs, TensorMap *outputs, LlamaAttentionWeights<T> *weights, LlamaAttentionDynamicParams *dynamic_params) {
        b200shim::StreamScope scope(active_stream);
        if (!qkv_buf || qkv_buf->shape[0] != dynamic_params->batch_size) allocateMemory(dynamic_params);
        Tensor *attention_input = inputs->at("attention_input");
        Tensor *attention_output = outputs->at("attention_output");
        Tensor *key_cache = outputs->at("all_k_cache");
        Tensor *value_cache = outputs->at("all_v_cache");
        Tensor *finished = inputs->at("finished");
        Tensor *step = inputs->at("step");
        Tensor *layer_id = inputs->at("layer_id");
        // the launchers take the head counts from the static params: make them agree with this layer's constructor arguments
        effective_params = *attention_static_params;
        effective_params.head_num = head_num, effective_params.kv_head_num = kv_head_num, effective_params.head_size = head_size;
        launchLinearGemm(attention_input->wrap<T>(), &weights->qkv, qkv_buf, cublas_wrapper, false, weights->qkv.is_transposed);
        launchRope(qkv_buf, step->wrap<int>(), &effective_params);
        launchDecoderMaskedMultiHeadAttention<T>(qkv_buf, &weights->qkv, layer_id->wrap<int>(), key_cache->wrap<T>(), value_cache->wrap<T>(),
                                                 finished->wrap<bool>(), step->wrap<int>(), mha_output, &effective_params);
        launchLinearGemm(mha_output, &weights->output, attention_output->wrap<T>(), cublas_wrapper, false, weights->output.is_transposed);
    }
};
