// Drop-in for the reference's src/layers/includes/context_decoder.h:15-94 + src/layers/context_decoder.cpp:58-199: prefill over
// all layers.
//   inputs  {"decoder_input" [T, hidden], "history_length", "input_length", "context_length" [bs] int (GPU), "layer_id" (CPU int)}
//           (+ "output_norm_weight", passed by the reference's callers and unused)
//   outputs {"decoder_output" [T, hidden], "all_k_cache", "all_v_cache" [L, bs, Hkv, S, d]}
// padding offsets -> causal mask -> per layer: RMSNorm -> context attention -> fused add-bias-residual-RMSNorm -> FFN ->
// add-residual.  Stream-ordered: no per-op malloc / free / synchronise; buffers come from a grow-only workspace.
// Unlike the reference (context_decoder.cpp:197 frees its sub-layers after the first call) the object can be reused.
#pragma once

#include <memory>
#include <vector>
#include "../../kernels/includes/build_causal_mask.cuh"
#include "../../kernels/includes/cal_padding_offset.cuh"
#include "../../kernels/includes/add_residual_and_rmsnorm.cuh"
#include "../../kernels/includes/add_residual.cuh"
#include "../../kernels/includes/rmsnorm.cuh"
#include "../../layers/includes/context_attention.h"
#include "../../layers/includes/ffn.h"
#include "../../weights/includes/llama_weights.h"
#include "../../utils/tensor.h"

template <typename T> class LlamaContextDecoder {
private:
    int head_num;
    int kv_head_num;
    int head_size;
    int intermediate_size;
    int num_layer;
    int hidden_units;
    float rmsnorm_eps;

    TensorWrapper<T> *attention_mask = nullptr;
    TensorWrapper<int> *padding_offset = nullptr;
    TensorWrapper<int> *cum_seqlens = nullptr;
    TensorWrapper<T> *decoder_residual = nullptr;

    cudaStream_t stream;
    CublasWrapper *cublas_wrapper;
    BaseAllocator *allocator;

    LlamaContextAttentionLayer<T> *context_attention = nullptr;
    LlamaFFNLayer<T> *ffn = nullptr;
    DataType data_type;
    b200shim::Workspace workspace;
    cudaStream_t active_stream = nullptr;
    bool build_mask = false;  // the fused attention derives the mask from the lengths; the tensor is only built on request
    int alloc_key[4] = {-1, -1, -1, -1};

    void dropViews() {
        delete attention_mask;
        delete padding_offset;
        delete cum_seqlens;
        delete decoder_residual;
        attention_mask = nullptr, padding_offset = cum_seqlens = nullptr, decoder_residual = nullptr;
    }

public:
    LlamaContextDecoder(const int &head_num, const int &kv_head_num, const int &head_size, const int &intermediate_size, const int &num_layer,
                        LlamaAttentionStaticParams *const &attention_static_params, const float &rmsnorm_eps, const cudaStream_t &stream,
                        CublasWrapper *const &cublas_wrapper, BaseAllocator *const &allocator)
        : head_num(head_num), kv_head_num(kv_head_num), head_size(head_size), intermediate_size(intermediate_size), num_layer(num_layer),
          hidden_units(head_num * head_size), rmsnorm_eps(rmsnorm_eps), stream(stream), cublas_wrapper(cublas_wrapper), allocator(allocator),
          data_type(getTensorType<T>()), workspace(allocator) {
        context_attention = new LlamaContextAttentionLayer<T>(head_num, kv_head_num, head_size, attention_static_params, stream, cublas_wrapper, allocator);
        ffn = new LlamaFFNLayer<T>(head_num, head_size, intermediate_size, stream, cublas_wrapper, allocator);
    }
    ~LlamaContextDecoder() {
        freeBuf();
        delete context_attention;
        delete ffn;
    }
    LlamaContextDecoder(const LlamaContextDecoder &) = delete;
    LlamaContextDecoder &operator=(const LlamaContextDecoder &) = delete;

    void setStream(cudaStream_t s) {
        active_stream = s;
        context_attention->setStream(s);
        ffn->setStream(s);
    }
    // true: run the reference's literal attention chain (materialised mask, repeated KV, [bs,H,Sq,Sk] scores)
    void setUnfused(bool on) {
        build_mask = on;
        context_attention->setUnfused(on);
        alloc_key[0] = -1;
    }

    void allocateMemory(LlamaAttentionDynamicParams *p) {
        const int T_ = p->num_tokens, bs = p->batch_size, mq = p->max_q_len, mk = p->max_k_len;
        if (alloc_key[0] == T_ && alloc_key[1] == bs && alloc_key[2] == mq && alloc_key[3] == mk && decoder_residual) return;
        using W = b200shim::Workspace;
        const size_t n_res = (size_t)T_ * hidden_units, n_mask = (size_t)bs * mq * mk, n_po = (size_t)bs * mq, n_cum = (size_t)bs + 1;
        workspace.reserve(W::padded(n_res, sizeof(T)) + W::padded(n_mask, sizeof(T)) + W::padded(n_po, sizeof(int)) + W::padded(n_cum, sizeof(int)));
        dropViews();
        decoder_residual = new TensorWrapper<T>(Device::GPU, data_type, {T_, hidden_units}, workspace.template take<T>(n_res));
        attention_mask = new TensorWrapper<T>(Device::GPU, data_type, {bs, mq, mk}, workspace.template take<T>(n_mask));
        padding_offset = new TensorWrapper<int>(Device::GPU, getTensorType<int>(), {bs, mq}, workspace.template take<int>(n_po));
        cum_seqlens = new TensorWrapper<int>(Device::GPU, getTensorType<int>(), {bs + 1}, workspace.template take<int>(n_cum));
        alloc_key[0] = T_, alloc_key[1] = bs, alloc_key[2] = mq, alloc_key[3] = mk;
    }
    void freeBuf() {
        dropViews();
        workspace.release();
        alloc_key[0] = -1;
    }

    void forward(TensorMap *input_tensors, std::vector<LlamaLayerWeight<T> *> *layer_weights, TensorMap *output_tensors,
                 LlamaAttentionDynamicParams *attention_dynamic_params) {
        b200shim::StreamScope scope(active_stream);
        allocateMemory(attention_dynamic_params);
        Tensor *seq_lens = input_tensors->at("input_length");
        Tensor *context_length = input_tensors->at("context_length");
        Tensor *history_length = input_tensors->at("history_length");
        Tensor *decoder_output = output_tensors->at("decoder_output");
        Tensor *all_k_cache = output_tensors->at("all_k_cache");
        Tensor *all_v_cache = output_tensors->at("all_v_cache");
        Tensor *layer_id = input_tensors->at("layer_id");
        Tensor *decoder_input = input_tensors->at("decoder_input");
        LLM_CHECK_WITH_INFO(decoder_input->wrap<T>()->data != nullptr, "The data pointer of tensor inserted into TensorMap is nullptr!");
        LLM_CHECK_WITH_INFO(history_length->wrap<int>()->data != nullptr, "The data pointer of tensor inserted into TensorMap is nullptr!");

        // 1. padding offsets  2. causal mask
        launchCalPaddingOffset(padding_offset, cum_seqlens, seq_lens->wrap<int>());
        if (build_mask) launchBuildCausalMasks<T>(attention_mask, seq_lens->wrap<int>(), context_length->wrap<int>());

        TensorMap context_attention_inputs{{"attention_input", decoder_input},   {"padding_offset", padding_offset}, {"history_length", history_length},
                                           {"input_length", seq_lens},           {"context_length", context_length}, {"attention_mask", attention_mask},
                                           {"layer_id", layer_id}};
        TensorMap context_attention_outputs{{"attention_output", decoder_output}, {"all_k_cache", all_k_cache}, {"all_v_cache", all_v_cache}};
        std::vector<int> ids(num_layer);
        std::vector<TensorWrapper<int> *> id_tensors;
        for (int l = 0; l < num_layer; ++l) {
            if (l > 0) {
                ids[l] = l;
                id_tensors.push_back(new TensorWrapper<int>(Device::CPU, getTensorType<int>(), std::vector<int>{1}, &ids[l]));
                context_attention_inputs.insert("layer_id", id_tensors.back());  // (key, value): overwrites -- a pair would keep layer 0 (tensor.h:241-243)
            }
            Tensor *x = context_attention_inputs.at("attention_input");
            launchRMSNorm(x->wrap<T>(), decoder_residual, &layer_weights->at(l)->attention_norm_weight, rmsnorm_eps);
            context_attention->forward(&context_attention_inputs, &context_attention_outputs, &layer_weights->at(l)->self_attention_weight,
                                       attention_dynamic_params, context_attention->getAttentionStaticParams());
            launchFusedAddBiasResidualAndRMSNorm(decoder_residual, decoder_output->wrap<T>(), &layer_weights->at(l)->self_attention_weight.output,
                                                 layer_weights->at(l)->ffn_norm_weight.gamma, rmsnorm_eps);
            TensorMap ffn_inputs{{"ffn_input", decoder_output}};
            TensorMap ffn_outputs{{"ffn_output", decoder_output}};
            attention_dynamic_params->is_context = true;  // reference context_decoder.cpp:172
            ffn->forward(&ffn_inputs, &ffn_outputs, &layer_weights->at(l)->ffn_weight, attention_dynamic_params);
            launchAddResidual(decoder_residual, decoder_output->wrap<T>());
            context_attention_inputs.insert("attention_input", decoder_output);
        }
        for (TensorWrapper<int> *t : id_tensors) delete t;
    }
};
