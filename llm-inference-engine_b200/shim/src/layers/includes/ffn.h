// Drop-in for the reference's src/layers/includes/ffn.h:15-53 + src/layers/ffn.cpp:76-144: SwiGLU feed-forward.
//   inputs  {"ffn_input"  [tokens, hidden]}      outputs {"ffn_output" [tokens, hidden]}  (may be the same buffer)
//   tokens = dynamic_params->num_tokens if > 0 else dynamic_params->batch_size (ffn.cpp:84-91)
//   gate_up = x * W_gate_and_up -> [tokens, 2, inter]; act = silu(gate) * up; out = act * W_down
#pragma once

#include "b200_layer_common.h"
#include "../../weights/includes/ffn_weights.h"

template <typename T> class LlamaFFNLayer {
private:
    const int head_num;
    const int head_size;
    const int intermediate_size;
    const int hidden_units;
    cudaStream_t stream;
    BaseAllocator *allocator;
    CublasWrapper *cublas_wrapper;
    b200shim::Workspace workspace;
    cudaStream_t active_stream = nullptr;
    TensorWrapper<T> *swiglu_input = nullptr;     // [tokens, 2, inter]
    TensorWrapper<T> *down_proj_input = nullptr;  // [tokens, inter]

public:
    mutable int count = -1;  // reference ffn.h:22 (debug call counter)

    LlamaFFNLayer(int head_num, int head_size, int intermediate_size, cudaStream_t stream, CublasWrapper *cublas_wrapper, BaseAllocator *allocator)
        : head_num(head_num), head_size(head_size), intermediate_size(intermediate_size), hidden_units(head_num * head_size), stream(stream),
          allocator(allocator), cublas_wrapper(cublas_wrapper), workspace(allocator) {}
    ~LlamaFFNLayer() { freeBuf(); }

    void setStream(cudaStream_t s) { active_stream = s; }

    void allocateMemory(const int &tokens) {
        const size_t gu = (size_t)tokens * 2 * intermediate_size, act = (size_t)tokens * intermediate_size;
        workspace.reserve(b200shim::Workspace::padded(gu, sizeof(T)) + b200shim::Workspace::padded(act, sizeof(T)));
        delete swiglu_input;
        delete down_proj_input;
        swiglu_input = new TensorWrapper<T>(Device::GPU, getTensorType<T>(), {tokens, 2, intermediate_size}, workspace.take<T>(gu));
        down_proj_input = new TensorWrapper<T>(Device::GPU, getTensorType<T>(), {tokens, intermediate_size}, workspace.take<T>(act));
    }
    void allocateMemory(LlamaAttentionDynamicParams *dynamic_params) {
        allocateMemory(dynamic_params->num_tokens > 0 ? dynamic_params->num_tokens : dynamic_params->batch_size);
    }
    void freeBuf() {
        delete swiglu_input;
        delete down_proj_input;
        swiglu_input = down_proj_input = nullptr;
        workspace.release();
    }

    void forward(TensorMap *inputs, TensorMap *outputs, LlamaFFNWeights<T> *weights, LlamaAttentionDynamicParams *dynamic_params) {
        b200shim::StreamScope scope(active_stream);
        ++count;
        const int tokens = dynamic_params->num_tokens > 0 ? dynamic_params->num_tokens : dynamic_params->batch_size;
        if (!swiglu_input || swiglu_input->shape[0] != tokens) allocateMemory(tokens);
        Tensor *ffn_input = inputs->at("ffn_input");
        Tensor *ffn_output = outputs->at("ffn_output");
        launchLinearGemm(ffn_input->wrap<T>(), &weights->gate_and_up, swiglu_input, cublas_wrapper, false, weights->gate_and_up.is_transposed);
        launchSiluAndMul(swiglu_input, down_proj_input);
        launchLinearGemm(down_proj_input, &weights->down, ffn_output->wrap<T>(), cublas_wrapper, false, weights->down.is_transposed);
    }
};
