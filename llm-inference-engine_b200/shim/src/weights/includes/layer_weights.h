// Drop-in for the reference's src/weights/includes/layer_weights.h:8-42 (+ src/weights/layer_weights.cpp): the weights of
// one decoder layer.  Same constructor, same public members, same two loaders:
//   loadWeightsFromFile(path, type)  reads the reference's per-tensor fp32 .bin files (layer_weights.cpp:50-80);
//   loadWeightsFromFile()            dummy init, values rand()%10000/100000 drawn in the reference's order (:82-156).
// New: packForB200(format) builds the packed [N,K] (bf16/fp16/fp32, FP8 or INT4) copies that the decode engine streams.
#pragma once

#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include "attention_weights.h"
#include "ffn_weights.h"
#include "norm_weights.h"
#include "pack.h"
#include "../../utils/weight_utils.h"

template <typename T> class LlamaLayerWeight {
private:
    int head_num;
    int kv_head_num;
    int head_size;
    int hidden_units;
    int intermediate_size;
    WeightType weight_type;
    int bit_size = 8 * (int)sizeof(T);
    bool attention_bias;

    int qkvCols() const { return (head_num + 2 * kv_head_num) * head_size; }
    static void dummyFill(T *dev, size_t n) {
        std::vector<T> h(n);
        for (size_t i = 0; i < n; ++i) h[i] = b200shim::from_float<T>((float)(rand() % 10000) / 100000.0f);
        CHECK(cudaMemcpy(dev, h.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    }

public:
    LlamaLayerWeight() = delete;
    LlamaLayerWeight(int head_num, int kv_head_num, int head_size, int intermediate_size, WeightType weight_type, bool attention_bias)
        : head_num(head_num), kv_head_num(kv_head_num), head_size(head_size), hidden_units(head_num * head_size),
          intermediate_size(intermediate_size), weight_type(weight_type), attention_bias(attention_bias) {
        GPUMalloc(&attention_norm_weight.gamma, hidden_units);
        GPUMalloc(&ffn_norm_weight.gamma, hidden_units);
        self_attention_weight.qkv.type = weight_type;
        self_attention_weight.qkv.shape = {qkvCols(), hidden_units};
        GPUMalloc(&self_attention_weight.qkv.data, (size_t)hidden_units * qkvCols());
        self_attention_weight.output.type = weight_type;
        self_attention_weight.output.shape = {hidden_units, hidden_units};
        GPUMalloc(&self_attention_weight.output.data, (size_t)hidden_units * hidden_units);
        if (attention_bias) {
            GPUMalloc(&self_attention_weight.qkv.bias, qkvCols());
            GPUMalloc(&self_attention_weight.output.bias, hidden_units);
            CHECK(cudaMemset(self_attention_weight.qkv.bias, 0, sizeof(T) * qkvCols()));
            CHECK(cudaMemset(self_attention_weight.output.bias, 0, sizeof(T) * hidden_units));
        }
        ffn_weight.gate_and_up.type = weight_type;
        ffn_weight.gate_and_up.shape = {2 * intermediate_size, hidden_units};
        GPUMalloc(&ffn_weight.gate_and_up.data, (size_t)hidden_units * 2 * intermediate_size);
        ffn_weight.down.type = weight_type;
        ffn_weight.down.shape = {hidden_units, intermediate_size};
        GPUMalloc(&ffn_weight.down.data, (size_t)hidden_units * intermediate_size);
        // the reference's layers read every weight as [K,N] memory under an {N,K} declared shape (layer_weights.cpp:143-155)
        self_attention_weight.qkv.is_transposed = true;
        self_attention_weight.output.is_transposed = false;
        ffn_weight.gate_and_up.is_transposed = true;
        ffn_weight.down.is_transposed = true;
    }

    ~LlamaLayerWeight() {
        GPUFree(attention_norm_weight.gamma);
        GPUFree(ffn_norm_weight.gamma);
        freeWeights(&self_attention_weight.qkv);
        freeWeights(&self_attention_weight.output);
        freeWeights(&ffn_weight.gate_and_up);
        freeWeights(&ffn_weight.down);
    }

    void loadWeightsFromFile(const std::string &weight_path, WeightType type) {
        (void)type;
        auto load = [&](const std::string &suffix, std::vector<int> shape, T *ptr) {
            loadWeightFromBin<T, float>::loadFromFileToDevice(ptr, shape, weight_path + suffix);
        };
        load(".input_layernorm.weight.bin", {hidden_units}, attention_norm_weight.gamma);
        load(".post_attention_layernorm.weight.bin", {hidden_units}, ffn_norm_weight.gamma);
        load(".self_attn.qkv.weight.bin", {qkvCols(), hidden_units}, self_attention_weight.qkv.data);
        load(".self_attn.o_proj.weight.bin", {hidden_units, hidden_units}, self_attention_weight.output.data);
        load(".mlp.gate_up_proj.weight.bin", {2 * intermediate_size, hidden_units}, ffn_weight.gate_and_up.data);
        load(".mlp.down_proj.weight.bin", {hidden_units, intermediate_size}, ffn_weight.down.data);
        if (attention_bias) {
            load(".attention.wqkv.bias.bin", {qkvCols()}, self_attention_weight.qkv.bias);
            load(".attention.wo.bias.bin", {hidden_units}, self_attention_weight.output.bias);
        }
    }

    // dummy weights: same value formula and the same draw ORDER as the reference (norm gammas, o bias, down bias, down,
    // gate_up, o, qkv), so a given srand() seed gives the same model in both code bases.
    void loadWeightsFromFile() {
        dummyFill(attention_norm_weight.gamma, hidden_units);
        dummyFill(ffn_norm_weight.gamma, hidden_units);
        if (!self_attention_weight.output.bias) GPUMalloc(&self_attention_weight.output.bias, hidden_units);
        dummyFill(self_attention_weight.output.bias, hidden_units);
        if (!ffn_weight.down.bias) GPUMalloc(&ffn_weight.down.bias, hidden_units);
        dummyFill(ffn_weight.down.bias, hidden_units);
        dummyFill(ffn_weight.down.data, (size_t)hidden_units * intermediate_size);
        dummyFill(ffn_weight.gate_and_up.data, (size_t)hidden_units * 2 * intermediate_size);
        dummyFill(self_attention_weight.output.data, (size_t)hidden_units * hidden_units);
        dummyFill(self_attention_weight.qkv.data, (size_t)hidden_units * qkvCols());
        if (self_attention_weight.qkv.bias) {  // the reference drops the qkv bias in the dummy model
            GPUFree(self_attention_weight.qkv.bias);
        }
    }

    void freeWeights(BaseWeight<T> *weights) {
        GPUFree(weights->data);
        GPUFree(weights->bias);
        freePackedForB200(weights);
    }

    // New: build the packed [N,K] copies (dense T by default, or FP8_W / INT4_W).
    void packForB200(WeightType target = WeightType::UNSUPPORTED_W, int group = 128) {
        ::packForB200(&self_attention_weight.qkv, hidden_units, qkvCols(), target, group);
        ::packForB200(&self_attention_weight.output, hidden_units, hidden_units, target, group);
        ::packForB200(&ffn_weight.gate_and_up, hidden_units, 2 * intermediate_size, target, group);
        ::packForB200(&ffn_weight.down, intermediate_size, hidden_units, target, group);
    }

    LayerNormWeight<T> attention_norm_weight;
    LayerNormWeight<T> ffn_norm_weight;
    LlamaAttentionWeights<T> self_attention_weight;
    LlamaFFNWeights<T> ffn_weight;
};
