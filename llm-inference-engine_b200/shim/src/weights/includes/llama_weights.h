// Drop-in for the reference's src/weights/includes/llama_weights.h:10-45 (+ src/weights/llama_weights.cpp): the weights of a
// whole model -- per-layer weights, the final RMSNorm gamma and the two embedding tables.  Loading follows the reference's
// file naming (llama_weights.cpp:49-75); on the hot path only the members are used.
#pragma once

#include <memory>
#include <string>
#include <vector>
#include "weight.h"
#include "base_weights.h"
#include "embedding_weights.h"
#include "layer_weights.h"

template <typename T> class LlamaWeight : public Weight {
private:
    int hidden_units = 0;
    int intermediate_size = 0;
    int vocab_size = 0;
    int vocab_size_padded = 0;
    int num_layer = 0;
    WeightType weight_type = WeightType::UNSUPPORTED_W;

public:
    std::vector<std::unique_ptr<LlamaLayerWeight<T>>> llama_layer_weight;
    LayerNormWeight<T> out_rmsnorm_weight;
    EmbeddingWeight<T> post_decoder_embedding_weight;
    EmbeddingWeight<T> pre_decoder_embedding_weight;

    LlamaWeight() = default;
    LlamaWeight(int head_num, int kv_head_num, int head_size, int intermediate_size, int vocab_size, int num_layer, bool attention_bias,
                WeightType weight_type)
        : hidden_units(head_num * head_size), intermediate_size(intermediate_size), vocab_size(vocab_size), vocab_size_padded(vocab_size),
          num_layer(num_layer), weight_type(weight_type) {
        llama_layer_weight.reserve(num_layer);
        for (int l = 0; l < num_layer; ++l)
            llama_layer_weight.push_back(std::make_unique<LlamaLayerWeight<T>>(head_num, kv_head_num, head_size, intermediate_size, weight_type, attention_bias));
        GPUMalloc(&out_rmsnorm_weight.gamma, hidden_units);
        GPUMalloc(&post_decoder_embedding_weight.data, (size_t)vocab_size * hidden_units);
        GPUMalloc(&pre_decoder_embedding_weight.data, (size_t)vocab_size * hidden_units);
        pre_decoder_embedding_weight.shape = {vocab_size, hidden_units};
        post_decoder_embedding_weight.shape = {vocab_size, hidden_units};
        pre_decoder_embedding_weight.type = weight_type;
        post_decoder_embedding_weight.type = weight_type;
    }
    ~LlamaWeight() override {
        GPUFree(pre_decoder_embedding_weight.data);
        GPUFree(out_rmsnorm_weight.gamma);
        GPUFree(post_decoder_embedding_weight.data);
    }

    void loadWeightsFromFile(const std::string &weight_path) override {
        loadWeightFromBin<T, float>::loadFromFileToDevice(pre_decoder_embedding_weight.data, std::vector<int>{vocab_size, hidden_units},
                                                          weight_path + "model.embed_tokens.weight.bin");
        loadWeightFromBin<T, float>::loadFromFileToDevice(out_rmsnorm_weight.gamma, std::vector<int>{hidden_units}, weight_path + "model.norm.weight.bin");
        loadWeightFromBin<T, float>::loadFromFileToDevice(post_decoder_embedding_weight.data, std::vector<int>{vocab_size, hidden_units},
                                                          weight_path + "lm_head.weight.bin");
        for (int l = 0; l < num_layer; ++l)
            llama_layer_weight[l]->loadWeightsFromFile(weight_path + "model.layers." + std::to_string(l), weight_type);
    }
    void loadWeightsFromDummy() {
        auto fill = [](T *dev, size_t n, float v) {
            std::vector<T> h(n, b200shim::from_float<T>(v));
            CHECK(cudaMemcpy(dev, h.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
        };
        fill(out_rmsnorm_weight.gamma, hidden_units, 1.0f);
        fill(pre_decoder_embedding_weight.data, (size_t)vocab_size * hidden_units, 1.0f);
        fill(post_decoder_embedding_weight.data, (size_t)vocab_size * hidden_units, 1.0f);
        for (int l = 0; l < num_layer; ++l) llama_layer_weight[l]->loadWeightsFromFile();
    }
};
