// The small weight carriers of the reference's src/weights/includes/{weight,norm_weights,embedding_weights,attention_weights,ffn_weights}.h
// in one place (those headers forward here so that the reference's include paths keep working).  Names and members are the API:
// the layer classes, the examples and the unit tests of the reference address them by name.
#pragma once
#include <string>
#include "base_weights.h"

// RMSNorm scale vector [hidden] (device memory)
template <typename T> struct LayerNormWeight {
    T *gamma = nullptr;
};

// token / LM-head table: BaseWeight::data = [vocab, hidden]
template <typename T> struct EmbeddingWeight : public BaseWeight<T> {};

// Linears of the attention block.  `data` is the reference's [K,N] memory:
//   qkv    K = hidden, N = (head_num + 2 kv_head_num) * head_size       output  K = head_num * head_size, N = hidden (.bias = o-proj bias)
template <typename T> struct LlamaAttentionWeights {
    BaseWeight<T> qkv, output;
};

// Linears of the FFN block, [K,N] memory: gate_and_up K = hidden, N = 2 * inter (gate columns, then up columns); down K = inter, N = hidden.
// `gate` and `up` are declared by the reference and unused by its layers.
template <typename T> struct LlamaFFNWeights {
    BaseWeight<T> gate, up, down, gate_and_up;
};

// what LlamaWeight<T> implements: load every tensor found under a path prefix
class Weight {
public:
    virtual void loadWeightsFromFile(const std::string &weight_path) = 0;
    virtual ~Weight() = default;
};
