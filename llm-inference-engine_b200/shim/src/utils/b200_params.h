// Parameter carriers shared by launchers, layers and models: the string->number maps of the reference's src/utils/params.h and the
// attention parameter structs of src/models/llama/llama_params.h:3-21 (both headers forward here).  Field names, order and defaults of
// the two Llama structs are the API: the reference's examples fill them member by member and with brace initialisers.
#pragma once
#include <string>
#include <unordered_map>

template <typename V> using MapStringTo = std::unordered_map<std::string, V>;
using MapStringToInt = MapStringTo<int>;      // launchSampling: {"vocab_size", "step", "end_id"}
using MapStringToFloat = MapStringTo<float>;

// fixed per model
struct LlamaAttentionStaticParams {
    int rotary_embedding_dim;
    float rotary_embedding_base;
    int max_position_embeddings;
    bool use_dynamic_ntk;
    int head_size = 128, head_num = 32, kv_head_num = 32;
};

// per forward() call
struct LlamaAttentionDynamicParams {
    int batch_size, num_tokens, max_q_len, max_k_len, num_layers;
    bool is_context = false;
};

// New (no reference counterpart): which shard of the heads / FFN columns this process holds.
struct TensorParallelParams {
    int world = 1, rank = 0;
    void *nccl_comm = nullptr;  // ncclComm_t
};
