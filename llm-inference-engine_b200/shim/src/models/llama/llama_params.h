// Drop-in for the reference's src/models/llama/llama_params.h:3-21 (field names, order and defaults kept so that
// designated initialisers in examples/cpp compile unchanged), plus the tensor-parallel descriptor that is new here.
#pragma once

struct LlamaAttentionStaticParams {
    int rotary_embedding_dim;
    float rotary_embedding_base;
    int max_position_embeddings;
    bool use_dynamic_ntk;
    int head_size = 128;
    int head_num = 32;
    int kv_head_num = 32;
};

struct LlamaAttentionDynamicParams {
    int batch_size;
    int num_tokens;
    int max_q_len;
    int max_k_len;
    int num_layers;
    bool is_context = false;
};

// New (no reference counterpart): which shard of the heads / FFN columns this process holds.
struct TensorParallelParams {
    int world = 1;
    int rank = 0;
    void *nccl_comm = nullptr;  // ncclComm_t
};
