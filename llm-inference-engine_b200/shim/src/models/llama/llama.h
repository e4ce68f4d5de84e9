// Drop-in for the reference's src/models/llama/llama.h:13-214 + llama.cpp:1-400: LlamaModel<T>, the whole-model class -- weights,
// tokenizer, KV-cache ownership and the generation loop (embedding -> context decoder -> final RMSNorm -> LM head -> top-k ->
// sampling, then one self-decoder step per token).  The reference's class is dead code (it does not compile and hard-codes a
// 13-token prompt); this one keeps its constructor and public calls and runs the loop through
// b200_generate (include/b200llm.h): no allocation, no device synchronisation and no host round trip of the token id inside the loop.
// The linears are packed once at load time (LlamaLayerWeight::packForB200).
//
// Differences a caller can see: response() tokenises input[0] with the loaded tokenizer (the reference overwrites it with 13
// hard-coded ids, llama.cpp:328-340 -- still what happens when no tokenizer was loaded); printRes() is called for every token after
// the loop has finished rather than while it runs; generateIds() is new (the token-level entry point).
#pragma once

#include <memory>
#include <string>
#include <vector>
#include "../basemodel.h"
#include "llama_params.h"
#include "../tokenizer.h"
#include "../../weights/includes/llama_weights.h"
#include "../../kernels/includes/b200_launchers.h"
#include "../../memory/allocator/cuda_allocator.h"

template <typename T> class LlamaModel : public BaseModel {
private:
    int head_num, kv_head_num, head_size, inter_size, num_layers, vocab_size, hidden_units, max_seq_len;
    LlamaAttentionStaticParams attn_params;
    float rmsnorm_eps = 1e-5f;
    int output_token_limit = 20;  // llama.h:27
    int eos_token_id = 2;
    int K = 4;  // llama.h:45
    std::string prompt = "";
    bool tokenizer_loaded = false;
    WeightType packed_format = WeightType::UNSUPPORTED_W;

    Tokenizer tokenizer;
    std::unique_ptr<LlamaWeight<T>> llama_weights;

    b200_decoder_t *engine = nullptr;
    void *scratch = nullptr, *k_cache = nullptr, *v_cache = nullptr, *workspace = nullptr;
    size_t workspace_bytes = 0;

    static int wformat(WeightType t) { return t == WeightType::FP8_W ? B200_W_FP8E4M3 : (t == WeightType::INT4_W ? B200_W_INT4 : B200_W_DENSE); }

    void buildEngine() {
        if (engine) return;
        B200_CALL(b200_workspace_ensure(0));  // split-K / stream-K partials of the prefill's tensor-core linears (library-owned)
        b200_decoder_config_t c = {};
        c.hidden = hidden_units, c.head_num = head_num, c.kv_head_num = kv_head_num, c.head_size = head_size;
        c.inter_size = inter_size, c.num_layers = num_layers, c.max_seq_len = max_seq_len, c.max_batch = 1;  // the reference is batch 1
        c.dtype = b200DType<T>();
        auto &first = llama_weights->llama_layer_weight[0]->self_attention_weight.qkv;
        LLM_CHECK_WITH_INFO(first.packed != nullptr, "LlamaModel: load the weights before generating");
        c.w_format = wformat(first.packed_type), c.group = first.group_size;
        c.rmsnorm_eps = rmsnorm_eps;
        c.rotary_dim = attn_params.rotary_embedding_dim < head_size ? attn_params.rotary_embedding_dim : head_size;
        c.rotary_base = attn_params.rotary_embedding_base;
        c.tp_world = 1, c.tp_rank = 0;
        engine = b200_decoder_create(&c);
        LLM_CHECK_WITH_INFO(engine != nullptr, std::string("b200_decoder_create: ") + b200_last_error_string());
        for (int l = 0; l < num_layers; ++l) {
            LlamaLayerWeight<T> *w = llama_weights->llama_layer_weight[l].get();
            auto lin = [](const BaseWeight<T> &b) { return b200_linear_weight_t{b.packed, b.packed_scales, b.packed_zeros}; };
            b200_layer_weights_t lw = {};
            lw.attn_norm_gamma = w->attention_norm_weight.gamma;
            lw.qkv = lin(w->self_attention_weight.qkv), lw.qkv_bias = w->self_attention_weight.qkv.bias;
            lw.o = lin(w->self_attention_weight.output), lw.o_bias = w->self_attention_weight.output.bias;
            lw.ffn_norm_gamma = w->ffn_norm_weight.gamma;
            lw.gate_up = lin(w->ffn_weight.gate_and_up), lw.down = lin(w->ffn_weight.down);
            B200_CALL(b200_decoder_set_layer(engine, l, &lw));
        }
        const size_t sb = b200_decoder_scratch_bytes(engine);
        CHECK(cudaMalloc(&scratch, sb));
        B200_CALL(b200_decoder_set_scratch(engine, scratch, sb));
        // KV-cache ownership: the model owns [L, 1, Hkv, max_seq_len, d] for K and V (llama.cpp:47-48)
        const size_t cache = sizeof(T) * (size_t)num_layers * kv_head_num * max_seq_len * head_size;
        CHECK(cudaMalloc(&k_cache, cache));
        CHECK(cudaMalloc(&v_cache, cache));
    }
    void packAll() {
        for (auto &l : llama_weights->llama_layer_weight) l->packForB200(packed_format);
    }

public:
    LlamaModel(int head_num, int kv_head_num, int head_size, int inter_size, int num_layers, int vocab_size,
               const LlamaAttentionStaticParams &attention_static_params, int max_seq_len, cudaStream_t stream, CublasWrapper *cublas_wrapper,
               BaseAllocator *allocator, CudaDeviceProp *cuda_device_prop = nullptr)
        : BaseModel(stream, cublas_wrapper, allocator, cuda_device_prop), head_num(head_num), kv_head_num(kv_head_num), head_size(head_size),
          inter_size(inter_size), num_layers(num_layers), vocab_size(vocab_size), hidden_units(head_num * head_size), max_seq_len(max_seq_len),
          attn_params(attention_static_params) {
        model_name = "llama";
        llama_weights = std::make_unique<LlamaWeight<T>>(head_num, kv_head_num, head_size, inter_size, vocab_size, num_layers, false,
                                                         getWeightType<T>());
    }
    ~LlamaModel() override {
        if (engine) b200_decoder_destroy(engine);
        cudaFree(scratch), cudaFree(k_cache), cudaFree(v_cache), cudaFree(workspace);
    }
    LlamaModel(const LlamaModel &) = delete;
    LlamaModel &operator=(const LlamaModel &) = delete;

    // New: stream FP8_W / INT4_W copies of the linears instead of T (call before loading the weights)
    void setWeightFormat(WeightType t) { packed_format = t; }
    void setOutputTokenLimit(int n) { output_token_limit = n; }
    void setTopK(int k) { K = k; }

    void loadTokenizer(const std::string &file) override {
        tokenizer.Initialize(file);
        tokenizer_loaded = true;
    }
    void loadWeights(const std::string &file) override {
        llama_weights->loadWeightsFromFile(file);
        packAll();
    }
    void loadWeightsFromDummy() override {
        llama_weights->loadWeightsFromDummy();
        packAll();
    }

    std::vector<std::string> makeInput(const std::string &history, int round, const std::string &input) const override {
        return {(round == 0 ? "" : history) + input, history, input};
    }
    std::string makeHistory(const std::string &history, int round, const std::string &input, const std::string &output) const override {
        return (round == 0 ? prompt : history) + input + output;
    }

    // New: the token-level loop.  Returns the generated ids (without the end-of-sequence id).
    std::vector<int> generateIds(const std::vector<int> &prompt_ids, int max_new_tokens) {
        LLM_CHECK_WITH_INFO(!prompt_ids.empty() && max_new_tokens >= 1, "LlamaModel::generateIds: empty prompt or no tokens requested");
        buildEngine();
        b200_generate_params_t gp = {};
        gp.embedding = llama_weights->pre_decoder_embedding_weight.data;
        gp.final_gamma = llama_weights->out_rmsnorm_weight.gamma;
        gp.lm_head = llama_weights->post_decoder_embedding_weight.data;  // [vocab, hidden]: the reference multiplies with trans_b = true
        gp.vocab = vocab_size, gp.top_k = K, gp.end_id = eos_token_id, gp.max_new_tokens = max_new_tokens, gp.check_every = 4;
        const int n_prompt = (int)prompt_ids.size();
        const size_t need = b200_generate_workspace_bytes(engine, &gp, 1, n_prompt);
        LLM_CHECK_WITH_INFO(need != 0, std::string("b200_generate: ") + b200_last_error_string());
        if (need > workspace_bytes) {
            cudaFree(workspace);
            CHECK(cudaMalloc(&workspace, need));
            workspace_bytes = need;
        }
        std::vector<int> out((size_t)max_new_tokens);
        int n = 0;
        B200_CALL(b200_generate(engine, &gp, prompt_ids.data(), 1, n_prompt, k_cache, v_cache, workspace, workspace_bytes, out.data(), &n, b200GetStream()));
        out.resize((size_t)n);
        return out;
    }

    // One conversation round, batch 1 (llama.cpp:322-398)
    std::string response(const std::vector<std::string> &input, CallBack printRes) override {
        std::vector<int> ids;
        if (tokenizer_loaded && !input.empty() && !input[0].empty()) ids = tokenizer.Encode(input[0]);
        if (ids.empty()) ids = {1, 18637, 29892, 526, 366, 19861, 29973, 1815, 366, 5193, 304, 592, 29973};  // llama.cpp:328
        for (int &id : ids) id %= vocab_size;
        const std::vector<int> gen = generateIds(ids, output_token_limit);
        std::string ret_string;
        for (size_t i = 0; i < gen.size(); ++i) {
            const std::string piece = tokenizer_loaded ? tokenizer.Decode(std::vector<int>{gen[i]}) : ("<" + std::to_string(gen[i]) + ">");
            ret_string += piece;
            if (printRes) printRes((int)i, piece.c_str());
        }
        if (printRes) printRes(-1, ret_string.c_str());
        return ret_string;
    }
};
