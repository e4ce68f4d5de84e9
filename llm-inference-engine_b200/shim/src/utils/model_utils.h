// Drop-in for the reference's src/utils/model_utils.h:14-94: the model factory the chat entry (user_entry.cpp) calls.
//   llm::createModelWithName<T>("llama")            the model described by llama_config.json
//   llm::createDummyLLMModel<T>(tokenizer_file)      + tokenizer + dummy weights
//   llm::createRealLLMModel<T>(model_dir, tokenizer) + tokenizer + the per-tensor .bin weights under model_dir
// The reference's version needs nlohmann/json (not vendored) and a hard-coded absolute config path, and leaks objects that die with the
// factory's scope (the cuBLAS wrapper, the allocator and cudaDeviceProp are destroyed while the model still points at them,
// model_utils.h:50-74).  Here: a 20-line reader for the flat JSON object the reference ships (src/models/llama/llama_config.json), the
// path taken from $LLAMA_CONFIG_JSON or ./src/models/llama/llama_config.json (Llama-2-7B values when neither exists), and the helper
// objects kept alive for the life of the process.
#pragma once

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include "../models/basemodel.h"
#include "../models/llama/llama.h"
#include "macro.h"

namespace llm {

// value of "key" in a flat JSON object of numbers / booleans; `fallback` when the key (or the file) is missing
inline double configNumber(const std::string &text, const std::string &key, double fallback) {
    const size_t at = text.find("\"" + key + "\"");
    if (at == std::string::npos) return fallback;
    size_t p = text.find(':', at);
    if (p == std::string::npos) return fallback;
    ++p;
    while (p < text.size() && (text[p] == ' ' || text[p] == '\t' || text[p] == '\n' || text[p] == '\r')) ++p;
    if (text.compare(p, 4, "true") == 0) return 1.0;
    if (text.compare(p, 5, "false") == 0) return 0.0;
    return std::strtod(text.c_str() + p, nullptr);
}

inline std::string readConfigText() {
    const char *env = std::getenv("LLAMA_CONFIG_JSON");
    for (const std::string &path : {std::string(env ? env : ""), std::string("src/models/llama/llama_config.json")}) {
        if (path.empty()) continue;
        std::ifstream f(path);
        if (f) {
            std::stringstream ss;
            ss << f.rdbuf();
            return ss.str();
        }
    }
    return "";
}

// everything the factory needs to know about the model, with Llama-2-7B defaults (src/models/llama/llama_config.json)
struct ModelShape {
    int head_num = 32, kv_head_num = 32, head_size = 128, inter_size = 11008, num_layers = 32, max_seq_len = 64, vocab_size = 32000;
    int rotary_embedding_dim = 128, max_position_embeddings = 4096;
    float rotary_embedding_base = 10000.0f;
    bool use_dynamic_ntk = false;

    static ModelShape fromJson(const std::string &text) {
        ModelShape m;
        struct Field {
            const char *key;
            int *dst;
        };
        const Field ints[] = {{"head_num", &m.head_num},       {"kv_head_num", &m.kv_head_num}, {"head_size", &m.head_size},
                              {"inter_size", &m.inter_size},   {"num_layers", &m.num_layers},   {"max_seq_len", &m.max_seq_len},
                              {"vocab_size", &m.vocab_size},   {"rotary_embedding_dim", &m.rotary_embedding_dim},
                              {"max_position_embeddings", &m.max_position_embeddings}};
        for (const Field &f : ints) *f.dst = (int)configNumber(text, f.key, *f.dst);
        m.rotary_embedding_base = (float)configNumber(text, "rotary_embedding_base", m.rotary_embedding_base);
        m.use_dynamic_ntk = configNumber(text, "use_dynamic_ntk", m.use_dynamic_ntk ? 1.0 : 0.0) != 0.0;
        return m;
    }
};

// helpers the model borrows for its whole life: created once per process (the reference lets them die with the factory call)
struct ModelRuntime {
    cublasHandle_t cublas = nullptr;
    cublasLtHandle_t cublaslt = nullptr;
    CublasWrapper *wrapper = nullptr;
    BaseAllocator *allocator = nullptr;
    cudaDeviceProp prop;

    static ModelRuntime &get() {
        static ModelRuntime rt;
        if (!rt.wrapper) {
            cublasCreate(&rt.cublas);
            cublasLtCreate(&rt.cublaslt);
            cublasSetMathMode(rt.cublas, CUBLAS_DEFAULT_MATH);
            rt.wrapper = new CublasWrapper(rt.cublas, rt.cublaslt);
            rt.wrapper->setFP32GemmConfig();
            rt.allocator = new CudaAllocator;
            cudaGetDeviceProperties(&rt.prop, 0);
        }
        return rt;
    }
};

template <typename T> BaseModel *createModelWithName(const std::string &model_name) {
    LLM_CHECK_WITH_INFO(model_name == "llama", "Currently, only llama models are supported!");
    const ModelShape m = ModelShape::fromJson(readConfigText());
    LlamaAttentionStaticParams rope = {};
    rope.rotary_embedding_dim = m.rotary_embedding_dim, rope.rotary_embedding_base = m.rotary_embedding_base;
    rope.max_position_embeddings = m.max_position_embeddings, rope.use_dynamic_ntk = m.use_dynamic_ntk;
    ModelRuntime &rt = ModelRuntime::get();
    return new LlamaModel<T>(m.head_num, m.kv_head_num, m.head_size, m.inter_size, m.num_layers, m.vocab_size, rope, m.max_seq_len, nullptr, rt.wrapper,
                             rt.allocator, &rt.prop);
}

template <typename T> BaseModel *createDummyLLMModel(const std::string &tokenizer_file) {
    std::unique_ptr<BaseModel> model(createModelWithName<T>("llama"));
    model->loadTokenizer(tokenizer_file);
    model->loadWeightsFromDummy();
    return model.release();
}

template <typename T> BaseModel *createRealLLMModel(const std::string &model_dir, const std::string &tokenizer_file) {
    std::unique_ptr<BaseModel> model(createModelWithName<T>("llama"));
    std::cout << "Start creating model..." << std::endl;
    model->loadTokenizer(tokenizer_file);
    model->loadWeights(model_dir);
    std::cout << "Finish creating model..." << std::endl;
    return model.release();
}

}  // namespace llm
