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

template <typename T> BaseModel *createModelWithName(const std::string &model_name) {
    LLM_CHECK_WITH_INFO(model_name == "llama", "Currently, only llama models are supported!");
    const std::string cfg = readConfigText();
    const int head_num = (int)configNumber(cfg, "head_num", 32), kv_head_num = (int)configNumber(cfg, "kv_head_num", 32);
    const int head_size = (int)configNumber(cfg, "head_size", 128), inter_size = (int)configNumber(cfg, "inter_size", 11008);
    const int num_layers = (int)configNumber(cfg, "num_layers", 32), max_seq_len = (int)configNumber(cfg, "max_seq_len", 64);
    const int vocab_size = (int)configNumber(cfg, "vocab_size", 32000);
    LlamaAttentionStaticParams attn_static_params = {};
    attn_static_params.rotary_embedding_dim = (int)configNumber(cfg, "rotary_embedding_dim", 128);
    attn_static_params.rotary_embedding_base = (float)configNumber(cfg, "rotary_embedding_base", 10000);
    attn_static_params.max_position_embeddings = (int)configNumber(cfg, "max_position_embeddings", 4096);
    attn_static_params.use_dynamic_ntk = configNumber(cfg, "use_dynamic_ntk", 0) != 0.0;

    // process-lifetime helpers: the model keeps raw pointers to them
    static cublasHandle_t cublas_handle = nullptr;
    static cublasLtHandle_t cublaslt_handle = nullptr;
    if (!cublas_handle) {
        cublasCreate(&cublas_handle);
        cublasLtCreate(&cublaslt_handle);
        cublasSetMathMode(cublas_handle, CUBLAS_DEFAULT_MATH);
    }
    static CublasWrapper *cublas_wrapper = new CublasWrapper(cublas_handle, cublaslt_handle);
    cublas_wrapper->setFP32GemmConfig();
    static BaseAllocator *allocator = new CudaAllocator;
    static cudaDeviceProp device_prop;
    cudaGetDeviceProperties(&device_prop, 0);
    return new LlamaModel<T>(head_num, kv_head_num, head_size, inter_size, num_layers, vocab_size, attn_static_params, max_seq_len, nullptr,
                             cublas_wrapper, allocator, &device_prop);
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
