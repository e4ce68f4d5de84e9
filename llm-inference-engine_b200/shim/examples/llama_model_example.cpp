// Whole-model smoke driver in the style of the reference's examples/cpp/*: builds a small LlamaModel<T> with dummy weights, runs one
// conversation round through response() (the generation loop, include/b200llm.h b200_generate) and prints the generated ids.
//   usage: llama_model_example [f32|f16] [token_limit]
#include <cstdio>
#include <cstring>
#include <string>
#include "../src/models/llama/llama.h"
#include "../src/utils/model_utils.h"

template <typename T> static int run(int limit) {
    const int head_num = 4, kv_head_num = 2, head_size = 128, inter_size = 768, num_layers = 2, vocab = 32000, max_seq_len = 64;
    LlamaAttentionStaticParams attn = {};
    attn.rotary_embedding_dim = 128, attn.rotary_embedding_base = 10000.0f, attn.max_position_embeddings = 2048, attn.use_dynamic_ntk = false;
    cudaStream_t stream = nullptr;
    cublasHandle_t cublas_handle;
    cublasLtHandle_t cublaslt_handle;
    cublasCreate(&cublas_handle);
    cublasLtCreate(&cublaslt_handle);
    CublasWrapper cublas(cublas_handle, cublaslt_handle);
    BaseAllocator *allocator = new CudaAllocator;
    srand(1234);
    LlamaModel<T> model(head_num, kv_head_num, head_size, inter_size, num_layers, vocab, attn, max_seq_len, stream, &cublas, allocator);
    model.loadWeightsFromDummy();
    model.setOutputTokenLimit(limit);
    model.setTopK(1);  // the dummy embedding / LM head are constant (as the reference's dummy loader): greedy picks id 0, never the end id
    int count = 0;
    std::string all = model.response(model.makeInput("", 0, "Hey, are you conscious? Can you talk to me?"), [&](int index, const char *text) {
        if (index >= 0) ++count;
        printf("token %d: %s\n", index, text);
    });
    std::vector<int> again = model.generateIds({1, 18637, 29892, 526, 366, 19861, 29973, 1815, 366, 5193, 304, 592, 29973}, limit);
    printf("generated %d tokens; second run generated %zu\n", count, again.size());
    const bool ok = count == (int)again.size() && count == limit;
    printf(ok ? "llama_model_example passed\n" : "llama_model_example FAILED\n");
    delete allocator;
    return ok ? 0 : 1;
}

// the reference's chat entry without the terminal: factory -> two rounds of MakeInput / Response / MakeHistory (user_entry.cpp:9-46)
static int run_factory(const char *config_json) {
    setenv("LLAMA_CONFIG_JSON", config_json, 1);
    srand(7);
    BaseModel *model = llm::createModelWithName<float>("llama");
    model->loadWeightsFromDummy();
    LlamaModel<float> *lm = dynamic_cast<LlamaModel<float> *>(model);
    lm->setTopK(1);
    lm->setOutputTokenLimit(6);
    std::string history = "";
    int tokens = 0;
    for (int round = 0; round < 2; ++round) {
        const std::string input = round == 0 ? "first question" : "second question";
        std::string ret = model->Response(model->MakeInput(history, round, input), [&](int index, const char *content) {
            if (index >= 0) ++tokens;
        });
        history = model->MakeHistory(history, round, input, ret);
        printf("round %d: %s\n", round, ret.c_str());
    }
    const bool ok = tokens == 12 && history.find("second question") != std::string::npos;
    printf(ok ? "chat factory passed\n" : "chat factory FAILED (%d tokens)\n", tokens);
    delete model;
    return ok ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc > 2 && !strcmp(argv[1], "factory")) return run_factory(argv[2]);
    const int limit = argc > 2 ? atoi(argv[2]) : 12;
    if (argc > 1 && !strcmp(argv[1], "f16")) return run<half>(limit);
    return run<float>(limit);
}
