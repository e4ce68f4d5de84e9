// Drop-in for the reference's src/models/basemodel.h:11-61: the abstract model interface (tokenizer / weight loading, prompt and
// history assembly, response()).  The reference's version does not compile (it includes a non-existent cublas_utils.h and names a
// member after its own type, basemodel.h:9,20); member names follow what src/models/llama/llama.h uses.
#pragma once

#include <functional>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../utils/tensor.h"
#include "../memory/allocator/base_allocator.h"
#include "../kernels/includes/cublas_utils.cuh"

// Callback printing the generated content of a conversation round: (token index, text); index -1 carries the whole answer.
using CallBack = std::function<void(int, const char *)>;
using CudaDeviceProp = cudaDeviceProp;

class BaseModel {
public:
    std::string model_name;
    cudaStream_t stream;
    CublasWrapper *cublas_wrapper;
    BaseAllocator *allocator;
    CudaDeviceProp *cuda_device_prop;

    BaseModel(cudaStream_t stream, CublasWrapper *cublas_wrapper, BaseAllocator *allocator, CudaDeviceProp *cuda_device_prop = nullptr)
        : stream(stream), cublas_wrapper(cublas_wrapper), allocator(allocator), cuda_device_prop(cuda_device_prop) {}
    virtual ~BaseModel() = default;

    virtual void loadTokenizer(const std::string &file) = 0;
    virtual void loadWeights(const std::string &file) = 0;
    virtual void loadWeightsFromDummy() = 0;
    // {history + input, history, input} of this round (llama.cpp:137-145)
    virtual std::vector<std::string> makeInput(const std::string &history, int round, const std::string &input) const = 0;
    virtual std::string makeHistory(const std::string &history, int round, const std::string &input, const std::string &output) const = 0;
    virtual std::string response(const std::vector<std::string> &input, CallBack printRes) = 0;

    // the spellings the reference's chat entry uses (user_entry.cpp:25-42)
    std::vector<std::string> MakeInput(const std::string &history, int round, const std::string &input) const { return makeInput(history, round, input); }
    std::string MakeHistory(const std::string &history, int round, const std::string &input, const std::string &output) const {
        return makeHistory(history, round, input, output);
    }
    std::string Response(const std::vector<std::string> &input, CallBack printRes) { return response(input, printRes); }
};
