// Abstract whole-model interface, drop-in for the reference's src/models/basemodel.h:11-61 (which does not compile as shipped: it
// includes a cublas_utils.h that does not exist and names a data member after its own type, basemodel.h:9,20).  Member names follow
// what src/models/llama/llama.h uses; the capitalised call spellings of user_entry.cpp:25-42 are provided next to the declared ones.
#pragma once

#include <cuda_runtime.h>
#include <functional>
#include <string>
#include <vector>

#include "../kernels/includes/cublas_utils.cuh"
#include "../memory/allocator/base_allocator.h"
#include "../utils/tensor.h"

using CudaDeviceProp = cudaDeviceProp;
// (token index, text) while a round is generated; index -1 delivers the complete answer
using CallBack = std::function<void(int, const char *)>;

class BaseModel {
public:
    using Strings = std::vector<std::string>;

    BaseModel(cudaStream_t s, CublasWrapper *w, BaseAllocator *a, CudaDeviceProp *p = nullptr) : stream(s), cublas_wrapper(w), allocator(a), cuda_device_prop(p) {}
    virtual ~BaseModel() = default;

    // ---- what a concrete model implements
    virtual void loadWeights(const std::string &dir_prefix) = 0;  // the per-tensor .bin files
    virtual void loadWeightsFromDummy() = 0;
    virtual void loadTokenizer(const std::string &vocabulary_file) = 0;
    virtual Strings makeInput(const std::string &history, int round, const std::string &input) const = 0;  // {history + input, history, input}
    virtual std::string makeHistory(const std::string &history, int round, const std::string &input, const std::string &output) const = 0;
    virtual std::string response(const Strings &input, CallBack printRes) = 0;  // one conversation round

    // ---- user_entry.cpp spells the three calls above with a capital letter
    Strings MakeInput(const std::string &h, int r, const std::string &in) const { return makeInput(h, r, in); }
    std::string MakeHistory(const std::string &h, int r, const std::string &in, const std::string &out) const { return makeHistory(h, r, in, out); }
    std::string Response(const Strings &in, CallBack cb) { return response(in, cb); }

    // ---- shared by every model: borrowed, never owned
    std::string model_name;
    cudaStream_t stream;
    CublasWrapper *cublas_wrapper;
    BaseAllocator *allocator;
    CudaDeviceProp *cuda_device_prop;
};
