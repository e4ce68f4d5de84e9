// Drop-in for the reference's src/utils/weight_utils.h (GPUMalloc / GPUFree / loadWeightFromBin): load-time helpers.
// The on-disk format is the reference's: one raw little-endian fp32 .bin per tensor (src/utils/weight_utils.cu:189-224).
#pragma once
#include <cstdio>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "macro.h"

template <typename T> void GPUMalloc(T **ptr, size_t size) {
    LLM_CHECK_WITH_INFO(size >= 0, "Ask cudaMalloc size " + std::to_string(size) + "< 0 is invalid.");
    CHECK(cudaMalloc((void **)(ptr), sizeof(T) * size));
}
template <typename T> void GPUFree(T *&ptr) {
    if (ptr != nullptr) {
        CHECK(cudaFree(ptr));
        ptr = nullptr;
    }
}
namespace b200shim {
template <typename T> inline T from_float(float v) { return static_cast<T>(v); }
template <> inline half from_float<half>(float v) { return __float2half(v); }
template <> inline __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
}  // namespace b200shim

// loadWeightFromBin<OutT, FileT>::loadFromFileToDevice(ptr, shape, path): read FileT elements, convert on the host, copy H2D.
template <typename OutT, typename FileT> struct loadWeightFromBin {
    static void loadFromFileToDevice(OutT *ptr, std::vector<size_t> shape, std::string filename) {
        size_t n = 1;
        for (size_t d : shape) n *= d;
        std::vector<FileT> host(n);
        std::ifstream in(filename, std::ios::in | std::ios::binary);
        if (!in.is_open()) {
            std::printf("file %s cannot be opened, loading model fails!\n", filename.c_str());
            return;
        }
        in.read(reinterpret_cast<char *>(host.data()), (std::streamsize)(n * sizeof(FileT)));
        if ((size_t)in.gcount() != n * sizeof(FileT)) {
            std::printf("file %s only has %ld of %zu bytes, loading model fails!\n", filename.c_str(), (long)in.gcount(), n * sizeof(FileT));
            return;
        }
        std::vector<OutT> conv(n);
        for (size_t i = 0; i < n; ++i) conv[i] = b200shim::from_float<OutT>(static_cast<float>(host[i]));
        CHECK(cudaMemcpy(ptr, conv.data(), sizeof(OutT) * n, cudaMemcpyHostToDevice));
    }
    static void loadFromFileToDevice(OutT *ptr, std::vector<int> shape, std::string filename) {
        loadFromFileToDevice(ptr, std::vector<size_t>(shape.begin(), shape.end()), filename);
    }
};
