// Drop-in for the reference's src/utils/debug_utils.h:17-119 (its SAVE_DATA tracing hook): saveTensor dumps a tensor's raw
// elements to a file so that two runs can be compared offline.  Same three overloads and the same rules:
//   * bytes written = sizeof(T) x the product of the dimensions for tensors of rank 2, 3 or 4, nothing for any other rank
//     (debug_utils.h:24-36: rows x columns is only formed for those ranks);
//   * (tensor, name)            -> <dir>/<name>
//   * (tensor, name, layer id)  -> <dir>/<id>_<name>, and only for layers 0..2 (id as a CPU TensorWrapper<int> or as an int);
//   * a line "Saving intermediate tensor in <name>" on stdout; a file that cannot be opened is skipped silently.
// <dir> is the reference's hard-coded "/home/data/" unless the environment variable LLM_SAVE_TENSOR_DIR names another directory.
// Host tensors (Device::CPU / CPU_PINNED) are written directly; device tensors are copied on the library's current stream first.
#pragma once

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "macro.h"
#include "tensor.h"
#include "../weights/includes/base_weights.h"

namespace b200shim {
inline std::string save_dir() {
    const char *env = std::getenv("LLM_SAVE_TENSOR_DIR");
    std::string dir = env && *env ? env : "/home/data/";
    if (dir.back() != '/') dir += '/';
    return dir;
}
template <typename T> inline void save_tensor_as(TensorWrapper<T> *input, const std::string &shown, const std::string &file_name) {
    size_t count = 0;
    const size_t rank = input->shape.size();
    if (rank >= 2 && rank <= 4) {
        count = 1;
        for (int d : input->shape) count *= (size_t)d;
    }
    std::vector<T> host(count);
    if (count > 0) {
        if (input->device == Device::GPU) CHECK(cudaMemcpy(host.data(), input->data, sizeof(T) * count, cudaMemcpyDeviceToHost));
        else std::copy(input->data, input->data + count, host.begin());
    }
    std::cout << "Saving intermediate tensor in " << shown << "\n";
    std::ofstream file(save_dir() + file_name, std::ofstream::binary);
    if (file) file.write(reinterpret_cast<const char *>(host.data()), (std::streamsize)(sizeof(T) * count));
}
}  // namespace b200shim

template <typename T> void saveTensor(TensorWrapper<T> *input, const std::string &filename) { b200shim::save_tensor_as(input, filename, filename); }

template <typename T> void saveTensor(TensorWrapper<T> *input, const std::string &filename, int layer_id) {
    if (layer_id > 2) return;  // the first three layers only (debug_utils.h:55-58, 95-97)
    b200shim::save_tensor_as(input, filename, std::to_string(layer_id) + "_" + filename);
}

template <typename T> void saveTensor(TensorWrapper<T> *input, const std::string &filename, TensorWrapper<int> *layer_id) {
    saveTensor(input, filename, layer_id->getVal());
}
