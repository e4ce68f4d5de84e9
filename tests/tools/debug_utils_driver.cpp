// Exercises the shim's tracing / printing aids (src/utils/debug_utils.h saveTensor, src/utils/output_utils.h print_tensor / print_weight;
// the reference's are debug_utils.h:17-119 and output_utils.h:9-33) on host tensors: the dump directory comes from LLM_SAVE_TENSOR_DIR.
// Built by __graft_entry__.build() into shim/_own_programs/debug_utils_shim; tests/test_tensor_carriers.py checks what it prints and writes.
#include <iostream>
#include "src/utils/debug_utils.h"
#include "src/utils/output_utils.h"
#include "src/models/common_params.h"

int main() {
    float *f = new float[24];
    for (int i = 0; i < 24; ++i) f[i] = 0.25f * i;
    int *layer = new int(1);
    auto *r2 = new TensorWrapper<float>(Device::CPU, DataType::FP32, {4, 6}, f);
    auto *r3 = new TensorWrapper<float>(Device::CPU, DataType::FP32, {2, 3, 4}, f);
    auto *r4 = new TensorWrapper<float>(Device::CPU, DataType::FP32, {1, 2, 3, 4}, f);
    auto *r1 = new TensorWrapper<float>(Device::CPU, DataType::FP32, {24}, f);
    auto *id = new TensorWrapper<int>(Device::CPU, DataType::INT32, {1}, layer);
    saveTensor(r2, "rank2.bin");
    saveTensor(r3, "rank3.bin", 0);
    saveTensor(r4, "rank4.bin", id);  // layer 1 -> 1_rank4.bin
    saveTensor(r1, "rank1.bin");      // rank 1: an empty file
    saveTensor(r2, "late.bin", 3);    // layers past 2 are not dumped
    *layer = 5;
    saveTensor(r2, "late2.bin", id);
    print_tensor(static_cast<const Tensor *>(r3));
    print_tensor(r4);
    BaseWeight<float> w;
    w.shape = {11008, 4096};
    print_weight(&w);
    return 0;
}
