// Drop-in for the reference's src/weights/includes/base_weights.h:7-33, extended for the B200 formats.
// Existing fields (type, shape, data, bias, is_transposed) keep their meaning: `data` is the reference's dense
// [K,N] row-major matrix.  New fields describe the PACKED form the B200 kernels stream:
// [N,K] row-major, optionally FP8-e4m3 (per-row fp32 scale) or INT4 (grouped scale + zero point).
#pragma once

#include <cstdint>
#include <type_traits>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

enum class WeightType { FP32_W, FP16_W, INT8_W, UNSUPPORTED_W, BF16_W, FP8_W, INT4_W };

template <typename T> inline WeightType getWeightType() {
    using U = typename std::remove_const<T>::type;
    if (std::is_same<U, float>::value) return WeightType::FP32_W;
    if (std::is_same<U, __half>::value) return WeightType::FP16_W;
    if (std::is_same<U, __nv_bfloat16>::value) return WeightType::BF16_W;
    if (std::is_same<U, int8_t>::value) return WeightType::INT8_W;
    return WeightType::UNSUPPORTED_W;
}

template <typename T> struct BaseWeight {
    WeightType type;
    std::vector<int> shape;
    T *data = nullptr, *bias = nullptr;
    bool is_transposed = false;

    // ---- B200 extension (all optional; filled by packForB200() in weights/includes/pack.h)
    void *packed = nullptr;          // [N,K] row-major: T, or uint8 (FP8: 1 byte/weight, INT4: 2 weights/byte)
    void *packed_scales = nullptr;   // FP8: float[N]; INT4: T[N, K/group]
    void *packed_zeros = nullptr;    // INT4: uint8[N, K/group]
    WeightType packed_type = WeightType::UNSUPPORTED_W;  // FP32_W/FP16_W/BF16_W (dense), FP8_W, INT4_W
    int group_size = 0;
    const void *packed_from = nullptr;  // the `data` pointer the packed copy was made from
};
