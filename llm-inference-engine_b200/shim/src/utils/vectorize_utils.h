// The reference's src/utils/vectorize_utils.h holds device-side float4/half2 helpers used only inside its kernels.
// It is included by the kernel headers, so the name is kept; the B200 kernels have their own helpers (csrc/common.cuh).
#pragma once
#include <cuda_fp16.h>
template <typename T> struct Vec {
    using Type = T;
    static constexpr int size = 1;
};
template <> struct Vec<float> {
    using Type = float4;
    static constexpr int size = 4;
};
template <> struct Vec<half> {
    using Type = half2;
    static constexpr int size = 2;
};
