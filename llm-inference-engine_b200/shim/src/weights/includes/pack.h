// pack.h (new, no reference counterpart) -- makes the packed B200 copy of a reference linear weight:
// dense [K,N] `data`  ->  [N,K] row-major, optionally FP8-e4m3 (per-row scale) or INT4 (group scale + zero point).
// Done once at load time (it allocates); afterwards launchLinearGemm() streams the packed copy.
#pragma once
#include <cuda_runtime.h>
#include "base_weights.h"
#include "../../utils/macro.h"
#include "../../utils/tensor.h"

// K, N: logical reduction / output sizes of the linear (memory of w->data is [K,N] whatever w->shape says).
template <typename T> void packForB200(BaseWeight<T> *w, int K, int N, WeightType target = WeightType::UNSUPPORTED_W, int group = 128) {
    if (w->packed && w->packed_from == w->data) return;
    LLM_CHECK_WITH_INFO(w->data != nullptr, "packForB200: weight has no data");
    if (target == WeightType::UNSUPPORTED_W) target = getWeightType<T>();
    T *nk = nullptr;
    CHECK(cudaMalloc(reinterpret_cast<void **>(&nk), sizeof(T) * (size_t)K * N));
    B200_CALL(b200_transpose2d(w->data, nk, K, N, b200DType<T>(), nullptr));
    if (target == WeightType::FP8_W) {
        void *q = nullptr;
        float *sc = nullptr;
        CHECK(cudaMalloc(&q, (size_t)K * N));
        CHECK(cudaMalloc(reinterpret_cast<void **>(&sc), sizeof(float) * N));
        B200_CALL(b200_quantize_fp8(nk, q, sc, N, K, b200DType<T>(), nullptr));
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaFree(nk));
        w->packed = q, w->packed_scales = sc, w->packed_zeros = nullptr;
    } else if (target == WeightType::INT4_W) {
        void *q = nullptr, *sc = nullptr, *z = nullptr;
        CHECK(cudaMalloc(&q, (size_t)K * N / 2));
        CHECK(cudaMalloc(&sc, sizeof(T) * (size_t)N * (K / group)));
        CHECK(cudaMalloc(&z, (size_t)N * (K / group)));
        B200_CALL(b200_quantize_int4(nk, q, sc, z, N, K, group, b200DType<T>(), nullptr));
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaFree(nk));
        w->packed = q, w->packed_scales = sc, w->packed_zeros = z, w->group_size = group;
    } else {
        CHECK(cudaDeviceSynchronize());
        w->packed = nk, w->packed_scales = nullptr, w->packed_zeros = nullptr;
    }
    w->packed_type = target;
    w->packed_from = w->data;
}

template <typename T> void freePackedForB200(BaseWeight<T> *w) {
    if (w->packed) cudaFree(w->packed);
    if (w->packed_scales) cudaFree(w->packed_scales);
    if (w->packed_zeros) cudaFree(w->packed_zeros);
    w->packed = w->packed_scales = w->packed_zeros = nullptr;
    w->packed_from = nullptr;
}
