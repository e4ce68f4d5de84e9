// Drop-in for the reference's src/utils/cuda_debug_utils.cuh:7-25 (its PRINT_DATA tracing hook): a one-thread kernel
// that prints elements 0 and 1 of a device buffer, and for a "target" buffer also elements 128..131 and 1024 -- launched as
// print_data<<<1, 1>>>(ptr[, true]) after a launcher while debugging.  Values are printed as "%f", whatever T is.
#pragma once

#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <typename T> __global__ void print_data(T *src1, bool is_target = false) {
    if (threadIdx.x != 0) return;
    const int first[2] = {0, 1}, more[5] = {128, 129, 130, 131, 1024};
    for (int i : first) printf("%dth = %f\n", i, (double)(float)src1[i]);
    if (is_target)
        for (int i : more) printf("%dth = %f\n", i, (double)(float)src1[i]);
}
