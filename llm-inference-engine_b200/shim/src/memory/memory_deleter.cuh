// Drop-in for the reference's src/memory/memory_deleter.cuh:7-25: deallocate(ptr, kind), kind in
// {"new", "new[]", "cudaMalloc", "malloc"}.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>
#include <cuda_runtime.h>

template <typename T> void deallocate(T *ptr, const std::string &alloc_type) {
    if (!ptr) return;
    if (alloc_type == "cudaMalloc") cudaFree(ptr);
    else if (alloc_type == "new") delete ptr;
    else if (alloc_type == "new[]") delete[] ptr;
    else if (alloc_type == "malloc") std::free(ptr);
    else std::cerr << "Unknown allocation type for deallocation." << std::endl;
}
