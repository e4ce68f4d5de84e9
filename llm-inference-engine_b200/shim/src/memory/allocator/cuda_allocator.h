// Drop-in for the reference's src/memory/allocator/cuda_allocator.h.  The reference's pooling allocator is never
// called (its call sites are commented out, base_allocator.h:15,27); this one is a plain pass-through.
#pragma once
#include "base_allocator.h"

class CudaAllocator : public BaseAllocator {
public:
    void unifyMalloc(void **ptr, size_t size, bool is_host = false) override {
        if (is_host) *ptr = std::malloc(size);
        else cudaMalloc(ptr, size);
    }
    void unifyFree(void *ptr, bool is_host = false) override {
        if (!ptr) return;
        if (is_host) std::free(ptr);
        else cudaFree(ptr);
    }
};
