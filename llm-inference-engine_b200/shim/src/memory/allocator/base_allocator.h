// Drop-in for the reference's src/memory/allocator/base_allocator.h: the allocator TYPE is kept so that layer
// constructors link; the B200 layers never allocate on the forward path (they keep a grow-only workspace).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>

class BaseAllocator {
public:
    BaseAllocator() = default;
    virtual ~BaseAllocator() = default;
    template <typename T> void malloc(T **ptr, size_t size, bool is_host) {
        if (is_host) *ptr = static_cast<T *>(std::malloc(size));
        else cudaMalloc(reinterpret_cast<void **>(ptr), size);
    }
    template <typename T> void free(T *ptr, bool is_host = false) {
        if (!ptr) return;
        if (is_host) std::free(static_cast<void *>(ptr));
        else cudaFree(static_cast<void *>(ptr));
    }
    virtual void unifyMalloc(void **ptr, size_t size, bool is_host = false) = 0;
    virtual void unifyFree(void *ptr, bool is_host = false) = 0;
};
