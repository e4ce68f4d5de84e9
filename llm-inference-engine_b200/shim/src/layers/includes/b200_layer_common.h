// Shared plumbing of the five layer classes: a grow-only device workspace obtained from the reference's BaseAllocator
// (allocated the first time a shape is seen, then reused -- the forward path itself never allocates, frees or synchronises),
// and the stream scope that makes every launcher of a forward() call run on the layer's stream.
#pragma once

#include <cstddef>
#include <vector>
#include <cuda_runtime.h>

#include "../../kernels/includes/b200_launchers.h"
#include "../../memory/allocator/base_allocator.h"
#include "../../memory/allocator/cuda_allocator.h"

namespace b200shim {

class Workspace {
public:
    explicit Workspace(BaseAllocator *allocator) : allocator_(allocator) {}
    ~Workspace() { release(); }
    Workspace(const Workspace &) = delete;
    Workspace &operator=(const Workspace &) = delete;

    // reserve(bytes) then take<T>(count) hands out 256-byte aligned slices; reserve() only reallocates when it must grow
    void reserve(size_t bytes) {
        cursor_ = 0;
        if (bytes <= capacity_) return;
        release();
        if (allocator_) allocator_->malloc(&base_, bytes, false);
        else cudaMalloc(reinterpret_cast<void **>(&base_), bytes);
        LLM_CHECK_WITH_INFO(base_ != nullptr, "layer workspace allocation failed");
        capacity_ = bytes;
    }
    template <typename T> T *take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        LLM_CHECK_WITH_INFO(cursor_ + bytes <= capacity_, "layer workspace overflow");
        T *p = reinterpret_cast<T *>(base_ + cursor_);
        cursor_ += bytes;
        return p;
    }
    static size_t padded(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }
    void release() {
        if (base_) {
            if (allocator_) allocator_->free(base_, false);
            else cudaFree(base_);
        }
        base_ = nullptr;
        capacity_ = cursor_ = 0;
    }

private:
    BaseAllocator *allocator_;
    char *base_ = nullptr;
    size_t capacity_ = 0, cursor_ = 0;
};

// The reference stores the stream handed to the constructor and never uses it; its examples even pass an uninitialised
// handle (self_decoder_example.cpp:51,171).  The shim therefore only adopts a stream the caller set explicitly through
// setStream(); otherwise everything stays on the legacy default stream, exactly like the reference.
class StreamScope {
public:
    explicit StreamScope(cudaStream_t s) : saved_(b200GetStream()) { b200SetStream(s); }
    ~StreamScope() { b200SetStream(saved_); }

private:
    cudaStream_t saved_;
};

}  // namespace b200shim
