// Drop-in for the reference's src/utils/tensor.h (Device, DataType, Tensor, TensorWrapper<T>, TensorMap; tensor.h:18-295):
// same public members and method names, so src/layers, examples/cpp and tests/unit_tests compile unchanged.
// Differences, all additive or defensive:
//   * DataType gains BF16 and FP8; getTensorType<__nv_bfloat16>() works;
//   * TensorWrapper is a non-owning VIEW: its destructor frees nothing (the reference frees `data` in the
//     destructor while callers free it too -- SURVEY D10); TensorMap does not delete its tensors either.
#pragma once

#include <cstdint>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <numeric>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "macro.h"
#include "string_utils.h"
#include "../memory/memory_deleter.cuh"

enum class Device { CPU_PINNED, CPU, GPU };

enum class DataType { FP32, FP16, INT8, INT32, BOOL, BYTES, UNSUPPORTED, BF16, FP8 };

template <typename T> inline DataType getTensorType() {
    using U = typename std::remove_const<T>::type;
    if (std::is_same<U, float>::value) return DataType::FP32;
    if (std::is_same<U, half>::value) return DataType::FP16;
    if (std::is_same<U, __nv_bfloat16>::value) return DataType::BF16;
    if (std::is_same<U, int8_t>::value) return DataType::INT8;
    if (std::is_same<U, int>::value) return DataType::INT32;
    if (std::is_same<U, bool>::value) return DataType::BOOL;
    if (std::is_same<U, char>::value) return DataType::BYTES;
    return DataType::UNSUPPORTED;
}

// element type -> b200_dtype_t of the C ABI (floating types only)
template <typename T> inline int b200DType() {
    static_assert(std::is_same<T, float>::value || std::is_same<T, half>::value || std::is_same<T, __nv_bfloat16>::value,
                  "libb200llm handles float, half and __nv_bfloat16");
    return std::is_same<T, float>::value ? B200_F32 : (std::is_same<T, half>::value ? B200_F16 : B200_BF16);
}

template <typename T> class TensorWrapper;

// Untyped base of every tensor handed to a launcher or stored in a TensorMap: where the data lives, what the elements are, and the extents.
// It carries no data pointer; callers down-cast with wrap<T>() (an unchecked static_cast, as in the reference -- the dtype field is what the
// launchers trust).  size() multiplies the extents in int arithmetic (so a tensor is limited to 2^31-1 elements, which the [L,B,Hkv,S,d]
// cache of the BASELINE configs respects per tensor) and is 0 for an empty shape.
class Tensor {
public:
    Device device;
    DataType dtype;
    std::vector<int> shape;

    Tensor() = default;
    Tensor(const Device &device, const DataType &dtype, const std::vector<int> &shape) : device(device), dtype(dtype), shape(shape) {}
    virtual ~Tensor() = default;

    virtual int size() const {
        if (shape.empty()) return 0;
        return std::accumulate(shape.begin(), shape.end(), 1, std::multiplies<int>());
    }

    template <typename T> TensorWrapper<T> *wrap() { return static_cast<TensorWrapper<T> *>(this); }

    std::string deviceString() const { return device == Device::GPU ? "GPU" : (device == Device::CPU ? "CPU" : "CPU_PINNED"); }

    static const char *typeString(DataType t) {
        switch (t) {
            case DataType::FP32: return "FP32";
            case DataType::FP16: return "FP16";
            case DataType::BF16: return "BF16";
            case DataType::FP8: return "FP8";
            case DataType::INT8: return "INT8";
            case DataType::INT32: return "INT32";
            case DataType::BOOL: return "BOOL";
            case DataType::BYTES: return "BYTES";
            default: return "UNSUPPORTED";
        }
    }

    virtual std::string toString() const {
        return fmtstr("Tensor[device = %s, type = %s, shape = %s]", deviceString().c_str(), typeString(dtype), vec2str(shape).c_str());
    }
};

// Typed view: Tensor + a borrowed pointer.  The constructor with data refuses a dtype that does not match T; size() is 0 while the pointer is
// null, which is how TensorMap recognises a tensor that was declared but never given storage.  getVal() reads host tensors only (step,
// layer_id): asking a device tensor for a value throws instead of dereferencing device memory on the host.
template <typename T> class TensorWrapper : public Tensor {
public:
    T *data = nullptr;

    TensorWrapper(const Device &device, const DataType &dtype, const std::vector<int> &shape) : Tensor(device, dtype, shape) {}
    TensorWrapper(const Device &device, const DataType &dtype, const std::vector<int> &shape, T *const &data)
        : Tensor(device, dtype, shape), data(data) {
        LLM_CHECK_WITH_INFO(getTensorType<T>() == dtype, "Passed in data type should be same as dtype in params");
    }
    ~TensorWrapper() override = default;  // a view: the caller owns `data`

    int size() const override { return data == nullptr ? 0 : Tensor::size(); }

    inline T getVal(const int &id) const {
        LLM_CHECK(device == Device::CPU);
        return data[id];
    }
    inline T getVal() const { return getVal(0); }
    inline T *getPtr() const { return data; }
    inline T *getPtrByOffset(const int &offset) const { return data + offset; }

    std::string toString() const override {
        return fmtstr("Tensor[device = %s, type = %s, shape = %s, data = %p]", deviceString().c_str(), typeString(dtype),
                      vec2str(shape).c_str(), (void *)data);
    }
};

// String-keyed bundle of tensors: the argument convention of every layer's forward().  Keys are the contract (SURVEY.md 8b lists them per
// layer); a lookup of a missing key throws and names the keys that are present.
class TensorMap {
public:
    std::unordered_map<std::string, Tensor *> tensor_map;

    TensorMap() = default;
    // brace-initialised maps are checked: an entry with size() == 0 (null data or empty shape) is an error (tensor.h:198-212)
    TensorMap(std::initializer_list<std::pair<std::string, Tensor *>> init) {
        for (const auto &kv : init) {
            LLM_CHECK_WITH_INFO(isValid(kv.second), fmtstr("%s is not a valid tensor, skipping insert into TensorMap", kv.first.c_str()));
            insert(kv.first, kv.second);
        }
    }
    // maps copied from an unordered_map silently drop such entries (tensor.h:214-220)
    TensorMap(const std::unordered_map<std::string, Tensor *> &init) {
        for (const auto &kv : init)
            if (isValid(kv.second)) insert(kv.first, kv.second);
    }
    ~TensorMap() = default;  // does not own its tensors

    inline size_t size() const { return tensor_map.size(); }
    inline bool isExist(const std::string &key) const { return tensor_map.find(key) != tensor_map.end(); }
    inline bool isValid(const Tensor *tensor) const { return tensor != nullptr && tensor->size() > 0; }

    // as in the reference (tensor.h:237-243): (key, value) overwrites and does not validate; a pair keeps an existing entry
    inline void insert(const std::string &key, Tensor *value) { tensor_map[key] = value; }
    inline void insert(const std::pair<std::string, Tensor *> &kv) { tensor_map.insert(kv); }

    inline Tensor *at(const std::string &key) const {
        LLM_CHECK_WITH_INFO(isExist(key), fmtstr("Cannot find a tensor of name %s in the tensor map (keys: %s)", key.c_str(),
                                                 vec2str(keys()).c_str()));
        return tensor_map.at(key);
    }
    inline Tensor *operator[](const std::string &key) const { return at(key); }

    std::vector<std::string> keys() const {
        std::vector<std::string> out;
        for (const auto &kv : tensor_map) out.push_back(kv.first);
        return out;
    }

    std::string toString() const {
        std::string s = "{";
        const std::vector<std::string> names = keys();
        for (size_t i = 0; i < names.size(); ++i) s += names[i] + ": " + at(names[i])->toString() + (i + 1 < names.size() ? ", " : "");
        return s + "}";
    }
};
