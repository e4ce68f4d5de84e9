// Drop-in for the reference's src/utils/output_utils.h:9-33: print the rank and the dimensions of a tensor / weight (two lines on stdout,
// every dimension followed by a blank), the debugging aid its layers include.
#pragma once

#include <iostream>
#include <vector>

#include "tensor.h"
#include "../weights/includes/base_weights.h"

namespace b200shim {
inline void print_dims(const std::vector<int> &shape) {
    std::cout << "number of dimensions: " << shape.size() << std::endl;
    for (int d : shape) std::cout << d << " ";
    std::cout << std::endl;
}
}  // namespace b200shim

inline void print_tensor(const Tensor *tensor) { b200shim::print_dims(tensor->shape); }
template <typename T> inline void print_tensor(const TensorWrapper<T> *tensor) { b200shim::print_dims(tensor->shape); }
template <typename T> inline void print_weight(const BaseWeight<T> *weight) { b200shim::print_dims(weight->shape); }
