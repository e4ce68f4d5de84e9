// Drop-in for the reference's src/weights/includes/ffn_weights.h.
#pragma once
#include "base_weights.h"
template <typename T> struct LlamaFFNWeights {
    BaseWeight<T> gate;
    BaseWeight<T> up;
    BaseWeight<T> down;         // [I, h] as [K,N]
    BaseWeight<T> gate_and_up;  // [h, 2 I] as [K,N]: gate columns then up columns
};
