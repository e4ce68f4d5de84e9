// Drop-in for the reference's src/weights/includes/attention_weights.h.
#pragma once
#include "base_weights.h"
template <typename T> struct LlamaAttentionWeights {
    BaseWeight<T> qkv;     // [h, (H + 2 Hkv) d] as [K,N]
    BaseWeight<T> output;  // [H d, h] as [K,N]; .bias = o-proj bias
};
