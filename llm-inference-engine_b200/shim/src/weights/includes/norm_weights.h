// Drop-in for the reference's src/weights/includes/norm_weights.h.
#pragma once
template <typename T> struct LayerNormWeight {
    T *gamma = nullptr;
};
