// Drop-in for the reference's src/weights/includes/embedding_weights.h.
#pragma once
#include "base_weights.h"
template <typename T> struct EmbeddingWeight : public BaseWeight<T> {};
