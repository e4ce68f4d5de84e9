// Drop-in for the reference's src/weights/includes/weight.h.
#pragma once
#include <string>
class Weight {
public:
    virtual ~Weight() = default;
    virtual void loadWeightsFromFile(const std::string &weight_path) = 0;
};
