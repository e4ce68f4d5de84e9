// Drop-in for the reference's src/utils/params.h.
#pragma once
#include <string>
#include <unordered_map>
using MapStringToInt = std::unordered_map<std::string, int>;
using MapStringToFloat = std::unordered_map<std::string, float>;
