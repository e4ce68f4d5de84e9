// Drop-in for the reference's src/utils/string_utils.h: fmtstr / vec2str / arr2str.
#pragma once
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

template <typename... Args> inline std::string fmtstr(const std::string &format, Args... args) {
    const int n = std::snprintf(nullptr, 0, format.c_str(), args...);
    if (n < 0) throw std::runtime_error("Error during formatting.");
    std::string out((size_t)n + 1, '\0');
    std::snprintf(&out[0], out.size(), format.c_str(), args...);
    out.resize((size_t)n);
    return out;
}
template <typename T> inline std::string arr2str(const T *const &arr, size_t size) {
    std::ostringstream ss;
    ss << "(";
    for (size_t i = 0; i < size; ++i) ss << (i ? ", " : "") << arr[i];
    ss << ")";
    return ss.str();
}
// The reference's vec2str (string_utils.h:24-36) writes every element followed by ", " and then the LAST element once more:
// {5, 7} prints as "(5, 7, 7)".  Kept, so that tensor descriptions and the key list of a failed TensorMap lookup read the same.
template <typename T> inline std::string vec2str(const std::vector<T> &vec) {
    std::ostringstream ss;
    ss << "(";
    for (size_t i = 0; i < vec.size(); ++i) ss << vec[i] << ", ";
    if (!vec.empty()) ss << vec.back();
    ss << ")";
    return ss.str();
}
