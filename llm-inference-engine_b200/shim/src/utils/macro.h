// Drop-in replacement for the reference's src/utils/macro.h (same macro names and error behaviour, macro.h:11-95):
//   CHECK(cuda call)               -> print + exit(1)
//   CHECK_CUBLAS(status)           -> throw std::runtime_error
//   DeviceSyncAndCheckCudaError()  -> cudaDeviceSynchronize + throw on a sticky error
//   LLM_CHECK / LLM_CHECK_WITH_INFO-> throw std::runtime_error("[oneLLM][ERROR] <info> Assertion fail: file:line")
// plus B200_CALL, which turns a libb200llm.so status code into the same exception type.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#include "b200llm.h"

namespace b200shim {
inline void cuda_or_exit(cudaError_t e, const char *file, int line) {
    if (e == cudaSuccess) return;
    std::fprintf(stderr, "CUDA Error:\n    File:       %s\n    Line:       %d\n    Error code: %d\n    Error text: %s\n", file, line, (int)e,
                 cudaGetErrorString(e));
    std::exit(1);
}
inline const char *status_text(cudaError_t e) { return cudaGetErrorString(e); }
inline const char *status_text(cublasStatus_t s) {
    static const char *names[] = {"CUBLAS_STATUS_SUCCESS", "CUBLAS_STATUS_NOT_INITIALIZED", "?", "CUBLAS_STATUS_ALLOC_FAILED"};
    switch (s) {
        case CUBLAS_STATUS_INVALID_VALUE: return "CUBLAS_STATUS_INVALID_VALUE";
        case CUBLAS_STATUS_ARCH_MISMATCH: return "CUBLAS_STATUS_ARCH_MISMATCH";
        case CUBLAS_STATUS_MAPPING_ERROR: return "CUBLAS_STATUS_MAPPING_ERROR";
        case CUBLAS_STATUS_EXECUTION_FAILED: return "CUBLAS_STATUS_EXECUTION_FAILED";
        case CUBLAS_STATUS_INTERNAL_ERROR: return "CUBLAS_STATUS_INTERNAL_ERROR";
        case CUBLAS_STATUS_NOT_SUPPORTED: return "CUBLAS_STATUS_NOT_SUPPORTED";
        case CUBLAS_STATUS_LICENSE_ERROR: return "CUBLAS_STATUS_LICENSE_ERROR";
        default: return (int)s >= 0 && (int)s <= 3 ? names[(int)s] : "<unknown>";
    }
}
[[noreturn]] inline void raise(const std::string &what, const char *file, int line) {
    throw std::runtime_error(what + " " + file + ":" + std::to_string(line) + " \n");
}
template <typename S> inline void throw_if_failed(S status, const char *file, int line) {
    if (status != 0) raise(std::string("[TM][ERROR] CUDA runtime error: ") + status_text(status), file, line);
}
inline void sync_and_throw(const char *file, int line) {
    cudaDeviceSynchronize();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) raise(std::string("[TM][ERROR] CUDA runtime error: ") + cudaGetErrorString(e), file, line);
}
inline void require(bool ok, const std::string &info, const char *file, int line) {
    if (!ok) raise("[oneLLM][ERROR] " + info + " Assertion fail:", file, line);
}
inline void b200_or_throw(int rc, const char *file, int line) {
    if (rc != B200_OK) raise(std::string("[oneLLM][ERROR] libb200llm: ") + b200_last_error_string() + " Assertion fail:", file, line);
}
}  // namespace b200shim

#define CHECK(call) ::b200shim::cuda_or_exit((call), __FILE__, __LINE__)
#define CHECK_CUBLAS(val) ::b200shim::throw_if_failed((val), __FILE__, __LINE__)
#define DeviceSyncAndCheckCudaError() ::b200shim::sync_and_throw(__FILE__, __LINE__)
#define LLM_CHECK(val) ::b200shim::require((val), "", __FILE__, __LINE__)
#define LLM_CHECK_WITH_INFO(val, info) ::b200shim::require((val), (info), __FILE__, __LINE__)
#define B200_CALL(rc) ::b200shim::b200_or_throw((rc), __FILE__, __LINE__)
