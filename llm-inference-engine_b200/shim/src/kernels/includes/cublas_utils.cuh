// Drop-in for the reference's src/kernels/includes/cublas_utils.cuh:17-75.  The class is kept as a TYPE (layer constructors
// and examples pass it around, and examples create the cuBLAS handles themselves), but no cuBLAS call is made on the
// decoder-layer path: gemm() / stridedBatchedGemm() run libb200llm's own kernels for the operand layout the reference
// layers use (OP_N, OP_N, packed leading dimensions, alpha = 1, beta = 0) and throw for anything else.
#pragma once

#include <cublasLt.h>
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <map>
#include <string>

#include "b200llm.h"
#include "../../utils/macro.h"

class CublasWrapper {
private:
    cublasHandle_t cublas_handle_;
    cublasLtHandle_t cublaslt_handle_;
    int dtype_ = B200_F32;  // element type of A, B and C

public:
    CublasWrapper(cublasHandle_t cublasHandle, cublasLtHandle_t cublasLtHandle) : cublas_handle_(cublasHandle), cublaslt_handle_(cublasLtHandle) {}
    ~CublasWrapper() = default;

    void setFP32GemmConfig() { dtype_ = B200_F32; }
    void setFP16GemmConfig() { dtype_ = B200_F16; }
    void setBF16GemmConfig() { dtype_ = B200_BF16; }  // new
    int b200Dtype() const { return dtype_; }

    // column-major C[m,n] = A[m,k] * B[k,n]  ==  row-major Y[n,m] = X[n,k] * W[k,m] with X = B, W = A
    void gemm(cublasOperation_t transa, cublasOperation_t transb, int m, int n, int k, const void *A, int lda, const void *B, int ldb, void *C,
              int ldc, float alpha = 1.0f, float beta = 0.0f) {
        LLM_CHECK_WITH_INFO(transa == CUBLAS_OP_N && transb == CUBLAS_OP_N && lda == m && ldb == k && ldc == m && alpha == 1.0f && beta == 0.0f,
                            "CublasWrapper::gemm: only the reference layers' configuration (OP_N, OP_N, packed, alpha=1, beta=0) is provided");
        B200_CALL(b200_workspace_ensure(0));
        B200_CALL(b200_linear(B, A, nullptr, nullptr, C, n, k, m, dtype_, B200_W_DENSE, B200_LAYOUT_KN, 0, nullptr));
    }

    void stridedBatchedGemm(cublasOperation_t transa, cublasOperation_t transb, int m, int n, int k, const void *A, int lda, int64_t strideA,
                            const void *B, int ldb, int64_t strideB, void *C, int ldc, int64_t strideC, int batchCount, float alpha = 1.0f,
                            float beta = 0.0f) {
        LLM_CHECK_WITH_INFO(transa == CUBLAS_OP_N && transb == CUBLAS_OP_N && lda == m && ldb == k && ldc == m && alpha == 1.0f && beta == 0.0f &&
                                strideA == (int64_t)m * k && strideB == (int64_t)k * n && strideC == (int64_t)m * n,
                            "CublasWrapper::stridedBatchedGemm: only packed OP_N/OP_N batches are provided");
        B200_CALL(b200_batched_gemm(B, A, C, batchCount, n, m, k, 0, dtype_, nullptr));
    }
};
