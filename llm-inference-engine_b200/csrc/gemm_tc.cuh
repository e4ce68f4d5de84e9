// gemm_tc.cuh -- tcgen05 / TMEM tensor-core GEMM for M > 8 tokens (prefill, batched decode).
#pragma once
#include "common.cuh"
namespace b200 {
// y[M,N] = x[M,K] * W^T with W packed [N,K], 16-bit dtype.  B200_ERR_UNSUPPORTED (no error text) if the
// shape / alignment cannot use the tensor-core kernel.
int launch_gemm_tc(const void *x, const void *w, void *y, int M, int N, int K, int dtype, cudaStream_t st);
// act[M,inter] = silu(x . Wgate^T) * (x . Wup^T), W packed [2*inter, K] (gate rows first): gate_up linear + SwiGLU in one kernel, M > 128.
int launch_gemm_tc_swiglu(const void *x, const void *w, void *act, int M, int inter, int K, int dtype, cudaStream_t st);
}
