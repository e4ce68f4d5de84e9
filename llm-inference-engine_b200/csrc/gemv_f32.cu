#include <stdlib.h>

#include "gemv_inst.cuh"
namespace b200 {
int launch_gemv_nk_f32(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_t<float, false>(a, fmt, swiglu, st); }
int launch_gemv_nk_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_nk_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_mma_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_mma_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_q_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_q_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_nk(const GemvArgs &a, int dtype, int fmt, bool swiglu, cudaStream_t st) {
    if (swiglu && (a.N != 2 * a.inter)) {
        set_error("gemv: SwiGLU epilogue needs N == 2*inter");
        return B200_ERR_INVALID_ARG;
    }
    const bool b16 = dtype == B200_BF16 || dtype == B200_F16;
    // 16-bit activations, 2..16 tokens, plain activations (no fused prologue): the tensor-core GEMV (gemv_mma.cuh), one pass over the
    // weights whatever the batch and the format
    // (quantised weights up to 4 tokens stay on gemv_q.cuh below, whose fused prologue saves the two norm launches per layer: measured on
    // the 7B step, INT4 / FP8: batch 2 3.01 / 2.57 ms there against 4.20 / 3.05 ms here, batch 8 5.05 / 4.64 against 4.95 / 3.47)
    if (b16 && a.M >= 2 && a.M <= 16 && (fmt == WF_DENSE || a.M > 4)) {
        const int rc = dtype == B200_BF16 ? launch_gemv_mma_bf16(a, fmt, swiglu, st) : launch_gemv_mma_f16(a, fmt, swiglu, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    // weight-only quantised, 16-bit activations, M <= 8 (up to 4 tokens, or shapes / fused prologues the kernel above does not take):
    // dequant into mma fragments with the SIMT GEMV's fused prologue (gemv_q.cuh)
    if (b16 && fmt != WF_DENSE && a.M <= 8) {
        const int rc = dtype == B200_BF16 ? launch_gemv_q_bf16(a, fmt, swiglu, st) : launch_gemv_q_f16(a, fmt, swiglu, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    if (a.M > 4) return B200_ERR_UNSUPPORTED;
    switch (dtype) {
        case B200_F32: return launch_gemv_nk_f32(a, fmt, swiglu, st);
        case B200_F16: return launch_gemv_nk_f16(a, fmt, swiglu, st);
        case B200_BF16: return launch_gemv_nk_bf16(a, fmt, swiglu, st);
    }
    set_error("gemv: unknown dtype %d", dtype);
    return B200_ERR_INVALID_ARG;
}
}
