#include "gemv_inst.cuh"
namespace b200 {
int launch_gemv_nk_f32(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_t<float, false>(a, fmt, swiglu, st); }
int launch_gemv_nk_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_nk_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st);
int launch_gemv_nk(const GemvArgs &a, int dtype, int fmt, bool swiglu, cudaStream_t st) {
    if (swiglu && (a.N != 2 * a.inter)) {
        set_error("gemv: SwiGLU epilogue needs N == 2*inter");
        return B200_ERR_INVALID_ARG;
    }
    switch (dtype) {
        case B200_F32: return launch_gemv_nk_f32(a, fmt, swiglu, st);
        case B200_F16: return launch_gemv_nk_f16(a, fmt, swiglu, st);
        case B200_BF16: return launch_gemv_nk_bf16(a, fmt, swiglu, st);
    }
    set_error("gemv: unknown dtype %d", dtype);
    return B200_ERR_INVALID_ARG;
}
}
