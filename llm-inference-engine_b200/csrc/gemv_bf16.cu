#include "gemv_inst.cuh"
namespace b200 {
int launch_gemv_nk_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_t<__nv_bfloat16, true>(a, fmt, swiglu, st); }
}
