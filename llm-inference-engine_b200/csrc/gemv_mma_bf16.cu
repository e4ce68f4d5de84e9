#include "gemv_mma.cuh"
namespace b200 {
int launch_gemv_mma_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_mma_t<__nv_bfloat16>(a, fmt, swiglu, st); }
}
