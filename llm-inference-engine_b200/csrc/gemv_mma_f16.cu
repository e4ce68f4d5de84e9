#include "gemv_mma.cuh"
namespace b200 {
int launch_gemv_mma_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_mma_t<__half>(a, fmt, swiglu, st); }
}
