#include "gemv_inst.cuh"
namespace b200 {
int launch_gemv_nk_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_t<__half, true>(a, fmt, swiglu, st); }
}
