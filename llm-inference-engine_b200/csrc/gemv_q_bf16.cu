#include "gemv_q.cuh"
namespace b200 {
int launch_gemv_q_bf16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_q_t<__nv_bfloat16>(a, fmt, swiglu, st); }
}
