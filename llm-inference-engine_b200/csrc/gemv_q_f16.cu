#include "gemv_q.cuh"
namespace b200 {
int launch_gemv_q_f16(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) { return launch_gemv_q_t<__half>(a, fmt, swiglu, st); }
}
