#include "gemm_tc.cuh"
namespace b200 {
int launch_gemm_tc(const void *, const void *, void *, int, int, int, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
}
