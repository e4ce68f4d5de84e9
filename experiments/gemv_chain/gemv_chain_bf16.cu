#include "gemv_chain_inst.cuh"
namespace b200 {
int launch_gemv_chain_bf16(ChainArgs &a, cudaStream_t st, bool dry) { return launch_gemv_chain_t<__nv_bfloat16>(a, st, dry); }
}
