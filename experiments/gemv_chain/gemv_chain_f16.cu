#include "gemv_chain_inst.cuh"
namespace b200 {
int launch_gemv_chain_f16(ChainArgs &a, cudaStream_t st, bool dry) { return launch_gemv_chain_t<__half>(a, st, dry); }
}
