#include "gemv_chain_inst.cuh"
namespace b200 {
int launch_gemv_chain_f32(ChainArgs &a, cudaStream_t st, bool dry) { return launch_gemv_chain_t<float>(a, st, dry); }
}
namespace b200 {
int launch_gemv_chain_bf16(ChainArgs &a, cudaStream_t st, bool dry);
int launch_gemv_chain_f16(ChainArgs &a, cudaStream_t st, bool dry);
int launch_gemv_chain(ChainArgs &a, int dtype, cudaStream_t st, bool dry) {
    switch (dtype) {
        case B200_F32: return launch_gemv_chain_f32(a, st, dry);
        case B200_F16: return launch_gemv_chain_f16(a, st, dry);
        case B200_BF16: return launch_gemv_chain_bf16(a, st, dry);
    }
    set_error("gemv_chain: unknown dtype %d", dtype);
    return B200_ERR_INVALID_ARG;
}
}
