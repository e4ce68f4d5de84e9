// gemv_chain_inst.cuh -- host-side geometry + launch of gemv_chain_kernel for one activation type (included by gemv_chain_<type>.cu).
#pragma once
#include <stdlib.h>

#include "gemv_chain.cuh"

namespace b200 {

template <typename T, int MB>
static int launch_chain_mb(const ChainArgs &a, size_t smem, cudaStream_t st) {
    auto kern = gemv_chain_kernel<T, MB>;
    static thread_local size_t cached_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return cuda_status("gemv_chain cudaFuncSetAttribute");
        cached_smem[dev] = smem;
    }
    // one CTA per SM, all co-resident (the kernel's grid barrier needs that): never more CTAs than SMs
    launch_pdl(kern, dim3(sm_count()), dim3(kGemvThreads), smem, st, true, a);
    return cuda_status("gemv_chain launch");
}

template <typename T>
static int launch_gemv_chain_t(ChainArgs &a, cudaStream_t st, bool dry) {
    constexpr int V = Elem<T>::kVec;
    if (a.n_phases < 1 || a.n_phases > kChainMaxPhases || a.M < 1 || a.M > 4 || !a.claim) return B200_ERR_UNSUPPORTED;
    const int MB = a.M <= 1 ? 1 : (a.M <= 2 ? 2 : 4);
    int stage_bytes = 0, max_k = 0;
    for (int p = 0; p < a.n_phases; ++p) {
        ChainPhase &P = a.ph[p];
        const size_t row_bytes = (size_t)P.K * sizeof(T);
        if (P.K < V || row_bytes % 16 != 0 || P.K % V != 0 || !aligned16(P.w)) return B200_ERR_UNSUPPORTED;
        if (P.x_ll ? !aligned16(P.x_ll) : (!P.x || !aligned16(P.x))) return B200_ERR_UNSUPPORTED;
        if (P.y_ll ? !aligned16(P.y_ll) : !P.y) return B200_ERR_UNSUPPORTED;
        if (P.res_ll && (!P.res_in || !aligned16(P.res_ll))) return B200_ERR_UNSUPPORTED;
        // an LL output is read back in whole 16-byte vectors of T; a pair of 16-bit rows shares one LL word
        const int n_out = P.swiglu ? P.inter : P.N;
        if (P.y_ll && (n_out % V != 0 || (!P.swiglu && P.N % 2 != 0))) return B200_ERR_UNSUPPORTED;
        if (P.norm && ((P.res_in && !aligned16(P.res_in)) || (P.res_out && !aligned16(P.res_out)) || (P.bias && !aligned16(P.bias)) ||
                       (P.gamma && !aligned16(P.gamma))))
            return B200_ERR_UNSUPPORTED;
        if (P.swiglu && P.N != 2 * P.inter) return B200_ERR_UNSUPPORTED;
        if (P.N < 1) return B200_ERR_UNSUPPORTED;
        // a row is cut into equal pieces of <= 8 KiB (multiples of 512 B = one warp-vector), one bulk copy each (as gemv_inst.cuh)
        P.pieces = (int)((row_bytes + kGemvPieceBytes - 1) / kGemvPieceBytes);
        P.piece_bytes = P.pieces == 1 ? (int)row_bytes : (int)(((row_bytes + P.pieces - 1) / P.pieces + 511) / 512 * 512);
        P.pieces = (int)((row_bytes + P.piece_bytes - 1) / P.piece_bytes);
        P.cw = ((P.piece_bytes / 16 + 31) / 32 + kGemvGW - 1) / kGemvGW;
        const int sb = kGemvRows * ((P.piece_bytes + 127) / 128 * 128);
        if (sb > stage_bytes) stage_bytes = sb;
        if (P.K > max_k) max_k = P.K;
        // register-resident activations: XV = 2 * pieces vectors per token per warp, at most 6 vectors x tokens in total
        P.xv = 0;
        if (P.cw <= 2) {
            if (P.pieces == 1 && MB <= 2) P.xv = 2;
            else if (MB == 1 && P.pieces == 2) P.xv = 4;
            else if (MB == 1 && P.pieces == 3) P.xv = 6;
        }
    }
    a.stage_bytes = stage_bytes;
    a.xs_elems = (max_k + V - 1) / V * V;
    size_t fixed = ((size_t)MB * a.xs_elems * sizeof(T) + 127) & ~(size_t)127;
    fixed += (size_t)kGemvGroups * (2 * kGemvMaxStages + 4) * 8;
    fixed += 128;  // unit ids of the ring stages and partial-sum slots
    static_assert(kGemvGroups * (kGemvMaxStages + 3) * sizeof(int) <= 128, "unit-id area");
    fixed += (size_t)kGemvGroups * kGemvGW * 2 * kGemvRows * MB * 32 * sizeof(float);
    const size_t budget = 224 * 1024;
    const size_t per_stage = (size_t)kGemvGroups * stage_bytes;
    if (fixed + 3 * per_stage > budget) return B200_ERR_UNSUPPORTED;
    a.stages = (int)((budget - fixed) / per_stage);
    if (a.stages > kGemvMaxStages) a.stages = kGemvMaxStages;
    const size_t smem = fixed + (size_t)a.stages * per_stage;
    if (dry) return B200_OK;
    if (MB == 1) return launch_chain_mb<T, 1>(a, smem, st);
    if (MB == 2) return launch_chain_mb<T, 2>(a, smem, st);
    return launch_chain_mb<T, 4>(a, smem, st);
}

}  // namespace b200
