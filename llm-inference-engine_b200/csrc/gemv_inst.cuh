// gemv_inst.cuh -- host-side launch of gemv_nk_kernel for one activation type (included by gemv_<type>.cu so that
// the three element types compile in parallel).
#pragma once
#include <stdlib.h>

#include "gemv.cuh"

namespace b200 {

template <typename T, int FMT, int MB, bool SW, int XV>
static int launch_gemv_geom(const GemvArgs &a, const GemvGeom &g, size_t smem, cudaStream_t st) {
    auto kern = gemv_nk_kernel<T, FMT, MB, SW, XV>;
    static thread_local size_t cached_smem[64] = {0};  // per device, per instantiation: opt in to large dynamic smem once
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return cuda_status("gemv cudaFuncSetAttribute");
        cached_smem[dev] = smem;
    }
    const int units = SW ? a.inter : (a.N + 1) / 2;
    int grid = sm_count();
    const int need = (units + g.groups - 1) / g.groups;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    launch_pdl(kern, dim3(grid), dim3((g.groups * kGemvGW + 2 * g.groups) * 32), smem, st, true, a, g);
    return cuda_status("gemv_nk launch");
}

template <typename T, int FMT, int MB, bool SW>
static int launch_gemv_inst(const GemvArgs &a, cudaStream_t st) {
    using WT = WTraits<T, FMT>;
    const size_t row_bytes = FMT == WF_DENSE ? (size_t)a.K * sizeof(T) : (FMT == WF_FP8 ? (size_t)a.K : (size_t)a.K / 2);
    GemvGeom g;
    // a row is cut into equal pieces of <= 8 KiB (multiples of 512 B = one warp-vector), one bulk copy each
    g.pieces = (int)((row_bytes + kGemvPieceBytes - 1) / kGemvPieceBytes);
    g.piece_bytes = g.pieces == 1 ? (int)row_bytes : (int)(((row_bytes + g.pieces - 1) / g.pieces + 511) / 512 * 512);
    g.pieces = (int)((row_bytes + g.piece_bytes - 1) / g.piece_bytes);
    g.stage_bytes = kGemvRows * ((g.piece_bytes + 127) / 128 * 128);
    g.cw = ((g.piece_bytes / 16 + 31) / 32 + kGemvGW - 1) / kGemvGW;
    // One CTA per SM, two groups of 8 compute warps.  (Half-size CTAs -- one group, two CTAs per SM, so that the next kernel's CTA
    // could take a slot and fill its ring while the other half still streams -- were measured on the B200 at two CTAs per SM: 2.71 ms
    // per 7B step against 2.61 ms; for the O projection alone, co-resident with one attention CTA per SM: 2.66 against 2.63.)
    g.groups = kGemvGroups;
    // shared memory: activations + rings + barriers + per-lane partial sums
    const int Kp = (a.K + WT::kBlock - 1) / WT::kBlock * WT::kBlock;
    size_t fixed = ((size_t)MB * Kp * sizeof(typename WT::XS) + 127) & ~(size_t)127;
    fixed += (size_t)g.groups * (2 * kGemvMaxStages + 4) * 8;
    fixed += (size_t)g.groups * kGemvGW * 2 * kGemvRows * MB * 32 * sizeof(float);
    const size_t budget = 224 * 1024;
    const size_t per_stage = (size_t)g.groups * g.stage_bytes;
    if (fixed + 3 * per_stage > budget) return B200_ERR_UNSUPPORTED;
    g.stages = (int)((budget - fixed) / per_stage);
    if (g.stages > kGemvMaxStages) g.stages = kGemvMaxStages;
    const size_t smem = fixed + (size_t)g.stages * per_stage;
    // register-resident activations: XV = 2 * pieces vectors per token per warp, at most 6 vectors x tokens in total
    if constexpr (FMT == WF_DENSE) {
        if (g.cw <= 2) {
            if (g.pieces == 1 && MB <= 2) return launch_gemv_geom<T, FMT, MB, SW, 2>(a, g, smem, st);
            if constexpr (MB == 1) {
                if (g.pieces == 2) return launch_gemv_geom<T, FMT, MB, SW, 4>(a, g, smem, st);
                if (g.pieces == 3) return launch_gemv_geom<T, FMT, MB, SW, 6>(a, g, smem, st);
            }
        }
    }
    return launch_gemv_geom<T, FMT, MB, SW, 0>(a, g, smem, st);
}

template <typename T, int FMT, bool SW>
static int launch_gemv_mb(const GemvArgs &a, cudaStream_t st) {
    if (a.M <= 1) return launch_gemv_inst<T, FMT, 1, SW>(a, st);
    if (a.M <= 2) {
        const int rc = launch_gemv_inst<T, FMT, 2, SW>(a, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    if (a.M <= 4) return launch_gemv_inst<T, FMT, 4, SW>(a, st);
    return B200_ERR_UNSUPPORTED;
}

template <typename T, bool kQuant>
static int launch_gemv_t(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) {
    // shape requirements of the bulk-copy pipeline: 16-byte aligned rows and vectors
    const size_t row_bytes = fmt == WF_DENSE ? (size_t)a.K * sizeof(T) : (fmt == WF_FP8 ? (size_t)a.K : (size_t)a.K / 2);
    if (row_bytes % 16 != 0 || a.K % Elem<T>::kVec != 0 || !aligned16(a.w) || !aligned16(a.x)) return B200_ERR_UNSUPPORTED;
    if (a.norm && ((a.res_in && !aligned16(a.res_in)) || (a.res_out && !aligned16(a.res_out)) ||
                   (a.bias && !aligned16(a.bias)) || (a.gamma && !aligned16(a.gamma))))
        return B200_ERR_UNSUPPORTED;
    if (fmt == WF_INT4 && (a.group % 32 != 0 || a.K % a.group != 0)) return B200_ERR_UNSUPPORTED;
    if (fmt == WF_DENSE) return swiglu ? launch_gemv_mb<T, WF_DENSE, true>(a, st) : launch_gemv_mb<T, WF_DENSE, false>(a, st);
    if constexpr (kQuant) {
        if (fmt == WF_FP8) return swiglu ? launch_gemv_mb<T, WF_FP8, true>(a, st) : launch_gemv_mb<T, WF_FP8, false>(a, st);
        if (fmt == WF_INT4) return swiglu ? launch_gemv_mb<T, WF_INT4, true>(a, st) : launch_gemv_mb<T, WF_INT4, false>(a, st);
    }
    return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
