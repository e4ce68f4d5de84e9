// gemv_inst.cuh -- host-side launch of gemv_nk_kernel for one activation type (included by gemv_<type>.cu so that
// the three element types compile in parallel).
#pragma once
#include "gemv.cuh"

namespace b200 {

template <typename T, int FMT> static size_t gemv_smem_bytes(int MB, int K) {
    using WT = WTraits<T, FMT>;
    const int Kp = (K + WT::kBlock - 1) / WT::kBlock * WT::kBlock;
    size_t b = ((size_t)MB * Kp * sizeof(typename WT::XS) + 127) & ~(size_t)127;
    b += (size_t)kGemvWarps * kGemvStages * kGemvStageBytes;
    b += (size_t)kGemvWarps * kGemvStages * 8;
    return b;
}

template <typename T, int FMT, int MB, bool SW>
static int launch_gemv_inst(const GemvArgs &a, cudaStream_t st) {
    const size_t smem = gemv_smem_bytes<T, FMT>(MB, a.K);
    if (smem > 227 * 1024) return B200_ERR_UNSUPPORTED;
    auto kern = gemv_nk_kernel<T, FMT, MB, SW>;
    // per-device, per-instantiation setup: opt in to large dynamic smem, find CTAs/SM for this footprint
    static thread_local size_t cached_smem[64] = {0};
    static thread_local int cached_occ[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] != smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return cuda_status("gemv cudaFuncSetAttribute");
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGemvThreads, smem) != cudaSuccess || occ < 1) occ = 1;
        cached_occ[dev] = occ > 2 ? 2 : occ;
        cached_smem[dev] = smem;
    }
    const int units = SW ? a.inter : (a.N + 1) / 2;
    int grid = sm_count() * cached_occ[dev];
    const int need = (units + kGemvWarps - 1) / kGemvWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    launch_pdl(kern, dim3(grid), dim3(kGemvThreads), smem, st, true, a);
    return cuda_status("gemv_nk launch");
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
