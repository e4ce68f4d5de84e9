// linear.cu -- b200_linear / b200_batched_gemm and the weight packing helpers.
//
// Dispatch of y[M,N] = x[M,K] * W (reference src/kernels/linear.cu:10-158):
//   M <= 4, W packed [N,K] (dense / FP8 / INT4)  -> gemv_nk_kernel  (gemv.cuh; TMA-bulk weight streaming)
//   M <= 4, W dense  [K,N] (the reference's own memory order) -> gemv_kn_kernel (this file)
//   M  > 4, 16-bit activations, W packed [N,K]   -> tcgen05 / TMEM tensor-core GEMM (gemm_tc.cu)
//   everything else                              -> gemm_simt_kernel (tiled fp32-accumulate fallback, this file)
#include <stdlib.h>

#include "gemv.cuh"
#include "gemm_tc.cuh"

namespace b200 {

// =====================================================================================================
// [K,N]-layout GEMV: columns are contiguous, so a lane owns 16 bytes of consecutive columns and the K
// dimension is split over warps and CTAs; partial sums go through the library workspace and the last CTA of a
// column tile (ticket) adds them in a fixed order -> deterministic.
// =====================================================================================================
constexpr int kKnWarps = 8;
constexpr int kKnUnroll = 8;

template <typename T, int MB>
__global__ void __launch_bounds__(kKnWarps * 32)
gemv_kn_kernel(const T *__restrict__ w, const T *__restrict__ x, T *__restrict__ y, float *partial, unsigned int *tickets,
               int M, int K, int N, int ksplit) {
    constexpr int V = Elem<T>::kVec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *xs = reinterpret_cast<float *>(smem_raw);                 // [MB][kslice]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x, ks = blockIdx.y;
    const int kslice = (K + ksplit - 1) / ksplit;
    const int k_begin = ks * kslice, k_end = min(K, k_begin + kslice);
    const int n0 = (tile * 32 + lane) * V;
    float *redbuf = xs + (size_t)MB * kslice;                        // [warps][MB][32*V]
    __shared__ bool is_last;

    pdl_wait();
    for (int i = threadIdx.x; i < MB * kslice; i += blockDim.x) {
        const int m = i / kslice, k = k_begin + i % kslice;
        xs[i] = (m < M && k < k_end) ? Elem<T>::to_f(x[(size_t)m * K + k]) : 0.0f;
    }
    __syncthreads();
    pdl_launch_dependents();

    float acc[MB][V];
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int j = 0; j < V; ++j) acc[m][j] = 0.0f;

    if (n0 < N) {
        for (int kb = k_begin + warp * kKnUnroll; kb < k_end; kb += kKnWarps * kKnUnroll) {
            uint4 wv[kKnUnroll];
#pragma unroll
            for (int u = 0; u < kKnUnroll; ++u)
                wv[u] = (kb + u < k_end) ? ld_stream_v4(w + (size_t)(kb + u) * N + n0) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < kKnUnroll; ++u) {
                if (kb + u < k_end) {
                    float wf[V];
                    unpack16<T>(wv[u], wf);
#pragma unroll
                    for (int m = 0; m < MB; ++m) {
                        const float xv = xs[m * kslice + (kb + u - k_begin)];
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[m][j] = fmaf(xv, wf[j], acc[m][j]);
                    }
                }
            }
        }
    }
    // cross-warp reduction in a fixed order
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int j = 0; j < V; ++j) redbuf[((size_t)warp * MB + m) * 32 * V + lane * V + j] = acc[m][j];
    __syncthreads();
    const int tile_cols = 32 * V;
    for (int i = threadIdx.x; i < MB * tile_cols; i += blockDim.x) {
        const int m = i / tile_cols, cidx = i % tile_cols;
        float s = 0.0f;
#pragma unroll
        for (int wq = 0; wq < kKnWarps; ++wq) s += redbuf[((size_t)wq * MB + m) * tile_cols + cidx];
        const int n = tile * tile_cols + cidx;
        if (n < N && m < M) {
            if (ksplit == 1) y[(size_t)m * N + n] = Elem<T>::from_f(s);
            else partial[((size_t)ks * MB + m) * N + n] = s;
        }
    }
    if (ksplit == 1) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(&tickets[tile], ksplit - 1) == (unsigned)(ksplit - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < MB * tile_cols; i += blockDim.x) {
        const int m = i / tile_cols, n = tile * tile_cols + i % tile_cols;
        if (n < N && m < M) {
            float s = 0.0f;
            for (int q = 0; q < ksplit; ++q) s += __ldcg(&partial[((size_t)q * MB + m) * N + n]);
            y[(size_t)m * N + n] = Elem<T>::from_f(s);
        }
    }
}

// =====================================================================================================
// Generic tiled GEMM (fallback): C[b][M,N] = A[b][M,K] * B[b], fp32 accumulate, any shape / alignment.
// B element (k, n) comes from a loader functor so that all layouts and weight formats share one kernel.
// =====================================================================================================
template <typename T> struct DenseB {
    const T *p;
    long long sk, sn, sb;  // element strides of k, n and batch
    __device__ float operator()(int b, int k, int n) const { return Elem<T>::to_f(p[(long long)b * sb + (long long)k * sk + (long long)n * sn]); }
};
__device__ __forceinline__ float e4m3_to_f(uint8_t v) {
    uint32_t h;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h) : "h"((unsigned short)v));
    return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu)));
}
struct Fp8B {
    const uint8_t *p;
    const float *scales;
    int K;
    __device__ float operator()(int, int k, int n) const { return e4m3_to_f(p[(size_t)n * K + k]) * scales[n]; }
};
template <typename T> struct Int4B {
    const uint8_t *p;
    const T *scales;
    const uint8_t *zeros;
    int K, group;
    __device__ float operator()(int, int k, int n) const {
        const size_t idx = (size_t)n * K + k;
        const int q = (idx & 1) ? (p[idx >> 1] >> 4) : (p[idx >> 1] & 15);
        const size_t gi = (size_t)n * (K / group) + k / group;
        return (float)(q - (int)zeros[gi]) * Elem<T>::to_f(scales[gi]);
    }
};

constexpr int kTile = 64, kTileK = 16;
template <typename T, typename BL>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T *__restrict__ A, long long lda, long long strideA, BL bl, T *__restrict__ C, long long ldc,
                 long long strideC, int M, int N, int K) {
    __shared__ float As[kTileK][kTile + 4];
    __shared__ float Bs[kTileK][kTile + 4];
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const T *Ab = A + (long long)b * strideA;
    float acc[4][4] = {};
    pdl_wait();
    for (int k0 = 0; k0 < K; k0 += kTileK) {
        for (int i = threadIdx.x; i < kTile * kTileK; i += 256) {
            const int kk = i % kTileK, mm = i / kTileK;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? Elem<T>::to_f(Ab[(long long)m * lda + k]) : 0.0f;
        }
        for (int i = threadIdx.x; i < kTile * kTileK; i += 256) {
            const int kk = i % kTileK, nn = i / kTileK;
            const int n = n0 + nn, k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? bl(b, k, n) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTileK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    T *Cb = C + (long long)b * strideC;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) Cb[(long long)m * ldc + n] = Elem<T>::from_f(acc[i][j]);
        }
}

template <typename T, typename BL>
static int launch_simt(const T *A, long long lda, long long sA, BL bl, T *C, long long ldc, long long sC, int batch, int M,
                       int N, int K, cudaStream_t st) {
    dim3 grid((N + kTile - 1) / kTile, (M + kTile - 1) / kTile, batch);
    launch_pdl(gemm_simt_kernel<T, BL>, grid, dim3(256), 0, st, true, A, lda, sA, bl, C, ldc, sC, M, N, K);
    return cuda_status("gemm_simt launch");
}

// =====================================================================================================
// quantisers / dequantiser / transpose (load-time packing; not on the per-token path)
// =====================================================================================================
template <typename T>
__global__ void quantize_fp8_kernel(const T *__restrict__ src, uint8_t *__restrict__ q, float *__restrict__ scales, int N, int K) {
    __shared__ float red[33];
    const int n = blockIdx.x;
    float mx = 0.0f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, fabsf(Elem<T>::to_f(src[(size_t)n * K + k])));
    mx = block_max(mx, red);
    const float sc = mx > 0.0f ? mx / 448.0f : 1.0f;
    if (threadIdx.x == 0) scales[n] = sc;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = Elem<T>::to_f(src[(size_t)n * K + k]) / sc;
        unsigned short r;
        asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(0.0f), "f"(v));
        q[(size_t)n * K + k] = (uint8_t)(r & 0xff);
    }
}

// one warp per (row, group); requires group % 64 == 0
template <typename T>
__global__ void quantize_int4_kernel(const T *__restrict__ src, uint8_t *__restrict__ q, T *__restrict__ scales,
                                     uint8_t *__restrict__ zeros, int N, int K, int group) {
    const int G = K / group;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)N * G) return;
    const int n = (int)(gw / G), g = (int)(gw % G);
    const T *p = src + (size_t)n * K + (size_t)g * group;
    float mn = INFINITY, mx = -INFINITY;
    for (int k = lane; k < group; k += 32) {
        const float v = Elem<T>::to_f(p[k]);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    mn = -warp_max(-mn);
    float sc = (mx - mn) / 15.0f;
    if (!(sc > 0.0f)) sc = 1.0f;
    sc = Elem<T>::to_f(Elem<T>::from_f(sc));  // the stored scale is the one used
    float zf = rintf(-mn / sc);
    zf = fminf(fmaxf(zf, 0.0f), 15.0f);
    if (lane == 0) {
        scales[(size_t)n * G + g] = Elem<T>::from_f(sc);
        zeros[(size_t)n * G + g] = (uint8_t)zf;
    }
    for (int k = lane * 2; k < group; k += 64) {
        float q0 = fminf(fmaxf(rintf(Elem<T>::to_f(p[k]) / sc + zf), 0.0f), 15.0f);
        float q1 = fminf(fmaxf(rintf(Elem<T>::to_f(p[k + 1]) / sc + zf), 0.0f), 15.0f);
        q[((size_t)n * K + (size_t)g * group + k) / 2] = (uint8_t)((int)q0 | ((int)q1 << 4));
    }
}

template <typename T, typename BL>
__global__ void dequant_kernel(BL bl, T *__restrict__ dst, int N, int K) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)N * K; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = Elem<T>::from_f(bl(0, (int)(i % K), (int)(i / K)));
}

// Vectorised dequantiser for the prefill path (FP8 / INT4 -> T, 8 weights per thread and iteration): the quantised prefill linears run as
// "dequantise into a scratch tensor, then the tcgen05 GEMM" -- at M = 2048 the extra pass over the weights (packed bytes in, 2 bytes per
// weight out) is ~15 % of the GEMM's time, against a SIMT fallback that never touched the tensor cores.
template <typename T, int FMT>
__global__ void __launch_bounds__(256)
dequant_vec_kernel(const uint8_t *__restrict__ w, const void *__restrict__ scales, const uint8_t *__restrict__ zeros, T *__restrict__ dst, int N, int K,
                   int group) {
    const size_t nvec = (size_t)N * K / 8;
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e0 = i * 8;
        const int n = (int)(e0 / K), k = (int)(e0 % K);
        float f[8];
        if constexpr (FMT == B200_W_FP8E4M3) {
            const uint2 q = *reinterpret_cast<const uint2 *>(w + e0);
            const float sc = reinterpret_cast<const float *>(scales)[n];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = e4m3_to_f((uint8_t)(((j < 4 ? q.x : q.y) >> (8 * (j & 3))) & 0xffu)) * sc;
        } else {
            const uint32_t q = *reinterpret_cast<const uint32_t *>(w + e0 / 2);
            const size_t gi = (size_t)n * (K / group) + k / group;  // 8 consecutive k never straddle a group (group % 8 == 0)
            const float sc = Elem<T>::to_f(reinterpret_cast<const T *>(scales)[gi]);
            const int z = zeros[gi];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (float)((int)((q >> (4 * j)) & 15u) - z) * sc;
        }
        if constexpr (sizeof(T) == 2) {
            st_v4(dst + e0, pack16<T>(f));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[e0 + j] = Elem<T>::from_f(f[j]);
        }
    }
}

// dst[N,K] of T <- dequantised weights; B200_ERR_UNSUPPORTED when the shape cannot be vectorised
int launch_dequant_vec(const void *w, const void *scales, const void *zeros, void *dst, int N, int K, int w_format, int group, int dtype,
                       cudaStream_t st) {
    if (K % 8 != 0 || !aligned16(w) || !aligned16(dst) || (w_format == B200_W_INT4 && (group % 8 != 0 || K % group != 0))) return B200_ERR_UNSUPPORTED;
    const int grid = sm_count() * 8;
    B200_DISPATCH_DTYPE(dtype, {
        if (w_format == B200_W_FP8E4M3)
            launch_pdl(dequant_vec_kernel<T, B200_W_FP8E4M3>, dim3(grid), dim3(256), 0, st, true, (const uint8_t *)w, scales, (const uint8_t *)zeros, (T *)dst, N, K, group);
        else if (w_format == B200_W_INT4)
            launch_pdl(dequant_vec_kernel<T, B200_W_INT4>, dim3(grid), dim3(256), 0, st, true, (const uint8_t *)w, scales, (const uint8_t *)zeros, (T *)dst, N, K, group);
        else
            return B200_ERR_UNSUPPORTED;
    });
    return cuda_status("dequant_vec launch");
}

template <typename T>
__global__ void transpose_kernel(const T *__restrict__ src, T *__restrict__ dst, int rows, int cols) {
    __shared__ T tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

static int elem_bytes(int dtype) { return dtype == B200_F32 ? 4 : 2; }

}  // namespace b200

using namespace b200;

extern "C" {

int b200_linear(const void *x, const void *w, const void *scales, const void *zeros, void *y, int M, int K, int N,
                int dtype, int w_format, int w_layout, int group, b200_stream_t stream) {
    B200_REQUIRE(x && w && y, "linear: null pointer");
    B200_REQUIRE(M >= 0 && K > 0 && N > 0, "linear: bad shape M=%d K=%d N=%d", M, K, N);
    B200_REQUIRE(dtype == B200_F32 || dtype == B200_F16 || dtype == B200_BF16, "linear: unknown dtype %d", dtype);
    B200_REQUIRE(w_format >= B200_W_DENSE && w_format <= B200_W_INT4, "linear: unknown weight format %d", w_format);
    B200_REQUIRE(w_layout == B200_LAYOUT_KN || w_layout == B200_LAYOUT_NK, "linear: unknown weight layout %d", w_layout);
    if (w_format != B200_W_DENSE) {
        B200_REQUIRE(w_layout == B200_LAYOUT_NK, "linear: quantised weights must be packed [N,K]");
        B200_REQUIRE(dtype != B200_F32, "linear: quantised weights need 16-bit activations");
        B200_REQUIRE(scales, "linear: quantised weights need scales");
        if (w_format == B200_W_INT4) {
            B200_REQUIRE(zeros, "linear: INT4 weights need zero points");
            B200_REQUIRE(group >= 32 && group % 32 == 0 && K % group == 0, "linear: INT4 group %d must be a multiple of 32 dividing K=%d", group, K);
        }
    }
    if (M == 0) return B200_OK;
    cudaStream_t st = as_stream(stream);

    // ---- decode-shaped: weight-streaming GEMV
    if (w_layout == B200_LAYOUT_NK) {
        int done = 0, rc = B200_OK;
        const bool b16 = dtype != B200_F32;
        // weight-streaming GEMV in passes of `step` tokens (one pass = one read of the weights): 16-bit activations take 16 tokens per pass on
        // the tensor-core GEMV (gemv_mma.cuh: dense, FP8, INT4), fp32 4 on the SIMT GEMV
        auto gemv_passes = [&](int step) -> int {
            for (int m0 = 0; m0 < M; m0 += step) {
                GemvArgs a = {};
                a.w = w, a.scales = scales, a.zeros = zeros;
                a.x = (const char *)x + (size_t)m0 * K * elem_bytes(dtype);
                a.y = (char *)y + (size_t)m0 * N * elem_bytes(dtype);
                a.M = M - m0 < step ? M - m0 : step, a.K = K, a.N = N, a.group = group;
                const int r = launch_gemv_nk(a, dtype, w_format, false, st);
                if (r != B200_OK) return r;
            }
            return B200_OK;
        };
        // decode batches: one pass
        if (M <= (b16 ? 16 : 4)) {
            rc = gemv_passes(b16 ? 16 : 4);
            if (rc == B200_OK) done = 1;
            else if (rc != B200_ERR_UNSUPPORTED) return rc;
        }
        // dense 16-bit beyond that (and shapes the GEMV cannot take): tcgen05 GEMM (swap-AB + stream-K for M <= 128, 128x256 tiles above)
        if (!done && M > 4 && b16 && w_format == B200_W_DENSE) {
            rc = launch_gemm_tc(x, w, y, M, N, K, dtype, st);
            if (rc == B200_OK) done = 1;
            else if (rc != B200_ERR_UNSUPPORTED) return rc;
        }
        // quantised weights up to 128 tokens: passes of 16 (4 passes over packed INT4 move what one pass over bf16 moves); smaller steps
        // for the shapes only the older kernels take.  (Prefill-sized quantised linears inside the engine: dequantise + tcgen05 GEMM,
        // decoder.cu prefill_linear.)
        for (int step = b16 ? 16 : 4; !done && step >= 4 && M <= 128; step = step == 16 ? 8 : step - 4) {
            if (M > (step == 16 ? 8 : 4) * step) continue;
            rc = gemv_passes(step);
            if (rc == B200_OK) done = 1;
            else if (rc != B200_ERR_UNSUPPORTED) return rc;
        }
        if (done) return B200_OK;
    } else if (M <= 4 && N % (16 / elem_bytes(dtype)) == 0 && aligned16(w)) {
        Workspace ws;
        if (!get_workspace(&ws)) return B200_ERR_WORKSPACE;
        const int V = 16 / elem_bytes(dtype);
        const int tiles = (N + 32 * V - 1) / (32 * V);
        int ksplit = (2 * sm_count() + tiles - 1) / tiles;
        if (ksplit > (K + 63) / 64) ksplit = (K + 63) / 64;
        if (ksplit < 1) ksplit = 1;
        const int MB = M <= 1 ? 1 : (M <= 2 ? 2 : 4);
        const int kslice = (K + ksplit - 1) / ksplit;
        const size_t smem = (size_t)MB * kslice * 4 + (size_t)kKnWarps * MB * 32 * V * 4;
        const size_t need = (size_t)ksplit * MB * N * 4;
        if (smem <= 200 * 1024 && need <= ws.scratch_bytes && (size_t)tiles <= ws.n_tickets) {
            float *partial = reinterpret_cast<float *>(ws.scratch);
            B200_DISPATCH_DTYPE(dtype, {
                auto run = [&](auto mb) {
                    constexpr int kMB = decltype(mb)::value;
                    cudaFuncSetAttribute(gemv_kn_kernel<T, kMB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    launch_pdl(gemv_kn_kernel<T, kMB>, dim3(tiles, ksplit), dim3(kKnWarps * 32), smem, st, true, (const T *)w,
                               (const T *)x, (T *)y, partial, ws.tickets, M, K, N, ksplit);
                };
                if (MB == 1) run(std::integral_constant<int, 1>());
                else if (MB == 2) run(std::integral_constant<int, 2>());
                else run(std::integral_constant<int, 4>());
            });
            return cuda_status("gemv_kn launch");
        }
    }

    // ---- generic fallback
    B200_DISPATCH_DTYPE(dtype, {
        if (w_format == B200_W_DENSE) {
            DenseB<T> bl{(const T *)w, w_layout == B200_LAYOUT_KN ? (long long)N : 1, w_layout == B200_LAYOUT_KN ? 1 : (long long)K, 0};
            return launch_simt<T>((const T *)x, K, 0, bl, (T *)y, N, 0, 1, M, N, K, st);
        } else if (w_format == B200_W_FP8E4M3) {
            Fp8B bl{(const uint8_t *)w, (const float *)scales, K};
            return launch_simt<T>((const T *)x, K, 0, bl, (T *)y, N, 0, 1, M, N, K, st);
        } else {
            Int4B<T> bl{(const uint8_t *)w, (const T *)scales, (const uint8_t *)zeros, K, group};
            return launch_simt<T>((const T *)x, K, 0, bl, (T *)y, N, 0, 1, M, N, K, st);
        }
    });
    return B200_OK;
}

int b200_batched_gemm(const void *a, const void *b, void *c, int batch, int M, int N, int K, int trans_b, int dtype,
                      b200_stream_t stream) {
    B200_REQUIRE(a && b && c, "batched_gemm: null pointer");
    B200_REQUIRE(batch >= 0 && M >= 0 && N > 0 && K > 0, "batched_gemm: bad shape");
    if (batch == 0 || M == 0) return B200_OK;
    B200_REQUIRE(batch <= 65535, "batched_gemm: batch %d > 65535", batch);
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DTYPE(dtype, {
        DenseB<T> bl{(const T *)b, trans_b ? 1 : (long long)N, trans_b ? (long long)K : 1, (long long)N * K};
        return launch_simt<T>((const T *)a, K, (long long)M * K, bl, (T *)c, N, (long long)M * N, batch, M, N, K, st);
    });
    return B200_OK;
}

int b200_quantize_fp8(const void *src, void *w_out, float *scales_out, int N, int K, int dtype, b200_stream_t stream) {
    B200_REQUIRE(src && w_out && scales_out && N > 0 && K > 0, "quantize_fp8: bad argument");
    B200_DISPATCH_DTYPE(dtype, quantize_fp8_kernel<T><<<N, 256, 0, as_stream(stream)>>>((const T *)src, (uint8_t *)w_out, scales_out, N, K));
    return cuda_status("quantize_fp8 launch");
}

int b200_quantize_int4(const void *src, void *w_out, void *scales_out, void *zeros_out, int N, int K, int group, int dtype,
                       b200_stream_t stream) {
    B200_REQUIRE(src && w_out && scales_out && zeros_out && N > 0 && K > 0, "quantize_int4: bad argument");
    B200_REQUIRE(group >= 64 && group % 64 == 0 && K % group == 0, "quantize_int4: group %d must be a multiple of 64 dividing K", group);
    const long long warps = (long long)N * (K / group);
    const long long blocks = (warps * 32 + 255) / 256;
    B200_DISPATCH_DTYPE(dtype, quantize_int4_kernel<T><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
                                   (const T *)src, (uint8_t *)w_out, (T *)scales_out, (uint8_t *)zeros_out, N, K, group));
    return cuda_status("quantize_int4 launch");
}

int b200_dequantize(const void *w, const void *scales, const void *zeros, void *dst, int N, int K, int w_format, int group,
                    int dtype, b200_stream_t stream) {
    B200_REQUIRE(w && scales && dst && N > 0 && K > 0, "dequantize: bad argument");
    const int grid = sm_count() * 8;
    B200_DISPATCH_DTYPE(dtype, {
        if (w_format == B200_W_FP8E4M3) {
            Fp8B bl{(const uint8_t *)w, (const float *)scales, K};
            dequant_kernel<T, Fp8B><<<grid, 256, 0, as_stream(stream)>>>(bl, (T *)dst, N, K);
        } else if (w_format == B200_W_INT4) {
            B200_REQUIRE(zeros && group > 0 && K % group == 0, "dequantize: INT4 needs zeros and a group dividing K");
            Int4B<T> bl{(const uint8_t *)w, (const T *)scales, (const uint8_t *)zeros, K, group};
            dequant_kernel<T, Int4B<T>><<<grid, 256, 0, as_stream(stream)>>>(bl, (T *)dst, N, K);
        } else {
            set_error("dequantize: format %d is not quantised", w_format);
            return B200_ERR_INVALID_ARG;
        }
    });
    return cuda_status("dequantize launch");
}

int b200_transpose2d(const void *src, void *dst, int rows, int cols, int dtype, b200_stream_t stream) {
    B200_REQUIRE(src && dst && rows > 0 && cols > 0, "transpose2d: bad argument");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    B200_REQUIRE(grid.y <= 65535, "transpose2d: too many rows");
    B200_DISPATCH_DTYPE(dtype, transpose_kernel<T><<<grid, dim3(32, 8), 0, as_stream(stream)>>>((const T *)src, (T *)dst, rows, cols));
    return cuda_status("transpose2d launch");
}

int b200_linear_swiglu(const void *x, const void *w_gate_up, void *act, int M, int K, int inter_size, int dtype, b200_stream_t stream) {
    B200_REQUIRE(x && w_gate_up && act, "linear_swiglu: null pointer");
    B200_REQUIRE(M >= 0 && K > 0 && inter_size > 0, "linear_swiglu: bad shape M=%d K=%d inter=%d", M, K, inter_size);
    B200_REQUIRE(dtype == B200_F16 || dtype == B200_BF16, "linear_swiglu: 16-bit activations only (dtype %d)", dtype);
    if (M == 0) return B200_OK;
    const int rc = launch_gemm_tc_swiglu(x, w_gate_up, act, M, inter_size, K, dtype, as_stream(stream));
    if (rc == B200_ERR_UNSUPPORTED) set_error("linear_swiglu: shape M=%d K=%d inter=%d not served by the fused tensor-core kernel (M > 128, K %% 8 == 0, inter >= 128)", M, K, inter_size);
    return rc;
}

}  // extern "C"
