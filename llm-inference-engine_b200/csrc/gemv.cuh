// gemv.cuh -- the decode-shaped linear (M <= 4 tokens) for sm_100a: y[M,N] = x[M,K] * W^T, W packed [N,K].
//
// HBM-bound by construction: every weight byte is read exactly once per call.  Design:
//   * persistent grid (a multiple of the SM count); each WARP owns work units of R = 2 weight rows,
//     dealt round-robin over CTAs first so that every SM streams the same number of bytes (+-1 unit);
//   * weight rows are contiguous in the [N,K] packing, so they are streamed with 1-D TMA bulk copies
//     (cp.async.bulk.shared.global, 2 KiB per row chunk) into a PER-WARP shared-memory ring of kGemvStages
//     stages, completion tracked by one mbarrier per stage: no registers are held by loads in flight and
//     a single lane issues a whole stage (SASS: UBLKCP + SYNCS);
//   * the ring is filled BEFORE griddepcontrol.wait, so under programmatic dependent launch the weight
//     stream of kernel i+1 is already in flight while kernel i drains;
//   * the activation rows live in shared memory (optionally produced by the fused
//     add-residual + bias + RMSNorm prologue, reference src/kernels/add_residual_and_rmsnorm.cu:43-121
//     and src/kernels/rmsnorm.cu:35-80), fp32 accumulation, warp-shuffle reduction, optional SwiGLU
//     epilogue (reference src/kernels/silu_and_mul.cu:6-41) when the unit's two rows are (gate_i, up_i);
//   * FP8-e4m3 (per-row fp32 scale) and INT4 (grouped scale + zero point) weights are dequantised in
//     registers; only their packed bytes cross HBM.
#pragma once
#include "common.cuh"

namespace b200 {

enum { WF_DENSE = 0, WF_FP8 = 1, WF_INT4 = 2 };

constexpr int kGemvWarps = 8;
constexpr int kGemvThreads = kGemvWarps * 32;
constexpr int kGemvStages = 3;
constexpr int kGemvRows = 2;           // weight rows per work unit
constexpr int kGemvChunkBytes = 2048;  // bytes of one row per pipeline stage (128 x 16 B)
constexpr int kGemvStageBytes = kGemvRows * kGemvChunkBytes;

struct GemvArgs {
    const void *w;       // [N, row_bytes]
    const void *scales;  // FP8: float[N]; INT4: T[N, K/group]
    const void *zeros;   // INT4: uint8[N, K/group]
    const void *x;       // [M, K] of T
    void *y;             // [M, n_out] of T (or float when y_f32)
    // prologue (norm != 0): o = x (+ res_in); res_out <- o; o += bias; xs = gamma * o * rsqrt(mean(o^2) + eps)
    // (gamma == NULL: xs = o).  res_out must not alias res_in or x: other CTAs are still reading them.
    const void *res_in;
    void *res_out;
    const void *bias;
    const void *gamma;
    float eps;
    int norm;
    int M, K, N;
    int group;
    int inter;  // SwiGLU: rows (i, inter + i) form a unit, n_out = inter
    int y_f32;
};

// ------------------------------------------------------------------ mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ------------------------------------------------------------------ weight formats
template <typename T, int FMT> struct WTraits;
template <typename T> struct WTraits<T, WF_DENSE> {
    static constexpr int kEPV = 16 / (int)sizeof(T);  // weights per 16-byte vector
    using XS = T;                                     // activation type staged in shared memory
    static constexpr int kBlock = 1;                  // natural layout
};
template <typename T> struct WTraits<T, WF_FP8> {
    static constexpr int kEPV = 16;
    using XS = float;
    static constexpr int kBlock = 32 * 16;  // xs is permuted inside blocks of 32 lanes x kEPV elements
};
template <typename T> struct WTraits<T, WF_INT4> {
    static constexpr int kEPV = 32;
    using XS = float;
    static constexpr int kBlock = 32 * 32;
};

// Quantised formats keep x as fp32 in shared memory, permuted so that the float4 read by lane l for
// sub-word q of its weight vector is at (q*32 + l)*4 inside the block: consecutive lanes read consecutive
// 16 bytes (conflict-free) although each lane's elements are kEPV apart in k.
template <int EPV> __device__ __forceinline__ int xs_perm(int k) {
    constexpr int B = 32 * EPV;
    const int blk = k / B, w = k % B;
    const int lane = w / EPV, q = (w % EPV) / 4, e = w % 4;
    return blk * B + (q * 32 + lane) * 4 + e;
}

__device__ __forceinline__ void load_x4(const float *p, float *f) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    f[0] = v.x, f[1] = v.y, f[2] = v.z, f[3] = v.w;
}

// acc[r][m] += sum_e w[r]_e * x[m][k0 + e] for one 16-byte weight vector per row.
// vb = index of the 32-lane vector block inside the row (k0 = (vb*32 + lane) * EPV).
// INT4: zpk[r] = f16x2 {1024 + z, 1024 + z}, sc[r] = group scale; result is scaled here.
template <typename T, int FMT, int MB>
__device__ __forceinline__ void dot_rows(const uint4 (&wv)[kGemvRows], const bool (&valid)[kGemvRows],
                                         const typename WTraits<T, FMT>::XS *xs, int Kp, int vb, int lane,
                                         const uint32_t (&zpk)[kGemvRows], const float (&sc)[kGemvRows],
                                         float (&acc)[kGemvRows][MB]) {
    constexpr int EPV = WTraits<T, FMT>::kEPV;
    if constexpr (FMT == WF_DENSE) {
        const int k0 = (vb * 32 + lane) * EPV;
        float wf[kGemvRows][EPV];
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) unpack16<T>(wv[r], wf[r]);
#pragma unroll
        for (int m = 0; m < MB; ++m) {
            float xf[EPV];
            unpack16<T>(*reinterpret_cast<const uint4 *>(xs + (size_t)m * Kp + k0), xf);
#pragma unroll
            for (int r = 0; r < kGemvRows; ++r)
                if (valid[r]) {
#pragma unroll
                    for (int e = 0; e < EPV; ++e) acc[r][m] = fmaf(wf[r][e], xf[e], acc[r][m]);
                }
        }
    } else if constexpr (FMT == WF_FP8) {
        const float *xb = xs + (size_t)vb * (32 * EPV) + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // word q holds elements 4q .. 4q+3
            float wf[kGemvRows][4];
#pragma unroll
            for (int r = 0; r < kGemvRows; ++r) {
                const uint32_t wd = q == 0 ? wv[r].x : (q == 1 ? wv[r].y : (q == 2 ? wv[r].z : wv[r].w));
                uint32_t h01, h23;
                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h01) : "h"((unsigned short)(wd & 0xffffu)));
                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h23) : "h"((unsigned short)(wd >> 16)));
                f16x2_to_f32(h01, wf[r][0], wf[r][1]);
                f16x2_to_f32(h23, wf[r][2], wf[r][3]);
            }
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                float xf[4];
                load_x4(xb + (size_t)m * Kp + q * 128, xf);
#pragma unroll
                for (int r = 0; r < kGemvRows; ++r)
                    if (valid[r]) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[r][m] = fmaf(wf[r][e], xf[e], acc[r][m]);
                    }
            }
        }
    } else {
        const float *xb = xs + (size_t)vb * (32 * EPV) + lane * 4;
        float part[kGemvRows][MB];
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r)
#pragma unroll
            for (int m = 0; m < MB; ++m) part[r][m] = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // word q holds elements 8q .. 8q+7, nibble i = element 8q + i
            float wf[kGemvRows][8];
#pragma unroll
            for (int r = 0; r < kGemvRows; ++r) {
                const uint32_t wd = q == 0 ? wv[r].x : (q == 1 ? wv[r].y : (q == 2 ? wv[r].z : wv[r].w));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // {1024 + n_i, 1024 + n_{i+4}} as f16x2 minus {1024 + z, 1024 + z}: exact (n - z)
                    const uint32_t h = ((wd >> (4 * i)) & 0x000f000fu) | 0x64006400u;
                    const __half2 hv = __hsub2(*reinterpret_cast<const __half2 *>(&h), *reinterpret_cast<const __half2 *>(&zpk[r]));
                    const float2 fv = __half22float2(hv);
                    wf[r][i] = fv.x;
                    wf[r][i + 4] = fv.y;
                }
            }
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                float xf[8];
                load_x4(xb + (size_t)m * Kp + (2 * q) * 128, xf);
                load_x4(xb + (size_t)m * Kp + (2 * q + 1) * 128, xf + 4);
#pragma unroll
                for (int r = 0; r < kGemvRows; ++r)
                    if (valid[r]) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) part[r][m] = fmaf(wf[r][e], xf[e], part[r][m]);
                    }
            }
        }
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r)
#pragma unroll
            for (int m = 0; m < MB; ++m) acc[r][m] = fmaf(sc[r], part[r][m], acc[r][m]);
    }
}

// ------------------------------------------------------------------ the kernel
// smem: [ xs : MB * Kp * sizeof(XS) | ring : warps * stages * stage_bytes | barriers : warps * stages * 8 ]
template <typename T, int FMT, int MB, bool kSwiGLU>
__global__ void __launch_bounds__(kGemvThreads)
gemv_nk_kernel(const GemvArgs a) {
    using WT = WTraits<T, FMT>;
    using XS = typename WT::XS;
    constexpr int EPV = WT::kEPV;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float red[33];

    const int K = a.K, N = a.N;
    const int Kp = (K + WT::kBlock - 1) / WT::kBlock * WT::kBlock;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row_bytes = FMT == WF_DENSE ? (size_t)K * sizeof(T) : (FMT == WF_FP8 ? (size_t)K : (size_t)K / 2);
    const int nvec_row = (int)(row_bytes / 16);
    const int chunks = (nvec_row + 127) / 128;  // pipeline items per unit
    const int units = kSwiGLU ? a.inter : (N + 1) / 2;
    const int ngroups = FMT == WF_INT4 ? K / a.group : 0;

    XS *xs = reinterpret_cast<XS *>(smem);
    size_t off = ((size_t)MB * Kp * sizeof(XS) + 127) & ~(size_t)127;
    unsigned char *ring = smem + off + (size_t)warp * kGemvStages * kGemvStageBytes;
    off += (size_t)kGemvWarps * kGemvStages * kGemvStageBytes;
    const uint32_t bars = smem_u32(smem + off) + warp * kGemvStages * 8;
    const uint32_t ring_u32 = smem_u32(ring);

    // unit u belongs to CTA u % grid, warp (u / grid) % warps: the remainder is spread over CTAs (= SMs) first
    const int gwarp = warp * gridDim.x + blockIdx.x;
    const int total_warps = gridDim.x * kGemvWarps;
    const int my_units = gwarp < units ? (units - gwarp + total_warps - 1) / total_warps : 0;
    const int my_items = my_units * chunks;

    auto unit_row = [&](int u, int r) -> int { return kSwiGLU ? u + r * a.inter : 2 * u + r; };
    const unsigned char *wbase = reinterpret_cast<const unsigned char *>(a.w);

    // producer: lane 0 arms the stage barrier and issues one bulk copy per row
    auto issue = [&](int item) {
        const int s = item % kGemvStages;
        const int u = gwarp + (item / chunks) * total_warps;
        const int c = item % chunks;
        const int vecs = min(128, nvec_row - c * 128);
        const uint32_t bytes = (uint32_t)vecs * 16u;
        const uint32_t bar = bars + s * 8;
        int nrows = 0;
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) nrows += unit_row(u, r) < N ? 1 : 0;
        mbar_expect_tx(bar, bytes * nrows);
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) {
            const int row = unit_row(u, r);
            if (row < N)
                bulk_g2s(ring_u32 + s * kGemvStageBytes + r * kGemvChunkBytes,
                         wbase + (size_t)row * row_bytes + (size_t)c * kGemvChunkBytes, bytes, bar);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < kGemvStages; ++s) mbar_init(bars + s * 8, 1);
        fence_mbar_init();
    }
    __syncwarp();
    // fill the ring before waiting on the previous kernel: the weights do not depend on it
    if (lane == 0) {
        for (int it = 0; it < kGemvStages && it < my_items; ++it) issue(it);
    }

    pdl_wait();

    // ---------------- stage the activations (with the fused add-residual / bias / RMSNorm prologue)
    {
        const T *xin = reinterpret_cast<const T *>(a.x);
        const T *rin = a.norm ? reinterpret_cast<const T *>(a.res_in) : nullptr;
        T *rout = a.norm ? reinterpret_cast<T *>(a.res_out) : nullptr;
        const T *bias = a.norm ? reinterpret_cast<const T *>(a.bias) : nullptr;
        const T *gamma = a.norm ? reinterpret_cast<const T *>(a.gamma) : nullptr;
        constexpr int V = Elem<T>::kVec;
        const int nv = K / V;
        auto prenorm = [&](int m, int i, float *f) {
            unpack16<T>(ld_v4(xin + (size_t)m * K + (size_t)i * V), f);
            if (rin) {
                float r[V];
                unpack16<T>(ld_v4(rin + (size_t)m * K + (size_t)i * V), r);
#pragma unroll
                for (int j = 0; j < V; ++j) f[j] = round_to<T>(f[j] + r[j]);
            }
        };
        auto add_bias = [&](int i, float *f) {
            if (bias) {
                float b[V];
                unpack16<T>(ld_v4(bias + (size_t)i * V), b);
#pragma unroll
                for (int j = 0; j < V; ++j) f[j] = round_to<T>(f[j] + b[j]);
            }
        };
        auto store_xs = [&](int m, int i, const float *f) {
            if constexpr (FMT == WF_DENSE) {
                *reinterpret_cast<uint4 *>(xs + (size_t)m * Kp + (size_t)i * V) = pack16<T>(f);
            } else {
                float g[V];
                unpack16<T>(pack16<T>(f), g);  // the un-fused reference hands the GEMM a tensor of T
#pragma unroll
                for (int j = 0; j < V; ++j) xs[(size_t)m * Kp + xs_perm<EPV>(i * V + j)] = g[j];
            }
        };
        if constexpr (FMT != WF_DENSE) {
            for (int i = K + threadIdx.x; i < Kp; i += kGemvThreads)
                for (int m = 0; m < MB; ++m) xs[(size_t)m * Kp + xs_perm<EPV>(i)] = 0.0f;
        }
        for (int m = 0; m < MB; ++m) {
            if (m >= a.M) {  // padding rows of the batch tile
                for (int i = threadIdx.x; i < Kp; i += kGemvThreads) xs[(size_t)m * Kp + i] = XS(0.0f);
                continue;
            }
            float ss = 0.0f;
            for (int i = threadIdx.x; i < nv; i += kGemvThreads) {
                float f[V];
                prenorm(m, i, f);
                if (rout && blockIdx.x == 0) st_v4(rout + (size_t)m * K + (size_t)i * V, pack16<T>(f));
                add_bias(i, f);
                if (gamma) {
#pragma unroll
                    for (int j = 0; j < V; ++j) ss += f[j] * f[j];
                } else {
                    store_xs(m, i, f);
                }
            }
            if (gamma) {
                ss = block_sum(ss, red);
                const float rs = rsqrtf(ss / (float)K + a.eps);
                // second pass over the (L1/L2-resident) inputs: recompute the pre-norm value and scale it
                for (int i = threadIdx.x; i < nv; i += kGemvThreads) {
                    float f[V], g[V];
                    prenorm(m, i, f);
                    add_bias(i, f);
                    unpack16<T>(ld_v4(gamma + (size_t)i * V), g);
#pragma unroll
                    for (int j = 0; j < V; ++j) f[j] = (f[j] * g[j]) * rs;
                    store_xs(m, i, f);
                }
            }
        }
        __syncthreads();
    }
    pdl_launch_dependents();

    // ---------------- consume
    float acc[kGemvRows][MB];
#pragma unroll
    for (int r = 0; r < kGemvRows; ++r)
#pragma unroll
        for (int m = 0; m < MB; ++m) acc[r][m] = 0.0f;

    for (int item = 0; item < my_items; ++item) {
        const int s = item % kGemvStages;
        const int u = gwarp + (item / chunks) * total_warps;
        const int c = item % chunks;
        bool valid[kGemvRows];
#pragma unroll
        for (int r = 0; r < kGemvRows; ++r) valid[r] = unit_row(u, r) < N;
        mbar_wait(bars + s * 8, (item / kGemvStages) & 1);
        const unsigned char *st = ring + s * kGemvStageBytes;
        const int vecs = min(128, nvec_row - c * 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = j * 32 + lane;
            if (v < vecs) {
                uint4 wv[kGemvRows];
                uint32_t zpk[kGemvRows];
                float sc[kGemvRows];
#pragma unroll
                for (int r = 0; r < kGemvRows; ++r) {
                    wv[r] = *reinterpret_cast<const uint4 *>(st + r * kGemvChunkBytes + v * 16);
                    zpk[r] = 0;
                    sc[r] = 0.0f;
                    if constexpr (FMT == WF_INT4) {
                        if (valid[r]) {
                            // a 32-element vector never straddles a group (group % 32 == 0)
                            const size_t gi = (size_t)unit_row(u, r) * ngroups + ((c * 128 + v) * EPV) / a.group;
                            sc[r] = Elem<T>::to_f(__ldg(reinterpret_cast<const T *>(a.scales) + gi));
                            const uint32_t z = 0x6400u | __ldg(reinterpret_cast<const uint8_t *>(a.zeros) + gi);
                            zpk[r] = z | (z << 16);
                        }
                    }
                }
                dot_rows<T, FMT, MB>(wv, valid, xs, Kp, c * 4 + j, lane, zpk, sc, acc);
            }
        }
        __syncwarp();
        // this stage is free again: refill it with the item kGemvStages ahead
        if (item + kGemvStages < my_items && lane == 0) {
            fence_proxy_async();
            issue(item + kGemvStages);
        }
        if (c == chunks - 1) {
            // ---------------- unit epilogue
            float out[kGemvRows][MB];
#pragma unroll
            for (int r = 0; r < kGemvRows; ++r)
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    out[r][m] = warp_sum(acc[r][m]);
                    acc[r][m] = 0.0f;
                }
            if (lane == 0) {
                if constexpr (FMT == WF_FP8) {
#pragma unroll
                    for (int r = 0; r < kGemvRows; ++r)
                        if (valid[r]) {
                            const float s8 = __ldg(reinterpret_cast<const float *>(a.scales) + unit_row(u, r));
#pragma unroll
                            for (int m = 0; m < MB; ++m) out[r][m] *= s8;
                        }
                }
                if constexpr (kSwiGLU) {
#pragma unroll
                    for (int m = 0; m < MB; ++m)
                        if (m < a.M) {
                            // the un-fused reference stores gate/up in T before SiLU reads them
                            const float g = Elem<T>::to_f(Elem<T>::from_f(out[0][m]));
                            const float up = Elem<T>::to_f(Elem<T>::from_f(out[1][m]));
                            const float v = (g / (1.0f + expf(-g))) * up;
                            if (a.y_f32) reinterpret_cast<float *>(a.y)[(size_t)m * a.inter + u] = v;
                            else reinterpret_cast<T *>(a.y)[(size_t)m * a.inter + u] = Elem<T>::from_f(v);
                        }
                } else {
#pragma unroll
                    for (int r = 0; r < kGemvRows; ++r)
                        if (valid[r]) {
                            const int row = unit_row(u, r);
#pragma unroll
                            for (int m = 0; m < MB; ++m)
                                if (m < a.M) {
                                    if (a.y_f32) reinterpret_cast<float *>(a.y)[(size_t)m * N + row] = out[r][m];
                                    else reinterpret_cast<T *>(a.y)[(size_t)m * N + row] = Elem<T>::from_f(out[r][m]);
                                }
                        }
                }
            }
        }
    }
}

// Host side (linear.cu).  launch_gemv_nk returns B200_ERR_UNSUPPORTED (without setting the error text) when the
// shape cannot use this kernel; the caller then falls back to the generic tiled kernel.
int launch_gemv_nk(const GemvArgs &a, int dtype, int fmt, bool swiglu, cudaStream_t st);

}  // namespace b200
