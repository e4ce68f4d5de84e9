// gemv.cuh -- the decode-shaped linear (M <= 4 tokens) for sm_100a: y[M,N] = x[M,K] * W^T, W packed [N,K].
//
// HBM-bound by construction: every weight byte is read exactly once per call.  Design:
//   * one persistent CTA per SM: 16 compute warps in 2 groups of 8, plus one TMA-producer warp and one reducer warp per
//     group (warp-specialised; the compute warps never wait on each other).  A GROUP owns work units of R = 2 weight rows,
//     dealt round-robin over SMs first; inside a unit the K dimension is split over the group's 8 warps, so a unit is
//     16 KB of traffic done by 8 warps (not by one) and the last wave of units costs < 1 us: every SM streams the same
//     number of bytes to within one unit.  Because a warp always works on the same K-slice, its slice of the
//     activations is held in REGISTERS (no shared-memory read or bf16 unpack of x in the hot loop);
//   * weight rows are contiguous in the [N,K] packing, so a unit (or an 8 KiB-per-row piece of it when rows are longer) is
//     fetched with ONE 1-D TMA bulk copy per row (cp.async.bulk.shared.global, up to 8 KiB each) into a group-shared ring
//     of stages with full/empty mbarriers: no registers are held by loads in flight and a single lane issues a whole
//     stage (SASS: UBLKCP + SYNCS);
//   * the ring is filled BEFORE griddepcontrol.wait, so under programmatic dependent launch the weight stream of kernel i+1
//     is already in flight while kernel i drains;
//   * the activation rows live in shared memory (optionally produced by the fused add-residual + bias + RMSNorm prologue,
//     reference src/kernels/add_residual_and_rmsnorm.cu:43-121 and src/kernels/rmsnorm.cu:35-80: one global read pass, the
//     row cached in registers across the block reduction), fp32 accumulation; the per-lane partial sums of a group's
//     warps meet in shared memory (double-buffered, ready/free mbarriers) where the reducer warp adds them in a fixed order
//     (deterministic), reduces across lanes and stores, with the optional SwiGLU epilogue (reference
//     src/kernels/silu_and_mul.cu:6-41) when the unit's two rows are (gate_i, up_i);
//   * FP8-e4m3 (per-row fp32 scale) and INT4 (grouped scale + zero point) weights are dequantised in registers; only their
//     packed bytes cross HBM.
#pragma once
#include "common.cuh"

namespace b200 {

enum { WF_DENSE = 0, WF_FP8 = 1, WF_INT4 = 2 };

constexpr int kGemvGroups = 2;            // groups of compute warps per CTA
constexpr int kGemvGW = 8;                // compute warps per group (the K split of a unit)
constexpr int kGemvWarps = kGemvGroups * kGemvGW;                      // compute warps
constexpr int kGemvThreads = (kGemvWarps + 2 * kGemvGroups) * 32;      // + one producer and one reducer warp per group
constexpr int kGemvMaxStages = 8;
constexpr int kGemvRows = 2;             // weight rows per work unit
constexpr int kGemvPieceBytes = 8192;    // bytes of one row per pipeline stage (one bulk copy)
constexpr int kGemvXCache = 2;           // 16-byte activation vectors cached per thread across the RMSNorm reduction

// per-launch geometry, computed on the host (gemv_inst.cuh)
struct GemvGeom {
    int stages;       // ring depth per group
    int piece_bytes;  // bytes of a row per stage (<= kGemvPieceBytes, multiple of 512 except for the only piece of a short row)
    int pieces;       // stages per unit = ceil(row_bytes / piece_bytes)
    int stage_bytes;  // kGemvRows * row stride inside a stage
    int cw;           // warp-vectors (32 lanes x 16 B) of a row piece per compute warp
    int groups;       // groups of 8 compute warps in a CTA: 2 (one CTA fills the SM) or 1 (half-size CTA: the next kernel's CTA can be
                      // co-resident and prefetch its weights while this one streams)
};

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
    TpExchange tp;  // world > 1: x is the sum over ranks of tp.peer_x[*] (fused one-shot all-reduce), `x` itself is ignored
    // push.n > 0: y is this rank's PARTIAL sum of a row-sharded linear; the reducer stores it into every rank's exchange buffer
    // (peer-mapped memory: posted NVLink writes of LL words, common.cuh) instead of `y`, so that the consumers read only local memory
    TpPush push;
};

// ------------------------------------------------------------------ mbarrier / bulk-copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// L2 policy for read-once streams (weights, KV tiles): evict-first, so that 13 GB of weights per step do not push the small
// re-used tensors (residual stream, activations, gammas, biases) out of L2 -- the prologues then hit L2 instead of queueing behind
// megabytes of outstanding weight requests in the DRAM queues.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(l2_evict_first_policy())
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

// ------------------------------------------------------------------ pre-norm value of one activation vector
// Vector i (V = 16 bytes of T) of token row m before the RMSNorm: x (or, under tensor parallelism, the fused all-reduce of every rank's
// partial -- TpExchange) (+ residual) -> T; residual_out <- that (CTA 0 only, when write_res); (+ bias) -> T.  The reference's kernels
// form these sums in T (add_residual_and_rmsnorm.cu:71-92): rounded at the same points.
template <typename T>
__device__ __forceinline__ void gemv_prenorm_vec(const GemvArgs &a, unsigned int tp_want, int m, int i, float *f, bool write_res) {
    constexpr int V = Elem<T>::kVec;
    const int K = a.K;
    const T *rin = a.norm ? reinterpret_cast<const T *>(a.res_in) : nullptr;
    T *rout = a.norm ? reinterpret_cast<T *>(a.res_out) : nullptr;
    const T *bias = a.norm ? reinterpret_cast<const T *>(a.bias) : nullptr;
    if (a.tp.world > 1) {
        // one-shot all-reduce: add every rank's partial in rank order, round to T as an all-reduced tensor of T would be
        tp_reduce_vec<T>(a.tp, tp_want, ((size_t)m * K + (size_t)i * V) * sizeof(T) / 4, f);
    } else {
        unpack16<T>(ld_v4(reinterpret_cast<const T *>(a.x) + (size_t)m * K + (size_t)i * V), f);
    }
    if (rin) {
        float r[V];
        unpack16<T>(ld_v4(rin + (size_t)m * K + (size_t)i * V), r);
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] += r[j];
        round_vec<T>(f);
    }
    if (write_res && rout && blockIdx.x == 0) st_v4(rout + (size_t)m * K + (size_t)i * V, pack16<T>(f));
    if (bias) {
        float b[V];
        unpack16<T>(ld_v4(bias + (size_t)i * V), b);
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] += b[j];
        round_vec<T>(f);
    }
}

// ------------------------------------------------------------------ activation staging shared by the GEMV kernels
// For every token row m < a.M and every 16-byte vector i of the row: x (or, under tensor parallelism, the fused all-reduce of every
// rank's partial -- TpExchange) (+ residual) -> T; residual_out <- that (CTA 0 only); (+ bias) -> T; RMSNorm with gamma when given
// (reference src/kernels/rmsnorm.cu:35-80, add_residual_and_rmsnorm.cu:43-121).  The V fp32 values of each vector are handed to
// store(m, i, f).  One global read pass when the row fits the per-thread register cache.  Every thread of the CTA must call;
// `red` = shared float[33].  The caller synchronises afterwards.
template <typename T, int MB, typename Store>
__device__ __forceinline__ void gemv_stage_activations(const GemvArgs &a, int n_threads, float *red, Store store) {
    constexpr int V = Elem<T>::kVec;
    const int K = a.K;
    const T *gamma = a.norm ? reinterpret_cast<const T *>(a.gamma) : nullptr;
    const int nv = K / V;
    const unsigned int tp_want = a.tp.world > 1 ? tp_flag(a.tp.epoch, a.tp.seq) : 0u;
    // pre-norm value of vector i of row m
    auto prenorm = [&](int m, int i, float *f, bool write_res) { gemv_prenorm_vec<T>(a, tp_want, m, i, f, write_res); };
    const bool cached = nv <= kGemvXCache * n_threads;  // the row fits the per-thread register cache: one global pass
    // MB == 1: a single trip known at compile time (keeps the B = 1 instantiation spill-free); otherwise a rolled run-time loop
#pragma unroll 1
    for (int m = 0; m < (MB == 1 ? 1 : a.M); ++m) {
        if (!gamma) {
            for (int i = threadIdx.x; i < nv; i += n_threads) {
                float f[V];
                prenorm(m, i, f, true);
                store(m, i, f);
            }
            continue;
        }
        float cache[kGemvXCache][V], gm[kGemvXCache][V];
        float ss = 0.0f;
        if (cached) {
#pragma unroll
            for (int c = 0; c < kGemvXCache; ++c) {
                const int i = threadIdx.x + c * n_threads;
                if (i < nv) {
                    unpack16<T>(ld_v4(gamma + (size_t)i * V), gm[c]);  // independent of the reduction: issue it now
                    prenorm(m, i, cache[c], true);
#pragma unroll
                    for (int j = 0; j < V; ++j) ss += cache[c][j] * cache[c][j];
                }
            }
        } else {
            for (int i = threadIdx.x; i < nv; i += n_threads) {
                float f[V];
                prenorm(m, i, f, true);
#pragma unroll
                for (int j = 0; j < V; ++j) ss += f[j] * f[j];
            }
        }
        ss = block_sum(ss, red);
        const float rs = rsqrtf(ss / (float)K + a.eps);
        if (cached) {
#pragma unroll
            for (int c = 0; c < kGemvXCache; ++c) {
                const int i = threadIdx.x + c * n_threads;
                if (i < nv) {
#pragma unroll
                    for (int j = 0; j < V; ++j) cache[c][j] = (cache[c][j] * gm[c][j]) * rs;
                    store(m, i, cache[c]);
                }
            }
        } else {
            // second pass over the (L1/L2-resident) inputs: recompute the pre-norm value and scale it
            for (int i = threadIdx.x; i < nv; i += n_threads) {
                float f[V], g[V];
                prenorm(m, i, f, false);
                unpack16<T>(ld_v4(gamma + (size_t)i * V), g);
#pragma unroll
                for (int j = 0; j < V; ++j) f[j] = (f[j] * g[j]) * rs;
                store(m, i, f);
            }
        }
    }
}

// ------------------------------------------------------------------ the kernel
// smem: [ xs : MB * Kp * sizeof(XS) | rings : groups * stages * stage_bytes | barriers : groups * (2 * kGemvMaxStages + 4) * 8 |
//         partial sums : groups * 2 * GW * R*MB * 32 floats ]
// XV > 0: every compute warp keeps its XV = pieces * cw activation vectors per token in registers (dense formats only).
template <typename T, int FMT, int MB, bool kSwiGLU, int XV>
__global__ void __launch_bounds__(kGemvThreads, 1)
gemv_nk_kernel(const GemvArgs a, const GemvGeom geo) {
    using WT = WTraits<T, FMT>;
    using XS = typename WT::XS;
    constexpr int EPV = WT::kEPV;
    constexpr int R = kGemvRows, GW = kGemvGW;
    constexpr int V = Elem<T>::kVec;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float red[33];

    const int K = a.K, N = a.N;
    const int Kp = (K + WT::kBlock - 1) / WT::kBlock * WT::kBlock;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row_bytes = FMT == WF_DENSE ? (size_t)K * sizeof(T) : (FMT == WF_FP8 ? (size_t)K : (size_t)K / 2);
    const int nvec_row = (int)(row_bytes / 16);
    const int units = kSwiGLU ? a.inter : (N + 1) / 2;
    const int ngroups_k = FMT == WF_INT4 ? K / a.group : 0;
    const int stages = geo.stages, pieces = geo.pieces, cw = geo.cw;
    const int piece_vecs = geo.piece_bytes / 16, row_stride = geo.stage_bytes / R;

    // ---- roles: warps [0, 16) compute (group = warp / 8), warp 16 + g produces for group g, warp 18 + g reduces for group g
    const int n_groups = geo.groups, n_compute = n_groups * GW, n_threads = (int)blockDim.x;
    const bool is_compute = warp < n_compute;
    const int grp = is_compute ? warp / GW : (warp - n_compute) % n_groups;
    const bool is_producer = !is_compute && warp < n_compute + n_groups;
    const int wg = warp % GW;
    // group `grp` of this CTA owns units gid, gid + total_groups, ...; a unit is `pieces` ring stages
    const int gid = grp * gridDim.x + blockIdx.x, total_groups = gridDim.x * n_groups;
    const int my_units = gid < units ? (units - gid + total_groups - 1) / total_groups : 0;
    const int my_items = my_units * pieces;

    XS *xs = reinterpret_cast<XS *>(smem);
    size_t off = ((size_t)MB * Kp * sizeof(XS) + 127) & ~(size_t)127;
    unsigned char *ring = smem + off + (size_t)grp * stages * geo.stage_bytes;
    off += (size_t)n_groups * stages * geo.stage_bytes;
    const uint32_t full0 = smem_u32(smem + off) + grp * (2 * kGemvMaxStages + 4) * 8, empty0 = full0 + kGemvMaxStages * 8;
    const uint32_t ready0 = empty0 + kGemvMaxStages * 8, free0 = ready0 + 16;
    off += (size_t)n_groups * (2 * kGemvMaxStages + 4) * 8;
    float *gred = reinterpret_cast<float *>(smem + off) + (size_t)grp * 2 * GW * (R * MB) * 32;  // [parity][warp][R*MB][lane]
    const uint32_t ring_u32 = smem_u32(ring);

    auto unit_row = [&](int u, int r) -> int { return kSwiGLU ? u + r * a.inter : 2 * u + r; };

    // ---- producer state machine (one lane): arm the stage's full barrier and issue one bulk copy per row
    int p_un = 0, p_pc = 0, p_item = 0, p_s = 0, p_ph = 0;
    auto issue_next = [&]() {
        const int u = gid + p_un * total_groups;
        const int v0 = p_pc * piece_vecs;
        const uint32_t bytes = (uint32_t)min(piece_vecs, nvec_row - v0) * 16u;
        const uint32_t bar = full0 + p_s * 8;
        int nrows = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) nrows += unit_row(u, r) < N ? 1 : 0;
        mbar_expect_tx(bar, bytes * nrows);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = unit_row(u, r);
            if (row < N)
                bulk_g2s(ring_u32 + p_s * geo.stage_bytes + r * row_stride,
                         reinterpret_cast<const unsigned char *>(a.w) + (size_t)row * row_bytes + (size_t)v0 * 16, bytes, bar);
        }
        ++p_item;
        if (++p_pc == pieces) p_pc = 0, ++p_un;
        if (++p_s == stages) p_s = 0, p_ph ^= 1;
    };
    if (is_producer && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + s * 8, 1);
            mbar_init(empty0 + s * 8, GW);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(ready0 + b * 8, GW);
            mbar_init(free0 + b * 8, 1);
        }
        fence_mbar_init();
        // fill the ring before waiting on the previous kernel: the weights do not depend on it
        while (p_item < stages && p_item < my_items) issue_next();
    }

    pdl_wait();

    // ---------------- stage the activations (with the fused add-residual / bias / RMSNorm prologue); all warps help
    {
        if constexpr (FMT != WF_DENSE) {
            for (int i = K + threadIdx.x; i < Kp; i += n_threads)
                for (int m = 0; m < MB; ++m) xs[(size_t)m * Kp + xs_perm<EPV>(i)] = 0.0f;
        }
        for (int m = a.M; m < MB; ++m)  // padding rows of the batch tile
            for (int i = threadIdx.x; i < Kp; i += n_threads) xs[(size_t)m * Kp + i] = XS(0.0f);
        gemv_stage_activations<T, MB>(a, n_threads, red, [&](int m, int i, const float *f) {
            if constexpr (FMT == WF_DENSE) {
                *reinterpret_cast<uint4 *>(xs + (size_t)m * Kp + (size_t)i * V) = pack16<T>(f);
            } else {
                float g[V];
                unpack16<T>(pack16<T>(f), g);  // the un-fused reference hands the GEMM a tensor of T
#pragma unroll
                for (int j = 0; j < V; ++j) xs[(size_t)m * Kp + xs_perm<EPV>(i * V + j)] = g[j];
            }
        });
        __syncthreads();  // xs complete; also publishes the producers' mbarrier initialisation
    }
    pdl_launch_dependents();

    if (is_producer) {
        // ================================================= TMA producer: refill a stage once all 8 warps have left it
        if (lane == 0) {
            int e_s = 0, e_ph = 0;  // stage / phase of the empty barrier to wait on next
            while (p_item < my_items) {
                mbar_wait(empty0 + e_s * 8, e_ph);
                if (++e_s == stages) e_s = 0, e_ph ^= 1;
                fence_proxy_async();
                issue_next();
            }
        }
    } else if (!is_compute) {
        // ================================================= reducer: add the group's per-lane partials in a fixed order
        const unsigned int push_flag = a.push.n > 0 ? tp_flag(a.push.epoch, a.push.seq) : 0u;
        for (int un = 0; un < my_units; ++un) {
            const int u = gid + un * total_groups;
            const int b = un & 1;
            mbar_wait(ready0 + b * 8, (un >> 1) & 1);
            const float *slot = gred + (size_t)b * GW * (R * MB) * 32;
            float out[R][MB];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    float t = 0.0f;
#pragma unroll
                    for (int w2 = 0; w2 < GW; ++w2) t += slot[(w2 * (R * MB) + r * MB + m) * 32 + lane];
                    out[r][m] = warp_sum(t);
                }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(free0 + b * 8);  // the slot may be overwritten
                bool valid[R];
#pragma unroll
                for (int r = 0; r < R; ++r) valid[r] = unit_row(u, r) < N;
                if constexpr (FMT == WF_FP8) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
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
                            const float g = round_to<T>(out[0][m]), up = round_to<T>(out[1][m]);
                            const float v = (g / (1.0f + expf(-g))) * up;
                            if (a.y_f32) reinterpret_cast<float *>(a.y)[(size_t)m * a.inter + u] = v;
                            else reinterpret_cast<T *>(a.y)[(size_t)m * a.inter + u] = Elem<T>::from_f(v);
                        }
                } else if (a.push.n > 0) {
                    // rows (2u, 2u + 1) of every token: one LL word (16-bit T) or two (fp32) per token and peer; N is even here
#pragma unroll
                    for (int m = 0; m < MB; ++m)
                        if (m < a.M) tp_push_pair<T>(a.push, push_flag, (size_t)m * N + 2 * u, out[0][m], out[1][m]);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (valid[r]) {
                            const int row = unit_row(u, r);
#pragma unroll
                            for (int m = 0; m < MB; ++m)
                                if (m < a.M) {
                                    if (a.y_f32) {
                                        reinterpret_cast<float *>(a.y)[(size_t)m * N + row] = out[r][m];
                                    } else {
                                        reinterpret_cast<T *>(a.y)[(size_t)m * N + row] = Elem<T>::from_f(out[r][m]);
                                    }
                                }
                        }
                }
            }
        }
    } else {
        // ================================================= compute warps
        // this warp's slice of the activations, in registers (XV > 0): vector (pc, j) covers k = ((pc*pv32 + wg*cw + j)*32 + lane)*V
        float xr[XV > 0 ? XV : 1][MB][V];
        if constexpr (XV > 0) {
#pragma unroll
            for (int i = 0; i < XV; ++i) {
                const int pc = i / 2, j = i % 2;  // XV path: cw <= 2, pieces <= XV / 2
                const int v = pc * piece_vecs + (wg * cw + j) * 32 + lane;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    if (j < cw && pc < pieces && v < nvec_row) {
                        unpack16<T>(*reinterpret_cast<const uint4 *>(xs + (size_t)m * Kp + (size_t)v * V), xr[i][m]);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; ++e) xr[i][m][e] = 0.0f;
                    }
                }
            }
        }
        float acc[R][MB][2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int m = 0; m < MB; ++m) acc[r][m][0] = acc[r][m][1] = 0.0f;

        int s = 0, ph = 0;
        const unsigned char *my_ring = ring + (size_t)(wg * cw * 32 + lane) * 16;
        for (int un = 0; un < my_units; ++un) {
            const int u = gid + un * total_groups;
            if constexpr (XV > 0) {
                // ---- hot path: fully unrolled over (piece, vector), x in registers, no predication inside the FMAs
#pragma unroll
                for (int pc = 0; pc < XV / 2; ++pc) {
                    if (pc < pieces) {
                        mbar_wait(full0 + s * 8, ph);
                        const unsigned char *st = my_ring + (size_t)s * geo.stage_bytes;
                        const int pv = min(piece_vecs, nvec_row - pc * piece_vecs);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j < cw && (wg * cw + j) * 32 + lane < pv) {
                                uint4 wv[R];
#pragma unroll
                                for (int r = 0; r < R; ++r) wv[r] = *reinterpret_cast<const uint4 *>(st + r * row_stride + j * 512);
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    float wf[V];
                                    unpack16<T>(wv[r], wf);
#pragma unroll
                                    for (int m = 0; m < MB; ++m)
#pragma unroll
                                        for (int e = 0; e < V; ++e) acc[r][m][j] = fmaf(wf[e], xr[pc * 2 + j][m][e], acc[r][m][j]);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty0 + s * 8);  // this warp has left the stage
                        if (++s == stages) s = 0, ph ^= 1;
                    }
                }
            } else {
                // ---- generic path: any K / format, x from shared memory
                const bool all_valid[R] = {true, true};
                for (int pc = 0; pc < pieces; ++pc) {
                    mbar_wait(full0 + s * 8, ph);
                    const unsigned char *st = ring + (size_t)s * geo.stage_bytes;
                    const int pv = min(piece_vecs, nvec_row - pc * piece_vecs);  // 16-byte vectors per row in this piece
                    const int wvb = pc * (piece_vecs / 32);                      // absolute warp-vector index of the piece start
                    for (int j = 0; j < cw; ++j) {
                        const int wv_i = wg * cw + j;
                        const int v = wv_i * 32 + lane;
                        if (v < pv) {
                            uint4 wv[R];
                            uint32_t zpk[R];
                            float sc[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                wv[r] = *reinterpret_cast<const uint4 *>(st + r * row_stride + v * 16);
                                zpk[r] = 0;
                                sc[r] = 0.0f;
                                if constexpr (FMT == WF_INT4) {
                                    const int row = min(unit_row(u, r), N - 1);
                                    // a 32-element vector never straddles a group (group % 32 == 0)
                                    const size_t gi = (size_t)row * ngroups_k + ((size_t)(pc * piece_vecs + v) * EPV) / a.group;
                                    sc[r] = Elem<T>::to_f(__ldg(reinterpret_cast<const T *>(a.scales) + gi));
                                    const uint32_t z = 0x6400u | __ldg(reinterpret_cast<const uint8_t *>(a.zeros) + gi);
                                    zpk[r] = z | (z << 16);
                                }
                            }
                            float acc1[R][MB];
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int m = 0; m < MB; ++m) acc1[r][m] = (j & 1) ? acc[r][m][1] : acc[r][m][0];
                            dot_rows<T, FMT, MB>(wv, all_valid, xs, Kp, wvb + wv_i, lane, zpk, sc, acc1);
                            // same two-accumulator order as the register path: a token's result does not depend on
                            // which instantiation (batch size) computed it
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int m = 0; m < MB; ++m) {
                                    if (j & 1) acc[r][m][1] = acc1[r][m];
                                    else acc[r][m][0] = acc1[r][m];
                                }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + s * 8);
                    if (++s == stages) s = 0, ph ^= 1;
                }
            }
            // ---- hand the per-lane partial sums to the reducer (double-buffered slot)
            const int b = un & 1;
            if (un >= 2) mbar_wait(free0 + b * 8, ((un >> 1) - 1) & 1);
            float *slot = gred + (size_t)b * GW * (R * MB) * 32;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool ok = unit_row(u, r) < N;  // rows past N were never copied: discard whatever the stage held
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    slot[(wg * (R * MB) + r * MB + m) * 32 + lane] = ok ? acc[r][m][0] + acc[r][m][1] : 0.0f;
                    acc[r][m][0] = acc[r][m][1] = 0.0f;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(ready0 + b * 8);
        }
    }
}

// Host side (linear.cu).  launch_gemv_nk returns B200_ERR_UNSUPPORTED (without setting the error text) when the
// shape cannot use this kernel; the caller then falls back to the generic tiled kernel.
int launch_gemv_nk(const GemvArgs &a, int dtype, int fmt, bool swiglu, cudaStream_t st);

}  // namespace b200
