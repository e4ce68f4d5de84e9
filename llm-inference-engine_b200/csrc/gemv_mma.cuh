// gemv_mma.cuh -- the decode-shaped linear for 2..16 tokens on the tensor cores: y[M,N] = x[M,K] * W^T, W packed [N,K], dense 16-bit,
// FP8-e4m3 (W8A16) or INT4-g128 (W4A16) weights.  Reference: launchLinearGemm (src/kernels/linear.cu:10-87) at the decode shapes of
// src/layers/self_attention.cpp:79-86,131-138 and src/layers/ffn.cpp:105-139, with the quantisation the README plans (README.md:36-39).
//
// HBM-bound by construction: every weight byte crosses HBM once per call WHATEVER the batch (<= 16), and the arithmetic runs on
// mma.sync.m16n8k16 (the SIMT GEMV of gemv.cuh is ALU-bound from 4 tokens on: 46 % of HBM at batch 4).  Orientation: the TOKENS are
// the mma M dimension (16 slots: batch 9..16 costs what batch 1..8 costs), a work unit is 8 WEIGHT ROWS (the mma N dimension; 16 for the
// quantised formats, whose activation fragments are then loaded once per two row groups), so small linears (N = 4096: 512 / 256 units)
// still spread over all SMs.  Structure:
//   * one persistent CTA per SM: 16 compute warps, one TMA-producer warp, one reducer warp.  Units are dealt round-robin over CTAs; the
//     16 warps split the k range of every ring stage (a stage = <= 4 KiB of each of the unit's 8 rows, one 1-D bulk copy per row), so
//     a unit is done by the whole CTA and the tail of a launch is one stage, not one unit;
//   * the ring is filled BEFORE griddepcontrol.wait (programmatic dependent launch), full / empty mbarriers, evict-first L2 policy;
//   * activations: a plain [M,K] tensor of T (the RMSNorm / residual / tensor-parallel reduce in front of a linear run as one small
//     kernel for 2+ tokens: fused into this prologue, every one of the 148 CTAs redid them for all tokens -- two sweeps of M * K / 512
//     dependent L2 loads per thread, 6-16 us per launch, twice the weight-streaming time at batch 8).  They are staged in shared memory
//     in K PARTS of <= 64 KiB whatever K is, which is what lets K = 8192 (70B-shaped) and K = 11008 (7B down projection) keep a >= 4-stage
//     ring at 8 and at 16 tokens.  Dense: a part is M bulk copies (one per token row) issued by one thread, no thread touches the data;
//     quantised: one sweep with all of a thread's loads in flight (bf16 -> f16, INT4 k order).  Between parts only the 16 compute warps
//     meet (named barrier); the producer keeps the ring full meanwhile;
//   * weights reach the B fragments straight from the stage rows: a lane reads the 8 consecutive k of (row g) that its two k-steps
//     use -- 16 bytes dense, 8 bytes FP8 (cvt.rn.f16x2.e4m3x2, exact), 4 bytes INT4 (0x6400 magic number, exact integers, the group's
//     scale applied to the accumulators once per 128 k) -- and the same 8 k of token g / g+8 for the A fragments; k is permuted
//     consistently inside each 32-k block for A and B, so no shuffles and no ldmatrix;
//   * per-warp 16x8 fp32 tiles meet in shared memory (double-buffered, ready / free mbarriers); the reducer warp adds them in a fixed
//     order (deterministic), across K parts too, applies the FP8 row scale / SwiGLU (reference src/kernels/silu_and_mul.cu:6-41; a
//     unit is then 4 gate rows interleaved with their 4 up rows) and stores -- or pushes the tensor-parallel partial (TpPush).
#pragma once
#include <type_traits>

#include "gemv.cuh"

namespace b200 {

constexpr int kMmaWarps = 16;                        // compute warps
constexpr int kMmaThreads = (kMmaWarps + 2) * 32;    // + producer warp + reducer warp
constexpr int kMmaRows = 8;                          // weight rows per row group = mma N; a unit is RG row groups
constexpr int kMmaMaxStages = 8;

struct MmaGeom {
    int piece_bytes;  // bytes of one weight row per ring stage (one bulk copy)
    int piece_k;      // k per stage
    int row_stride;   // bytes between the rows of a stage (piece + pad: conflict-free fragment loads)
    int stage_bytes;
    int stages;
    int part_k;       // k per activation part (a multiple of piece_k)
    int parts;
    int xs_stride;    // elements between the token rows of the staged part
    int tok;          // token slots of the mma tile in use: 8 or 16
    int xs_rows;      // token rows staged: M (+ one zero row that every slot past the batch reads)
    int max_units;    // units of the busiest CTA (size of the cross-part accumulator)
    int rg;           // row groups of 8 weight rows per unit (the A fragments of a k block are loaded once and used rg times)
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t mma_e4m3x2_to_f16x2(uint32_t two_bytes) {
    uint32_t h;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h) : "h"((unsigned short)two_bytes));
    return h;
}
// The 8 nibbles of a 32-bit word (nibble i = element i) as four f16x2 pairs {n_i - z, n_{i+4} - z}, i = 0..3, exact integers:
//   i even: ((w' & 0x000f000f) | 0x64006400) = {1024 + n, 1024 + n} minus {1024 + z};
//   i odd : ((w' & 0x00f000f0) | 0x64006400) = {1024 + 16 n, ...}: one HFMA2 with 1/16 and -(64 + z) (exact: 64 + n is an f16 integer);
//   w' = w for i = 0, 1 and w >> 8 for i = 2, 3.
// zpk = f16x2 {1024 + z, 1024 + z}, nz = f16x2 {-(64 + z), -(64 + z)}: prepared once per quantisation group (mma_int4_consts)
__device__ __forceinline__ void mma_int4_consts(uint32_t zq, uint32_t &zpk, uint32_t &nz) {
    const uint32_t z16 = 0x6400u | zq;
    zpk = z16 | (z16 << 16);
    const __half2 n2 = __float2half2_rn(-(64.0f + (float)zq));
    nz = *reinterpret_cast<const uint32_t *>(&n2);
}
__device__ __forceinline__ void mma_deq_int4(uint32_t w, uint32_t zpk, uint32_t nz32, uint32_t (&d)[4]) {
    const uint32_t m_lo = 0x000f000fu, m_hi = 0x00f000f0u, magic = 0x64006400u, sixteenth = 0x2c002c00u;  // f16 1/16 = 0x2c00
    const uint32_t w8 = w >> 8;
    uint32_t h0, h1, h2, h3;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(h0) : "r"(w), "r"(m_lo), "r"(magic));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(h1) : "r"(w), "r"(m_hi), "r"(magic));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(h2) : "r"(w8), "r"(m_lo), "r"(magic));
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(h3) : "r"(w8), "r"(m_hi), "r"(magic));
    const __half2 z = *reinterpret_cast<const __half2 *>(&zpk), nz = *reinterpret_cast<const __half2 *>(&nz32);
    const __half2 s16 = *reinterpret_cast<const __half2 *>(&sixteenth);
    const __half2 r0 = __hsub2(*reinterpret_cast<const __half2 *>(&h0), z), r2 = __hsub2(*reinterpret_cast<const __half2 *>(&h2), z);
    const __half2 r1 = __hfma2(*reinterpret_cast<const __half2 *>(&h1), s16, nz), r3 = __hfma2(*reinterpret_cast<const __half2 *>(&h3), s16, nz);
    d[0] = *reinterpret_cast<const uint32_t *>(&r0), d[1] = *reinterpret_cast<const uint32_t *>(&r1);
    d[2] = *reinterpret_cast<const uint32_t *>(&r2), d[3] = *reinterpret_cast<const uint32_t *>(&r3);
}

// smem: [ xs : tok * xs_stride * 2 B | ring : stages * stage_bytes | barriers : (2 * kMmaMaxStages + 4) * 8 | tiles : 2 * 16 warps * 128
//         floats | cross-part accumulator : max_units * 128 floats (parts > 1) ]
// NT = 1: M <= 8 (token slots 8..15 are zero A fragments), NT = 2: M <= 16.
// RG = row groups (of 8 weight rows) per unit: 2 for FP8, whose k per weight byte is twice the dense one -- with one row group the
// activation fragments re-read from shared memory for every 8 rows load the shared-memory pipe as much as the weights do.
template <typename T, int FMT, bool kSwiGLU, int NT, int RG>
__global__ void __launch_bounds__(kMmaThreads, 1)
gemv_mma_kernel(const GemvArgs a, const MmaGeom geo) {
    constexpr int V = Elem<T>::kVec;
    static_assert(V == 8, "gemv_mma_kernel: 16-bit activation types only");
    constexpr bool kDense = FMT == WF_DENSE;
    constexpr int R = kMmaRows * RG;  // weight rows per unit
    constexpr int kMmaTile = 16 * R;  // fp32 values of one warp's 16-token x R-row tile
    static_assert(R <= 16, "one producer lane per row of a unit");
    using XT = typename std::conditional<kDense, T, __half>::type;  // staged activation type (quantised weights dequantise to f16)
    extern __shared__ __align__(128) unsigned char smem[];

    const int K = a.K, N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const size_t row_bytes = kDense ? (size_t)K * 2 : (FMT == WF_FP8 ? (size_t)K : (size_t)K / 2);
    const int units = kSwiGLU ? (a.inter + R / 2 - 1) / (R / 2) : (N + R - 1) / R;
    const int stages = geo.stages, parts = geo.parts;
    const bool is_compute = warp < kMmaWarps, is_producer = warp == kMmaWarps;
    const int gid = blockIdx.x, total = gridDim.x;
    const int my_units = gid < units ? (units - gid + total - 1) / total : 0;
    auto part_pieces = [&](int p) -> int { return (min(geo.part_k, K - p * geo.part_k) + geo.piece_k - 1) / geo.piece_k; };
    int my_items = 0;
    for (int p = 0; p < parts; ++p) my_items += part_pieces(p) * my_units;

    XT *xs = reinterpret_cast<XT *>(smem);
    size_t off = ((size_t)geo.xs_rows * geo.xs_stride * sizeof(XT) + 127) & ~(size_t)127;
    unsigned char *ring = smem + off;
    off += (size_t)stages * geo.stage_bytes;
    const uint32_t full0 = smem_u32(smem + off), empty0 = full0 + kMmaMaxStages * 8, ready0 = empty0 + kMmaMaxStages * 8, free0 = ready0 + 16;
    const uint32_t xbar = free0 + 16;  // dense: the staged part has landed (bulk copies)
    off += (size_t)(2 * kMmaMaxStages + 6) * 8;
    float *tiles = reinterpret_cast<float *>(smem + off);  // [parity][warp][16 tokens][8 rows]
    off += (size_t)2 * kMmaWarps * kMmaTile * sizeof(float);
    float *outacc = reinterpret_cast<float *>(smem + off);  // [unit][16][8], parts > 1 only
    off += parts > 1 ? (size_t)geo.max_units * kMmaTile * sizeof(float) : 0;
    const uint32_t ring_u32 = smem_u32(ring);

    // row r (0 .. R-1) of unit u: plain R u + r; SwiGLU: gate row j = (R/2) u + r/2 (r even) and its up row inter + j (r odd)
    auto unit_row = [&](int u, int r) -> int { return kSwiGLU ? ((r & 1) ? a.inter : 0) + (R / 2) * u + (r >> 1) : R * u + r; };
    auto row_ok = [&](int u, int r) -> bool { return kSwiGLU ? ((R / 2) * u + (r >> 1) < a.inter) : (R * u + r < N); };

    // ---- producer WARP: order (part, unit, piece); lane 0 arms the stage barrier, lanes 0 .. R-1 issue one bulk copy per row
    int p_part = 0, p_un = 0, p_pc = 0, p_item = 0, p_s = 0;
    auto issue_next = [&]() {  // all 32 lanes of the producer warp
        const int u = gid + p_un * total;
        const int k0 = p_part * geo.part_k + p_pc * geo.piece_k;
        const uint32_t bytes = (uint32_t)((size_t)min(geo.piece_k, K - k0) * row_bytes / (size_t)K);
        const uint32_t bar = full0 + p_s * 8;
        const bool mine = lane < R && row_ok(u, lane);
        const unsigned nrows = __popc(__ballot_sync(0xffffffffu, mine));
        if (lane == 0) mbar_expect_tx(bar, bytes * nrows);
        __syncwarp();
        if (mine)
            bulk_g2s(ring_u32 + p_s * geo.stage_bytes + lane * geo.row_stride,
                     reinterpret_cast<const unsigned char *>(a.w) + (size_t)unit_row(u, lane) * row_bytes + (size_t)k0 * row_bytes / (size_t)K, bytes, bar);
        ++p_item;
        if (++p_pc == part_pieces(p_part)) {
            p_pc = 0;
            if (++p_un == my_units) p_un = 0, ++p_part;
        }
        if (++p_s == stages) p_s = 0;
    };
    if (is_producer) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(full0 + s * 8, 1);
                mbar_init(empty0 + s * 8, kMmaWarps);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(ready0 + b * 8, kMmaWarps);
                mbar_init(free0 + b * 8, 1);
            }
            mbar_init(xbar, 1);
            fence_mbar_init();
        }
        __syncwarp();
        while (p_item < stages && p_item < my_items) issue_next();  // the weights do not depend on the previous kernel
    }
    __syncthreads();  // publishes the mbarrier initialisation to the consumers
    pdl_wait();

    if (is_producer) {
        // ================================================= TMA producer: refill a stage once all 16 warps have left it
        int e_s = 0, e_ph = 0;
        while (p_item < my_items) {
            mbar_wait(empty0 + e_s * 8, e_ph);
            if (++e_s == stages) e_s = 0, e_ph ^= 1;
            fence_proxy_async();
            issue_next();
        }
    } else if (!is_compute) {
        // ================================================= reducer: per pass, lane l finishes 4 consecutive rows of one token of the tile
        const unsigned int push_flag = a.push.n > 0 ? tp_flag(a.push.epoch, a.push.seq) : 0u;
        int h = 0;  // hand-off counter: order (part, unit)
        for (int p = 0; p < parts; ++p) {
            for (int un = 0; un < my_units; ++un, ++h) {
                const int u = gid + un * total;
                const int b = h & 1;
                float s8[RG][4];
                if constexpr (FMT == WF_FP8) {  // requested before the wait: the L2 latency hides behind the compute warps
#pragma unroll
                    for (int ps = 0; ps < RG; ++ps) {
                        const int r0 = (ps * 128 + lane * 4) % R;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            s8[ps][i] = (p + 1 == parts && row_ok(u, r0 + i)) ? __ldg(reinterpret_cast<const float *>(a.scales) + unit_row(u, r0 + i)) : 1.0f;
                    }
                }
                mbar_wait(ready0 + b * 8, (h >> 1) & 1);
                float4 v[RG];
#pragma unroll
                for (int ps = 0; ps < RG; ++ps) {
                    const float *slot = tiles + (size_t)b * kMmaWarps * kMmaTile + ps * 128 + lane * 4;
                    v[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w2 = 0; w2 < kMmaWarps; ++w2) {  // fixed order: deterministic
                        const float4 q = *reinterpret_cast<const float4 *>(slot + (size_t)w2 * kMmaTile);
                        v[ps].x += q.x, v[ps].y += q.y, v[ps].z += q.z, v[ps].w += q.w;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(free0 + b * 8);
#pragma unroll
                for (int ps = 0; ps < RG; ++ps) {
                    const int e0 = ps * 128 + lane * 4, m = e0 / R, r0 = e0 % R;  // token, first of this lane's 4 rows
                    if (parts > 1) {  // accumulate across K parts (part order: deterministic)
                        float4 *acc = reinterpret_cast<float4 *>(outacc + (size_t)un * kMmaTile + e0);
                        if (p > 0) {
                            const float4 o = *acc;
                            v[ps].x += o.x, v[ps].y += o.y, v[ps].z += o.z, v[ps].w += o.w;
                        }
                        if (p + 1 < parts) {
                            *acc = v[ps];
                            continue;
                        }
                    }
                    if (m >= a.M) continue;
                    float o[4] = {v[ps].x, v[ps].y, v[ps].z, v[ps].w};
                    if constexpr (FMT == WF_FP8) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i] *= s8[ps][i];
                    }
                    if constexpr (kSwiGLU) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int j = (R / 2) * u + (r0 >> 1) + i;  // rows (r0 + 2i, r0 + 2i + 1) = (gate_j, up_j)
                            if (j < a.inter) {
                                // the un-fused reference stores gate / up in T before SiLU reads them
                                const float gt = round_to<T>(o[2 * i]), up = round_to<T>(o[2 * i + 1]);
                                const float sv = (gt / (1.0f + expf(-gt))) * up;
                                if (a.y_f32) reinterpret_cast<float *>(a.y)[(size_t)m * a.inter + j] = sv;
                                else reinterpret_cast<T *>(a.y)[(size_t)m * a.inter + j] = Elem<T>::from_f(sv);
                            }
                        }
                    } else if (a.push.n > 0) {
                        // tensor-parallel partial: this lane's four consecutive rows are two LL words = one 16-byte store per peer
                        if (row_ok(u, r0 + 3) && (N & 3) == 0) {
                            tp_push_quad<T>(a.push, push_flag, (size_t)m * N + R * u + r0, o);
                        } else {
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (row_ok(u, r0 + 2 * i + 1)) tp_push_pair<T>(a.push, push_flag, (size_t)m * N + R * u + r0 + 2 * i, o[2 * i], o[2 * i + 1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (row_ok(u, r0 + i)) {
                                const size_t idx = (size_t)m * N + R * u + r0 + i;
                                if (a.y_f32) reinterpret_cast<float *>(a.y)[idx] = o[i];
                                else reinterpret_cast<T *>(a.y)[idx] = Elem<T>::from_f(o[i]);
                            }
                    }
                }
            }
        }
    } else {
        // ================================================= compute warps
        const int tid = threadIdx.x;  // 0 .. 511
        constexpr int NC = kMmaWarps * 32;
        auto cbar = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(NC) : "memory"); };
        const T *xin = reinterpret_cast<const T *>(a.x);
        // token slots past the batch all read one row of zeros (row M), written once
        for (int i = tid; i < (geo.xs_rows - a.M) * geo.xs_stride / 8; i += NC)
            reinterpret_cast<uint4 *>(xs + (size_t)a.M * geo.xs_stride)[i] = make_uint4(0u, 0u, 0u, 0u);
        cbar();

        const int wk = geo.piece_k / kMmaWarps;  // k per warp per stage (a multiple of 32; 128 for INT4 groups)
        const int ngroups_k = FMT == WF_INT4 ? K / a.group : 0;
        const int wbytes = geo.piece_bytes / kMmaWarps;  // bytes of a stage row this warp owns
        int s = 0, ph = 0, h = 0;
        for (int p = 0; p < parts; ++p) {
            const int pk0 = p * geo.part_k, pk1 = min(K, pk0 + geo.part_k);
            // ---- stage this part of the activations (the ring keeps filling meanwhile)
            if (p > 0) cbar();  // every warp has finished reading the previous part
            if constexpr (kDense) {
                // M bulk copies, one per token row, issued by one thread; everybody waits on the mbarrier
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)(pk1 - pk0) * 2u;
                    fence_proxy_async();
                    mbar_expect_tx(xbar, bytes * (uint32_t)a.M);
                    for (int m = 0; m < a.M; ++m) {
                        // activations are re-read by every CTA: default L2 policy (not evict-first)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         smem_u32(xs + (size_t)m * geo.xs_stride)),
                                     "l"(xin + (size_t)m * K + pk0), "r"(bytes), "r"(xbar)
                                     : "memory");
                    }
                }
                mbar_wait(xbar, p & 1);
            } else {
                // one sweep: thread (column c) takes vector c of every token row, all loads in flight before the first conversion
                const int nvp = (pk1 - pk0) / V;
                for (int c = tid; c < nvp; c += NC) {
#pragma unroll 1
                    for (int m0 = 0; m0 < a.M; m0 += 8) {  // eight rows' loads in flight at a time
                        uint4 raw[8];
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            if (m0 + m < a.M) raw[m] = ld_v4(xin + (size_t)(m0 + m) * K + pk0 + (size_t)c * V);
#pragma unroll
                        for (int m = 0; m < 8; ++m)
                            if (m0 + m < a.M) {
                                float q[V], o[V];
                                unpack16<T>(raw[m], q);
#pragma unroll
                                for (int j = 0; j < 8; ++j) o[FMT == WF_INT4 ? ((j & 3) * 2 + (j >> 2)) : j] = q[j];  // INT4: [k0 k4 k1 k5 k2 k6 k3 k7]
                                const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
                                const __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
                                uint4 pk;
                                pk.x = *reinterpret_cast<const uint32_t *>(&h0), pk.y = *reinterpret_cast<const uint32_t *>(&h1);
                                pk.z = *reinterpret_cast<const uint32_t *>(&h2), pk.w = *reinterpret_cast<const uint32_t *>(&h3);
                                *reinterpret_cast<uint4 *>(xs + (size_t)(m0 + m) * geo.xs_stride + (size_t)c * V) = pk;
                            }
                    }
                }
                cbar();
            }
            if (p == 0) pdl_launch_dependents();

            const XT *x0 = xs + (size_t)min(g, a.M) * geo.xs_stride + 8 * t;
            const XT *x1 = xs + (size_t)min(g + 8, a.M) * geo.xs_stride + 8 * t;
            const int npc = part_pieces(p);
            // INT4: zero points of rows g + 8 rg and scales of rows 2t, 2t+1 (+ 8 rg) for the (<= 2) quantisation groups of this warp's
            // slice of a stage.  They are requested ONE STAGE AHEAD (a stage is 0.4-0.7 us of this SM's share of HBM, an L2 round trip is longer)
            struct QParams {
                uint32_t z;         // byte [2 rg + q]
                float s0[RG][2], s1[RG][2];
            };
            auto load_qparams = [&](int un_, int pc_) -> QParams {
                QParams q = {};
                if constexpr (FMT == WF_INT4) {
                    const int u_ = gid + un_ * total;
                    const int kw_ = pk0 + pc_ * geo.piece_k + warp * wk, kend_ = min(kw_ + wk, pk1);
                    const int ng = kend_ > kw_ ? (kend_ - kw_ + a.group - 1) / a.group : 0;
#pragma unroll
                    for (int rg = 0; rg < RG; ++rg) {
                        const int zrow = min(unit_row(u_, 8 * rg + g), N - 1), srow0 = min(unit_row(u_, 8 * rg + 2 * t), N - 1),
                                  srow1 = min(unit_row(u_, 8 * rg + 2 * t + 1), N - 1);
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            if (j < ng) {
                                const size_t gi = (size_t)(kw_ / a.group + j);
                                q.z |= (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(a.zeros) + (size_t)zrow * ngroups_k + gi) << (8 * (2 * rg + j));
                                q.s0[rg][j] = Elem<T>::to_f(__ldg(reinterpret_cast<const T *>(a.scales) + (size_t)srow0 * ngroups_k + gi));
                                q.s1[rg][j] = Elem<T>::to_f(__ldg(reinterpret_cast<const T *>(a.scales) + (size_t)srow1 * ngroups_k + gi));
                            }
                    }
                }
                return q;
            };
            QParams qcur = my_units > 0 ? load_qparams(0, 0) : QParams{};
            for (int un = 0; un < my_units; ++un, ++h) {
                float acc[RG][4], acc2[RG][4];
#pragma unroll
                for (int rg = 0; rg < RG; ++rg)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[rg][i] = acc2[rg][i] = 0.0f;
                for (int pc = 0; pc < npc; ++pc) {
                    const int kw = pk0 + pc * geo.piece_k + warp * wk;  // first k of this warp's slice of the stage
                    const int kend = min(kw + wk, pk1);
                    mbar_wait(full0 + s * 8, ph);
                    QParams qnxt = {};
                    if constexpr (FMT == WF_INT4) {
                        const int pc2 = pc + 1 < npc ? pc + 1 : 0, un2 = pc + 1 < npc ? un : un + 1;
                        if (un2 < my_units) qnxt = load_qparams(un2, pc2);
                    }
                    // row g of row group rg, this warp's bytes of the stage row
                    const unsigned char *wrow = ring + (size_t)s * geo.stage_bytes + (size_t)g * geo.row_stride + (size_t)warp * wbytes;
                    const size_t rg_stride = (size_t)8 * geo.row_stride;
                    const XT *xa = x0 + (kw - pk0), *xb = x1 + (kw - pk0);
                    if constexpr (FMT == WF_INT4) {
#pragma unroll
                        for (int q = 0; q < 2; ++q) {  // 128-k quantisation groups of the slice (unrolled: the parameters stay in registers)
                            if (kw + q * 128 >= kend) continue;
                            float ag[RG][4];
                            uint32_t zpk[RG], nz[RG];
#pragma unroll
                            for (int rg = 0; rg < RG; ++rg) {
                                ag[rg][0] = ag[rg][1] = ag[rg][2] = ag[rg][3] = 0.0f;
                                mma_int4_consts((qcur.z >> (8 * (2 * rg + q))) & 0xffu, zpk[rg], nz[rg]);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {  // 32-k blocks of the group: the A fragments are loaded once for all row groups
                                const int kk = q * 128 + j * 32;
                                const uint4 av = *reinterpret_cast<const uint4 *>(xa + kk);
                                uint4 bv = make_uint4(0u, 0u, 0u, 0u);
                                if constexpr (NT == 2) bv = *reinterpret_cast<const uint4 *>(xb + kk);
#pragma unroll
                                for (int rg = 0; rg < RG; ++rg) {
                                    const uint32_t wv = *reinterpret_cast<const uint32_t *>(wrow + rg * rg_stride + (kk + 8 * t) / 2);
                                    uint32_t d[4];
                                    mma_deq_int4(wv, zpk[rg], nz[rg], d);
                                    mma_f16f16(ag[rg], av.x, bv.x, av.y, bv.y, d[0], d[1]);
                                    mma_f16f16(ag[rg], av.z, bv.z, av.w, bv.w, d[2], d[3]);
                                }
                            }
#pragma unroll
                            for (int rg = 0; rg < RG; ++rg) {
                                acc[rg][0] = fmaf(qcur.s0[rg][q], ag[rg][0], acc[rg][0]), acc[rg][1] = fmaf(qcur.s1[rg][q], ag[rg][1], acc[rg][1]);
                                acc[rg][2] = fmaf(qcur.s0[rg][q], ag[rg][2], acc[rg][2]), acc[rg][3] = fmaf(qcur.s1[rg][q], ag[rg][3], acc[rg][3]);
                            }
                        }
                        qcur = qnxt;
                    } else {
                        // 32-k blocks (two mma k-steps each), four at a time: all fragment loads first, then two independent accumulator
                        // chains per row group (a rolled loop serialised load -> mma -> mma per block: 4 x ~110 cycles per stage and warp)
                        using WQ = typename std::conditional<kDense, uint4, uint2>::type;  // the 8 weights of (a row, this lane's 8 k)
                        auto block = [&](float (&c)[4], const uint4 &av, const uint4 &bv, const WQ &wq) {
                            if constexpr (kDense) {
                                if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                                    mma_bf16(c, av.x, bv.x, av.y, bv.y, wq.x, wq.y);
                                    mma_bf16(c, av.z, bv.z, av.w, bv.w, wq.z, wq.w);
                                } else {
                                    mma_f16f16(c, av.x, bv.x, av.y, bv.y, wq.x, wq.y);
                                    mma_f16f16(c, av.z, bv.z, av.w, bv.w, wq.z, wq.w);
                                }
                            } else {  // FP8: wq.x / wq.y hold the 8 bytes of the block
                                mma_f16f16(c, av.x, bv.x, av.y, bv.y, mma_e4m3x2_to_f16x2(wq.x), mma_e4m3x2_to_f16x2(wq.x >> 16));
                                mma_f16f16(c, av.z, bv.z, av.w, bv.w, mma_e4m3x2_to_f16x2(wq.y), mma_e4m3x2_to_f16x2(wq.y >> 16));
                            }
                        };
                        auto load_w = [&](int rg, int kk) -> WQ {
                            if constexpr (kDense) return *reinterpret_cast<const uint4 *>(wrow + rg * rg_stride + (size_t)(kk + 8 * t) * 2);
                            else return *reinterpret_cast<const uint2 *>(wrow + rg * rg_stride + (kk + 8 * t));
                        };
                        int kk = 0;
                        for (; kw + kk + 128 <= kend; kk += 128) {
                            WQ wq[RG][4];
                            uint4 av[4], bv[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
#pragma unroll
                                for (int rg = 0; rg < RG; ++rg) wq[rg][j] = load_w(rg, kk + 32 * j);
                                av[j] = *reinterpret_cast<const uint4 *>(xa + kk + 32 * j);
                                bv[j] = make_uint4(0u, 0u, 0u, 0u);
                                if constexpr (NT == 2) bv[j] = *reinterpret_cast<const uint4 *>(xb + kk + 32 * j);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int rg = 0; rg < RG; ++rg) block((j & 1) ? acc2[rg] : acc[rg], av[j], bv[j], wq[rg][j]);
                        }
                        for (; kw + kk < kend; kk += 32) {  // short slices (K not a multiple of the stage width)
                            const uint4 av = *reinterpret_cast<const uint4 *>(xa + kk);
                            uint4 bv = make_uint4(0u, 0u, 0u, 0u);
                            if constexpr (NT == 2) bv = *reinterpret_cast<const uint4 *>(xb + kk);
#pragma unroll
                            for (int rg = 0; rg < RG; ++rg) block(acc[rg], av, bv, load_w(rg, kk));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + s * 8);  // this warp has left the stage
                    if (++s == stages) s = 0, ph ^= 1;
                }
                // ---- hand this warp's 16 x R tile of the (part, unit) to the reducer (double-buffered slot)
                const int b = h & 1;
                if (h >= 2) mbar_wait(free0 + b * 8, ((h >> 1) - 1) & 1);
                float *slot = tiles + ((size_t)b * kMmaWarps + warp) * kMmaTile;
#pragma unroll
                for (int rg = 0; rg < RG; ++rg) {
                    *reinterpret_cast<float2 *>(slot + g * R + 8 * rg + 2 * t) = make_float2(acc[rg][0] + acc2[rg][0], acc[rg][1] + acc2[rg][1]);
                    *reinterpret_cast<float2 *>(slot + (g + 8) * R + 8 * rg + 2 * t) = make_float2(acc[rg][2] + acc2[rg][2], acc[rg][3] + acc2[rg][3]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(ready0 + b * 8);
            }
        }
    }
}

// host side ----------------------------------------------------------------------------------------------------------
// Geometry for (M, K, format): false when the shape cannot use this kernel.
inline size_t gemv_mma_fixed_smem(const MmaGeom &g) {
    size_t fixed = ((size_t)g.xs_rows * g.xs_stride * 2 + 127) & ~(size_t)127;
    const size_t tile = (size_t)16 * kMmaRows * g.rg * 4;  // bytes of one warp's fp32 tile
    return fixed + (size_t)(2 * kMmaMaxStages + 6) * 8 + (size_t)2 * kMmaWarps * tile + (g.parts > 1 ? (size_t)g.max_units * tile : 0);
}
inline size_t gemv_mma_smem(const MmaGeom &g) { return gemv_mma_fixed_smem(g) + (size_t)g.stages * g.stage_bytes; }

// Geometry for (M, K, format): false when the shape cannot use this kernel.  Among the (bytes per stage row, number of activation parts)
// that leave a ring of >= 3 stages, the best-scoring one wins (see the score below).
// row groups per unit: 2 for FP8 (measured at batch 16: 5.32 -> 4.98 ms per 7B step); INT4 stays at 1 (with 2 its dequantisation
// registers spill: 6.06 -> 8.15 ms)
inline int gemv_mma_row_groups(int fmt) { return fmt == WF_FP8 ? 2 : 1; }
inline bool gemv_mma_geometry(int M, int K, int N_units_max, int fmt, MmaGeom *out) {
    if (M < 1 || M > 16 || K % 128 != 0) return false;
    const bool dense = fmt == WF_DENSE;
    const size_t budget = 226 * 1024;
    const int k_round = (K + 511) / 512 * 512;  // what 16 warps x 32 k can split
    long best = 0;
    bool found = false;
    // a part boundary (one named barrier + one re-staging + one more hand-off per unit) is worth 48 KiB of ring at <= 8 tokens and 8 KiB at
    // 9..16, where the alternative to more parts is 2 KiB copies (measured, 7B step: batch 8 3.78 ms with 1 part against 3.90 with 2;
    // batch 16 4.82 ms with 2 / 6 parts of 4 KiB pieces against 5.16 with 2 / 3 parts and 2 KiB pieces for the down projection)
    const long part_penalty = M <= 8 ? 49152L : 8192L;
    for (int piece_bytes = fmt == WF_INT4 ? 2048 : 4096; piece_bytes >= 1024; piece_bytes /= 2) {
        MmaGeom g = {};
        g.tok = M <= 8 ? 8 : 16;
        g.xs_rows = M == g.tok ? g.tok : M + 1;  // token slots past the batch all read ONE zero row
        g.max_units = N_units_max;
        g.rg = gemv_mma_row_groups(fmt);
        g.piece_bytes = piece_bytes;
        g.piece_k = dense ? g.piece_bytes / 2 : (fmt == WF_FP8 ? g.piece_bytes : g.piece_bytes * 2);
        if (g.piece_k > k_round) {  // short rows: one piece
            g.piece_k = k_round;
            g.piece_bytes = dense ? g.piece_k * 2 : (fmt == WF_FP8 ? g.piece_k : g.piece_k / 2);
        }
        // a warp's slice of a stage must hold whole quantisation groups, and at most two of them (their parameters live in registers)
        if (fmt == WF_INT4 && ((g.piece_k / kMmaWarps) % 128 != 0 || g.piece_k / kMmaWarps > 256)) continue;
        // pad so that the per-lane fragment loads of 8 rows fall on different banks: 16 B (dense) / 8 B (FP8) / 4 B (INT4) per lane, 4 lanes per row
        g.row_stride = g.piece_bytes + (dense ? 64 : (fmt == WF_FP8 ? 32 : 16));
        g.stage_bytes = kMmaRows * g.rg * g.row_stride;
        const int pieces_total = (K + g.piece_k - 1) / g.piece_k;
        for (g.parts = 1; g.parts <= pieces_total; ++g.parts) {
            const int ppp = (pieces_total + g.parts - 1) / g.parts;  // pieces per part
            if ((pieces_total + ppp - 1) / ppp != g.parts) continue;  // part counts that would leave an empty part
            g.part_k = ppp * g.piece_k;
            g.xs_stride = g.part_k + 32;  // token rows 64 B apart modulo 128: conflict-free 16-byte loads of (token g, 8 k of lane t)
            const size_t fixed = gemv_mma_fixed_smem(g);
            if (fixed + (size_t)3 * g.stage_bytes > budget) continue;
            g.stages = (int)((budget - fixed) / g.stage_bytes);
            if (g.stages > kMmaMaxStages) g.stages = kMmaMaxStages;
            // bytes in flight (what keeps HBM busy; beyond ~190 KiB nothing is gained), 4 KiB bulk copies preferred over 2 KiB ones (the
            // copy engine is request-rate bound: round 1 measured 1 KiB copies capping a kernel near 3.3 TB/s), minus the part penalty
            const long inflight = (long)g.stages * g.stage_bytes;
            const long score = (inflight < 196608 ? inflight : 196608) + (g.piece_bytes >= 4096 ? 49152 : (g.piece_bytes >= 2048 ? 0 : -49152)) -
                               part_penalty * (g.parts - 1);
            if (!found || score > best) best = score, *out = g, found = true;
        }
    }
    return found;
}

template <typename T, int FMT, bool SW, int NT>
static int launch_gemv_mma_inst(const GemvArgs &a, const MmaGeom &g, int grid, cudaStream_t st) {
    auto kern = gemv_mma_kernel<T, FMT, SW, NT, FMT == WF_FP8 ? 2 : 1>;
    const size_t smem = gemv_mma_smem(g);
    static thread_local size_t cached_smem[64] = {0};  // per device, per instantiation
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cuda_status("gemv_mma cudaFuncSetAttribute");
        cached_smem[dev] = smem;
    }
    launch_pdl(kern, dim3(grid), dim3(kMmaThreads), smem, st, true, a, g);
    return cuda_status("gemv_mma launch");
}

// Returns B200_ERR_UNSUPPORTED (no error text) when the shape cannot use the tensor-core GEMV.
template <typename T>
static int launch_gemv_mma_t(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) {
    if (a.M < 1 || a.M > 16 || a.K % 128 != 0 || !aligned16(a.w) || !aligned16(a.x)) return B200_ERR_UNSUPPORTED;
    if (a.norm || a.tp.world > 1) return B200_ERR_UNSUPPORTED;  // plain activations only (see the header: the prologue work runs once, in front)
    if (fmt == WF_INT4 && a.group != 128) return B200_ERR_UNSUPPORTED;
    if (swiglu && a.inter % 8 != 0) return B200_ERR_UNSUPPORTED;
    if (a.push.n > 0 && a.N % 2 != 0) return B200_ERR_UNSUPPORTED;
    const size_t row_bytes = fmt == WF_DENSE ? (size_t)a.K * 2 : (fmt == WF_FP8 ? (size_t)a.K : (size_t)a.K / 2);
    if (row_bytes % 16 != 0) return B200_ERR_UNSUPPORTED;
    const int R = kMmaRows * gemv_mma_row_groups(fmt);  // weight rows per unit
    const int units = swiglu ? (a.inter + R / 2 - 1) / (R / 2) : (a.N + R - 1) / R;
    int grid = sm_count();
    if (grid > units) grid = units;
    if (grid < 1) grid = 1;
    MmaGeom g = {};
    if (!gemv_mma_geometry(a.M, a.K, (units + grid - 1) / grid, fmt, &g)) return B200_ERR_UNSUPPORTED;
#define B200_MMA_GO(FMTC)                                                                                                     \
    do {                                                                                                                      \
        if (swiglu) return g.tok == 8 ? launch_gemv_mma_inst<T, FMTC, true, 1>(a, g, grid, st) : launch_gemv_mma_inst<T, FMTC, true, 2>(a, g, grid, st); \
        return g.tok == 8 ? launch_gemv_mma_inst<T, FMTC, false, 1>(a, g, grid, st) : launch_gemv_mma_inst<T, FMTC, false, 2>(a, g, grid, st);          \
    } while (0)
    if (fmt == WF_DENSE) B200_MMA_GO(WF_DENSE);
    if (fmt == WF_FP8) B200_MMA_GO(WF_FP8);
    if (fmt == WF_INT4) B200_MMA_GO(WF_INT4);
#undef B200_MMA_GO
    return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
