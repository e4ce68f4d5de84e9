// gemv_q.cuh -- weight-only quantised decode linear (FP8-e4m3 W8A16, INT4-g128 W4A16) for M <= 8 tokens on sm_100a:
// y[M,N] = x[M,K] * dequant(W)^T, W packed [N,K].
//
// Only the packed bytes cross HBM, and the dot products run on the tensor cores: a work unit is 16 weight rows, the
// mma.sync.m16n8k16 (f16 x f16 -> fp32) M dimension; the (<= 8) tokens are its N dimension.  The weights are dequantised in
// registers straight into A fragments -- e4m3 pairs with cvt.rn.f16x2.e4m3x2 (exact), int4 nibbles with the 0x6400 magic number and
// one hsub2 against the group's zero point (exact integers) -- so the per-weight ALU work is 1 (FP8) or ~1.4 (INT4) instructions
// instead of the ~4 of a SIMT dequant + FMA, which left the first version of the quantised path slower than bf16.  Activations are
// staged as f16 (exact for bf16 / fp16 inputs in f16's normal range).  Everything else is the structure of gemv_nk_kernel
// (gemv.cuh): persistent CTA per SM, 2 groups x 8 compute warps splitting K inside a unit, a TMA-bulk producer warp and a reducer
// warp per group, ring filled before griddepcontrol.wait, fused residual / bias / RMSNorm / tensor-parallel-exchange prologue,
// optional SwiGLU epilogue (unit = 8 gate rows + the 8 matching up rows), deterministic fixed-order reductions.
#pragma once
#include "gemv.cuh"

namespace b200 {

constexpr int kQRows = 16;         // weight rows per unit = mma M
constexpr int kQTok = 8;           // token slots = mma N
constexpr int kQPieceBytes = 2048;  // preferred bytes of a row per ring stage = one bulk copy (1 KiB copies cap the kernel near 3.3 TB/s:
                                    // the copy engine is request-rate bound); 1024 when shared memory is short (long rows, many tokens)
constexpr int kQRowPad = 16;       // bytes added to a stage row: keeps ldmatrix rows on different banks

struct GemvQGeom {
    int piece_bytes;  // bytes of a row per ring stage: 2048 or 1024; each of the 8 warps of a group owns 1/8 of them
    int stages;
    int pieces;       // stages per unit = ceil(K / kQPieceK)
    int row_stride;   // bytes between rows inside a stage
    int stage_bytes;
    int xs_stride;    // halves between token rows of the staged activations
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// (a & b) | c in one instruction (the compiler emits two LOP3 when b and c are immediates)
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// The 8 nibbles of a 32-bit word as four f16x2 pairs {n_i - z, n_{i+4} - z}, i = 0..3, exact integers:
//   i even: ((w' & 0x000f000f) | 0x64006400) = {1024 + n, 1024 + n} minus {1024 + z};
//   i odd : ((w' & 0x00f000f0) | 0x64006400) = {1024 + 16 n, ...}: one HFMA2 with 1/16 and -(64 + z) (exact: 64 + n is an f16 integer);
//   w' = w for i = 0, 1 and w >> 8 for i = 2, 3: one shift per word instead of three.
struct Int4Consts {
    uint32_t m_lo, m_hi, magic, sixteenth;  // 0x000f000f, 0x00f000f0, 0x64006400, f16x2 {1/16, 1/16}
};
__device__ __forceinline__ void deq_int4_word(uint32_t w, uint32_t zpk, uint32_t nz64, const Int4Consts &c, uint32_t (&d)[4]) {
    const uint32_t w8 = w >> 8;
    const uint32_t h0 = lop3_and_or(w, c.m_lo, c.magic), h1 = lop3_and_or(w, c.m_hi, c.magic);
    const uint32_t h2 = lop3_and_or(w8, c.m_lo, c.magic), h3 = lop3_and_or(w8, c.m_hi, c.magic);
    const __half2 z = *reinterpret_cast<const __half2 *>(&zpk), nz = *reinterpret_cast<const __half2 *>(&nz64);
    const __half2 s16 = *reinterpret_cast<const __half2 *>(&c.sixteenth);
    const __half2 r0 = __hsub2(*reinterpret_cast<const __half2 *>(&h0), z), r2 = __hsub2(*reinterpret_cast<const __half2 *>(&h2), z);
    const __half2 r1 = __hfma2(*reinterpret_cast<const __half2 *>(&h1), s16, nz), r3 = __hfma2(*reinterpret_cast<const __half2 *>(&h3), s16, nz);
    d[0] = *reinterpret_cast<const uint32_t *>(&r0), d[1] = *reinterpret_cast<const uint32_t *>(&r1);
    d[2] = *reinterpret_cast<const uint32_t *>(&r2), d[3] = *reinterpret_cast<const uint32_t *>(&r3);
}
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint32_t two_bytes) {
    uint32_t h;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h) : "h"((unsigned short)two_bytes));
    return h;
}

// smem: [ xs : M * xs_stride halves | rings : groups * stages * stage_bytes | barriers | partials : groups * 2 * GW * 16*8 floats ]
template <typename T, int FMT, bool kSwiGLU>
__global__ void __launch_bounds__(kGemvThreads, 1)
gemv_q_kernel(const GemvArgs a, const GemvQGeom geo) {
    static_assert(FMT == WF_FP8 || FMT == WF_INT4, "quantised formats only");
    constexpr int GW = kGemvGW, R = kQRows;
    constexpr int V = Elem<T>::kVec;
    constexpr int kBytesPerK8 = FMT == WF_FP8 ? 8 : 4;            // bytes of 8 consecutive k of one row
    constexpr int kBlkBytes = 128 * kBytesPerK8 / 8;               // bytes of one 128-k block of a row: 128 / 64
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float red[33];

    const int K = a.K, N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row_bytes = FMT == WF_FP8 ? (size_t)K : (size_t)K / 2;
    const int units = kSwiGLU ? (a.inter + 7) / 8 : (N + R - 1) / R;
    const int stages = geo.stages, pieces = geo.pieces;
    const int n_threads = (int)blockDim.x;
    const int kQPieceK = geo.piece_bytes * 8 / kBytesPerK8;   // k per stage
    const int kQWarpK = kQPieceK / kGemvGW;                   // k per warp per stage (a multiple of 128)
    const int kWarpBytes = geo.piece_bytes / kGemvGW;         // bytes of a row a warp owns per stage

    const bool is_compute = warp < kGemvWarps;
    const int grp = is_compute ? warp / GW : (warp - kGemvWarps) % kGemvGroups;
    const bool is_producer = !is_compute && warp < kGemvWarps + kGemvGroups;
    const int wg = warp % GW;
    const int gid = grp * gridDim.x + blockIdx.x, total_groups = gridDim.x * kGemvGroups;
    const int my_units = gid < units ? (units - gid + total_groups - 1) / total_groups : 0;
    const int my_items = my_units * pieces;

    __half *xs = reinterpret_cast<__half *>(smem);
    size_t off = ((size_t)a.M * geo.xs_stride * sizeof(__half) + 127) & ~(size_t)127;  // token slots >= M read as zero
    unsigned char *ring = smem + off + (size_t)grp * stages * geo.stage_bytes;
    off += (size_t)kGemvGroups * stages * geo.stage_bytes;
    const uint32_t full0 = smem_u32(smem + off) + grp * (2 * kGemvMaxStages + 4) * 8, empty0 = full0 + kGemvMaxStages * 8;
    const uint32_t ready0 = empty0 + kGemvMaxStages * 8, free0 = ready0 + 16;
    off += (size_t)kGemvGroups * (2 * kGemvMaxStages + 4) * 8;
    float *gred = reinterpret_cast<float *>(smem + off) + (size_t)grp * 2 * GW * (R * kQTok);  // [parity][warp][16 rows][8 tokens]
    const uint32_t ring_u32 = smem_u32(ring);

    // row r (0..15) of unit u: plain: 16u + r; SwiGLU: gate rows 8u + r (r < 8), up rows inter + 8u + (r - 8)
    auto unit_row = [&](int u, int r) -> int { return kSwiGLU ? (r < 8 ? 8 * u + r : a.inter + 8 * u + (r - 8)) : R * u + r; };
    auto row_ok = [&](int u, int r) -> bool { return kSwiGLU ? (8 * u + (r & 7) < a.inter) : (R * u + r < N); };

    // ---- producer WARP: lane 0 arms the stage barrier, then lanes 0..15 each issue the bulk copy of one row (a single thread
    //      issuing 16 small copies per stage cannot keep up with HBM)
    int p_un = 0, p_pc = 0, p_item = 0, p_s = 0;
    auto issue_next = [&]() {  // called by all 32 lanes of the producer warp
        const int u = gid + p_un * total_groups;
        const int k0 = p_pc * kQPieceK;
        const uint32_t bytes = (uint32_t)(min(kQPieceK, K - k0) * kBytesPerK8 / 8);
        const uint32_t bar = full0 + p_s * 8;
        const bool mine = lane < R && row_ok(u, lane);
        const unsigned nrows = __popc(__ballot_sync(0xffffffffu, mine));
        if (lane == 0) mbar_expect_tx(bar, bytes * nrows);
        __syncwarp();
        if (mine)
            bulk_g2s(ring_u32 + p_s * geo.stage_bytes + lane * geo.row_stride,
                     reinterpret_cast<const unsigned char *>(a.w) + (size_t)unit_row(u, lane) * row_bytes + (size_t)k0 * kBytesPerK8 / 8, bytes, bar);
        ++p_item;
        if (++p_pc == pieces) p_pc = 0, ++p_un;
        if (++p_s == stages) p_s = 0;
    };
    if (is_producer) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(full0 + s * 8, 1);
                mbar_init(empty0 + s * 8, GW);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(ready0 + b * 8, GW);
                mbar_init(free0 + b * 8, 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        while (p_item < stages && p_item < my_items) issue_next();  // the weights do not depend on the previous kernel
    }

    pdl_wait();

    // ---------------- stage the activations as f16 (a.M token rows); INT4 stores every 8 k in the order
    //                  [k0 k4 k1 k5 k2 k6 k3 k7] -- the order in which the nibble pairs come out of a 32-bit word
    {
        auto store = [&](int m, int i, const float *f) {
            float g[V];
            unpack16<T>(pack16<T>(f), g);  // the un-fused reference hands the GEMM a tensor of T
            // one 16-byte store per vector of 8 k (V == 8 for the 16-bit types this kernel is instantiated for)
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int dstj = FMT == WF_FP8 ? j : ((j & 3) * 2 + (j >> 2));
                o[dstj] = g[j];
            }
            const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
            const __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
            uint4 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&h0), pk.y = *reinterpret_cast<const uint32_t *>(&h1);
            pk.z = *reinterpret_cast<const uint32_t *>(&h2), pk.w = *reinterpret_cast<const uint32_t *>(&h3);
            *reinterpret_cast<uint4 *>(xs + (size_t)m * geo.xs_stride + (size_t)i * V) = pk;
        };
        static_assert(V == 8, "gemv_q_kernel: 16-bit activation types only");
        if (a.M == 1) gemv_stage_activations<T, 1>(a, n_threads, red, store);  // single-trip instantiation: no spills in the B = 1 prologue
        else gemv_stage_activations<T, 0>(a, n_threads, red, store);
        __syncthreads();
    }
    pdl_launch_dependents();

    if (is_producer) {
        int e_s = 0, e_ph = 0;
        while (p_item < my_items) {
            mbar_wait(empty0 + e_s * 8, e_ph);
            if (++e_s == stages) e_s = 0, e_ph ^= 1;
            fence_proxy_async();
            issue_next();
        }
    } else if (!is_compute) {
        // ================================================= reducer: lane (g, t) finishes rows g, g+8 x tokens 2t, 2t+1
        const int g = lane >> 2, t = lane & 3;
        const unsigned int push_flag = a.push.n > 0 ? tp_flag(a.push.epoch, a.push.seq) : 0u;
        for (int un = 0; un < my_units; ++un) {
            const int u = gid + un * total_groups;
            const int b = un & 1;
            mbar_wait(ready0 + b * 8, (un >> 1) & 1);
            const float *slot = gred + (size_t)b * GW * (R * kQTok);
            float o[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [row g / g+8][token 2t / 2t+1]
#pragma unroll
            for (int w2 = 0; w2 < GW; ++w2) {  // fixed order: deterministic
                const float2 lo = *reinterpret_cast<const float2 *>(slot + (size_t)w2 * (R * kQTok) + g * kQTok + 2 * t);
                const float2 hi = *reinterpret_cast<const float2 *>(slot + (size_t)w2 * (R * kQTok) + (g + 8) * kQTok + 2 * t);
                o[0][0] += lo.x, o[0][1] += lo.y, o[1][0] += hi.x, o[1][1] += hi.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(free0 + b * 8);
            if constexpr (FMT == WF_FP8) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (row_ok(u, g + 8 * h)) {
                        const float s8 = __ldg(reinterpret_cast<const float *>(a.scales) + unit_row(u, g + 8 * h));
                        o[h][0] *= s8, o[h][1] *= s8;
                    }
            }
            if (!kSwiGLU && a.push.n > 0) {
                // tensor-parallel partial: LL words hold two consecutive rows.  Lanes g and g ^ 1 (lane ^ 4) swap one value each: the even
                // lane ends up with rows (g, g + 1), the odd one with rows (g + 7, g + 8); N is a multiple of 16 rows per unit here
                const bool odd = g & 1;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float give = odd ? o[0][e] : o[1][e];
                    const float got = __shfl_xor_sync(0xffffffffu, give, 4);
                    const int m = 2 * t + e;
                    if (m < a.M) {
                        const int r0 = odd ? g + 7 : g;  // first row of the pair inside the unit
                        if (row_ok(u, r0 + 1))
                            tp_push_pair<T>(a.push, push_flag, (size_t)m * N + unit_row(u, r0), odd ? got : o[0][e], odd ? o[1][e] : got);
                    }
                }
                continue;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = 2 * t + e;
                if (m >= a.M) continue;
                if constexpr (kSwiGLU) {
                    if (row_ok(u, g)) {
                        // the un-fused reference stores gate/up in T before SiLU reads them
                        const float gt = round_to<T>(o[0][e]), up = round_to<T>(o[1][e]);
                        const float v = (gt / (1.0f + expf(-gt))) * up;
                        const size_t idx = (size_t)m * a.inter + 8 * u + g;
                        if (a.y_f32) reinterpret_cast<float *>(a.y)[idx] = v;
                        else reinterpret_cast<T *>(a.y)[idx] = Elem<T>::from_f(v);
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (row_ok(u, g + 8 * h)) {
                            const size_t idx = (size_t)m * N + unit_row(u, g + 8 * h);
                            if (a.y_f32) {
                                reinterpret_cast<float *>(a.y)[idx] = o[h][e];
                            } else {
                                reinterpret_cast<T *>(a.y)[idx] = Elem<T>::from_f(o[h][e]);
                            }
                        }
                }
            }
        }
    } else {
        // ================================================= compute warps: this warp's 128 k of every piece, 16 rows x 8 tokens
        const int g = lane >> 2, t = lane & 3;
        // ldmatrix.x4 row addresses: lanes 0-7 rows 0-7 (first 16 bytes), 8-15 rows 8-15, 16-23 rows 0-7 (+16 bytes), 24-31 rows 8-15 (+16)
        const uint32_t lm_off = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * geo.row_stride + (uint32_t)(lane >> 4) * 16 + (uint32_t)wg * kWarpBytes;
        // B fragments: token g; slots past the batch read token 0 (their accumulator columns are never reduced)
        const __half *xrow = xs + (size_t)(g < a.M ? g : 0) * geo.xs_stride;
        const int ngroups_k = FMT == WF_INT4 ? K / a.group : 0;
        float acc[4];
        Int4Consts i4c;
        i4c.m_lo = 0x000f000fu, i4c.m_hi = 0x00f000f0u, i4c.magic = 0x64006400u, i4c.sixteenth = 0x2c002c00u;  // f16 1/16 = 0x2c00
        (void)i4c;
        int s = 0, ph = 0;
        for (int un = 0; un < my_units; ++un) {
            const int u = gid + un * total_groups;
            acc[0] = acc[1] = acc[2] = acc[3] = 0.0f;
            for (int pc = 0; pc < pieces; ++pc) {
                const int kw = pc * kQPieceK + wg * kQWarpK;  // first k of this warp's slice
                // INT4: the group scales / zero points of this warp's (<= 4) consecutive 128-k blocks, for rows g and g+8, requested BEFORE
                // waiting for the stage and kept packed (4 x T in a uint2, 4 x u8 in a uint32): fetched block by block inside the loop
                // they put a ~600-cycle dependent L2 load in front of every block, which bounded the INT4 kernel near 1.8 TB/s
                uint2 sraw[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
                uint32_t zraw[2] = {0u, 0u};
                if constexpr (FMT == WF_INT4) {
                    const int nblk = min(kQWarpK, max(K - kw, 0)) / 128;  // blocks of this slice that exist
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int row = min(unit_row(u, g + 8 * h), N - 1);
                        const size_t gi = (size_t)row * ngroups_k + kw / a.group;
                        const T *sp = reinterpret_cast<const T *>(a.scales) + gi;
                        const uint8_t *zp = reinterpret_cast<const uint8_t *>(a.zeros) + gi;
                        if (nblk == 4 && (ngroups_k & 3) == 0) {  // 8-byte / 4-byte aligned: one load each
                            sraw[h] = __ldg(reinterpret_cast<const uint2 *>(sp));
                            zraw[h] = __ldg(reinterpret_cast<const uint32_t *>(zp));
                        } else {
                            uint32_t s16[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                            for (int bk = 0; bk < 4; ++bk)
                                if (bk < nblk) {
                                    s16[bk] = __ldg(reinterpret_cast<const unsigned short *>(sp) + bk);
                                    zraw[h] |= (uint32_t)__ldg(zp + bk) << (8 * bk);
                                }
                            sraw[h] = make_uint2(s16[0] | (s16[1] << 16), s16[2] | (s16[3] << 16));
                        }
                    }
                }
                mbar_wait(full0 + s * 8, ph);
                const uint32_t st = ring_u32 + (uint32_t)s * geo.stage_bytes + lm_off;
                for (int bk = 0; bk < kQWarpK / 128; ++bk) {  // 128-k blocks of the slice (K is a multiple of 128)
                    const int kb = kw + bk * 128;
                    if (kb >= K) break;
                    if constexpr (FMT == WF_FP8) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {  // 32 k per ldmatrix.x4 = two mma k-steps
                            uint32_t r[4];
                            ldmatrix_x4(r, st + bk * kBlkBytes + j * 32);
                            // r0: row g, fp8 k = 4t..4t+3 of the first 16 k; r1: row g+8; r2 / r3: the next 16 k.
                            // k is permuted consistently for A and B: mma slots (2t, 2t+1) <- k (4t, 4t+1), slots (2t+8, 2t+9) <- (4t+2, 4t+3)
                            uint2 b01 = *reinterpret_cast<const uint2 *>(xrow + kb + j * 32 + 4 * t);
                            uint2 b23 = *reinterpret_cast<const uint2 *>(xrow + kb + j * 32 + 16 + 4 * t);
                            mma_f16(acc, e4m3x2_to_f16x2(r[0]), e4m3x2_to_f16x2(r[1]), e4m3x2_to_f16x2(r[0] >> 16), e4m3x2_to_f16x2(r[1] >> 16), b01.x, b01.y);
                            mma_f16(acc, e4m3x2_to_f16x2(r[2]), e4m3x2_to_f16x2(r[3]), e4m3x2_to_f16x2(r[2] >> 16), e4m3x2_to_f16x2(r[3] >> 16), b23.x, b23.y);
                        }
                    } else {
                        // a 128-k block is exactly one quantisation group (group == 128) of every row
                        float sc[2];
                        uint32_t zpk[2], nz64[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const uint32_t pair = bk < 2 ? sraw[h].x : sraw[h].y;
                            const unsigned short s16 = (unsigned short)((bk & 1) ? (pair >> 16) : (pair & 0xffffu));
                            sc[h] = Elem<T>::to_f(*reinterpret_cast<const T *>(&s16));
                            const uint32_t zq = (zraw[h] >> (8 * bk)) & 0xffu;
                            const uint32_t z = 0x6400u | zq;
                            zpk[h] = z | (z << 16);  // f16x2 {1024 + z, 1024 + z}
                            const __half2 n2 = __float2half2_rn(-(64.0f + (float)zq));
                            nz64[h] = *reinterpret_cast<const uint32_t *>(&n2);  // f16x2 {-(64 + z), -(64 + z)}
                        }
                        float ag[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < 2; ++j) {  // 64 k per ldmatrix.x4 = four mma k-steps
                            uint32_t r[4];
                            ldmatrix_x4(r, st + bk * kBlkBytes + j * 32);
                            // r0: row g, nibbles k = 8t..8t+7 of the first 32 k; r1: row g+8; r2 / r3: the next 32 k
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                uint32_t d0[4], d1[4];
                                deq_int4_word(r[2 * q], zpk[0], nz64[0], i4c, d0);
                                deq_int4_word(r[2 * q + 1], zpk[1], nz64[1], i4c, d1);
                                // token slots past the batch read token 0's activations: their accumulator columns are never reduced
                                const uint4 xb = *reinterpret_cast<const uint4 *>(xrow + kb + j * 64 + q * 32 + 8 * t);  // (k0,k4)(k1,k5)(k2,k6)(k3,k7)
                                mma_f16(ag, d0[0], d1[0], d0[1], d1[1], xb.x, xb.y);
                                mma_f16(ag, d0[2], d1[2], d0[3], d1[3], xb.z, xb.w);
                            }
                        }
                        acc[0] = fmaf(sc[0], ag[0], acc[0]), acc[1] = fmaf(sc[0], ag[1], acc[1]);
                        acc[2] = fmaf(sc[1], ag[2], acc[2]), acc[3] = fmaf(sc[1], ag[3], acc[3]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + s * 8);
                if (++s == stages) s = 0, ph ^= 1;
            }
            // ---- hand the 16 x 8 partial tile to the reducer (double-buffered slot); rows never copied are discarded
            const int b = un & 1;
            if (un >= 2) mbar_wait(free0 + b * 8, ((un >> 1) - 1) & 1);
            float *slot = gred + (size_t)b * GW * (R * kQTok) + (size_t)wg * (R * kQTok);
            const bool ok0 = row_ok(u, g), ok1 = row_ok(u, g + 8);
            *reinterpret_cast<float2 *>(slot + g * kQTok + 2 * t) = ok0 ? make_float2(acc[0], acc[1]) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(slot + (g + 8) * kQTok + 2 * t) = ok1 ? make_float2(acc[2], acc[3]) : make_float2(0.f, 0.f);
            __syncwarp();
            if (lane == 0) mbar_arrive(ready0 + b * 8);
        }
    }
}

// host side ----------------------------------------------------------------------------------------------------------
template <typename T, int FMT, bool SW>
static int launch_gemv_q_inst(const GemvArgs &a, cudaStream_t st) {
    GemvQGeom g;
    const int bytes_per_k8 = FMT == WF_FP8 ? 8 : 4;
    g.xs_stride = a.K + (FMT == WF_FP8 ? 16 : 32);  // token rows shifted by 8 (64-bit B loads) / 16 (128-bit) banks: conflict-free
    size_t fixed = ((size_t)a.M * g.xs_stride * sizeof(__half) + 127) & ~(size_t)127;
    fixed += (size_t)kGemvGroups * (2 * kGemvMaxStages + 4) * 8;
    fixed += (size_t)kGemvWarps * 2 * kQRows * kQTok * sizeof(float);
    const size_t budget = 224 * 1024;
    size_t per_stage = 0;
    for (g.piece_bytes = kQPieceBytes; g.piece_bytes >= 1024; g.piece_bytes /= 2) {  // 2 KiB pieces if >= 3 stages fit, else 1 KiB
        g.row_stride = g.piece_bytes + kQRowPad;
        g.stage_bytes = kQRows * g.row_stride;
        per_stage = (size_t)kGemvGroups * g.stage_bytes;
        if (fixed + 3 * per_stage <= budget) break;
    }
    if (g.piece_bytes < 1024) return B200_ERR_UNSUPPORTED;
    const int piece_k = g.piece_bytes * 8 / bytes_per_k8;
    g.pieces = (a.K + piece_k - 1) / piece_k;
    g.stages = (int)((budget - fixed) / per_stage);
    if (g.stages > kGemvMaxStages) g.stages = kGemvMaxStages;
    const size_t smem = fixed + (size_t)g.stages * per_stage;
    auto kern = gemv_q_kernel<T, FMT, SW>;
    static thread_local size_t cached_smem[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cached_smem[dev] < smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return cuda_status("gemv_q cudaFuncSetAttribute");
        cached_smem[dev] = smem;
    }
    const int units = SW ? (a.inter + 7) / 8 : (a.N + kQRows - 1) / kQRows;
    int grid = sm_count();
    const int need = (units + kGemvGroups - 1) / kGemvGroups;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    launch_pdl(kern, dim3(grid), dim3(kGemvThreads), smem, st, true, a, g);
    return cuda_status("gemv_q launch");
}

// Returns B200_ERR_UNSUPPORTED (no error text) when the shape cannot use the tensor-core quantised kernel.
template <typename T>
static int launch_gemv_q_t(const GemvArgs &a, int fmt, bool swiglu, cudaStream_t st) {
    if (a.M < 1 || a.M > kQTok || a.K % 128 != 0 || !aligned16(a.w) || !aligned16(a.x)) return B200_ERR_UNSUPPORTED;
    if (a.norm && ((a.res_in && !aligned16(a.res_in)) || (a.res_out && !aligned16(a.res_out)) || (a.bias && !aligned16(a.bias)) ||
                   (a.gamma && !aligned16(a.gamma))))
        return B200_ERR_UNSUPPORTED;
    if (fmt == WF_INT4 && a.group != 128) return B200_ERR_UNSUPPORTED;
    if (swiglu && a.inter % 8 != 0) return B200_ERR_UNSUPPORTED;
    if (fmt == WF_FP8) return swiglu ? launch_gemv_q_inst<T, WF_FP8, true>(a, st) : launch_gemv_q_inst<T, WF_FP8, false>(a, st);
    if (fmt == WF_INT4) return swiglu ? launch_gemv_q_inst<T, WF_INT4, true>(a, st) : launch_gemv_q_inst<T, WF_INT4, false>(a, st);
    return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
