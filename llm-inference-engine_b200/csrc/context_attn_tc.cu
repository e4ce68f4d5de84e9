// context_attn_tc.cu -- prefill ("context") attention on tcgen05 / TMEM for sm_100a, head_size 128, 16-bit activations.
//
// Semantics: the reference chain src/layers/context_attention.cpp:221-289 with a true QK^T (launchRepeatKVCache ->
// launchLinearStridedBatchGemm -> launchFusedScaleMaskAndSoftmax -> launchLinearStridedBatchGemm ->
// launchFusedTransposeAndRemovePadding), i.e. per (batch b, head h, query row i):
//   s_k = scale * q_i.k_k for keys k < context_len[b] with k <= i + context_len[b] - input_len[b]   (others carry the additive
//   -10000 of the mask and vanish: expf(-10000 + ...) == 0 in fp32), p = exp(s - max(max_k s, FLT_MIN)) / (sum + 1e-6), out = p.V
// without materialising [B,H,Sq,Sk] scores or the GQA-repeated K/V, output already un-padded [T, H, d].
//
// One CTA = 128 query rows of one (b, h).  576 threads, warp-specialised:
//   warp 16  TMA producer: Q tile once, then K and V tiles of 128 keys (cp.async.bulk.tensor.2d, 128-byte swizzle); K and V have their
//            own 2-stage rings and barriers, requested in the order the MMAs consume them (K0 K1 V0 K2 V1 ...): a K slot is free as
//            soon as ITS S = Q.K^T has been computed, a whole tile step before the V slot of the same tile
//   warp 17  MMA issuer:  S = Q.K^T (128x128x128, tcgen05.mma kind::f16, fp32 in TMEM, two S buffers so the logits of tile j+1 are
//            computed while the softmax threads work on tile j) and O += P_j.V_j (V is the MN-major B operand); the output
//            accumulator STAYS in TMEM for the whole key loop
//   warps 0-15 softmax + epilogue: warp w owns TMEM lanes 32*(w%4).. (query rows) and columns 32*(w/4).. : four threads share a
//            row, 32 keys / 32 output dims each (16 warps keep the SM's four schedulers busy; with one thread per row the softmax
//            ran at one warp per scheduler and took 6x longer than the MMAs).  tcgen05.ld the logits once into registers, mask
//            (diagonal / last tile only), exchange the row max through shared memory (named barrier of the row's 4 warps),
//            p = ex2(s - m) -> shared memory (bf16/fp16, the 128-byte-swizzled K-major layout the second MMA reads).
//            LAZY RESCALE: m is the max the row's accumulator was last scaled to, not the running max; only when the running max
//            has grown by more than 2^8 is the accumulator pulled out of TMEM, scaled and put back (tcgen05.ld / .st) -- after the
//            first tiles that almost never happens, so no thread waits for a P.V product inside the loop (round 1 read every
//            tile's product back into registers: the softmax threads sat behind the MMA they had just fed, 4.5 us per tile step).
//            The final output is O * c / (l * c + 1e-6) with c = 2^(m - running max): the reference's +1e-6 is relative to the
//            sum taken at the true max.
// Causal structure: q tiles visit only the key tiles up to their diagonal; heavy (late) q tiles are scheduled first.
#include "common.cuh"

#include <cuda.h>
#include <float.h>
#include <mutex>

namespace b200 {

namespace catc {

constexpr int kRows = 128;   // query rows per CTA = UMMA M
constexpr int kKeys = 128;   // keys per tile = UMMA N of the first MMA, K extent of the second
constexpr int kD = 128;      // head size
constexpr int kSoftmaxWarps = 16;
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;  // 4 threads per query row
constexpr int kThreads = kSoftmaxThreads + 64;      // + producer warp + MMA warp
constexpr int kHalfBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 elements (128 B)
constexpr int kTileBytes = 2 * kHalfBytes;
constexpr int kKvStages = 2;

struct Params {
    void *out;                  // [T, H, d]
    const int *seq_off;         // [B] padded-slot minus token index of each sequence
    const int *input_len, *context_len;
    int head_num, kv_head_num, max_q_len, max_seq_len;
    float scale;
    int is_bf16;
    // paged cache (optional): K / V are page pools [num_pages, Hkv, 64, d] (tensor maps with boxes of 64 rows); the keys [64 i, 64 i + 64)
    // of batch row b live in page block_table[b * max_pages + i].  A 128-key tile is then two pages = four boxes.
    const int *block_table;
    int max_pages;
};

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// Shared-memory matrix descriptors (sm_100 format: version 1 at bit 46, layout type in bits 61..63, 2 = SWIZZLE_128B).
// K-major operand (rows of 128 bytes along K, 8-row atoms of 1024 bytes): SBO = 1024, LBO unused (1).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major operand (V: rows = keys (the MMA K dimension), 128 bytes = 64 head dims contiguous along N): 8 keys = one 1024-byte
// group (SBO), the next 64 head dims live in the second TMA box, kHalfBytes further (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(kHalfBytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_f16(bool bf16, int m, int n, bool b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                   // D: F32
    d |= (bf16 ? 1u : 0u) << 7;     // A format
    d |= (bf16 ? 1u : 0u) << 10;    // B format
    d |= (b_mn_major ? 1u : 0u) << 16;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;
}

// 2^x for x <= ~8 on the FMA / ALU pipes (no MUFU): n = round(x) by the 1.5 * 2^23 trick, 2^(x - n) by a degree-4 polynomial on [-0.5, 0.5]
// (max relative error 3.7e-6, far inside the 2^-9 rounding of the 16-bit probabilities), 2^n added into the exponent field.  One in four
// exponentials of the softmax goes this way: the loop's ex2 burst is bound by the XU pipe (16 results per clock per SM) while the FMA pipe
// idles -- the split FlashAttention-4 uses.  x = -inf (masked key) gives ~1e-38: below every sum it is added to.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(0.009560510516f, f, 0.05591703951f);
    p = fmaf(p, f, 0.2402498126f);
    p = fmaf(p, f, 0.6931219697f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef B200_CTX_POLY_EVERY
#define B200_CTX_POLY_EVERY 0
#endif
constexpr int kPolyEvery = B200_CTX_POLY_EVERY;  // every n-th exponential on the polynomial (0: all on MUFU)
__host__ __device__ constexpr bool poly_lane(int u) { return kPolyEvery > 0 && u % (kPolyEvery > 0 ? kPolyEvery : 1) == kPolyEvery - 1; }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleStep = 8.0f;  // log2 domain: the accumulator is rescaled when the running max has grown by more than 2^8

// smem: [Q 32K][K ring 2x32K][V ring 2x32K][P 32K][barriers][tmem slot][row-max / row-sum exchange 2x128x4 floats]
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
context_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sQ = smem, *sK = sQ + kTileBytes, *sV = sK + kKvStages * kTileBytes, *sP = sV + kKvStages * kTileBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + kTileBytes);
    // barriers: q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_free[2], p_ready, pv_done
    const uint32_t q_full = s_u32(bars), k_full0 = q_full + 8, k_empty0 = k_full0 + 16, v_full0 = k_empty0 + 16, v_empty0 = v_full0 + 16,
                   s_full0 = v_empty0 + 16, s_free0 = s_full0 + 16, p_ready = s_free0 + 16, pv_done = p_ready + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    float *red = reinterpret_cast<float *>(bars + 18);  // [2][128][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = gridDim.x - 1 - blockIdx.x;  // heavy (late) tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * kRows;

    if (warp == kSoftmaxWarps && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
        bar_init(q_full, 1);
        for (int s = 0; s < kKvStages; ++s) {
            bar_init(k_full0 + 8 * s, 1);
            bar_init(k_empty0 + 8 * s, 1);   // released by tcgen05.commit of the Q.K^T MMA
            bar_init(v_full0 + 8 * s, 1);
            bar_init(v_empty0 + 8 * s, 1);   // released by tcgen05.commit of the P.V MMA
            bar_init(s_full0 + 8 * s, 1);
            bar_init(s_free0 + 8 * s, kSoftmaxWarps);     // all softmax warps have read S[s] (one arrival per warp)
        }
        bar_init(p_ready, kSoftmaxWarps);    // one arrival per softmax warp: 512 arrivals on one mbarrier serialise (measured)
        bar_init(pv_done, 1);                // tcgen05.commit of the P.V MMA: P may be overwritten, O holds tiles 0..j
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kSoftmaxWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // one K or V tile of 128 keys: two boxes (the two 64-column halves) of 128 rows from the contiguous cache, or -- paged -- per half
    // one box of 64 rows from each of the tile's two pages (a page past the context is not requested: its rows are masked / zeroed)
    auto load_kv_tile = [&](const CUtensorMap &tm, unsigned char *dst_tile, uint32_t bar, int j, int kv_row0, int kvh, int klen) {
        const uint32_t d0 = s_u32(dst_tile);
        if (p.block_table) {
            const int *bt = p.block_table + (size_t)b * p.max_pages;
            const int npg = (j * kKeys + 64 < klen) ? 2 : 1;
            bar_expect_tx(bar, (uint32_t)npg * (kTileBytes / 2));
            for (int g = 0; g < npg; ++g) {
                const int row = (__ldg(bt + 2 * j + g) * p.kv_head_num + kvh) * 64;
                tma_load_2d(d0 + g * 64 * 128, &tm, 0, row, bar);
                tma_load_2d(d0 + kHalfBytes + g * 64 * 128, &tm, 64, row, bar);
            }
        } else {
            bar_expect_tx(bar, kTileBytes);
            tma_load_2d(d0, &tm, 0, kv_row0 + j * kKeys, bar);
            tma_load_2d(d0 + kHalfBytes, &tm, 64, kv_row0 + j * kKeys, bar);
        }
    };
    // ---- the producer does not wait for the TMEM allocation: Q and the first two K tiles are requested before the CTA-wide barrier
    //      (its own thread initialised the mbarriers), so their latency overlaps the rest of the prologue
    int ntiles_early = 0;
    if (warp == kSoftmaxWarps && lane == 0) {
        pdl_wait();  // q and the cache rows come from the kernel in front of this one
        const int qlen = p.input_len[b], klen = p.context_len[b];
        const int q_hi = min(q0 + kRows, qlen) - 1;
        const int k_hi = q0 < qlen ? min(klen, q_hi + (klen - qlen) + 1) : 0;
        ntiles_early = (k_hi + kKeys - 1) / kKeys;
        if (ntiles_early > 0) {
            const int kvh = h / (p.head_num / p.kv_head_num);
            const int q_row0 = (b * p.head_num + h) * p.max_q_len + q0, kv_row0 = (b * p.kv_head_num + kvh) * p.max_seq_len;
            bar_expect_tx(q_full, kTileBytes);
            tma_load_2d(s_u32(sQ), &tmQ, 0, q_row0, q_full);
            tma_load_2d(s_u32(sQ) + kHalfBytes, &tmQ, 64, q_row0, q_full);
            for (int j = 0; j < min(ntiles_early, kKvStages); ++j) load_kv_tile(tmK, sK + j * kTileBytes, k_full0 + 8 * j, j, kv_row0, kvh, klen);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;  // columns: S0 [0,128) S1 [128,256) O [256,384)

    pdl_wait();
    pdl_launch_dependents();

    const int qlen = p.input_len[b], klen = p.context_len[b];
    const bool active = q0 < qlen;  // CTA-uniform
    // last key any row of this tile may see (exclusive)
    const int q_hi = min(q0 + kRows, qlen) - 1;
    const int k_hi = active ? min(klen, q_hi + (klen - qlen) + 1) : 0;
    const int ntiles = (k_hi + kKeys - 1) / kKeys;
    const int kvh = h / (p.head_num / p.kv_head_num);
    const int kv_row0 = (b * p.kv_head_num + kvh) * p.max_seq_len;

    if (warp == kSoftmaxWarps) {
        // ================================================= TMA producer (K0, K1 are on their way): V0 | K2 V1 | K3 V2 | ... | V(n-1)
        if (lane == 0) {
            for (int j = 1; j <= ntiles; ++j) {
                if (j >= kKvStages && j < ntiles) {
                    const int s = j % kKvStages;
                    bar_wait(k_empty0 + 8 * s, ((j / kKvStages) & 1) ^ 1);
                    load_kv_tile(tmK, sK + s * kTileBytes, k_full0 + 8 * s, j, kv_row0, kvh, klen);
                }
                const int jv = j - 1, s = jv % kKvStages;
                bar_wait(v_empty0 + 8 * s, ((jv / kKvStages) & 1) ^ 1);
                load_kv_tile(tmV, sV + s * kTileBytes, v_full0 + 8 * s, jv, kv_row0, kvh, klen);
            }
        }
    } else if (warp == kSoftmaxWarps + 1) {
        // ================================================= MMA issuer
        if (ntiles > 0) {
            const uint32_t id_qk = idesc_f16(p.is_bf16 != 0, kRows, kKeys, false);
            const uint32_t id_pv = idesc_f16(p.is_bf16 != 0, kRows, kD, true);
            auto issue_qk = [&](int j) {  // S[j % 2] = Q . K_j^T
                const int s = j % kKvStages, sb = j & 1;
                bar_wait(k_full0 + 8 * s, (j / kKvStages) & 1);
                bar_wait(s_free0 + 8 * sb, ((j >> 1) & 1) ^ 1);  // the softmax threads are done with what S[sb] held (tile j - 2)
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t aq = s_u32(sQ), bk = s_u32(sK + s * kTileBytes);
#pragma unroll
                    for (int k = 0; k < kD / 16; ++k) {
                        const uint32_t off = (uint32_t)(k / 4) * kHalfBytes + (uint32_t)(k % 4) * 32;
                        tc_mma_f16(tmem + (uint32_t)(sb * kKeys), desc_kmajor(aq + off), desc_kmajor(bk + off), id_qk, k > 0 ? 1u : 0u);
                    }
                    tc_commit(s_full0 + 8 * sb);
                    tc_commit(k_empty0 + 8 * s);  // K_j may be overwritten
                }
                __syncwarp();
            };
            bar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < ntiles; ++j) {
                if (j + 1 < ntiles) issue_qk(j + 1);  // next tile's logits while the softmax threads work on tile j
                const int s = j % kKvStages;
                bar_wait(v_full0 + 8 * s, (j / kKvStages) & 1);
                bar_wait(p_ready, j & 1);                       // P_j is in shared memory, O has been rescaled if it had to be
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t ap = s_u32(sP), bv = s_u32(sV + s * kTileBytes);
#pragma unroll
                    for (int k = 0; k < kKeys / 16; ++k) {
                        const uint32_t aoff = (uint32_t)(k / 4) * kHalfBytes + (uint32_t)(k % 4) * 32;  // P: K-major over keys
                        const uint32_t boff = (uint32_t)k * 16 * 128;                                    // V: 16 keys = 2048 bytes further
                        tc_mma_f16(tmem + 2 * kKeys, desc_kmajor(ap + aoff), desc_mnmajor(bv + boff), id_pv, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(v_empty0 + 8 * s);  // V_j may be overwritten
                    tc_commit(pv_done);           // P_j may be overwritten; O holds tiles 0..j
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================= softmax + epilogue: 4 threads per query row, 32 columns each
        const int quarter = warp & 3, cg = warp >> 2;
        const int row = quarter * 32 + lane;   // == TMEM lane
        const int qi = q0 + row;
        const int lim = qi + (klen - qlen);    // keys <= lim are visible to this row
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cg * 32);
        const uint32_t o_addr = lane_addr + (uint32_t)(2 * kKeys);
        const float sl2 = p.scale * kLog2e;  // logits in the log2 domain: exp(x) = ex2(x * log2 e)
        // the reference's running max starts at FLT_MIN (scale_and_mask_and_softmax.cu:86-126)
        float m_run = FLT_MIN * kLog2e;  // running max of the row (log2 domain)
        float m_use = m_run;             // the max the accumulator and the partial sum are scaled to (m_use <= m_run <= m_use + 8)
        float l_part = 0.0f;
        uint32_t r[32];
        for (int j = 0; j < ntiles; ++j) {
            const int sb = j & 1, k0 = j * kKeys + cg * 32;
            bar_wait(s_full0 + 8 * sb, (j >> 1) & 1);
            tc_fence_after();
            tc_ld32(lane_addr + (uint32_t)(sb * kKeys), r);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) bar_arrive(s_free0 + 8 * sb);  // the warp's logits are in registers: S[sb] may be overwritten by tile j + 2
            // tiles entirely below the diagonal and inside the context need no mask (CTA-uniform: row 0 of the tile sees the fewest keys)
            const bool edge = j * kKeys + kKeys - 1 > q0 + (klen - qlen) || j * kKeys + kKeys > klen;
            float mloc = -INFINITY;
            if (edge) {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int kg = k0 + e;
                    if (!(kg < klen && kg <= lim)) r[e] = 0xff800000u;  // -inf
                    mloc = fmaxf(mloc, __uint_as_float(r[e]));
                }
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) mloc = fmaxf(mloc, __uint_as_float(r[e]));
            }
            float *rx = red + ((size_t)(j & 1) * kRows + row) * 4;
            rx[cg] = mloc * sl2;  // scale > 0: the max commutes with the scaling
            asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(128) : "memory");  // the 4 warps that share these 32 rows
            const float4 m4 = *reinterpret_cast<const float4 *>(rx);
            m_run = fmaxf(fmaxf(m_run, fmaxf(m4.x, m4.y)), fmaxf(m4.z, m4.w));
            // ---- lazy rescale (identical decision in the row's 4 threads: they see the same m_run and m_use)
            if (j == 0) {
                m_use = m_run;  // tile 0 overwrites the accumulator and the sum is still 0: adopting the first tile's max is free
            } else {
                const bool need = m_run - m_use > kRescaleStep;
                if (__any_sync(0xffffffffu, need)) {
                    const float corr = need ? ex2f(m_use - m_run) : 1.0f;
                    if (need) m_use = m_run;
                    l_part *= corr;
                    uint32_t acc[32];
                    bar_wait(pv_done, (j - 1) & 1);  // every P.V issued so far has landed in O
                    tc_fence_after();
                    tc_ld32(o_addr, acc);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc[e] = __float_as_uint(__uint_as_float(acc[e]) * corr);
                    tc_st32(o_addr, acc);
                    tc_wait_st();
                    tc_fence_before();
                }
            }
            // p = ex2(s - m_use): masked keys are -inf -> 0.  Rounded to T for the MMA (the reference's probabilities are a T tensor);
            // the row sum uses the unrounded values.  Packed two at a time (F2FP on the ALU pipe: a scalar cvt to bf16 is an XU
            // instruction like ex2 itself and doubled the load on the pipe this loop is bound by).
            // Shared memory: K-major, 128-byte swizzle (16-byte chunk index XOR (row % 8)).
            float psum = 0.0f;
            uint4 packed[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float pf[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float x = fmaf(__uint_as_float(r[8 * q + u]), sl2, -m_use);
                    pf[u] = poly_lane(u) ? ex2_poly(x) : ex2f(x);
                    psum += pf[u];
                }
                packed[q] = pack16<T>(pf);
            }
            l_part += psum;
            if (j > 0) bar_wait(pv_done, (j - 1) & 1);  // the previous P.V has read P out of shared memory
            unsigned char *prow = sP + (cg >> 1) * kHalfBytes + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int chunk = ((cg & 1) * 4 + q) ^ (row & 7);
                *reinterpret_cast<uint4 *>(prow + chunk * 16) = packed[q];
            }
            // rows of V past the context must not reach the tensor core: 0 * NaN would poison the output (the cache beyond
            // context_len is not initialised by anybody).  Row r of the tile is key j*128 + r; each of the row's 4 threads clears
            // its quarter of the 256 bytes.
            if (j * kKeys + row >= klen) {
                bar_wait(v_full0 + 8 * (j % kKvStages), (j / kKvStages) & 1);
                unsigned char *vrow = sV + (j % kKvStages) * kTileBytes + (cg >> 1) * kHalfBytes + row * 128 + (cg & 1) * 64;
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4 *>(vrow + q * 16) = make_uint4(0, 0, 0, 0);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of P / V visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) bar_arrive(p_ready);
        }
        // row sum across the row's 4 threads; the accumulator comes out of TMEM once
        float *rx = red + ((size_t)(ntiles & 1) * kRows + row) * 4;
        rx[cg] = l_part;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "r"(128) : "memory");
        if (ntiles > 0) {  // CTA-uniform
            bar_wait(pv_done, (ntiles - 1) & 1);
            tc_fence_after();
            tc_ld32(o_addr, r);
            tc_wait_ld();
            if (qi < qlen) {
                const float4 l4 = *reinterpret_cast<const float4 *>(rx);
                const float c = ex2f(m_use - m_run);  // the sum and the accumulator are relative to m_use; the reference's are relative to the max
                const float inv = c / (((l4.x + l4.y) + (l4.z + l4.w)) * c + 1e-6f);
                const int t = b * p.max_q_len + qi - p.seq_off[b];  // un-padded token index
                T *dst = reinterpret_cast<T *>(p.out) + ((size_t)t * p.head_num + h) * kD + cg * 32;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    float f[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(r[e + u]) * inv;
                    st_v4(dst + e, pack16<T>(f));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kSoftmaxWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        else
            cudaGetLastError();
    });
    return fn;
}
// rows x 128 matrix of a 16-bit type, boxes of 128 rows x 64 columns, 128-byte swizzle
static bool make_map(CUtensorMap *map, const void *ptr, size_t rows, bool bf16, int box_rows = 128) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kD * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace catc

// q [B, H, max_q_len, 128]; k/v: LAYER base of the cache [B, Hkv, S, 128]; out [T, H, 128]; seq_off [B] (device).
// block_table != NULL: k/v are the layer base of a page pool [num_pages, Hkv, 64, 128] and block_table [B, max_pages] (device) maps
// 64-key blocks to pages (max_seq_len is then ignored).
// Returns B200_ERR_UNSUPPORTED (no error text) when the shape / type cannot use the tensor-core kernel.
int launch_context_attention_tc(const void *q, const void *k_layer, const void *v_layer, void *out, const int *seq_off, const int *input_len,
                                const int *context_len, int batch, int head_num, int kv_head_num, int max_q_len, int max_seq_len, int head_size,
                                float scale, int dtype, cudaStream_t st, const int *block_table, int max_pages, int num_pages) {
    using namespace catc;
    if ((dtype != B200_BF16 && dtype != B200_F16) || head_size != kD) return B200_ERR_UNSUPPORTED;
    if (!aligned16(q) || !aligned16(k_layer) || !aligned16(v_layer) || !aligned16(out)) return B200_ERR_UNSUPPORTED;
    const bool bf16 = dtype == B200_BF16;
    CUtensorMap tmQ, tmK, tmV;
    const size_t kv_rows = block_table ? (size_t)num_pages * kv_head_num * 64 : (size_t)batch * kv_head_num * max_seq_len;
    const int kv_box = block_table ? 64 : 128;
    if (!make_map(&tmQ, q, (size_t)batch * head_num * max_q_len, bf16) || !make_map(&tmK, k_layer, kv_rows, bf16, kv_box) ||
        !make_map(&tmV, v_layer, kv_rows, bf16, kv_box)) {
        set_error("context_attention: cuTensorMapEncodeTiled failed");
        return B200_ERR_CUDA;
    }
    Params p = {};
    p.out = out, p.seq_off = seq_off, p.input_len = input_len, p.context_len = context_len;
    p.head_num = head_num, p.kv_head_num = kv_head_num, p.max_q_len = max_q_len, p.max_seq_len = max_seq_len;
    p.scale = scale, p.is_bf16 = bf16 ? 1 : 0;
    p.block_table = block_table, p.max_pages = max_pages;
    const size_t smem = (size_t)(2 + 2 * kKvStages) * kTileBytes + 1024 + 18 * 8 + 16 + 2 * kRows * 4 * sizeof(float);
    dim3 grid((max_q_len + kRows - 1) / kRows, head_num, batch);
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_pdl(kern, grid, dim3(kThreads), smem, st, true, tmQ, tmK, tmV, p);
    };
    if (bf16) go(context_attn_tc_kernel<__nv_bfloat16>);
    else go(context_attn_tc_kernel<__half>);
    return cuda_status("context_attention (tcgen05) launch");
}

}  // namespace b200
