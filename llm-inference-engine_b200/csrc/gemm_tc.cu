// gemm_tc.cu -- tcgen05 / TMEM tensor-core GEMM for sm_100a: D[i,j] = sum_k A[i,k] * B[j,k], both operands K-major.
//
// Serves the two GEMM-shaped cases of launchLinearGemm (reference src/kernels/linear.cu:10-87) with W packed [N,K]:
//   * prefill / context linears (M = tokens >= 129):  A = x [M,K],  B = W [N,K],  C[i,j]      (tensor-pipe bound)
//   * batched decode linears   (5 <= M <= 128):       A = W [N,K],  B = x [M,K],  C[j,i]      ("swap-AB": the weight rows
//     fill the 128-row MMA M dimension, the few tokens are the MMA N dimension; HBM bound -> stream-K: the tiles x k-blocks
//     space is cut into one equal contiguous range per SM so that every SM streams the same number of weight bytes)
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (128B swizzle) of A and B into a ring of smem stages, mbarrier tx counts
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16 per instruction, fp32
//            accumulate in TMEM), tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 2-5 epilogue: tcgen05.ld the accumulator (each warp owns its 32-lane TMEM quarter), convert, store; double-buffered
//            accumulators so the epilogue of tile t overlaps the MMAs of tile t+1
// Tiles shared by several CTAs (stream-K) exchange fp32 partials through the library workspace; the last CTA to arrive at a tile
// (self-resetting ticket) adds the slots in a fixed order (deterministic, no atomics on data).
#include "gemm_tc.cuh"

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link dependency)
#include <mutex>

namespace b200 {

constexpr int kBM = 128;   // UMMA M
constexpr int kBK = 64;    // K elements per stage = one 128-byte swizzle row of a 16-bit type
constexpr int kUmmaK = 16; // K per tcgen05.mma for 16-bit inputs
constexpr int kTcThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBM * kBK * 2;

struct GemmTcParams {
    void *C;
    float *partial;
    unsigned int *tickets;
    int ldc;
    int rowsA, rowsB, K;
    int tilesA, tiles, kb_total;
    int streamk;   // 0: whole tiles dealt round-robin; 1: the tiles x k-blocks space cut into gridDim.x equal contiguous ranges
    int maxslots;  // stream-K: partial slots reserved per tile
    int bn;        // UMMA N: rows of B per tile, multiple of 16, <= 256
    int stages;
    int swap;      // 0: C[i*ldc + j]   1: C[j*ldc + i]
    int is_bf16;
    int acc_bufs;  // TMEM accumulators (2 when 2*bn <= 512)
    int acc_stride;  // TMEM columns between the accumulators: bn, or 256 for the 192..240-column tiles (power-of-two placement)
    unsigned int tmem_cols;
    // SwiGLU epilogue (gate_up linear of the prefill path, launchLinearGemm + launchSiluAndMul in one kernel): B = W [2I, K] with the gate
    // rows first; tile tb multiplies gate rows [hw tb, +hw) AND up rows [I + hw tb, +hw) as one N = 2 hw tile (two TMA boxes into one
    // B stage), the epilogue reads both halves of the accumulator and writes silu(gate) * up: C [rowsA, I], rowsB = I.  0: plain GEMM.
    int swiglu_inter;
    int swiglu_hw;  // gate (and up) columns per tile: bn = 2 * swiglu_hw (128, 112 or 96: whichever wastes least of the last wave)
};

// One contiguous piece of work of a CTA: k-blocks [kb0, kb1) of one output tile.  nslots > 1: the tile is shared with
// other CTAs; this CTA's partial goes to slot `slot` and the last arriver (ticket) adds the slots in order.
struct Seg {
    int tile, kb0, kb1, slot, nslots;
};
struct SegIter {
    const GemmTcParams &p;
    long long g, gend, total;
    int item, G;
    __device__ SegIter(const GemmTcParams &pp) : p(pp) {
        G = gridDim.x;
        total = (long long)p.tiles * p.kb_total;
        item = blockIdx.x;
        g = start(blockIdx.x);
        gend = start(blockIdx.x + 1);
    }
    __device__ long long start(int c) const { return (long long)c * total / G; }
    __device__ int owner(long long gg) const { return (int)(((gg + 1) * G + total - 1) / total) - 1; }  // max c: start(c) <= gg
    __device__ bool next(Seg &s) {
        if (!p.streamk) {
            if (item >= p.tiles) return false;
            s.tile = item, s.kb0 = 0, s.kb1 = p.kb_total, s.slot = 0, s.nslots = 1;
            item += G;
            return true;
        }
        if (g >= gend) return false;
        s.tile = (int)(g / p.kb_total);
        s.kb0 = (int)(g - (long long)s.tile * p.kb_total);
        const long long len = min((long long)(p.kb_total - s.kb0), gend - g);
        s.kb1 = s.kb0 + (int)len;
        const int first = owner((long long)s.tile * p.kb_total), last = owner((long long)(s.tile + 1) * p.kb_total - 1);
        s.slot = (int)blockIdx.x - first;
        s.nslots = last - first + 1;
        g += len;
        return true;
    }
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t = lane base + t)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), descriptor version 1 (sm_100),
// layout type 2 (SWIZZLE_128B) in bits 61..63; LBO is unused for swizzled K-major layouts (canonical value 1).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B both K-major, no negate / sparsity / saturation.
__device__ __forceinline__ uint32_t instr_desc_f16(bool bf16, int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;                      // D format: F32
    d |= (bf16 ? 1u : 0u) << 7;        // A format
    d |= (bf16 ? 1u : 0u) << 10;       // B format
    d |= (uint32_t)(n >> 3) << 17;     // N / 8
    d |= (uint32_t)(m >> 4) << 24;     // M / 16
    return d;
}

template <typename T> __device__ __forceinline__ void store_row16(T *dst, const uint32_t (&r)[16], int valid) {
    // dst: 16 consecutive outputs of one row (32 bytes for 16-bit T); valid = number of in-range columns
    if (valid >= 16 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(r[e]);
        st_v4(dst, pack16<T>(f));
        st_v4(dst + 8, pack16<T>(f + 8));
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
            if (e < valid) dst[e] = Elem<T>::from_f(__uint_as_float(r[e]));
    }
}

template <typename T>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: stages x [A tile | B tile] (1024-byte aligned), then barriers, then the TMEM base address
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = p.bn * kBK * 2;
    const int stage_bytes = kATileBytes + b_tile_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    const uint32_t full0 = s_u32(bars), empty0 = full0 + 8 * kMaxStages;
    const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);
    int *flag_slot = reinterpret_cast<int *>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) {
            bar_init(full0 + 8 * s, 1);
            bar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            bar_init(tfull0 + 8 * a, 1);
            bar_init(tempty0 + 8 * a, 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    pdl_wait();               // everything below reads (x) or overwrites (C) tensors of the previous kernel
    pdl_launch_dependents();

    if (warp == 0) {
        // ================================================= TMA producer
        if (elect_one()) {
            SegIter si(p);
            Seg sg;
            int it = 0;
            while (si.next(sg)) {
                const int ta = sg.tile % p.tilesA, tb = sg.tile / p.tilesA;
                for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
                    const int s = it % p.stages;
                    bar_wait(empty0 + 8 * s, ((it / p.stages) & 1) ^ 1);
                    const uint32_t sa = s_u32(smem + (size_t)s * stage_bytes);
                    bar_expect_tx(full0 + 8 * s, (uint32_t)stage_bytes);
                    tma_load_2d(sa, &tmA, kb * kBK, ta * kBM, full0 + 8 * s);
                    if (p.swiglu_inter) {  // box = hw rows: the gate rows, then the up rows of the same columns
                        tma_load_2d(sa + kATileBytes, &tmB, kb * kBK, tb * p.swiglu_hw, full0 + 8 * s);
                        tma_load_2d(sa + kATileBytes + p.swiglu_hw * kBK * 2, &tmB, kb * kBK, p.swiglu_inter + tb * p.swiglu_hw, full0 + 8 * s);
                    } else {
                        tma_load_2d(sa + kATileBytes, &tmB, kb * kBK, tb * p.bn, full0 + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================= MMA issuer
        const uint32_t idesc = instr_desc_f16(p.is_bf16 != 0, kBM, p.bn);
        SegIter si(p);
        Seg sg;
        int it = 0, t = 0;
        for (; si.next(sg); ++t) {
            const int acc = t % p.acc_bufs;
            bar_wait(tempty0 + 8 * acc, ((t / p.acc_bufs) & 1) ^ 1);  // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.acc_stride);
            for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
                const int s = it % p.stages;
                bar_wait(full0 + 8 * s, (it / p.stages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = s_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in the (>>4) start-address field
                        tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > sg.kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty0 + 8 * s);                          // smem stage reusable once these MMAs have read it
                    if (kb == sg.kb1 - 1) tc_commit(tfull0 + 8 * acc);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================= epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int q = warp & 3;
        const int ep_tid = (warp - 2) * 32 + lane;  // 0..127
        T *C = reinterpret_cast<T *>(p.C);
        SegIter si(p);
        Seg sg;
        int t = 0;
        for (; si.next(sg); ++t) {
            const int ta = sg.tile % p.tilesA, tb = sg.tile / p.tilesA;
            const int acc = t % p.acc_bufs;
            bar_wait(tfull0 + 8 * acc, (t / p.acc_bufs) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.acc_stride);
            const int i = ta * kBM + q * 32 + lane;  // row of A this thread owns
            const int j0 = tb * p.bn;
            if (p.swiglu_inter) {
                // gate in accumulator columns [0, hw), up in [hw, 2 hw): both rounded to T first (the reference's linear writes a T
                // tensor that launchSiluAndMul reads back, src/kernels/silu_and_mul.cu:6-41), same expression as silu_and_mul_kernel
                const int hw = p.swiglu_hw;
                for (int c = 0; c < hw; c += 16) {
                    uint32_t g[16], u[16];
                    tc_ld16(taddr + c, g);
                    tc_ld16(taddr + hw + c, u);
                    tc_wait_ld();
                    const int valid = min(16, p.rowsB - (tb * hw + c));
                    if (i < p.rowsA && valid > 0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float gf = round_to<T>(__uint_as_float(g[e])), uf = round_to<T>(__uint_as_float(u[e]));
                            g[e] = __float_as_uint((gf / (1.0f + expf(-gf))) * uf);
                        }
                        store_row16<T>(C + (size_t)i * p.ldc + tb * hw + c, g, valid);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(tempty0 + 8 * acc);
            } else if (sg.nslots == 1) {
                for (int c = 0; c < p.bn; c += 16) {
                    uint32_t r[16];
                    tc_ld16(taddr + c, r);
                    tc_wait_ld();
                    if (i < p.rowsA) {
                        if (!p.swap) {
                            const int valid = min(16, p.rowsB - (j0 + c));
                            if (valid > 0) store_row16<T>(C + (size_t)i * p.ldc + j0 + c, r, valid);
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (j0 + c + e < p.rowsB) C[(size_t)(j0 + c + e) * p.ldc + i] = Elem<T>::from_f(__uint_as_float(r[e]));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(tempty0 + 8 * acc);
            } else {
                // shared tile: publish this CTA's fp32 partial [bn][128] (coalesced over lanes); the last arriver reduces
                const size_t tile_floats = (size_t)p.bn * kBM;
                float *part = p.partial + ((size_t)sg.tile * p.maxslots + sg.slot) * tile_floats;
                for (int c = 0; c < p.bn; c += 16) {
                    uint32_t r[16];
                    tc_ld16(taddr + c, r);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) __stcg(part + (size_t)(c + e) * kBM + q * 32 + lane, __uint_as_float(r[e]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) bar_arrive(tempty0 + 8 * acc);  // the accumulator is free: the MMA warp may start the next segment
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (ep_tid == 0) *flag_slot = atomicInc(&p.tickets[sg.tile], (unsigned)(sg.nslots - 1)) == (unsigned)(sg.nslots - 1);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const bool last = *flag_slot != 0;
                asm volatile("bar.sync 1, 128;" ::: "memory");  // flag_slot may be rewritten by the next segment
                if (last) {
                    __threadfence();
                    // 128 threads x float4: element f = (column c, 4 consecutive A rows); slots added in order (deterministic)
                    const float4 *pt = reinterpret_cast<const float4 *>(p.partial + (size_t)sg.tile * p.maxslots * tile_floats);
                    const int nf = p.bn * (kBM / 4), slot_f4 = (int)(tile_floats / 4);
                    constexpr int U = 4;
                    for (int f0 = ep_tid; f0 < nf; f0 += 128 * U) {
                        float4 sum[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) sum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int k2 = 0; k2 < sg.nslots; ++k2) {
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int f = f0 + u * 128;
                                if (f < nf) {
                                    const float4 v = __ldcg(pt + (size_t)k2 * slot_f4 + f);
                                    sum[u].x += v.x, sum[u].y += v.y, sum[u].z += v.z, sum[u].w += v.w;
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int f = f0 + u * 128;
                            if (f >= nf) continue;
                            const int c = f / (kBM / 4), ii = ta * kBM + (f % (kBM / 4)) * 4;
                            if (j0 + c >= p.rowsB) continue;
                            const float v[4] = {sum[u].x, sum[u].y, sum[u].z, sum[u].w};
                            if (p.swap) {
                                T *dst = C + (size_t)(j0 + c) * p.ldc + ii;
                                if (ii + 3 < p.rowsA && (reinterpret_cast<uintptr_t>(dst) & 7) == 0 && sizeof(T) == 2) {
                                    uint2 pk;
                                    const T a0 = Elem<T>::from_f(v[0]), a1 = Elem<T>::from_f(v[1]), a2 = Elem<T>::from_f(v[2]), a3 = Elem<T>::from_f(v[3]);
                                    pk.x = (uint32_t)(*reinterpret_cast<const unsigned short *>(&a0)) | ((uint32_t)(*reinterpret_cast<const unsigned short *>(&a1)) << 16);
                                    pk.y = (uint32_t)(*reinterpret_cast<const unsigned short *>(&a2)) | ((uint32_t)(*reinterpret_cast<const unsigned short *>(&a3)) << 16);
                                    *reinterpret_cast<uint2 *>(dst) = pk;
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        if (ii + e < p.rowsA) dst[e] = Elem<T>::from_f(v[e]);
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    if (ii + e < p.rowsA) C[(size_t)(ii + e) * p.ldc + j0 + c] = Elem<T>::from_f(v[e]);
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        else
            cudaGetLastError();
    });
    return fn;
}

// 2-D K-major tensor map: dims {K, rows}, row pitch K * 2 bytes, box {64, box_rows}, 128-byte swizzle, zero fill out of bounds.
static bool make_map(CUtensorMap *map, const void *ptr, int rows, int K, int box_rows, bool bf16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides,
              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static unsigned int pow2_cols(int c) {
    unsigned int v = 32;
    while ((int)v < c) v <<= 1;
    return v;
}

static int launch_gemm_tc_impl(const void *x, const void *w, void *y, int M, int N, int K, int dtype, int swiglu_inter, cudaStream_t st) {
    if (dtype != B200_BF16 && dtype != B200_F16) return B200_ERR_UNSUPPORTED;
    if (M < 1 || K % 8 != 0 || K < kBK || !aligned16(x) || !aligned16(w) || !y) return B200_ERR_UNSUPPORTED;
    const bool bf16 = dtype == B200_BF16;
    const bool swap = M <= 128;
    if (swiglu_inter && (swap || swiglu_inter < 128)) return B200_ERR_UNSUPPORTED;  // the fused epilogue exists for prefill-sized token counts
    GemmTcParams p = {};
    p.C = y, p.ldc = N, p.K = K, p.swap = swap ? 1 : 0, p.is_bf16 = bf16 ? 1 : 0;
    p.swiglu_inter = swiglu_inter;
    p.kb_total = (K + kBK - 1) / kBK;
    const void *a_ptr, *b_ptr;
    int rows_b_map;  // rows of the B tensor as the TMA sees it
    if (swap) {
        a_ptr = w, b_ptr = x;
        p.rowsA = N, p.rowsB = M;
        p.bn = (M + 15) / 16 * 16;
        rows_b_map = M;
    } else {
        a_ptr = x, b_ptr = w;
        p.rowsA = M, p.rowsB = N;
        p.bn = N >= 256 ? 256 : (N + 15) / 16 * 16;
        rows_b_map = N;
    }
    const int sms = sm_count();
    const int tiles_a = ((swap ? N : M) + kBM - 1) / kBM;
    // Tile width of the prefill shapes: the widest tile is the most efficient one, but the number of waves is an integer -- 768 tiles of
    // 256 columns on 148 SMs (7B QKV projection at 2048 tokens) take 6 waves for 5.19 waves of work.  Choose the width whose
    // waves x width is smallest (ties: the wider one): 224 for that shape (880 tiles, 5.95 waves), 240 for N = 4096.  Measured gain 2-3 %
    // per GEMM, not the 6-12 % of the wave count: a narrower MMA re-reads the 128-row A tile more often per flop.
    auto waves_cost = [&](int cols, int width) { return (long long)((tiles_a * ((cols + width - 1) / width) + sms - 1) / sms) * width; };
    if (swiglu_inter) {
        // measured (7B, 2048 tokens): 2 x 112 columns (11 waves instead of 10 of 2 x 128) is 6 % SLOWER -- the narrower MMA re-reads the
        // A tile more often per flop than the better wave count saves; the half width stays a parameter, the choice is 128
        const int hw = 128;
        p.swiglu_hw = hw;
        p.bn = 2 * hw, p.rowsB = swiglu_inter, p.ldc = swiglu_inter;  // N = I output columns; the MMA tile is hw gate + hw up rows
        rows_b_map = 2 * swiglu_inter;
    } else if (!swap && p.bn == 256) {
        if (tiles_a * ((N + 255) / 256) * 2 <= sms) {
            p.bn = 128;  // under half a wave: more, smaller tiles
        } else {
            int bn = 256;
            for (int cand : {240, 224, 208, 192})
                if (waves_cost(N, cand) < waves_cost(N, bn)) bn = cand;
            p.bn = bn;
        }
    }
    p.tilesA = (p.rowsA + kBM - 1) / kBM;
    const int tilesB = swiglu_inter ? (swiglu_inter + p.swiglu_hw - 1) / p.swiglu_hw : (p.rowsB + p.bn - 1) / p.bn;
    p.tiles = p.tilesA * tilesB;
    const long long total = (long long)p.tiles * p.kb_total;
    int grid = p.tiles < sms ? p.tiles : sms;
    // Few tiles relative to the SM count (HBM-bound decode shapes): stream-K -- the tiles x k-blocks space is cut into one equal
    // contiguous range per SM, so every SM streams the same number of weight bytes.
    p.streamk = 0;
    if (swap && p.tiles < 4 * sms && total >= 2 * sms) {
        const int g2 = (int)(total < sms ? total : sms);
        const long long per = total / g2;  // >= 2 k-blocks per CTA
        const int maxslots = (int)((p.kb_total + per - 1) / per) + 1;
        Workspace ws;
        if (!get_workspace(&ws)) return B200_ERR_WORKSPACE;
        const size_t need = (size_t)p.tiles * maxslots * p.bn * kBM * sizeof(float);
        if (need <= ws.scratch_bytes && (size_t)p.tiles <= ws.n_tickets) {
            p.streamk = 1, p.maxslots = maxslots, grid = g2;
            p.partial = reinterpret_cast<float *>(ws.scratch);
            p.tickets = ws.tickets;
        }
    }
    p.acc_bufs = 2 * p.bn <= 512 ? 2 : 1;
    p.acc_stride = p.bn > 128 ? 256 : p.bn;
    p.tmem_cols = pow2_cols(p.acc_bufs == 2 ? p.acc_stride + p.bn : p.bn);
    const int stage_bytes = kATileBytes + p.bn * kBK * 2;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return B200_ERR_UNSUPPORTED;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 /*alignment slack*/ + (2 * kMaxStages + 4) * 8 + 16;

    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, a_ptr, p.rowsA, K, kBM, bf16) || !make_map(&tmB, b_ptr, rows_b_map, K, swiglu_inter ? p.swiglu_hw : p.bn, bf16)) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed (rows %d/%d, K %d)", p.rowsA, p.rowsB, K);
        return B200_ERR_CUDA;
    }
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, st, true, tmA, tmB, p);
    };
    if (bf16) launch(gemm_tc_kernel<__nv_bfloat16>);
    else launch(gemm_tc_kernel<__half>);
    return cuda_status("gemm_tc launch");
}

int launch_gemm_tc(const void *x, const void *w, void *y, int M, int N, int K, int dtype, cudaStream_t st) {
    return launch_gemm_tc_impl(x, w, y, M, N, K, dtype, 0, st);
}

// act[M, inter] = silu(x . Wgate^T) * (x . Wup^T), w = [2 * inter, K] (gate rows, then up rows): the gate_up linear and launchSiluAndMul in one
// kernel (prefill sizes, M > 128); the [M, 2 * inter] intermediate is never written.  B200_ERR_UNSUPPORTED: use the two launchers.
int launch_gemm_tc_swiglu(const void *x, const void *w, void *act, int M, int inter, int K, int dtype, cudaStream_t st) {
    return launch_gemm_tc_impl(x, w, act, M, inter, K, dtype, inter, st);
}

}  // namespace b200
