// attention_decode.cu -- decode-time RoPE and masked multi-head attention over the KV cache for sm_100a.
//
// Reference semantics: src/kernels/rope.cu:4-43 (+ rope_utils.cuh:6-19) and
// src/kernels/decoder_self_attention.cu:56-188 (bias after RoPE, K/V appended at step-1, logits q.k*rsqrt(d),
// softmax with +1e-6 in the denominator and the max-with-0 quirk when step < head_size, P.V).
//
// B200 design (HBM-bound: the K and V rows of [0, step) are read exactly once):
//   * split-KV ("flash-decoding"): grid = (splits, Hkv, B), about two CTAs per SM, so that B*Hkv = 32 still fills 148 SMs;
//     every CTA serves ALL q heads of its kv head (GQA: K/V rows are loaded once per group, not once per q head);
//   * the K and V rows of a head are contiguous in the [L,B,Hkv,S,d] cache, so a tile of 64 positions is TWO 16 KiB 1-D TMA
//     bulk copies (cp.async.bulk.shared.global) into a 3-stage shared-memory ring filled by a dedicated producer warp.  Inside the
//     fused engine the first stages are requested BEFORE griddepcontrol.wait: cached positions do not depend on the QKV
//     linear that precedes this kernel, so their HBM latency hides behind its tail (programmatic dependent launch);
//   * one pass over the positions: a K/V row of 128 elements is 16 lanes x 16 bytes (16-bit) or 32 lanes x 16 bytes (fp32);
//     each row group keeps a private online-softmax state (running max / sum / output slice) -- no block-wide barrier in
//     the loop; the row groups are merged through shared memory at the end, in a fixed order;
//   * RoPE + bias + cache append of the new token are fused here (the new K/V row never round-trips through HBM
//     before being used);
//   * partial (max, sum, out) per split go to scratch; the last CTA of a (b, kv head) -- elected with a
//     self-resetting ticket -- merges them in split order (deterministic) and applies the reference's final
//     normalisation.
#include "attention_decode.cuh"

#include <stdlib.h>

namespace b200 {

// theta denominators exactly as the reference computes them on the host side of its unit tests:
// powf(base, zid / rot_dim) with a correctly rounded result (the device powf may be 4 ulp off, which at
// position ~1000 moves cos/sin by 1e-4).
__device__ __forceinline__ float rope_denominator(float base, int zid, int rot_dim) {
    const float e = (float)zid / (float)rot_dim;
    return (float)pow((double)base, (double)e);
}
__device__ __forceinline__ float2 rope_cos_sin(float pos, float denom) {
    const float th = pos / denom;
    return make_float2(cosf(th), sinf(th));
}
__device__ __forceinline__ void rope_apply(float &x0, float &x1, const float2 cs) {
    const float a = x0, b = x1;
    x0 = a * cs.x - b * cs.y;
    x1 = b * cs.x + a * cs.y;
}
__device__ __forceinline__ void rope_rotate(float &x0, float &x1, float pos, float denom) { rope_apply(x0, x1, rope_cos_sin(pos, denom)); }

// (cos, sin) table: the same expressions as above, evaluated once per (position, pair) instead of once per head per step
__global__ void rope_table_kernel(float2 *table, int positions, int rot_dim, float base) {
    const int half = rot_dim / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < positions * half; i += gridDim.x * blockDim.x)
        table[i] = rope_cos_sin((float)(i / half), rope_denominator(base, 2 * (i % half), rot_dim));
}
int launch_rope_table(float2 *table, int positions, int rot_dim, float rot_base, cudaStream_t st) {
    if (positions <= 0 || rot_dim < 2) return B200_OK;
    rope_table_kernel<<<sm_count() * 4, 256, 0, st>>>(table, positions, rot_dim, rot_base);
    return cuda_status("rope_table launch");
}

// ------------------------------------------------------------------ standalone decode RoPE (launchRope)
template <typename T>
__global__ void rope_decode_kernel(T *qkv, int head_num, int kv_head_num, int head_size, int step, int rot_dim, float base) {
    const int b = blockIdx.y, h = blockIdx.x;  // h over q heads then k heads
    T *p = qkv + ((size_t)b * (head_num + 2 * kv_head_num) + h) * head_size;
    pdl_wait();
    for (int i = threadIdx.x; i < rot_dim / 2; i += blockDim.x) {
        float x0 = Elem<T>::to_f(p[i]), x1 = Elem<T>::to_f(p[i + head_size / 2]);
        rope_rotate(x0, x1, (float)(step - 1), rope_denominator(base, 2 * i, rot_dim));
        p[i] = Elem<T>::from_f(x0);
        p[i + head_size / 2] = Elem<T>::from_f(x1);
    }
}

// ------------------------------------------------------------------ fused split-KV decode attention, head_size 128
constexpr int kAttnWarps = 8;                          // compute warps
constexpr int kAttnThreads = (kAttnWarps + 1) * 32;    // + one TMA producer warp
constexpr int kAttnD = 128;
constexpr int kAttnTileBytes = 16384;                  // bytes of K (and of V) per ring stage: 64 positions (16-bit) / 32 (fp32)
constexpr int kAttnStages = 3;

__device__ __forceinline__ uint32_t a_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void a_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void a_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void a_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    // read-once per step: evict-first in L2 (see gemv.cuh l2_evict_first_policy)
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}

template <typename T> __device__ __forceinline__ float round_t(float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); }

// final normalisation shared by the single-split path and the merge: reference decoder_self_attention.cu:145-165
__device__ __forceinline__ float final_max(float m, int step, int head_size) { return (step < head_size && m < 0.0f) ? 0.0f : m; }

// GC = q heads served by one CTA (the whole GQA group, or half of it when the group has 8 heads: 8 x (q + output) slices do not
// fit the register budget of a 288-thread CTA -- the two halves read the same K/V rows, the second read hits L2).
// kPaged: the cache is a page pool addressed through a block table (attention_decode.cuh); a tile of TP positions never straddles a page
// (TP divides kAttnPageSize, tiles start at multiples of TP), so a stage is still two contiguous bulk copies.
template <typename T, int GC, bool kPaged>
__global__ void __launch_bounds__(kAttnThreads, GC == 4 ? 1 : 2)  // 4 heads per CTA: 135+ registers, one CTA per SM (as before the merge moved in)
decode_attn_kernel(const DecodeAttnArgs a) {
    constexpr int D = kAttnD;
    constexpr int G = GC;
    constexpr int V = Elem<T>::kVec;                   // elements per 16-byte vector
    constexpr int LPR = D / V;                         // lanes per K/V row: 16 (16-bit) or 32 (fp32)
    constexpr int RG = kAttnWarps * 32 / LPR;          // row groups per CTA: 16 or 8
    constexpr int TP = kAttnTileBytes / (D * (int)sizeof(T));  // positions per tile: 64 or 32
    constexpr int RPG = TP / RG;                       // rows of a tile per row group: 4
    constexpr int kStageBytes = 2 * kAttnTileBytes;
    constexpr int PS = kAttnPartStride;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;                                                   // [stages][K tile | V tile]
    float *merge = reinterpret_cast<float *>(smem_raw);                               // aliases the ring after the loop: [RG][G][D+2]
    float *qs = reinterpret_cast<float *>(smem_raw + kAttnStages * kStageBytes);      // [G][D]
    float *knew = qs + G * D;                                                         // [D]
    float *vnew = knew + D;                                                           // [D]
    float *wts = vnew + D;                                                            // [RG][G] merge weights, then [G] max, [G] sum
    const uint32_t full0 = a_smem_u32(wts + RG * G + 2 * G + 2), empty0 = full0 + 8 * kAttnStages;
    __shared__ bool is_last;
    static_assert(kAttnPageSize % TP == 0, "a tile must not straddle pages");

    const int split = blockIdx.x, b = blockIdx.z;
    const int H = a.head_num, Hkv = a.kv_head_num;
    const int Gtot = H / Hkv, gsplit = Gtot / G;       // CTAs per (b, kv head, split): 1, or 2 half-groups
    const int kvh = blockIdx.y / gsplit, g0 = (blockIdx.y % gsplit) * G;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int qkv_heads = H + 2 * Hkv;
    // ragged batches: every row has its own step (read from device memory: written several kernels ago, visible before griddepcontrol.wait)
    const int step = a.steps ? min(max(__ldg(a.steps + b), 1), a.step) : a.step;
    const int cached = step - 1;                       // positions [0, cached) are in the cache; position `cached` is the new token
    const int p0 = split * a.chunk, p1 = min(cached, p0 + a.chunk);
    const int ntiles = p1 > p0 ? (p1 - p0 + TP - 1) / TP : 0;
    // the row's last split also serves the token being appended (from shared memory, not the cache); the grid is planned for the longest
    // row: splits past a shorter row's end contribute an empty partial (max = -inf, sum = 0)
    const int nsplit_b = cached > 0 ? (cached + a.chunk - 1) / a.chunk : 1;
    const bool has_new = split == nsplit_b - 1;
    const T *qkv = reinterpret_cast<const T *>(a.qkv) + (size_t)b * qkv_heads * D;
    const T *bias = reinterpret_cast<const T *>(a.bias);
    // contiguous: the rows of (b, kv head); paged: the pool's layer base (rows are found through the block table)
    T *kc = reinterpret_cast<T *>(a.k_cache) + (kPaged ? (size_t)0 : ((size_t)b * Hkv + kvh) * a.max_seq_len * D);
    T *vc = reinterpret_cast<T *>(a.v_cache) + (kPaged ? (size_t)0 : ((size_t)b * Hkv + kvh) * a.max_seq_len * D);
    const int *bt = kPaged ? a.block_table + (size_t)b * a.max_pages : nullptr;
    // element offset of position `pos` of this (b, kv head) from kc / vc
    auto row_off = [&](int pos) -> size_t {
        if constexpr (kPaged) return (((size_t)__ldg(bt + pos / kAttnPageSize) * Hkv + kvh) * kAttnPageSize + pos % kAttnPageSize) * D;
        else return (size_t)pos * D;
    };
    const bool producer = warp == kAttnWarps && (tid & 31) == 0;

    // ---- producer: one stage = the K rows and the V rows of TP positions (two contiguous byte ranges)
    int p_t = 0, p_s = 0;
    size_t p_off = 0;  // paged: offset of the NEXT tile, looked up one tile ahead so that the table read hides behind the ring wait
    if (kPaged && producer && ntiles > 0) p_off = row_off(p0);
    auto issue_next = [&]() {
        const int base = p0 + p_t * TP;
        const uint32_t bytes = (uint32_t)min(TP, p1 - base) * (uint32_t)(D * sizeof(T));
        const uint32_t bar = full0 + 8 * p_s, dst = a_smem_u32(ring + (size_t)p_s * kStageBytes);
        const size_t off = kPaged ? p_off : (size_t)base * D;
        a_mbar_expect_tx(bar, 2 * bytes);
        a_bulk_g2s(dst, kc + off, bytes, bar);
        a_bulk_g2s(dst + kAttnTileBytes, vc + off, bytes, bar);
        ++p_t;
        if (kPaged && p_t < ntiles) p_off = row_off(base + TP);
        if (++p_s == kAttnStages) p_s = 0;
    };
    if (producer) {
        for (int s = 0; s < kAttnStages; ++s) {
            a_mbar_init(full0 + 8 * s, 1);
            a_mbar_init(empty0 + 8 * s, kAttnWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // cached positions do not depend on the kernel before this one (the QKV linear): request them now
        if (a.prefetch)
            while (p_t < kAttnStages && p_t < ntiles) issue_next();
    }

    pdl_wait();

    // ---- q (this CTA's G heads of the group), and the new k / v row: RoPE at step-1, then bias (reference order).  One pass over
    //      (G + 1) * D/2 rotary pairs and the D values of v, so that every global load of the prologue is in flight at once (a second
    //      loop for v cost the CTA that serves the new token a second round trip)
    constexpr int kPairs = (G + 1) * (D / 2);
    for (int i = tid; i < kPairs + D; i += kAttnThreads) {
        if (i >= kPairs) {  // the new v row
            if (!has_new) continue;
            const int j = i - kPairs, head = H + Hkv + kvh;
            float v = Elem<T>::to_f(qkv[(size_t)head * D + j]);
            if (bias) v = round_t<T>(v + Elem<T>::to_f(bias[(size_t)head * D + j]));
            vnew[j] = v;
            continue;
        }
        const int g = i / (D / 2), j = i % (D / 2);  // g == G: the k head
        if (g == G && !has_new) continue;
        const int head = g < G ? kvh * Gtot + g0 + g : H + kvh;
        float x0 = Elem<T>::to_f(qkv[(size_t)head * D + j]), x1 = Elem<T>::to_f(qkv[(size_t)head * D + j + D / 2]);
        if (a.apply_rope && j < a.rot_dim / 2) {
            const float2 cs = a.rope_cs ? __ldg(a.rope_cs + (size_t)(step - 1) * (a.rot_dim / 2) + j)
                                        : rope_cos_sin((float)(step - 1), rope_denominator(a.rot_base, 2 * j, a.rot_dim));
            rope_apply(x0, x1, cs);
            x0 = round_t<T>(x0), x1 = round_t<T>(x1);  // launchRope writes T back before the MHA kernel reads it
        }
        if (bias) {
            x0 = round_t<T>(x0 + Elem<T>::to_f(bias[(size_t)head * D + j]));
            x1 = round_t<T>(x1 + Elem<T>::to_f(bias[(size_t)head * D + j + D / 2]));
        }
        float *dst = g < G ? qs + g * D : knew;
        dst[j] = x0;
        dst[j + D / 2] = x1;
    }
    __syncthreads();  // q / knew / vnew complete; also publishes the producer's mbarrier initialisation
    if (has_new && g0 == 0) {  // cache append (decoder_self_attention.cu:126,172)
        const size_t off = row_off(step - 1);
        for (int j = tid; j < D; j += kAttnThreads) {
            kc[off + j] = Elem<T>::from_f(knew[j]);
            vc[off + j] = Elem<T>::from_f(vnew[j]);
        }
    }
    pdl_launch_dependents();

    if (warp == kAttnWarps) {
        // ================================================= TMA producer
        if (producer) {
            if (!a.prefetch)
                while (p_t < kAttnStages && p_t < ntiles) issue_next();
            int e_s = 0, e_ph = 0;
            while (p_t < ntiles) {
                a_mbar_wait(empty0 + 8 * e_s, e_ph);
                if (++e_s == kAttnStages) e_s = 0, e_ph ^= 1;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_next();
            }
        }
    } else {
        // ================================================= compute warps: private online softmax per row group
        const int rg = tid / LPR, l = tid % LPR;
        const float scale = rsqrtf((float)D);
        float qf[G][V];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int e = 0; e < V; ++e) qf[g][e] = qs[g * D + l * V + e];
        float m[G], sum[G], of[G][V];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            m[g] = -INFINITY, sum[g] = 0.0f;
#pragma unroll
            for (int e = 0; e < V; ++e) of[g][e] = 0.0f;
        }
        int s = 0, ph = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int base = p0 + t * TP;
            a_mbar_wait(full0 + 8 * s, ph);
            const unsigned char *kt = ring + (size_t)s * kStageBytes, *vt = kt + kAttnTileBytes;
            // ---- logits of this row group's RPG rows (rows past the chunk end read as -inf)
            float lg[G][RPG];
            bool ok[RPG];
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
                const int r = rg + i * RG, p = base + r;
                ok[i] = p < p1;
                float kf[V];
                unpack16<T>(*reinterpret_cast<const uint4 *>(kt + (size_t)r * (D * sizeof(T)) + l * 16), kf);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float d = 0.0f;
#pragma unroll
                    for (int e = 0; e < V; ++e) d = fmaf(qf[g][e], kf[e], d);
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    lg[g][i] = ok[i] ? d * scale : -INFINITY;
                }
            }
            // ---- online softmax update: lane i < RPG of the row computes exp(l_i - m_new), lane RPG the rescale factor
            float pw[G][RPG];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float mt = m[g];
#pragma unroll
                for (int i = 0; i < RPG; ++i) mt = fmaxf(mt, lg[g][i]);
                const float ms = mt == -INFINITY ? 0.0f : mt;
                float mine = m[g];
#pragma unroll
                for (int i = 0; i < RPG; ++i) mine = l == i ? lg[g][i] : mine;
                const float ex = expf(mine - ms);  // lanes >= RPG: the rescale factor exp(m_old - m_new)
                const unsigned lbase = (unsigned)((tid & 31) - l);
                float ps = 0.0f;
#pragma unroll
                for (int i = 0; i < RPG; ++i) {
                    pw[g][i] = __shfl_sync(0xffffffffu, ex, lbase + i);
                    ps += pw[g][i];
                }
                const float rs = __shfl_sync(0xffffffffu, ex, lbase + RPG);
                sum[g] = fmaf(sum[g], rs, ps);
                m[g] = mt;
#pragma unroll
                for (int e = 0; e < V; ++e) of[g][e] *= rs;
            }
            // ---- P.V
#pragma unroll
            for (int i = 0; i < RPG; ++i) {
                const int r = rg + i * RG;
                float vf[V];
                if (ok[i]) {
                    unpack16<T>(*reinterpret_cast<const uint4 *>(vt + (size_t)r * (D * sizeof(T)) + l * 16), vf);
                } else {
#pragma unroll
                    for (int e = 0; e < V; ++e) vf[e] = 0.0f;  // never multiply stale shared memory (could hold NaN bits)
                }
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int e = 0; e < V; ++e) of[g][e] = fmaf(pw[g][i], vf[e], of[g][e]);
            }
            __syncwarp();
            if ((tid & 31) == 0) a_mbar_arrive(empty0 + 8 * s);
            if (++s == kAttnStages) s = 0, ph ^= 1;
        }
        // ---- the token being appended: one more row, held in shared memory (never read back from the cache), row group 0 only
        if (has_new && rg == 0) {
            float kf[V], vf[V];
#pragma unroll
            for (int e = 0; e < V; ++e) kf[e] = knew[l * V + e], vf[e] = vnew[l * V + e];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float d = 0.0f;
#pragma unroll
                for (int e = 0; e < V; ++e) d = fmaf(qf[g][e], kf[e], d);
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(LPR == 32 ? 0xffffffffu : 0x0000ffffu, d, o);
                const float lgn = d * scale, mt = fmaxf(m[g], lgn);
                const float rs = expf(m[g] - mt), pn = expf(lgn - mt);  // m == -inf (no cached row in this group): rs = 0
                sum[g] = fmaf(sum[g], rs, pn);
                m[g] = mt;
#pragma unroll
                for (int e = 0; e < V; ++e) of[g][e] = fmaf(pn, vf[e], of[g][e] * rs);
            }
        }
        // ---- publish the row group's state (the ring is dead: every tile has been consumed by every warp after the barrier)
        asm volatile("bar.sync 1, %0;" ::"r"(kAttnWarps * 32) : "memory");
#pragma unroll
        for (int g = 0; g < G; ++g) {
            float *mp = merge + ((size_t)rg * G + g) * (D + 2);
#pragma unroll
            for (int e = 0; e < V; ++e) mp[l * V + e] = of[g][e];
            if (l == 0) mp[D] = m[g], mp[D + 1] = sum[g];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kAttnWarps * 32) : "memory");
        // merge weights exp(m_rg - M) per (rg, g); M and the merged sum per g: warp g, lane = row group (fixed shuffle order)
        if (warp < G) {
            const int g = warp, r = tid & 31;
            const float *mp = merge + ((size_t)(r < RG ? r : 0) * G + g) * (D + 2);
            const float mr = r < RG ? mp[D] : -INFINITY, sr = r < RG ? mp[D + 1] : 0.0f;
            const float M = warp_max(mr);
            const float w = M == -INFINITY ? 0.0f : expf(mr - M);  // a row group without rows has m = -inf: weight 0
            const float S = warp_sum(sr * w);
            if (r < RG) wts[r * G + g] = w;
            if (r == 0) wts[RG * G + g] = M, wts[RG * G + G + g] = S;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kAttnWarps * 32) : "memory");
    }
    __syncthreads();

    // ---- combine the row groups; single split: finish here, else hand the partial to the merge
    T *out = reinterpret_cast<T *>(a.out) + ((size_t)b * H + (size_t)kvh * Gtot + g0) * D;
    if (a.nsplit == 1) {
        for (int i = tid; i < G * D; i += kAttnThreads) {
            const int g = i / D, d = i % D;
            float o = 0.0f;
#pragma unroll
            for (int r = 0; r < RG; ++r) o = fmaf(merge[((size_t)r * G + g) * (D + 2) + d], wts[r * G + g], o);
            const float mm = wts[RG * G + g], mf = final_max(mm, step, D), c = expf(mm - mf);
            out[(size_t)g * D + d] = Elem<T>::from_f(o * c / (wts[RG * G + G + g] * c + 1e-6f));
        }
        return;
    }
    if (a.ll_merge) {
        // ---- flagged words [b, kv head][split][q head][o[D], max, sum, pad, pad] x {value, flag}: an aligned 8-byte store is never torn,
        //      so a word whose flag reads 1 carries its value.  The region is zero between launches: the merger clears what it consumed
        //      (the next launch that touches these words is several kernels away in the stream).
        uint2 *words = reinterpret_cast<uint2 *>(a.partials);
        const size_t rec0 = (((size_t)b * Hkv + kvh) * a.nsplit * Gtot + g0) * (size_t)PS;  // record of split 0, head g0
        const size_t sstride = (size_t)Gtot * PS;                                            // words between consecutive splits
        const int merger = a.nsplit - 1;  // launched after every split it waits for (blockIdx.x is the fastest grid dimension)
        if (split != merger) {
            uint2 *mine = words + rec0 + (size_t)split * sstride;
            for (int i = tid; i < G * D; i += kAttnThreads) {
                const int g = i / D, d = i % D;
                float o = 0.0f;
#pragma unroll
                for (int r = 0; r < RG; ++r) o = fmaf(merge[((size_t)r * G + g) * (D + 2) + d], wts[r * G + g], o);
                __stcg(mine + (size_t)g * PS + d, make_uint2(__float_as_uint(o), 1u));
            }
            if (tid < 2 * G) {
                const int g = tid >> 1, which = tid & 1;  // 0: max, 1: sum
                __stcg(mine + (size_t)g * PS + D + which, make_uint2(__float_as_uint(wts[RG * G + which * G + g]), 1u));
            }
            return;
        }
        // ---- the merger: its own partial never leaves shared memory; the others are polled (all loads of a pass in flight, then only the
        //      words that had not landed), merged in split order with the arithmetic of the ticket path below -- bit-identical results
        constexpr int kMaxOther = kAttnMaxSplits - 1;
        const long long t_end = clock64() + 2000000000ll;  // ~1 s: a split that never publishes must not hang the GPU
        // the n words p[j * stride], j < n, as floats: every load of a pass in flight, then only the words that had not landed yet
        auto poll = [&](const uint2 *p, int n, size_t stride, float (&out)[kMaxOther]) {
            unsigned pending = (1u << n) - 1u;
            while (pending) {
                uint2 v[kMaxOther];
#pragma unroll
                for (int j = 0; j < kMaxOther; ++j)
                    if (pending >> j & 1u) asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[j].x), "=r"(v[j].y) : "l"(p + (size_t)j * stride));
#pragma unroll
                for (int j = 0; j < kMaxOther; ++j)
                    if ((pending >> j & 1u) && v[j].y == 1u) out[j] = __uint_as_float(v[j].x), pending &= ~(1u << j);
                if (pending && clock64() > t_end) {
#pragma unroll
                    for (int j = 0; j < kMaxOther; ++j)
                        if (pending >> j & 1u) out[j] = __int_as_float(0x7fc00000);  // NaN: obvious, not plausible
                    break;
                }
            }
        };
        const int nother = a.nsplit - 1;  // splits 0 .. nsplit-2, in this order
        const uint2 *others = words + rec0;
        for (int i = tid; i < G * D; i += kAttnThreads) {
            const int g = i / D, d = i % D;
            float o_own = 0.0f;
#pragma unroll
            for (int r = 0; r < RG; ++r) o_own = fmaf(merge[((size_t)r * G + g) * (D + 2) + d], wts[r * G + g], o_own);
            const float m_own = wts[RG * G + g], s_own = wts[RG * G + G + g];
            float cw[kMaxOther], val[kMaxOther];  // three passes over one buffer: max -> merge weights, then the sums, then the outputs
            poll(others + (size_t)g * PS + D, nother, sstride, cw);
            float mm = m_own;
#pragma unroll
            for (int j = 0; j < kMaxOther; ++j)
                if (j < nother) mm = fmaxf(mm, cw[j]);
            mm = final_max(mm, step, D);
#pragma unroll
            for (int j = 0; j < kMaxOther; ++j)
                if (j < nother) cw[j] = expf(cw[j] - mm);
            float ssum = 0.0f, o = 0.0f;
            poll(others + (size_t)g * PS + D + 1, nother, sstride, val);
#pragma unroll
            for (int j = 0; j < kMaxOther; ++j)
                if (j < nother) ssum = fmaf(val[j], cw[j], ssum);
            poll(others + (size_t)g * PS + d, nother, sstride, val);
#pragma unroll
            for (int j = 0; j < kMaxOther; ++j)
                if (j < nother) o = fmaf(val[j], cw[j], o);
            {
                const float c = expf(m_own - mm);  // the merger is the last split
                ssum = fmaf(s_own, c, ssum);
                o = fmaf(o_own, c, o);
            }
            out[(size_t)g * D + d] = Elem<T>::from_f(o / (ssum + 1e-6f));
            for (int j = 0; j < nother; ++j) __stcg(const_cast<uint2 *>(others) + (size_t)j * sstride + (size_t)g * PS + d, make_uint2(0u, 0u));
        }
        __syncthreads();  // every thread has read the (max, sum) words of its head
        for (int i = tid; i < 2 * G * nother; i += kAttnThreads) {
            const int j = i / (2 * G), g = (i % (2 * G)) >> 1, which = i & 1;
            __stcg(const_cast<uint2 *>(others) + (size_t)j * sstride + (size_t)g * PS + D + which, make_uint2(0u, 0u));
        }
        return;
    }
    // ---- partials: [b, kv head][split][q head of the group][o[D], max, sum, pad, pad]
    float *part = a.partials + ((((size_t)b * Hkv + kvh) * a.nsplit + split) * Gtot + g0) * (size_t)PS;
    for (int i = tid; i < G * D; i += kAttnThreads) {
        const int g = i / D, d = i % D;
        float o = 0.0f;
#pragma unroll
        for (int r = 0; r < RG; ++r) o = fmaf(merge[((size_t)r * G + g) * (D + 2) + d], wts[r * G + g], o);
        part[(size_t)g * PS + d] = o;
    }
    if (tid < G) {
        part[(size_t)tid * PS + D] = wts[RG * G + tid];
        part[(size_t)tid * PS + D + 1] = wts[RG * G + G + tid];
    }
    // ---- the last CTA of the (b, kv head, half group) (self-resetting ticket) merges
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicInc(&a.tickets[(b * Hkv + kvh) * gsplit + (int)(blockIdx.y % gsplit)], a.nsplit - 1) == (unsigned)(a.nsplit - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const float *pbase = a.partials + (((size_t)b * Hkv + kvh) * a.nsplit * Gtot + g0) * (size_t)PS;
    const size_t sstride = (size_t)Gtot * PS;  // floats between the records of consecutive splits
    // Every split's (max, sum, o[d]) is requested before the first one is used: the loads are independent L2 round trips (~0.6 us each
    // under load), and a rolled `for (s2 < nsplit)` loop serialises them.  Fixed split order: deterministic.
    constexpr int kBatch = 16;
    for (int i = tid; i < G * D; i += kAttnThreads) {
        const int g = i / D, d = i % D;
        const float *pg = pbase + (size_t)g * PS;
        float mm = -INFINITY;
        for (int s0 = 0; s0 < a.nsplit; s0 += kBatch) {
            float mv[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) mv[j] = s0 + j < a.nsplit ? __ldcg(pg + (size_t)(s0 + j) * sstride + D) : -INFINITY;
#pragma unroll
            for (int j = 0; j < kBatch; ++j) mm = fmaxf(mm, mv[j]);
        }
        mm = final_max(mm, step, D);
        float ssum = 0.0f, o = 0.0f;
        for (int s0 = 0; s0 < a.nsplit; s0 += kBatch) {
            float mv[kBatch], sv[kBatch], ov[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                mv[j] = -INFINITY, sv[j] = 0.0f, ov[j] = 0.0f;
                if (s0 + j < a.nsplit) {
                    const float *ps = pg + (size_t)(s0 + j) * sstride;
                    mv[j] = __ldcg(ps + D), sv[j] = __ldcg(ps + D + 1), ov[j] = __ldcg(ps + d);
                }
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j)
                if (s0 + j < a.nsplit) {
                    const float c = expf(mv[j] - mm);
                    ssum = fmaf(sv[j], c, ssum);
                    o = fmaf(ov[j], c, o);
                }
        }
        out[(size_t)g * D + d] = Elem<T>::from_f(o / (ssum + 1e-6f));
    }
}

// ------------------------------------------------------------------ any head size (toy shapes of the reference's examples)
template <typename T>
__global__ void decode_attn_generic_kernel(const DecodeAttnArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = a.head_size;
    float *q = reinterpret_cast<float *>(smem_raw);  // [D]
    float *knew = q + D, *vnew = knew + D;           // [D] each
    float *ls = vnew + D;                            // [step]
    __shared__ float red[33];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int H = a.head_num, Hkv = a.kv_head_num, rep = H / Hkv, kvh = h / rep;
    const int step = a.steps ? min(max(__ldg(a.steps + b), 1), a.step) : a.step;
    const T *qkv = reinterpret_cast<const T *>(a.qkv) + (size_t)b * (H + 2 * Hkv) * D;
    const T *bias = reinterpret_cast<const T *>(a.bias);
    T *kc = reinterpret_cast<T *>(a.k_cache) + ((size_t)b * Hkv + kvh) * a.max_seq_len * D;
    T *vc = reinterpret_cast<T *>(a.v_cache) + ((size_t)b * Hkv + kvh) * a.max_seq_len * D;
    pdl_wait();
    for (int j = tid; j < D; j += blockDim.x) {
        const int half = D / 2;
        const int qh = h, kh = H + kvh, vh = H + Hkv + kvh;
        float qv = Elem<T>::to_f(qkv[(size_t)qh * D + j]), kv = Elem<T>::to_f(qkv[(size_t)kh * D + j]);
        if (a.apply_rope && (j % half) < a.rot_dim / 2 && j < 2 * half) {
            const int i = j % half;
            const float den = rope_denominator(a.rot_base, 2 * i, a.rot_dim), pos = (float)(step - 1);
            float q0 = Elem<T>::to_f(qkv[(size_t)qh * D + i]), q1 = Elem<T>::to_f(qkv[(size_t)qh * D + i + half]);
            float k0 = Elem<T>::to_f(qkv[(size_t)kh * D + i]), k1 = Elem<T>::to_f(qkv[(size_t)kh * D + i + half]);
            rope_rotate(q0, q1, pos, den);
            rope_rotate(k0, k1, pos, den);
            qv = round_t<T>(j < half ? q0 : q1);
            kv = round_t<T>(j < half ? k0 : k1);
        }
        float vv = Elem<T>::to_f(qkv[(size_t)vh * D + j]);
        if (bias) {
            qv = round_t<T>(qv + Elem<T>::to_f(bias[(size_t)qh * D + j]));
            kv = round_t<T>(kv + Elem<T>::to_f(bias[(size_t)kh * D + j]));
            vv = round_t<T>(vv + Elem<T>::to_f(bias[(size_t)vh * D + j]));
        }
        q[j] = qv, knew[j] = kv, vnew[j] = vv;
        if (h % rep == 0) {
            kc[(size_t)(step - 1) * D + j] = Elem<T>::from_f(kv);
            vc[(size_t)(step - 1) * D + j] = Elem<T>::from_f(vv);
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)D);
    float m = -INFINITY;
    for (int p = tid; p < step; p += blockDim.x) {
        float s = 0.0f;
        for (int j = 0; j < D; ++j) s = fmaf(q[j], p == step - 1 ? knew[j] : Elem<T>::to_f(kc[(size_t)p * D + j]), s);
        s *= scale;
        ls[p] = s;
        m = fmaxf(m, s);
    }
    m = final_max(block_max(m, red), step, D);
    float sum = 0.0f;
    for (int p = tid; p < step; p += blockDim.x) {
        const float e = expf(ls[p] - m);
        ls[p] = e;
        sum += e;
    }
    sum = block_sum(sum, red) + 1e-6f;
    T *out = reinterpret_cast<T *>(a.out) + ((size_t)b * H + h) * D;
    for (int j = tid; j < D; j += blockDim.x) {
        float o = 0.0f;
        for (int p = 0; p < step; ++p) o = fmaf(ls[p], p == step - 1 ? vnew[j] : Elem<T>::to_f(vc[(size_t)p * D + j]), o);
        out[j] = Elem<T>::from_f(o / sum);
    }
}

size_t decode_attn_partials_floats(int batch, int head_num, int kv_head_num, int head_size, int max_splits) {
    return (size_t)batch * kv_head_num * max_splits * (head_num / kv_head_num) * attn_part_stride(head_size);  // ll_merge: twice as many
}

int decode_attn_plan(int batch, int kv_head_num, int step, int *chunk) {
    // The cached positions [0, step-1) are cut into chunks of whole 64-position tiles, about ONE CTA per SM (measured on the 7B
    // step, ctx 1024: 4 splits x 32 heads = 128 CTAs 2.535 / 2.616 ms, 8 splits = 256 CTAs 2.550 / 2.625 ms on two boxes -- a CTA with
    // a 3-stage ring streams its share of the cache as fast as two, and the merge reads half as many partials) and at most kAttnMaxSplits = 8 splits
    // (the merge requests the partials in one batch); the last split also serves the token being appended, which never comes from
    // the cache.
    const int cached = step - 1;
    int want = sm_count() / (batch * kv_head_num);
    if (want < 1) want = 1;
    if (want > kAttnMaxSplits) want = kAttnMaxSplits;
    int c = (cached + want - 1) / want;
    c = (c + 63) & ~63;
    if (c < 64) c = 64;
    *chunk = c;
    const int n = (cached + c - 1) / c;
    return n < 1 ? 1 : n;
}

template <typename T>
static int launch_decode_attn_t(DecodeAttnArgs a, cudaStream_t st) {
    const int G = a.head_num / a.kv_head_num;
    const bool fast = a.head_size == kAttnD && (G == 1 || G == 2 || G == 4 || G == 8) && aligned16(a.k_cache) && aligned16(a.v_cache);
    if (fast) {
        const int GC = G == 8 ? 4 : G;  // q heads per CTA
        const int RG = kAttnWarps * 32 / (kAttnD / Elem<T>::kVec);
        const size_t smem = (size_t)kAttnStages * 2 * kAttnTileBytes + sizeof(float) * ((size_t)GC * kAttnD + 2 * kAttnD + (size_t)RG * GC + 2 * GC + 2) +
                            2 * kAttnStages * 8 + 16;
        dim3 grid(a.nsplit, a.kv_head_num * (G / GC), a.batch);
        auto go = [&](auto kern, int slot) {
            // the three instantiations share one function-pointer type: the opt-in to large dynamic shared memory is per kernel and device
            static thread_local bool attr_set[6][64] = {{false}};
            int dev = 0;
            cudaGetDevice(&dev);
            dev &= 63;
            if (!attr_set[slot][dev]) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                attr_set[slot][dev] = true;
            }
            launch_pdl(kern, grid, dim3(kAttnThreads), smem, st, true, a);
        };
        if (a.block_table) {
            switch (GC) {
                case 1: go(decode_attn_kernel<T, 1, true>, 3); break;
                case 2: go(decode_attn_kernel<T, 2, true>, 4); break;
                default: go(decode_attn_kernel<T, 4, true>, 5); break;
            }
        } else {
            switch (GC) {
                case 1: go(decode_attn_kernel<T, 1, false>, 0); break;
                case 2: go(decode_attn_kernel<T, 2, false>, 1); break;
                default: go(decode_attn_kernel<T, 4, false>, 2); break;
            }
        }
        return cuda_status("decode_attn launch");
    }
    if (a.block_table) return B200_ERR_UNSUPPORTED;  // the paged cache is served by the head-size-128 kernel only
    const size_t smem = sizeof(float) * ((size_t)3 * a.head_size + a.step);
    if (smem > 200 * 1024) {
        set_error("decode_mha: step %d too long for the generic head-size path", a.step);
        return B200_ERR_UNSUPPORTED;
    }
    cudaFuncSetAttribute(decode_attn_generic_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(decode_attn_generic_kernel<T>, dim3(a.head_num, a.batch), dim3(128), smem, st, true, a);
    return cuda_status("decode_attn_generic launch");
}

int launch_decode_attn(const DecodeAttnArgs &a, int dtype, cudaStream_t st) {
    B200_DISPATCH_DTYPE(dtype, return launch_decode_attn_t<T>(a, st));
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_rope_decode(void *qkv, int batch, int head_num, int kv_head_num, int head_size, int step, int rotary_dim,
                     float rotary_base, int dtype, b200_stream_t stream) {
    B200_REQUIRE(qkv, "rope_decode: null qkv");
    B200_REQUIRE(batch >= 0 && head_num > 0 && kv_head_num > 0 && head_size > 0 && step >= 1, "rope_decode: bad shape");
    B200_REQUIRE(rotary_dim >= 0 && rotary_dim <= head_size && rotary_dim % 2 == 0, "rope_decode: bad rotary_dim %d", rotary_dim);
    if (batch == 0 || rotary_dim == 0) return B200_OK;
    dim3 grid(head_num + kv_head_num, batch);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(rope_decode_kernel<T>, grid, dim3(64), 0, as_stream(stream), true, (T *)qkv, head_num,
                                          kv_head_num, head_size, step, rotary_dim, rotary_base));
    return cuda_status("rope_decode launch");
}

static int decode_mha_impl(const void *qkv, const void *qkv_bias, void *k_cache, void *v_cache, void *out, const int *steps,
                           const int *block_table, int max_pages, int num_pages, int batch, int head_num, int kv_head_num, int head_size,
                           int max_seq_len, int step, int layer, int apply_rope, int rotary_dim, float rotary_base, int dtype,
                           b200_stream_t stream) {
    B200_REQUIRE(qkv && k_cache && v_cache && out, "decode_mha: null pointer");
    B200_REQUIRE(batch >= 0 && head_num > 0 && kv_head_num > 0 && head_size > 0, "decode_mha: bad shape");
    B200_REQUIRE(head_num % kv_head_num == 0, "decode_mha: head_num %d not a multiple of kv_head_num %d", head_num, kv_head_num);
    B200_REQUIRE(step >= 1 && step <= max_seq_len, "decode_mha: step %d outside [1, max_seq_len=%d]", step, max_seq_len);
    B200_REQUIRE(layer >= 0, "decode_mha: negative layer");
    B200_REQUIRE(!apply_rope || (rotary_dim >= 0 && rotary_dim <= head_size && rotary_dim % 2 == 0), "decode_mha: bad rotary_dim");
    if (batch == 0) return B200_OK;
    Workspace ws;
    if (!get_workspace(&ws)) return B200_ERR_WORKSPACE;
    const size_t eb = dtype == B200_F32 ? 4 : 2;
    // contiguous cache [L, B, Hkv, S, d]; paged pool [L, num_pages, Hkv, kAttnPageSize, d]
    const size_t layer_off = block_table ? (size_t)layer * num_pages * kv_head_num * kAttnPageSize * head_size * eb
                                         : (size_t)layer * batch * kv_head_num * max_seq_len * head_size * eb;
    DecodeAttnArgs a = {};
    a.qkv = qkv, a.bias = qkv_bias;
    a.k_cache = (char *)k_cache + layer_off, a.v_cache = (char *)v_cache + layer_off;
    a.out = out;
    a.batch = batch, a.head_num = head_num, a.kv_head_num = kv_head_num, a.head_size = head_size;
    a.max_seq_len = max_seq_len, a.step = step, a.steps = steps;
    a.block_table = block_table, a.max_pages = max_pages;
    a.apply_rope = apply_rope, a.rot_dim = rotary_dim, a.rot_base = rotary_base;
    a.nsplit = decode_attn_plan(batch, kv_head_num, step, &a.chunk);
    a.partials = reinterpret_cast<float *>(ws.scratch);
    a.tickets = ws.tickets;
    B200_REQUIRE((size_t)batch * kv_head_num * 2 <= ws.n_tickets, "decode_mha: batch*kv_head_num exceeds the ticket pool");
    B200_REQUIRE(decode_attn_partials_floats(batch, head_num, kv_head_num, head_size, a.nsplit) * 4 <= ws.scratch_bytes,
                 "decode_mha: library workspace too small for %d splits", a.nsplit);
    const int rc = launch_decode_attn(a, dtype, as_stream(stream));
    if (rc == B200_ERR_UNSUPPORTED && block_table) set_error("decode_mha_paged: head size %d / group %d not served by the paged kernel (head size 128, group 1/2/4/8)", head_size, head_num / kv_head_num);
    return rc;
}

int b200_decode_mha(const void *qkv, const void *qkv_bias, void *k_cache, void *v_cache, void *out, const uint8_t *finished,
                    int batch, int head_num, int kv_head_num, int head_size, int max_seq_len, int step, int layer,
                    int apply_rope, int rotary_dim, float rotary_base, int dtype, b200_stream_t stream) {
    (void)finished;  // unused by the reference kernel as well
    return decode_mha_impl(qkv, qkv_bias, k_cache, v_cache, out, nullptr, nullptr, 0, 0, batch, head_num, kv_head_num, head_size, max_seq_len, step,
                           layer, apply_rope, rotary_dim, rotary_base, dtype, stream);
}

int b200_decode_mha_ragged(const void *qkv, const void *qkv_bias, void *k_cache, void *v_cache, void *out, const int *steps, int batch,
                           int head_num, int kv_head_num, int head_size, int max_seq_len, int max_step, int layer, int apply_rope,
                           int rotary_dim, float rotary_base, int dtype, b200_stream_t stream) {
    B200_REQUIRE(steps, "decode_mha_ragged: null steps");
    return decode_mha_impl(qkv, qkv_bias, k_cache, v_cache, out, steps, nullptr, 0, 0, batch, head_num, kv_head_num, head_size, max_seq_len,
                           max_step, layer, apply_rope, rotary_dim, rotary_base, dtype, stream);
}

int b200_decode_mha_paged(const void *qkv, const void *qkv_bias, void *k_pool, void *v_pool, void *out, const int *block_table,
                          const int *steps, int batch, int head_num, int kv_head_num, int head_size, int num_pages, int max_pages_per_seq,
                          int max_step, int layer, int apply_rope, int rotary_dim, float rotary_base, int dtype, b200_stream_t stream) {
    B200_REQUIRE(block_table && steps, "decode_mha_paged: null block table / steps");
    B200_REQUIRE(num_pages >= 1 && max_pages_per_seq >= 1, "decode_mha_paged: bad pool (num_pages %d, max_pages_per_seq %d)", num_pages, max_pages_per_seq);
    return decode_mha_impl(qkv, qkv_bias, k_pool, v_pool, out, steps, block_table, max_pages_per_seq, num_pages, batch, head_num, kv_head_num,
                           head_size, max_pages_per_seq * kAttnPageSize, max_step, layer, apply_rope, rotary_dim, rotary_base, dtype, stream);
}

}  // extern "C"
