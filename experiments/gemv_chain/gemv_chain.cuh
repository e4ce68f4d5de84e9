// gemv_chain.cuh -- several dependent decode-shaped linears in ONE persistent kernel (dense weights, M <= 4):
//
//     O projection -> [residual add + bias + RMSNorm] gate/up + SwiGLU -> down projection -> [residual add + RMSNorm] next layer's QKV
//
// (reference src/layers/self_attention.cpp:131-138, src/layers/ffn.cpp:105-140, src/layers/self_decoder.cpp:69-119,
// src/kernels/add_residual_and_rmsnorm.cu:43-121, rmsnorm.cu:35-80, silu_and_mul.cu:6-41.)
//
// Why: between two weight-streaming GEMV launches HBM idles (CTA relaunch, first-byte latency, and -- measured with the
// %globaltimer trace below -- 4-7 us of skew between the fastest and the slowest SM of a statically partitioned launch).  Here
//   * the TMA producer warps never wait for a phase boundary: the weights of phase p+1 do not depend on phase p, so the ring
//     (192 KB per SM) keeps filling while the last outputs of phase p are reduced, published and re-staged;
//   * the unit id travels with the ring stage (a small shared array next to the mbarriers) and an empty "end of phase" stage closes
//     a phase, so the consumers do not need to know the partition.  (Claiming units from ONE global counter was tried: same-address
//     atomics retire at ~1 per 5 ns on B200, which made the claims themselves the bottleneck -- 329 vs 374 tok/s.)
//   * there is no grid barrier.  Activations that cross a phase boundary travel as 8-byte {payload, flag} words (the "LL" scheme of
//     collective libraries): the reducer stores value(s) and flag in ONE 8-byte store, the stagers of the next phase poll the words
//     they need.  No fence, no counter, no second round trip: the measured fence + atomic + poll of a counter barrier was 2.2 us.
//     The buffers are per layer and zeroed (flags and claim counters, one memset) at the start of every step.
//   * the residual stream is not exchanged between CTAs inside a launch: phase 3 recomputes  res1 = y_attn + res0  from the LL copy
//     of y_attn and the residual the previous launch left, exactly the value phase 1 formed; CTA 0 writes the residual stream back
//     for the next launch (a kernel boundary orders that).
//
// Same skeleton as gemv_nk_kernel (gemv.cuh): one CTA per SM, 2 groups x 8 compute warps, one producer and one reducer warp per
// group, units of 2 weight rows, K split over the 8 warps, per-lane partials summed by the reducer in a fixed order.  All CTAs
// must be co-resident (a stager polls data other CTAs produce): grid <= number of SMs, one CTA per SM; a poll that does not
// complete within ~2 s traps instead of hanging the GPU.
#pragma once
#include "gemv.cuh"

namespace b200 {

constexpr int kChainMaxPhases = 4;
constexpr int kChainStagers = (kGemvWarps + kGemvGroups) * 32;  // compute + reducer warps
constexpr unsigned int kChainFlag = 1u;                         // LL flag value (buffers are zeroed every step)

// elements of T per 8-byte LL word {payload, flag}
template <typename T> struct ChainLL {
    static constexpr int kEPW = sizeof(T) == 2 ? 2 : 1;
};

struct ChainPhase {
    const void *w;        // [N, K] of T
    const void *x;        // plain [M, K] input                                  (when x_ll == NULL)
    const uint2 *x_ll;    // LL input written by an earlier phase of this launch: [M][K / EPW] words
    void *y;              // plain output [M, n_out]                             (when y_ll == NULL)
    uint2 *y_ll;          // LL output [M][n_out / EPW] words, consumed by a later phase of this launch
    // prologue (norm != 0): r = res_in (or round(ll(res_ll) + res_in) when res_ll); o = round(x + r); res_out <- o (CTA 0);
    //                       o = round(o + bias); xs = gamma * o * rsqrt(mean(o^2) + eps)   (gamma == NULL: xs = o)
    const void *res_in;
    const uint2 *res_ll;  // LL tensor [M][K / EPW] of an earlier phase (complete by the time it is read; polled like any LL word)
    void *res_out;
    const void *bias;
    const void *gamma;
    int norm;
    int K, N;
    int inter;   // swiglu: rows (i, inter + i) form a unit, n_out = inter
    int swiglu;
    int pieces, piece_bytes, cw;  // ring geometry of this phase (as GemvGeom)
    int xv;                       // register-resident activation vectors per token per warp (0: read x from shared memory)
};

struct ChainArgs {
    ChainPhase ph[kChainMaxPhases];
    int n_phases;
    int M;
    float eps;
    int stages, stage_bytes;    // ring geometry common to all phases (stage_bytes = 2 x the largest aligned piece)
    int xs_elems;               // elements of T reserved per token row in shared memory (>= max K)
    unsigned int *claim;        // reserved (zeroed with the LL buffers)
    int poll_ns;                // back-off between two polls of an LL word
    int l2_ahead;               // ring stages' worth of weights the L2 prefetcher keeps requested in front of the ring (0: off)
    unsigned long long *trace;  // optional [gridDim.x][kChainMaxPhases][8] globaltimer stamps (diagnostics), or NULL
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint4 ld_cg_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void *p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_ll(uint2 *p, uint32_t payload) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(kChainFlag) : "memory");
}

// One 16-byte vector's worth of T (V elements = 4 LL words = 32 bytes) from an LL tensor: spins until all four flags are set.
template <typename T>
__device__ __forceinline__ void ll_read_vec(const uint2 *words, float *f, unsigned backoff_ns = 40) {
    uint4 a, b;
    const long long t0 = clock64();
    for (;;) {
        a = ld_relaxed_v4(words);
        b = ld_relaxed_v4(words + 2);
        if (a.y == kChainFlag && a.w == kChainFlag && b.y == kChainFlag && b.w == kChainFlag) break;
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: the producing CTA never ran (grid not co-resident)
        __nanosleep(backoff_ns);
    }
    if constexpr (sizeof(T) == 2) {
        unpack16<T>(make_uint4(a.x, a.z, b.x, b.z), f);
    } else {
        f[0] = __uint_as_float(a.x), f[1] = __uint_as_float(a.z), f[2] = __uint_as_float(b.x), f[3] = __uint_as_float(b.z);
    }
}

// sum over the kChainStagers staging threads, broadcast; red = shared float[33]
__device__ __forceinline__ float chain_stager_sum(float v, float *red, int swarp, int lane) {
    constexpr int NW = kChainStagers / 32;
    v = warp_sum(v);
    named_bar_sync(1, kChainStagers);  // protect `red` from a previous use
    if (lane == 0) red[swarp] = v;
    named_bar_sync(1, kChainStagers);
    if (swarp == 0) {
        float t = lane < NW ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    named_bar_sync(1, kChainStagers);
    return red[32];
}

// Activation staging of one phase by the staging threads (stid in [0, kChainStagers)): same arithmetic as gemv_stage_activations.
template <typename T, int MB>
__device__ __forceinline__ void chain_stage(const ChainPhase &P, int M, float eps, T *xs, int xs_stride, float *red, int stid, unsigned poll_ns) {
    constexpr int V = Elem<T>::kVec;
    constexpr int EPW = ChainLL<T>::kEPW;
    constexpr int NT = kChainStagers;
    const int K = P.K, nv = K / V;
    const int swarp = stid >> 5, lane = stid & 31;
    const T *xin = reinterpret_cast<const T *>(P.x);
    const T *rin = P.norm ? reinterpret_cast<const T *>(P.res_in) : nullptr;
    const uint2 *rll = P.norm ? P.res_ll : nullptr;
    T *rout = P.norm ? reinterpret_cast<T *>(P.res_out) : nullptr;
    const T *bias = P.norm ? reinterpret_cast<const T *>(P.bias) : nullptr;
    const T *gamma = P.norm ? reinterpret_cast<const T *>(P.gamma) : nullptr;
    const size_t ll_row = (size_t)K / EPW;  // LL words per token row
    auto prenorm = [&](int m, int i, float *f, bool write_res) {
        // operands that do not depend on this launch's earlier phases first: their latency overlaps the poll below
        float r[V];
        if (rin) unpack16<T>(ld_cg_v4(rin + (size_t)m * K + (size_t)i * V), r);
        float b[V];
        if (bias) unpack16<T>(ld_v4(bias + (size_t)i * V), b);
        if (rin && rll) {
            float g[V];
            ll_read_vec<T>(rll + (size_t)m * ll_row + (size_t)i * 4, g, poll_ns);
#pragma unroll
            for (int j = 0; j < V; ++j) r[j] = round_to<T>(g[j] + r[j]);
        }
        if (P.x_ll) ll_read_vec<T>(P.x_ll + (size_t)m * ll_row + (size_t)i * 4, f, poll_ns);
        else unpack16<T>(ld_cg_v4(xin + (size_t)m * K + (size_t)i * V), f);
        if (rin) {
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = round_to<T>(f[j] + r[j]);
        }
        if (write_res && rout && blockIdx.x == 0) st_v4(rout + (size_t)m * K + (size_t)i * V, pack16<T>(f));
        if (bias) {
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = round_to<T>(f[j] + b[j]);
        }
    };
    auto store = [&](int m, int i, const float *f) { *reinterpret_cast<uint4 *>(xs + (size_t)m * xs_stride + (size_t)i * V) = pack16<T>(f); };
    const bool cached = nv <= kGemvXCache * NT;
#pragma unroll 1
    for (int m = 0; m < (MB == 1 ? 1 : M); ++m) {
        if (!gamma) {
            for (int i = stid; i < nv; i += NT) {
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
                const int i = stid + c * NT;
                if (i < nv) {
                    unpack16<T>(ld_v4(gamma + (size_t)i * V), gm[c]);
                    prenorm(m, i, cache[c], true);
#pragma unroll
                    for (int j = 0; j < V; ++j) ss += cache[c][j] * cache[c][j];
                }
            }
        } else {
            for (int i = stid; i < nv; i += NT) {
                float f[V];
                prenorm(m, i, f, true);
#pragma unroll
                for (int j = 0; j < V; ++j) ss += f[j] * f[j];
            }
        }
        ss = chain_stager_sum(ss, red, swarp, lane);
        const float rs = rsqrtf(ss / (float)K + eps);
        if (cached) {
#pragma unroll
            for (int c = 0; c < kGemvXCache; ++c) {
                const int i = stid + c * NT;
                if (i < nv) {
#pragma unroll
                    for (int j = 0; j < V; ++j) cache[c][j] = (cache[c][j] * gm[c][j]) * rs;
                    store(m, i, cache[c]);
                }
            }
        } else {
            for (int i = stid; i < nv; i += NT) {
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

// per-CTA pipeline state shared by the role functions
struct ChainCtx {
    uint32_t full0, empty0, ready0, free0;  // mbarrier addresses of this warp's group
    const unsigned char *ring;              // this group's ring
    float *gred;                            // this group's partial-sum slots [2][GW][R*MB][32]
    volatile int *stage_unit;               // [kGemvMaxStages] unit id of the data in each ring stage (-1: end of phase), this group
    volatile int *slot_unit;                // [2] unit id of each partial-sum slot (-1: end of phase), this group
    int stages, stage_bytes, row_stride;
    int wg, lane;
};

// compute warps: one phase.  s / ph: ring stage and parity; ug: slots this group has handed to the reducer so far (slot parity).
template <typename T, int MB, int XV>
__device__ __forceinline__ void chain_compute(const ChainPhase &P, const ChainCtx &c, const T *xs, int xs_stride, int &s, int &ph, int &ug) {
    constexpr int R = kGemvRows, GW = kGemvGW;
    constexpr int V = Elem<T>::kVec;
    const int N = P.N;
    const int nvec_row = (int)((size_t)P.K * sizeof(T) / 16);
    const int pieces = P.pieces, cw = P.cw, piece_vecs = P.piece_bytes / 16;
    const int wg = c.wg, lane = c.lane;
    auto unit_row = [&](int u, int r) -> int { return P.swiglu ? u + r * P.inter : 2 * u + r; };

    float xr[XV > 0 ? XV : 1][MB][V];
    if constexpr (XV > 0) {
#pragma unroll
        for (int i = 0; i < XV; ++i) {
            const int pc = i / 2, j = i % 2;
            const int v = pc * piece_vecs + (wg * cw + j) * 32 + lane;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                if (j < cw && pc < pieces && v < nvec_row) {
                    unpack16<T>(*reinterpret_cast<const uint4 *>(xs + (size_t)m * xs_stride + (size_t)v * V), xr[i][m]);
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

    const unsigned char *my_ring = c.ring + (size_t)(wg * cw * 32 + lane) * 16;
    for (;;) {
        // the first stage of a unit carries its id (or the end-of-phase mark)
        mbar_wait(c.full0 + s * 8, ph);
        const int u = c.stage_unit[s];
        if (u < 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(c.empty0 + s * 8);
            if (++s == c.stages) s = 0, ph ^= 1;
            break;
        }
        if constexpr (XV > 0) {
#pragma unroll
            for (int pc = 0; pc < XV / 2; ++pc) {
                if (pc < pieces) {
                    if (pc > 0) mbar_wait(c.full0 + s * 8, ph);
                    const unsigned char *st = my_ring + (size_t)s * c.stage_bytes;
                    const int pv = min(piece_vecs, nvec_row - pc * piece_vecs);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (j < cw && (wg * cw + j) * 32 + lane < pv) {
                            uint4 wv[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) wv[r] = *reinterpret_cast<const uint4 *>(st + r * c.row_stride + j * 512);
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
                    if (lane == 0) mbar_arrive(c.empty0 + s * 8);
                    if (++s == c.stages) s = 0, ph ^= 1;
                }
            }
        } else {
            const bool all_valid[R] = {true, true};
            const uint32_t zpk[R] = {0, 0};
            const float sc[R] = {0.0f, 0.0f};
            for (int pc = 0; pc < pieces; ++pc) {
                if (pc > 0) mbar_wait(c.full0 + s * 8, ph);
                const unsigned char *st = c.ring + (size_t)s * c.stage_bytes;
                const int pv = min(piece_vecs, nvec_row - pc * piece_vecs);
                const int wvb = pc * (piece_vecs / 32);
                for (int j = 0; j < cw; ++j) {
                    const int wv_i = wg * cw + j;
                    const int v = wv_i * 32 + lane;
                    if (v < pv) {
                        uint4 wv[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) wv[r] = *reinterpret_cast<const uint4 *>(st + r * c.row_stride + v * 16);
                        float acc1[R][MB];
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int m = 0; m < MB; ++m) acc1[r][m] = (j & 1) ? acc[r][m][1] : acc[r][m][0];
                        dot_rows<T, WF_DENSE, MB>(wv, all_valid, xs, xs_stride, wvb + wv_i, lane, zpk, sc, acc1);
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
                if (lane == 0) mbar_arrive(c.empty0 + s * 8);
                if (++s == c.stages) s = 0, ph ^= 1;
            }
        }
        // hand the per-lane partial sums to the reducer (double-buffered slot, counted over the whole chain)
        const int b = ug & 1;
        if (ug >= 2) mbar_wait(c.free0 + b * 8, ((ug >> 1) - 1) & 1);
        float *slot = c.gred + (size_t)b * GW * (R * MB) * 32;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = unit_row(u, r) < N;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
                slot[(wg * (R * MB) + r * MB + m) * 32 + lane] = ok ? acc[r][m][0] + acc[r][m][1] : 0.0f;
                acc[r][m][0] = acc[r][m][1] = 0.0f;
            }
        }
        if (wg == 0 && lane == 0) c.slot_unit[b] = u;
        __syncwarp();
        if (lane == 0) mbar_arrive(c.ready0 + b * 8);
        ++ug;
    }
    // end of phase: tell the reducer through the slot protocol
    const int b = ug & 1;
    if (ug >= 2) mbar_wait(c.free0 + b * 8, ((ug >> 1) - 1) & 1);
    if (wg == 0 && lane == 0) c.slot_unit[b] = -1;
    __syncwarp();
    if (lane == 0) mbar_arrive(c.ready0 + b * 8);
    ++ug;
}

// reducer warp: one phase
template <typename T, int MB>
__device__ __forceinline__ void chain_reduce(const ChainPhase &P, const ChainCtx &c, int M, int &ug) {
    constexpr int R = kGemvRows, GW = kGemvGW;
    constexpr int EPW = ChainLL<T>::kEPW;
    const int N = P.N, lane = c.lane;
    auto unit_row = [&](int u, int r) -> int { return P.swiglu ? u + r * P.inter : 2 * u + r; };
    float held[MB];  // SwiGLU, 16-bit: value of the even unit of a pair, stored together with the odd one in one LL word
#pragma unroll
    for (int m = 0; m < MB; ++m) held[m] = 0.0f;
    for (;;) {
        const int b = ug & 1;
        mbar_wait(c.ready0 + b * 8, (ug >> 1) & 1);
        ++ug;
        const int u = c.slot_unit[b];
        if (u < 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(c.free0 + b * 8);
            break;
        }
        const float *slot = c.gred + (size_t)b * GW * (R * MB) * 32;
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
            mbar_arrive(c.free0 + b * 8);
            if (P.swiglu) {
#pragma unroll
                for (int m = 0; m < MB; ++m)
                    if (m < M) {
                        // the un-fused reference stores gate/up in T before SiLU reads them
                        const float g = round_to<T>(out[0][m]), up = round_to<T>(out[1][m]);
                        const float v = (g / (1.0f + expf(-g))) * up;
                        if (!P.y_ll) {
                            reinterpret_cast<T *>(P.y)[(size_t)m * P.inter + u] = Elem<T>::from_f(v);
                        } else if constexpr (EPW == 1) {
                            st_ll(P.y_ll + (size_t)m * P.inter + u, __float_as_uint(v));
                        } else {
                            // chunks are pairs of consecutive units run back to back by this group: (even, odd)
                            if ((u & 1) == 0 && u + 1 < P.inter) {
                                held[m] = v;
                            } else {
                                const float pr[2] = {(u & 1) ? held[m] : v, (u & 1) ? v : 0.0f};
                                float f8[8] = {pr[0], pr[1], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                st_ll(P.y_ll + (size_t)m * (P.inter / 2) + (u >> 1), pack16<T>(f8).x);
                            }
                        }
                    }
            } else {
#pragma unroll
                for (int m = 0; m < MB; ++m)
                    if (m < M) {
                        if (!P.y_ll) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int row = unit_row(u, r);
                                if (row < N) reinterpret_cast<T *>(P.y)[(size_t)m * N + row] = Elem<T>::from_f(out[r][m]);
                            }
                        } else if constexpr (EPW == 1) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int row = unit_row(u, r);
                                if (row < N) st_ll(P.y_ll + (size_t)m * N + row, __float_as_uint(out[r][m]));
                            }
                        } else {
                            float f8[8] = {out[0][m], out[1][m], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // rows 2u, 2u+1 (N is even)
                            st_ll(P.y_ll + (size_t)m * (N / 2) + u, pack16<T>(f8).x);
                        }
                    }
            }
        }
    }
}

// smem: [ xs : MB * xs_elems * sizeof(T) | rings : 2 * stages * stage_bytes | barriers : 2 * (2 * kGemvMaxStages + 4) * 8 |
//         unit ids : 2 * (kGemvMaxStages + 2) ints (128 B) | partial sums : 2 * 2 * GW * R*MB * 32 floats ]
template <typename T, int MB>
__global__ void __launch_bounds__(kGemvThreads, 1) gemv_chain_kernel(const ChainArgs a) {
    constexpr int R = kGemvRows, GW = kGemvGW, NG = kGemvGroups;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ float red[33];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int n_compute = NG * GW;
    const bool is_compute = warp < n_compute;
    const int grp = is_compute ? warp / GW : (warp - n_compute) % NG;
    const bool is_producer = !is_compute && warp < n_compute + NG;
    const int stages = a.stages;

    T *xs = reinterpret_cast<T *>(smem);
    size_t off = ((size_t)MB * a.xs_elems * sizeof(T) + 127) & ~(size_t)127;
    unsigned char *ring = smem + off + (size_t)grp * stages * a.stage_bytes;
    off += (size_t)NG * stages * a.stage_bytes;
    ChainCtx c;
    c.full0 = smem_u32(smem + off) + grp * (2 * kGemvMaxStages + 4) * 8;
    c.empty0 = c.full0 + kGemvMaxStages * 8;
    c.ready0 = c.empty0 + kGemvMaxStages * 8;
    c.free0 = c.ready0 + 16;
    off += (size_t)NG * (2 * kGemvMaxStages + 4) * 8;
    int *ids = reinterpret_cast<int *>(smem + off) + grp * (kGemvMaxStages + 3);
    c.stage_unit = ids, c.slot_unit = ids + kGemvMaxStages;
    volatile int *progress = ids + kGemvMaxStages + 2;  // data stages the producer has issued so far (read by the L2 prefetcher lane)
    off += 128;
    c.gred = reinterpret_cast<float *>(smem + off) + (size_t)grp * 2 * GW * (R * MB) * 32;
    c.ring = ring;
    c.stages = stages, c.stage_bytes = a.stage_bytes, c.row_stride = a.stage_bytes / R;
    c.wg = warp % GW, c.lane = lane;

    if (is_producer) {
        // ================================================= TMA producer: claims chunks of units and streams their rows, phase after phase
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) {
                mbar_init(c.full0 + s * 8, 1);
                mbar_init(c.empty0 + s * 8, GW);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(c.ready0 + b * 8, GW);
                mbar_init(c.free0 + b * 8, 1);
            }
            fence_mbar_init();
            *progress = 0;
        }
        __syncthreads();  // (A) publish the mbarrier initialisation to the consumers
        // the producer only reads weights, which no kernel writes: it never needs griddepcontrol.wait
        pdl_launch_dependents();
        if (lane == 0) {
            const uint32_t ring_u32 = smem_u32(ring);
            int p_s = 0, issued = 0;  // ring stage to fill next; stages filled so far
            int e_s = 0, e_ph = 0;    // stage / parity of the empty barrier to wait on next
            auto acquire_stage = [&]() {
                if (issued >= stages) {  // refill a stage once all 8 warps have left it
                    mbar_wait(c.empty0 + e_s * 8, e_ph);
                    if (++e_s == stages) e_s = 0, e_ph ^= 1;
                    fence_proxy_async();
                }
            };
            int data_items = 0;  // data stages issued (the end-of-phase marks do not count)
            auto advance = [&]() {
                ++issued;
                if (++p_s == stages) p_s = 0;
            };
            const int gid = grp * gridDim.x + blockIdx.x, total_groups = gridDim.x * NG;
            unsigned long long *ptr = (a.trace && grp == 0) ? a.trace + (size_t)blockIdx.x * kChainMaxPhases * 8 : nullptr;
            for (int p = 0; p < a.n_phases; ++p) {
                if (ptr) ptr[p * 8 + 1] = global_ns();  // producer of group 0 starts issuing phase p
                const ChainPhase &P = a.ph[p];
                const size_t row_bytes = (size_t)P.K * sizeof(T);
                const int nvec_row = (int)(row_bytes / 16), piece_vecs = P.piece_bytes / 16;
                const int units = P.swiglu ? P.inter : (P.N + 1) / 2;
                // chunks of consecutive units are dealt round-robin over the groups of the grid; a SwiGLU phase with an LL output pairs
                // units (2c, 2c+1) so that its reducer can publish two 16-bit values in one LL word
                const int chunk_units = (P.swiglu && P.y_ll && sizeof(T) == 2) ? 2 : 1;
                const int nchunks = (units + chunk_units - 1) / chunk_units;
                for (int chunk = gid; chunk < nchunks; chunk += total_groups) {
                    for (int k = 0; k < chunk_units; ++k) {
                        const int u = chunk * chunk_units + k;
                        if (u >= units) break;
                        for (int pc = 0; pc < P.pieces; ++pc) {
                            acquire_stage();
                            const int v0 = pc * piece_vecs;
                            const uint32_t bytes = (uint32_t)min(piece_vecs, nvec_row - v0) * 16u;
                            const uint32_t bar = c.full0 + p_s * 8;
                            c.stage_unit[p_s] = u;
                            int nrows = 0;
#pragma unroll
                            for (int r = 0; r < R; ++r) nrows += (P.swiglu ? u + r * P.inter : 2 * u + r) < P.N ? 1 : 0;
                            mbar_expect_tx(bar, bytes * nrows);
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int row = P.swiglu ? u + r * P.inter : 2 * u + r;
                                if (row < P.N)
                                    bulk_g2s(ring_u32 + p_s * a.stage_bytes + r * c.row_stride,
                                             reinterpret_cast<const unsigned char *>(P.w) + (size_t)row * row_bytes + (size_t)v0 * 16, bytes, bar);
                            }
                            advance();
                            *progress = ++data_items;
                        }
                    }
                }
                // end of phase: an empty stage carrying the mark
                if (ptr) ptr[p * 8 + 6] = global_ns();  // ... has issued the last copy of phase p
                acquire_stage();
                c.stage_unit[p_s] = -1;
                mbar_arrive(c.full0 + p_s * 8);
                advance();
            }
        } else if (lane == 1 && a.l2_ahead > 0) {
            // ---- L2 prefetcher: walks the same sequence of row pieces `l2_ahead` stages in front of the ring.  The ring alone only holds
            // requests that are already in flight when the compute side stalls at a phase boundary; the prefetch window keeps HBM
            // streaming (into L2) through the stall, and the ring then refills from L2.
            const int gid = grp * gridDim.x + blockIdx.x, total_groups = gridDim.x * NG;
            int item = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                const ChainPhase &P = a.ph[p];
                const size_t row_bytes = (size_t)P.K * sizeof(T);
                const int nvec_row = (int)(row_bytes / 16), piece_vecs = P.piece_bytes / 16;
                const int units = P.swiglu ? P.inter : (P.N + 1) / 2;
                const int chunk_units = (P.swiglu && P.y_ll && sizeof(T) == 2) ? 2 : 1;
                const int nchunks = (units + chunk_units - 1) / chunk_units;
                for (int chunk = gid; chunk < nchunks; chunk += total_groups) {
                    for (int k = 0; k < chunk_units; ++k) {
                        const int u = chunk * chunk_units + k;
                        if (u >= units) break;
                        for (int pc = 0; pc < P.pieces; ++pc, ++item) {
                            if (item < stages) continue;  // the first ring fill is requested directly
                            while (item >= *progress + stages + a.l2_ahead) __nanosleep(100);
                            const int v0 = pc * piece_vecs;
                            const uint32_t bytes = (uint32_t)min(piece_vecs, nvec_row - v0) * 16u;
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                const int row = P.swiglu ? u + r * P.inter : 2 * u + r;
                                if (row < P.N)
                                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char *>(P.w) +
                                                                                                   (size_t)row * row_bytes + (size_t)v0 * 16),
                                                 "r"(bytes)
                                                 : "memory");
                            }
                        }
                    }
                }
            }
        }
        return;
    }

    __syncthreads();  // (A)
    pdl_wait();

    // ================================================= compute + reducer warps: phase by phase
    const int stid = is_compute ? (int)threadIdx.x : (int)threadIdx.x - NG * 32;  // index among the staging threads
    int s = 0, ph = 0, ug = 0;
    unsigned long long *tr = a.trace ? a.trace + (size_t)blockIdx.x * kChainMaxPhases * 8 : nullptr;
    for (int p = 0; p < a.n_phases; ++p) {
        const ChainPhase &P = a.ph[p];
        if (tr && stid == 0) {
            tr[p * 8 + 0] = global_ns();  // compute warp 0 is done with phase p-1 (or past griddepcontrol.wait)
            unsigned int smid;
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            tr[p * 8 + 7] = smid;
        }
        named_bar_sync(1, kChainStagers);                  // every warp of this CTA has left phase p-1: xs may be overwritten
        if (tr && stid == 0) tr[p * 8 + 2] = global_ns();
        chain_stage<T, MB>(P, a.M, a.eps, xs, a.xs_elems, red, stid, (unsigned)a.poll_ns);
        if (MB > 1) {  // padding rows of the batch tile (read by the FMAs, results discarded)
            constexpr int V = Elem<T>::kVec;
            for (int m = a.M; m < MB; ++m)
                for (int i = stid; i < P.K / V; i += kChainStagers)
                    *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xs_elems + (size_t)i * V) = make_uint4(0, 0, 0, 0);
        }
        named_bar_sync(1, kChainStagers);  // xs complete
        if (tr && stid == 0) tr[p * 8 + 3] = global_ns();  // staging done
        if (p == 0) pdl_launch_dependents();

        if (is_compute) {
            switch (P.xv) {
                case 2:
                    if constexpr (MB <= 2) {
                        chain_compute<T, MB, 2>(P, c, xs, a.xs_elems, s, ph, ug);
                        break;
                    }
                case 4:
                    if constexpr (MB == 1) {
                        chain_compute<T, MB, 4>(P, c, xs, a.xs_elems, s, ph, ug);
                        break;
                    }
                case 6:
                    if constexpr (MB == 1) {
                        chain_compute<T, MB, 6>(P, c, xs, a.xs_elems, s, ph, ug);
                        break;
                    }
                default: chain_compute<T, MB, 0>(P, c, xs, a.xs_elems, s, ph, ug);
            }
            if (tr && stid == 0) tr[p * 8 + 4] = global_ns();  // compute warp 0 finished its units of phase p
        } else {
            chain_reduce<T, MB>(P, c, a.M, ug);
            if (tr && lane == 0 && grp == 0) tr[p * 8 + 5] = global_ns();  // reducer of group 0 stored its last output
        }
    }
}

// Host side (gemv_chain_inst.cuh / gemv_chain_f32.cu).  Fills the geometry fields of `a` and launches; B200_ERR_UNSUPPORTED (no error
// text) when some phase cannot use this kernel -- the caller then runs the phases as separate GEMV launches.
// dry: only check the shapes / fill the geometry, launch nothing.
int launch_gemv_chain(ChainArgs &a, int dtype, cudaStream_t st, bool dry = false);

}  // namespace b200
