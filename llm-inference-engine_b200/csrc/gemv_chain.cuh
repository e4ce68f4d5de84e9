// gemv_chain.cuh -- several dependent decode-shaped linears in ONE persistent kernel (dense weights, M <= 4):
//
//     O projection -> [residual add + bias + RMSNorm] gate/up + SwiGLU -> down projection -> [residual add + RMSNorm] next layer's QKV
//
// (reference src/layers/self_attention.cpp:131-138, src/layers/ffn.cpp:105-140, src/layers/self_decoder.cpp:69-119,
// src/kernels/add_residual_and_rmsnorm.cu:43-121, rmsnorm.cu:35-80, silu_and_mul.cu:6-41.)
//
// Why: a kernel boundary between two weight-streaming GEMVs idles HBM for ~2 us (CTA launch + first-byte latency; a 213 KB CTA and
// its successor cannot be co-resident, so programmatic dependent launch cannot hide it).  Here the boundary is a grid-wide barrier
// that only the COMPUTE side observes: the TMA producer warps never wait on it -- the weights of phase p+1 do not depend on phase p --
// so the ring (192 KB per SM = ~4.7 us of stream) keeps filling while the last partial sums of phase p are reduced, published, and
// the activations of phase p+1 are staged (x from L2, RMSNorm, registers).  The compute warps then drain the backlog from shared
// memory at several times the HBM rate.
//
// Same skeleton as gemv_nk_kernel (gemv.cuh): one CTA per SM, 2 groups x 8 compute warps, one producer and one reducer warp per
// group, units of 2 weight rows dealt round-robin over groups, K split over the 8 warps, per-lane partials summed by the reducer in
// a fixed order (deterministic).  Differences: the ring / slot state (stage index, mbarrier parities, unit counter) runs on across
// phases; the staging is done by the 18 non-producer warps (named barrier 1); activations written by other CTAs in this launch are
// read with ld.global.cg after an acquire on the phase counter.
//
// Grid barrier: every reducer warp, after its last store of phase p, does fence + atomicAdd(sync[p]); the stagers of phase p+1 wait
// for gridDim.x * 2 arrivals.  All CTAs must be co-resident: grid <= number of SMs, one CTA per SM (212+ KB of shared memory); a
// barrier that does not complete within ~2 s traps instead of hanging the GPU.
#pragma once
#include "gemv.cuh"

namespace b200 {

constexpr int kChainMaxPhases = 4;
constexpr int kChainStagers = (kGemvWarps + kGemvGroups) * 32;  // compute + reducer warps

struct ChainPhase {
    const void *w;  // [N, K] of T
    const void *x;  // [M, K] of T
    void *y;        // [M, n_out] of T
    // prologue (norm != 0): o = x + res_in; res_out <- o; o += bias; xs = gamma * o * rsqrt(mean(o^2) + eps)  (gamma == NULL: xs = o)
    const void *res_in;
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
    int stages, stage_bytes;  // ring geometry common to all phases (stage_bytes = 2 x the largest aligned piece)
    int xs_elems;             // elements of T reserved per token row in shared memory (>= max K)
    unsigned int *sync;       // [n_phases] arrival counters, zero when the kernel starts
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
__device__ __forceinline__ void chain_stage(const ChainPhase &P, int M, float eps, T *xs, int xs_stride, float *red, int stid) {
    constexpr int V = Elem<T>::kVec;
    constexpr int NT = kChainStagers;
    const int K = P.K, nv = K / V;
    const int swarp = stid >> 5, lane = stid & 31;
    const T *xin = reinterpret_cast<const T *>(P.x);
    const T *rin = P.norm ? reinterpret_cast<const T *>(P.res_in) : nullptr;
    T *rout = P.norm ? reinterpret_cast<T *>(P.res_out) : nullptr;
    const T *bias = P.norm ? reinterpret_cast<const T *>(P.bias) : nullptr;
    const T *gamma = P.norm ? reinterpret_cast<const T *>(P.gamma) : nullptr;
    auto prenorm = [&](int m, int i, float *f, bool write_res) {
        unpack16<T>(ld_cg_v4(xin + (size_t)m * K + (size_t)i * V), f);
        if (rin) {
            float r[V];
            unpack16<T>(ld_cg_v4(rin + (size_t)m * K + (size_t)i * V), r);
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = round_to<T>(f[j] + r[j]);
        }
        if (write_res && rout && blockIdx.x == 0) st_v4(rout + (size_t)m * K + (size_t)i * V, pack16<T>(f));
        if (bias) {
            float b[V];
            unpack16<T>(ld_v4(bias + (size_t)i * V), b);
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
    int stages, stage_bytes, row_stride;
    int gid, total_groups;
    int wg, lane;
};

// compute warps: one phase.  s / ph: ring stage and parity; un0: units this group has handed to the reducer before this phase.
template <typename T, int MB, int XV>
__device__ __forceinline__ void chain_compute(const ChainPhase &P, const ChainCtx &c, const T *xs, int xs_stride, int my_units, int &s, int &ph,
                                              int un0) {
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
    for (int un = 0; un < my_units; ++un) {
        const int u = c.gid + un * c.total_groups;
        if constexpr (XV > 0) {
#pragma unroll
            for (int pc = 0; pc < XV / 2; ++pc) {
                if (pc < pieces) {
                    mbar_wait(c.full0 + s * 8, ph);
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
                mbar_wait(c.full0 + s * 8, ph);
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
        const int ug = un0 + un;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(c.ready0 + b * 8);
    }
}

// reducer warp: one phase
template <typename T, int MB>
__device__ __forceinline__ void chain_reduce(const ChainPhase &P, const ChainCtx &c, int M, int my_units, int un0) {
    constexpr int R = kGemvRows, GW = kGemvGW;
    const int N = P.N, lane = c.lane;
    auto unit_row = [&](int u, int r) -> int { return P.swiglu ? u + r * P.inter : 2 * u + r; };
    for (int un = 0; un < my_units; ++un) {
        const int u = c.gid + un * c.total_groups;
        const int ug = un0 + un;
        const int b = ug & 1;
        mbar_wait(c.ready0 + b * 8, (ug >> 1) & 1);
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
                        const float g = round_to<T>(out[0][m]), up = round_to<T>(out[1][m]);
                        const float v = (g / (1.0f + expf(-g))) * up;
                        reinterpret_cast<T *>(P.y)[(size_t)m * P.inter + u] = Elem<T>::from_f(v);
                    }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int row = unit_row(u, r);
                    if (row < N) {
#pragma unroll
                        for (int m = 0; m < MB; ++m)
                            if (m < M) reinterpret_cast<T *>(P.y)[(size_t)m * N + row] = Elem<T>::from_f(out[r][m]);
                    }
                }
            }
        }
    }
}

// smem: [ xs : MB * xs_elems * sizeof(T) | rings : 2 * stages * stage_bytes | barriers : 2 * (2 * kGemvMaxStages + 4) * 8 |
//         partial sums : 2 * 2 * GW * R*MB * 32 floats ]
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
    c.gred = reinterpret_cast<float *>(smem + off) + (size_t)grp * 2 * GW * (R * MB) * 32;
    c.ring = ring;
    c.stages = stages, c.stage_bytes = a.stage_bytes, c.row_stride = a.stage_bytes / R;
    c.gid = grp * gridDim.x + blockIdx.x, c.total_groups = gridDim.x * NG;
    c.wg = warp % GW, c.lane = lane;

    auto units_of = [&](const ChainPhase &P) -> int {
        const int units = P.swiglu ? P.inter : (P.N + 1) / 2;
        return c.gid < units ? (units - c.gid + c.total_groups - 1) / c.total_groups : 0;
    };

    if (is_producer) {
        // ================================================= TMA producer: streams every phase's weights back to back
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
        }
        __syncthreads();  // (A) publish the mbarrier initialisation to the consumers
        // the producer only reads weights, which no kernel writes: it never needs griddepcontrol.wait
        pdl_launch_dependents();
        if (lane == 0) {
            const uint32_t ring_u32 = smem_u32(ring);
            int p_s = 0, issued = 0;  // ring stage to fill next; items issued so far
            int e_s = 0, e_ph = 0;    // stage / parity of the empty barrier to wait on next
            for (int p = 0; p < a.n_phases; ++p) {
                const ChainPhase &P = a.ph[p];
                const size_t row_bytes = (size_t)P.K * sizeof(T);
                const int nvec_row = (int)(row_bytes / 16), piece_vecs = P.piece_bytes / 16;
                const int my_units = units_of(P);
                for (int un = 0; un < my_units; ++un) {
                    const int u = c.gid + un * c.total_groups;
                    for (int pc = 0; pc < P.pieces; ++pc) {
                        if (issued >= stages) {  // refill a stage once all 8 warps have left it
                            mbar_wait(c.empty0 + e_s * 8, e_ph);
                            if (++e_s == stages) e_s = 0, e_ph ^= 1;
                            fence_proxy_async();
                        }
                        const int v0 = pc * piece_vecs;
                        const uint32_t bytes = (uint32_t)min(piece_vecs, nvec_row - v0) * 16u;
                        const uint32_t bar = c.full0 + p_s * 8;
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
                        ++issued;
                        if (++p_s == stages) p_s = 0;
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
    int s = 0, ph = 0, un0 = 0;
    const unsigned int arrivals = gridDim.x * NG;
    unsigned long long *tr = a.trace ? a.trace + (size_t)blockIdx.x * kChainMaxPhases * 8 : nullptr;
    for (int p = 0; p < a.n_phases; ++p) {
        const ChainPhase &P = a.ph[p];
        if (tr && stid == 0) tr[p * 8 + 0] = global_ns();  // this warp is done with phase p-1 (or past griddepcontrol.wait)
        if (p > 0) {
            // grid barrier: phase p-1's outputs (y, and the residual stream CTA 0 rewrote) are complete everywhere
            if (stid == 0) {
                const unsigned int *cnt = a.sync + (p - 1);
                const long long t0 = clock64();
                for (;;) {
                    unsigned int v;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                    if (v >= arrivals) break;
                    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a CTA of this grid never became resident
                }
            }
            __syncwarp();
        }
        if (tr && stid == 0) tr[p * 8 + 1] = global_ns();  // grid barrier observed
        named_bar_sync(1, kChainStagers);  // every compute warp has left phase p-1 (xs may be overwritten); barrier observed
        if (tr && stid == 0) tr[p * 8 + 2] = global_ns();  // all stagers of this CTA are here
        chain_stage<T, MB>(P, a.M, a.eps, xs, a.xs_elems, red, stid);
        if (MB > 1) {  // padding rows of the batch tile (read by the FMAs, results discarded)
            constexpr int V = Elem<T>::kVec;
            for (int m = a.M; m < MB; ++m)
                for (int i = stid; i < P.K / V; i += kChainStagers)
                    *reinterpret_cast<uint4 *>(xs + (size_t)m * a.xs_elems + (size_t)i * V) = make_uint4(0, 0, 0, 0);
        }
        named_bar_sync(1, kChainStagers);  // xs complete
        if (tr && stid == 0) tr[p * 8 + 3] = global_ns();  // staging done
        if (p == 0) pdl_launch_dependents();

        const int my_units = units_of(P);
        if (is_compute) {
            switch (P.xv) {
                case 2:
                    if constexpr (MB <= 2) {
                        chain_compute<T, MB, 2>(P, c, xs, a.xs_elems, my_units, s, ph, un0);
                        break;
                    }
                case 4:
                    if constexpr (MB == 1) {
                        chain_compute<T, MB, 4>(P, c, xs, a.xs_elems, my_units, s, ph, un0);
                        break;
                    }
                case 6:
                    if constexpr (MB == 1) {
                        chain_compute<T, MB, 6>(P, c, xs, a.xs_elems, my_units, s, ph, un0);
                        break;
                    }
                default: chain_compute<T, MB, 0>(P, c, xs, a.xs_elems, my_units, s, ph, un0);
            }
        } else {
            chain_reduce<T, MB>(P, c, a.M, my_units, un0);
            if (tr && lane == 0 && grp == 0) tr[p * 8 + 5] = global_ns();  // reducer of group 0 stored its last output
            if (lane == 0 && p + 1 < a.n_phases) {
                __threadfence();
                atomicAdd(a.sync + p, 1u);
            }
            if (tr && lane == 0 && grp == 0) tr[p * 8 + 6] = global_ns();  // ... and arrived at the grid barrier
        }
        if (tr && stid == 0) tr[p * 8 + 4] = global_ns();  // compute warp 0 finished its units of phase p
        un0 += my_units;
    }
}

// Host side (gemv_chain_inst.cuh / linear.cu).  Fills the geometry fields of `a` and launches; B200_ERR_UNSUPPORTED (no error text)
// when some phase cannot use this kernel -- the caller then runs the phases as separate GEMV launches.
// dry: only check the shapes / fill the geometry, launch nothing.
int launch_gemv_chain(ChainArgs &a, int dtype, cudaStream_t st, bool dry = false);

}  // namespace b200
