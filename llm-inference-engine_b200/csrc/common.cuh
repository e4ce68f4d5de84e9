// common.cuh -- device helpers shared by every kernel of the decoder-layer hot path (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/b200llm.h"

namespace b200 {

// ------------------------------------------------------------------ error plumbing (host)
void set_error(const char *fmt, ...);
int cuda_status(const char *what);  // cudaGetLastError() -> B200_OK / B200_ERR_CUDA (+ message)
int sm_count();

#define B200_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::b200::set_error(__VA_ARGS__);     \
            return B200_ERR_INVALID_ARG;        \
        }                                       \
    } while (0)

// NVTX ranges around the host-side phases (prefill pass, decode step, layers, generation loop, batcher iteration): visible in any
// NVTX-aware tool (the reference has no tracing beyond printf); the header-only NVTX v3 costs one predicted branch when no
// tool is attached.  -DB200_NO_NVTX compiles them out.
#ifndef B200_NO_NVTX
struct NvtxRange {
    explicit NvtxRange(const char *name);
    ~NvtxRange();
};
#else
struct NvtxRange {
    explicit NvtxRange(const char *) {}
};
#endif

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(b200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Library workspace (device memory, caller- or library-owned): [tickets | scratch].
struct Workspace {
    unsigned int *tickets;  // zero-initialised, self-resetting counters
    size_t n_tickets;
    char *scratch;
    size_t scratch_bytes;
};
// Returns false (and sets the error) if no workspace is available.
bool get_workspace(Workspace *ws);

// ------------------------------------------------------------------ tensor-parallel exchange over NVLink peer memory
// One-shot all-reduce fused into the producer AND the consumer of a row-sharded linear, with no handshake at all.  Every rank owns an
// exchange buffer that all other ranks of the node map through CUDA IPC.  PUSH: the producing linear's epilogue stores this rank's
// partial sum [B,h] straight into EVERY rank's buffer (slot [source rank]; posted NVLink writes, no round trip) as 8-byte words
// {payload: 32 bits of the tensor, flag: 32 bits} -- the "LL" layout: an aligned 8-byte store is never torn, so a word whose flag holds
// the expected value carries valid data.  The first kernel of the next block polls the P partials in its OWN memory until every word
// shows the flag of this block, and adds them in rank order (identical on every rank, deterministic).  No NCCL call, no extra launch,
// no remote read, no fence, no signal: the data is its own arrival notice, so the exchange costs one NVLink one-way latency behind
// the producer's last store.  (Round 1 signalled with per-rank flag words AFTER the consumer's griddepcontrol.wait: producer drained
// -> st.release.sys to 7 peers -> every CTA polled ld.acquire.sys -> only then the data was read: 8-18 us per exchange, and TP-8 ran
// slower than TP-4.)
// Flag values grow monotonically and are never zero (the buffers start zeroed): epoch (bumped once per decode step by a one-thread
// kernel) * 4096 + block sequence number.  Two slots alternate; a slot is rewritten by block seq + 2, which a rank can only produce after
// it has consumed block seq + 1 from everybody, i.e. after every rank has finished reading block seq.
constexpr int kTpMaxWorld = 8;
struct TpExchange {
    const void *peer_x[kTpMaxWorld];  // the P partials of this slot (LL words), one per source rank, all in THIS rank's exchange buffer
    const unsigned int *epoch;        // this rank's step counter (device memory)
    unsigned int *error;              // set to 1 if a peer's data never arrived (time-out instead of a hang); sticky
    int world, rank, seq;             // world <= 1: exchange disabled
};
// producer side of the same protocol: where this rank's partial of block `seq` goes in every rank's buffer
struct TpPush {
    void *dst[kTpMaxWorld];
    const unsigned int *epoch;
    int n, seq;                       // n == 0: plain store into y
};

// out = gamma * (o + bias) * rsqrt(mean((o+bias)^2)+eps), o = in (+ rin); rout <- o.  in == NULL: in place on out.
// tp (optional, world > 1): `in` is replaced by the sum over ranks of tp->peer_x[*] (see TpExchange).
int launch_norm_tp(int dtype, const void *in, void *out, const void *rin, void *rout, const void *bias, const void *gamma, float eps,
                   int tokens, int hidden, const TpExchange *tp, cudaStream_t st);
// gamma == NULL: out = o + bias.  (norm.cu)
int launch_norm_any(int dtype, const void *in, void *out, const void *rin, void *rout, const void *bias, const void *gamma,
                    float eps, int tokens, int hidden, cudaStream_t st);

// Launch with the programmatic-dependent-launch attribute so that the prologue of kernel i+1
// (weight prefetch, smem carve-up) overlaps the tail of kernel i.  Kernels call pdl_wait() before
// touching anything the previous kernel wrote and pdl_launch_dependents() as early as they can.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                      unsigned cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_x > 1) {  // thread-block cluster along x (distributed shared memory between the CTAs of a cluster)
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    return launch_pdl_cluster(kernel, grid, block, smem, stream, pdl, 1u, args...);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ PDL (device)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ------------------------------------------------------------------ element traits
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int kDtype = B200_F32;
    static constexpr int kVec = 4;  // elements per 16-byte vector
    __device__ __forceinline__ static float to_f(float v) { return v; }
    __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct Elem<__half> {
    static constexpr int kDtype = B200_F16;
    static constexpr int kVec = 8;
    __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
    __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int kDtype = B200_BF16;
    static constexpr int kVec = 8;
    __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// The value a tensor of T holds for the fp32 value v.  The reference's kernels are templates over T and form their
// intermediate sums in T (out += residual; out += bias; q += bias ...): the fused kernels round at the same points so that
// fusing does not change which values are representable downstream.
template <typename T> __device__ __forceinline__ float round_to(float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); }

// 16-byte vector of T unpacked to fp32 lanes.
template <typename T> struct Vec16 {
    static constexpr int N = Elem<T>::kVec;
};

// streaming (read-once) 128-bit load: bypass L1 allocation
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// plain coherent 128-bit load (data another kernel / CTA may have just written)
__device__ __forceinline__ uint4 ld_v4(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st_v4(void *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }

__device__ __forceinline__ void bf16x2_to_f32(uint32_t u, float &lo, float &hi) {
    lo = __uint_as_float(u << 16);
    hi = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ void f16x2_to_f32(uint32_t u, float &lo, float &hi) {
    __half2 h = *reinterpret_cast<__half2 *>(&u);
    float2 f = __half22float2(h);
    lo = f.x;
    hi = f.y;
}

// unpack one 16-byte vector of T into Elem<T>::kVec floats
template <typename T> __device__ __forceinline__ void unpack16(const uint4 &v, float *f);
template <> __device__ __forceinline__ void unpack16<float>(const uint4 &v, float *f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4 &v, float *f) {
    bf16x2_to_f32(v.x, f[0], f[1]);
    bf16x2_to_f32(v.y, f[2], f[3]);
    bf16x2_to_f32(v.z, f[4], f[5]);
    bf16x2_to_f32(v.w, f[6], f[7]);
}
template <> __device__ __forceinline__ void unpack16<__half>(const uint4 &v, float *f) {
    f16x2_to_f32(v.x, f[0], f[1]);
    f16x2_to_f32(v.y, f[2], f[3]);
    f16x2_to_f32(v.z, f[4], f[5]);
    f16x2_to_f32(v.w, f[6], f[7]);
}

// pack Elem<T>::kVec floats into one 16-byte vector of T (round-to-nearest-even)
template <typename T> __device__ __forceinline__ uint4 pack16(const float *f);
template <> __device__ __forceinline__ uint4 pack16<float>(const float *f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float *f) {
    uint4 r;
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
    r.x = *reinterpret_cast<uint32_t *>(&a);
    r.y = *reinterpret_cast<uint32_t *>(&b);
    r.z = *reinterpret_cast<uint32_t *>(&c);
    r.w = *reinterpret_cast<uint32_t *>(&d);
    return r;
}
template <> __device__ __forceinline__ uint4 pack16<__half>(const float *f) {
    uint4 r;
    __half2 a = __floats2half2_rn(f[0], f[1]);
    __half2 b = __floats2half2_rn(f[2], f[3]);
    __half2 c = __floats2half2_rn(f[4], f[5]);
    __half2 d = __floats2half2_rn(f[6], f[7]);
    r.x = *reinterpret_cast<uint32_t *>(&a);
    r.y = *reinterpret_cast<uint32_t *>(&b);
    r.z = *reinterpret_cast<uint32_t *>(&c);
    r.w = *reinterpret_cast<uint32_t *>(&d);
    return r;
}

// round_to<T> over the Elem<T>::kVec lanes of one 16-byte vector, two lanes per conversion: a scalar cvt.rn.bf16.f32 is an F2F on the XU
// pipe (one warp instruction per 8 cycles per scheduler -- the pipe of rsqrt / ex2), the packed cvt.rn.bf16x2.f32 an F2FP on the ALU pipe.
// Same rounding (nearest even), bit-identical results.
template <typename T> __device__ __forceinline__ void round_vec(float *f) {
    if constexpr (sizeof(T) == 2) unpack16<T>(pack16<T>(f), f);
}

// ------------------------------------------------------------------ tensor-parallel exchange (device)
__device__ __forceinline__ unsigned int tp_flag(const unsigned int *epoch, int seq) {
    unsigned int e;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(e) : "l"(epoch));
    return e * 4096u + (unsigned int)seq;
}
__device__ __forceinline__ uint4 tp_ld_v4(const void *p) {  // uncached: the words change under this kernel's feet
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void tp_st_v2(void *p, unsigned int payload, unsigned int flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(payload), "r"(flag) : "memory");
}
__device__ __forceinline__ void tp_st_v4(void *p, unsigned int p0, unsigned int p1, unsigned int flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%2};" ::"l"(p), "r"(p0), "r"(flag), "r"(p1) : "memory");
}
// LL words are indexed by 32-bit payload word: element e of a tensor of T lives in word e * sizeof(T) / 4.
// PUSH two consecutive elements (e0 even) of this rank's partial to every rank.
template <typename T>
__device__ __forceinline__ void tp_push_pair(const TpPush &p, unsigned int flag, size_t e0, float v0, float v1) {
    if constexpr (sizeof(T) == 2) {
        const T a = Elem<T>::from_f(v0), b = Elem<T>::from_f(v1);
        const unsigned int w = (unsigned int)*reinterpret_cast<const unsigned short *>(&a) | ((unsigned int)*reinterpret_cast<const unsigned short *>(&b) << 16);
        for (int r = 0; r < p.n; ++r) tp_st_v2(reinterpret_cast<char *>(p.dst[r]) + (e0 / 2) * 8, w, flag);
    } else {
        for (int r = 0; r < p.n; ++r) tp_st_v4(reinterpret_cast<char *>(p.dst[r]) + e0 * 8, __float_as_uint(v0), __float_as_uint(v1), flag);
    }
}
// PUSH four consecutive elements (e0 a multiple of 4) of a 16-bit tensor: two LL words in ONE 16-byte store per peer (NVLink moves a
// 16-byte store for the price of an 8-byte one: the packet overhead dominates at these sizes)
template <typename T>
__device__ __forceinline__ void tp_push_quad(const TpPush &p, unsigned int flag, size_t e0, const float (&v)[4]) {
    static_assert(sizeof(T) == 2, "tp_push_quad: 16-bit element types");
    const T a = Elem<T>::from_f(v[0]), b = Elem<T>::from_f(v[1]), c = Elem<T>::from_f(v[2]), d = Elem<T>::from_f(v[3]);
    const unsigned int w0 = (unsigned int)*reinterpret_cast<const unsigned short *>(&a) | ((unsigned int)*reinterpret_cast<const unsigned short *>(&b) << 16);
    const unsigned int w1 = (unsigned int)*reinterpret_cast<const unsigned short *>(&c) | ((unsigned int)*reinterpret_cast<const unsigned short *>(&d) << 16);
    for (int r = 0; r < p.n; ++r) tp_st_v4(reinterpret_cast<char *>(p.dst[r]) + (e0 / 2) * 8, w0, w1, flag);
}
// PUSH one 16-byte vector of T (4 payload words, first element e0 a multiple of the vector length)
__device__ __forceinline__ void tp_push_vec(const TpPush &p, unsigned int flag, size_t word0, const uint4 &v) {
    for (int r = 0; r < p.n; ++r) {
        char *d = reinterpret_cast<char *>(p.dst[r]) + word0 * 8;
        tp_st_v4(d, v.x, v.y, flag);
        tp_st_v4(d + 16, v.z, v.w, flag);
    }
}
// CONSUME one 16-byte vector of T (payload words [word0, word0 + 4)) of every rank's partial: f[j] = round_T(sum over ranks in rank
// order).  Polls until all flags show `want`; a peer that never delivers trips the (sticky) error word after ~2 s and the result is NaN.
// RB = ranks whose loads are in flight together: 4 inside the fused GEMV prologues (register budget), 8 where registers are free.
template <typename T, int RB = 4>
__device__ __forceinline__ void tp_reduce_vec(const TpExchange &t, unsigned int want, size_t word0, float *f) {
    constexpr int V = Elem<T>::kVec;
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] = 0.0f;
    bool poisoned = false;
#pragma unroll
    for (int r0 = 0; r0 < kTpMaxWorld; r0 += RB) {  // RB ranks' loads in flight at a time
        if (r0 >= t.world) break;
        uint4 a[RB], b[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j)
            if (r0 + j < t.world) {
                const char *src = reinterpret_cast<const char *>(t.peer_x[r0 + j]) + word0 * 8;
                a[j] = tp_ld_v4(src), b[j] = tp_ld_v4(src + 16);
            }
#pragma unroll
        for (int j = 0; j < RB; ++j)
            if (r0 + j < t.world) {
                if (a[j].y != want || a[j].w != want || b[j].y != want || b[j].w != want) {
                    const char *src = reinterpret_cast<const char *>(t.peer_x[r0 + j]) + word0 * 8;
                    const long long t0 = clock64();
                    for (;;) {
                        a[j] = tp_ld_v4(src), b[j] = tp_ld_v4(src + 16);
                        if (a[j].y == want && a[j].w == want && b[j].y == want && b[j].w == want) break;
                        unsigned int err;
                        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(err) : "l"(t.error));
                        if (err || clock64() - t0 > 4000000000LL) {  // ~2 s: a peer died; leave a mark instead of hanging the GPU
                            atomicExch(t.error, 1u);
                            poisoned = true;
                            break;
                        }
                        __nanosleep(32);
                    }
                }
                float g[V];
                unpack16<T>(make_uint4(a[j].x, a[j].z, b[j].x, b[j].z), g);
#pragma unroll
                for (int k = 0; k < V; ++k) f[k] += g[k];
            }
    }
    round_vec<T>(f);  // rounded to T as an all-reduced tensor of T would be
#pragma unroll
    for (int j = 0; j < V; ++j) f[j] = poisoned ? __int_as_float(0x7fc00000) : f[j];
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum broadcast to every thread.  `red` = shared float[33].  All threads must call.
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nwarp ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
__device__ __forceinline__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nwarp ? red[lane] : -INFINITY;
        t = warp_max(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}
#endif  // __CUDACC__

// dtype dispatch on the host: calls f(T{}) with the element type
#define B200_DISPATCH_DTYPE(dtype, ...)                                  \
    switch (dtype) {                                                     \
        case B200_F32: {                                                 \
            using T = float;                                             \
            __VA_ARGS__;                                                 \
        } break;                                                         \
        case B200_F16: {                                                 \
            using T = __half;                                            \
            __VA_ARGS__;                                                 \
        } break;                                                         \
        case B200_BF16: {                                                \
            using T = __nv_bfloat16;                                     \
            __VA_ARGS__;                                                 \
        } break;                                                         \
        default:                                                         \
            ::b200::set_error("unknown dtype %d", (int)(dtype));         \
            return B200_ERR_INVALID_ARG;                                 \
    }

}  // namespace b200
