// norm.cu -- RMSNorm, fused add-bias-residual-RMSNorm and add-residual for sm_100a.
//
// One CTA per token row; the row is read ONCE with 128-bit loads, kept in registers across the
// sum-of-squares reduction (warp shuffles + one smem hop) and written back once.  HBM-bound:
// algorithmic bytes per token = hidden * e * (reads + writes) -- see DESIGN.md.
// Semantics: reference src/kernels/rmsnorm.cu:35-80, add_residual_and_rmsnorm.cu:43-121, add_residual.cu:8-49.
#include "common.cuh"

namespace b200 {

constexpr int kNormThreads = 256;
constexpr int kNormMaxVec = 8;  // 16-byte vectors cached per thread

// residual_out <- o (pre-bias), out <- gamma * (o + bias) * rsqrt(mean((o+bias)^2) + eps), o = in (+ residual_in).
// kCopyOnly (launchRMSNorm): residual_out <- x, no add.
// kThreads x kMaxVec: 256 x 8 for many rows (prefill); 512 x 4 for the few rows of a decode batch, where a row is ONE dependent chain of
// loads (under tensor parallelism 2 x world polled loads per vector) and more threads mean fewer of them per thread.
// kTp: the row comes from the tensor-parallel exchange (a separate instance: the polled 2 x world loads per vector cost ~100 registers)
template <typename T, bool kVec, int kThreads, int kMaxVec, bool kTp>
__global__ void __launch_bounds__(kThreads)
norm_kernel(const T *in, T *out, const T *residual_in, T *residual_out, const T *__restrict__ bias,
            const T *__restrict__ gamma, float eps, int hidden, const TpExchange tp) {
    constexpr int V = kVec ? Elem<T>::kVec : 1;
    __shared__ float red[33];
    const int row = blockIdx.x;
    const int nvec = hidden / V;
    T *o = out + (size_t)row * hidden;
    const T *x = (in ? in : out) + (size_t)row * hidden;
    const T *rin = residual_in ? residual_in + (size_t)row * hidden : nullptr;
    T *rout = residual_out ? residual_out + (size_t)row * hidden : nullptr;

    pdl_wait();
    const unsigned int tp_want = kTp ? tp_flag(tp.epoch, tp.seq) : 0u;
    // pre-norm value of vector i: o (+ residual); residual_out <- that; (+ bias)
    auto prenorm = [&](int i, float *f) {
        if constexpr (kVec) {
            if constexpr (kTp) {  // fused one-shot all-reduce of the row-sharded linear's partial sums (rank order), LL words
                tp_reduce_vec<T, kTpMaxWorld>(tp, tp_want, ((size_t)row * hidden + (size_t)i * V) * sizeof(T) / 4, f);
            } else {
                unpack16<T>(ld_v4(x + (size_t)i * V), f);
            }
            if (rin) {
                float r[V];
                unpack16<T>(ld_v4(rin + (size_t)i * V), r);
#pragma unroll
                for (int j = 0; j < V; ++j) f[j] += r[j];
        round_vec<T>(f);
            }
            if (rout) st_v4(rout + (size_t)i * V, pack16<T>(f));
            if (bias) {
                float b[V];
                unpack16<T>(ld_v4(bias + (size_t)i * V), b);
#pragma unroll
                for (int j = 0; j < V; ++j) f[j] += b[j];
        round_vec<T>(f);
            }
        } else {
            f[0] = Elem<T>::to_f(x[i]);
            if (rin) f[0] = round_to<T>(f[0] + Elem<T>::to_f(rin[i]));
            if (rout) rout[i] = Elem<T>::from_f(f[0]);
            if (bias) f[0] = round_to<T>(f[0] + Elem<T>::to_f(bias[i]));
        }
    };
    auto finish = [&](int i, float *f, float r) {
        if (gamma) {
            float g[V];
            if constexpr (kVec) unpack16<T>(ld_v4(gamma + (size_t)i * V), g);
            else g[0] = Elem<T>::to_f(gamma[i]);
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = (f[j] * g[j]) * r;
        }
        if constexpr (kVec) st_v4(o + (size_t)i * V, pack16<T>(f));
        else o[i] = Elem<T>::from_f(f[0]);
    };

    float cache[kMaxVec][V];
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
        const int i = threadIdx.x + c * kThreads;
        if (i < nvec) {
            prenorm(i, cache[c]);
#pragma unroll
            for (int j = 0; j < V; ++j) ss += cache[c][j] * cache[c][j];
        }
    }
    // rows longer than the register cache: park the pre-norm value in `out` and re-read it below
    for (int i = threadIdx.x + kMaxVec * kThreads; i < nvec; i += kThreads) {
        float f[V];
        prenorm(i, f);
#pragma unroll
        for (int j = 0; j < V; ++j) ss += f[j] * f[j];
        if constexpr (kVec) st_v4(o + (size_t)i * V, pack16<T>(f));
        else o[i] = Elem<T>::from_f(f[0]);
    }
    pdl_launch_dependents();
    ss = block_sum(ss, red);
    const float r = rsqrtf(ss / (float)hidden + eps);
#pragma unroll
    for (int c = 0; c < kMaxVec; ++c) {
        const int i = threadIdx.x + c * kThreads;
        if (i < nvec) finish(i, cache[c], r);
    }
    for (int i = threadIdx.x + kMaxVec * kThreads; i < nvec; i += kThreads) {
        float f[V];
        if constexpr (kVec) unpack16<T>(ld_v4(o + (size_t)i * V), f);
        else f[0] = Elem<T>::to_f(o[i]);
        finish(i, f, r);
    }
}

template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
add_residual_kernel(T *__restrict__ out, const T *__restrict__ residual, size_t n_items) {
    constexpr int V = kVec ? Elem<T>::kVec : 1;
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (size_t)gridDim.x * blockDim.x) {
        if constexpr (kVec) {
            float a[V], b[V];
            unpack16<T>(ld_v4(out + i * V), a);
            unpack16<T>(ld_stream_v4(residual + i * V), b);
#pragma unroll
            for (int j = 0; j < V; ++j) a[j] += b[j];
            st_v4(out + i * V, pack16<T>(a));
        } else {
            out[i] = Elem<T>::from_f(Elem<T>::to_f(out[i]) + Elem<T>::to_f(residual[i]));
        }
    }
}

template <typename T>
static int launch_norm(const T *in, T *out, const T *rin, T *rout, const T *bias, const T *gamma, float eps, int tokens,
                       int hidden, cudaStream_t st, const TpExchange *tpx = nullptr) {
    const bool vec = hidden % Elem<T>::kVec == 0 && aligned16(out) && (!in || aligned16(in)) && (!rin || aligned16(rin)) &&
                     (!rout || aligned16(rout)) && (!bias || aligned16(bias)) && (!gamma || aligned16(gamma));
    TpExchange tp = {};
    if (tpx) tp = *tpx;
    if (tp.world > 1 && !vec) {
        set_error("norm: the fused tensor-parallel exchange needs 16-byte aligned rows");
        return B200_ERR_UNSUPPORTED;
    }
    cudaError_t e;
    const int nvec = hidden / (16 / (int)sizeof(T));
    // a decode batch (<= 16 rows): 512 threads per row, one to four vectors each -- (nearly) every load of the row in flight at once
    const bool wide = vec && tokens <= 16 && nvec <= 4 * 512;
    auto go = [&](auto kern, int threads) { return launch_pdl(kern, dim3(tokens), dim3(threads), 0, st, true, in, out, rin, rout, bias, gamma, eps, hidden, tp); };
    if (tp.world > 1) {
        if (wide) e = go(norm_kernel<T, true, 512, 4, true>, 512);
        else e = go(norm_kernel<T, true, kNormThreads, kNormMaxVec, true>, kNormThreads);
    } else if (wide) {
        e = go(norm_kernel<T, true, 512, 4, false>, 512);
    } else if (vec && nvec <= 2 * kNormThreads) {
        // many rows (prefill): the register cache is sized to the row (2 vectors per thread at hidden 4096, 16-bit) so that several CTAs
        // fit an SM -- one HBM pass needs loads in flight, and round 1's single 152-register instance ran ONE CTA per SM (27 us for
        // 2048 x 4096 where the bytes take 10)
        e = go(norm_kernel<T, true, kNormThreads, 2, false>, kNormThreads);
    } else if (vec && nvec <= 4 * kNormThreads) {
        e = go(norm_kernel<T, true, kNormThreads, 4, false>, kNormThreads);
    } else if (vec) {
        e = go(norm_kernel<T, true, kNormThreads, kNormMaxVec, false>, kNormThreads);
    } else {
        e = go(norm_kernel<T, false, kNormThreads, kNormMaxVec, false>, kNormThreads);
    }
    (void)e;
    return cuda_status("norm kernel launch");
}

int launch_norm_tp(int dtype, const void *in, void *out, const void *rin, void *rout, const void *bias, const void *gamma, float eps,
                   int tokens, int hidden, const TpExchange *tp, cudaStream_t st) {
    B200_DISPATCH_DTYPE(dtype, return launch_norm<T>((const T *)in, (T *)out, (const T *)rin, (T *)rout, (const T *)bias,
                                                     (const T *)gamma, eps, tokens, hidden, st, tp));
    return B200_OK;
}

int launch_norm_any(int dtype, const void *in, void *out, const void *rin, void *rout, const void *bias, const void *gamma,
                    float eps, int tokens, int hidden, cudaStream_t st) {
    return launch_norm_tp(dtype, in, out, rin, rout, bias, gamma, eps, tokens, hidden, nullptr, st);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_rmsnorm(void *x, void *residual, const void *gamma, float eps, int tokens, int hidden, int dtype,
                 b200_stream_t stream) {
    B200_REQUIRE(tokens >= 0 && hidden > 0, "rmsnorm: bad shape tokens=%d hidden=%d", tokens, hidden);
    if (tokens == 0) return B200_OK;
    B200_REQUIRE(x && gamma, "rmsnorm: x and gamma must be non-null");
    B200_DISPATCH_DTYPE(dtype, return launch_norm<T>(nullptr, (T *)x, nullptr, (T *)residual, nullptr, (const T *)gamma, eps,
                                                     tokens, hidden, as_stream(stream)));
    return B200_OK;
}

int b200_fused_add_bias_residual_rmsnorm(void *residual, void *out, const void *bias, const void *gamma, float eps,
                                         int tokens, int hidden, int dtype, b200_stream_t stream) {
    B200_REQUIRE(out, "fused_add_bias_residual_rmsnorm: out must be non-null");
    B200_REQUIRE(tokens >= 0 && hidden > 0, "fused_add_bias_residual_rmsnorm: bad shape tokens=%d hidden=%d", tokens, hidden);
    if (tokens == 0) return B200_OK;
    B200_DISPATCH_DTYPE(dtype, return launch_norm<T>(nullptr, (T *)out, (const T *)residual, (T *)residual, (const T *)bias,
                                                     (const T *)gamma, eps, tokens, hidden, as_stream(stream)));
    return B200_OK;
}

int b200_add_residual(const void *residual, void *out, int tokens, int hidden, int dtype, b200_stream_t stream) {
    B200_REQUIRE(residual && out, "add_residual: null pointer");
    B200_REQUIRE(tokens >= 0 && hidden > 0, "add_residual: bad shape");
    const size_t n = (size_t)tokens * hidden;
    if (n == 0) return B200_OK;
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DTYPE(dtype, {
        const bool vec = n % Elem<T>::kVec == 0 && aligned16(out) && aligned16(residual);
        const size_t items = vec ? n / Elem<T>::kVec : n;
        const int grid = (int)((items + 255) / 256 < (size_t)sm_count() * 8 ? (items + 255) / 256 : (size_t)sm_count() * 8);
        if (vec) launch_pdl(add_residual_kernel<T, true>, dim3(grid), dim3(256), 0, st, true, (T *)out, (const T *)residual, items);
        else launch_pdl(add_residual_kernel<T, false>, dim3(grid), dim3(256), 0, st, true, (T *)out, (const T *)residual, items);
    });
    return cuda_status("add_residual kernel launch");
}

}  // extern "C"
