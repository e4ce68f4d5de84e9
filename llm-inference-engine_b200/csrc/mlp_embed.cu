// mlp_embed.cu -- SwiGLU activation and embedding gather (stand-alone launchers; the decode engine fuses SwiGLU into
// the gate/up GEMV epilogue).  Reference: src/kernels/silu_and_mul.cu:6-82, src/kernels/input_embedding.cu:4-51.
#include "common.cuh"

namespace b200 {

template <typename T, bool kVec>
__global__ void __launch_bounds__(256)
silu_and_mul_kernel(const T *__restrict__ in, T *__restrict__ out, int tokens, int inter) {
    constexpr int V = kVec ? Elem<T>::kVec : 1;
    const int per_row = inter / V;
    const size_t total = (size_t)tokens * per_row;
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i / per_row, c = i % per_row;
        const T *g = in + t * 2 * inter + c * V;
        const T *u = g + inter;
        float gf[V], uf[V];
        if constexpr (kVec) {
            unpack16<T>(ld_stream_v4(g), gf);
            unpack16<T>(ld_stream_v4(u), uf);
        } else {
            gf[0] = Elem<T>::to_f(*g), uf[0] = Elem<T>::to_f(*u);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) gf[j] = (gf[j] / (1.0f + expf(-gf[j]))) * uf[j];
        if constexpr (kVec) st_v4(out + t * inter + c * V, pack16<T>(gf));
        else out[t * inter + c] = Elem<T>::from_f(gf[0]);
    }
}

// one CTA per token row: a bit-exact copy of table[ids[t], :]
template <typename T>
__global__ void __launch_bounds__(256)
embedding_kernel(const int *__restrict__ ids, const T *__restrict__ table, T *__restrict__ out, int hidden, bool vec) {
    const int t = blockIdx.x;
    pdl_wait();
    const T *src = table + (size_t)ids[t] * hidden;
    T *dst = out + (size_t)t * hidden;
    if (vec) {
        constexpr int V = Elem<T>::kVec;
        for (int i = threadIdx.x; i < hidden / V; i += blockDim.x) st_v4(dst + (size_t)i * V, ld_stream_v4(src + (size_t)i * V));
    } else {
        for (int i = threadIdx.x; i < hidden; i += blockDim.x) dst[i] = src[i];
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_silu_and_mul(const void *in, void *out, int tokens, int inter_size, int dtype, b200_stream_t stream) {
    B200_REQUIRE(in && out, "silu_and_mul: null pointer");
    B200_REQUIRE(tokens >= 0 && inter_size > 0, "silu_and_mul: bad shape");
    if (tokens == 0) return B200_OK;
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DTYPE(dtype, {
        const bool vec = inter_size % Elem<T>::kVec == 0 && aligned16(in) && aligned16(out);
        const size_t items = (size_t)tokens * (vec ? inter_size / Elem<T>::kVec : inter_size);
        size_t g = (items + 255) / 256, cap = (size_t)sm_count() * 8;
        const int grid = (int)(g < cap ? g : cap);
        if (vec) launch_pdl(silu_and_mul_kernel<T, true>, dim3(grid), dim3(256), 0, st, true, (const T *)in, (T *)out, tokens, inter_size);
        else launch_pdl(silu_and_mul_kernel<T, false>, dim3(grid), dim3(256), 0, st, true, (const T *)in, (T *)out, tokens, inter_size);
    });
    return cuda_status("silu_and_mul launch");
}

int b200_input_embedding(const int *ids, const void *table, void *out, int tokens, int hidden, int dtype, b200_stream_t stream) {
    B200_REQUIRE(ids && table && out, "input_embedding: null pointer");
    B200_REQUIRE(tokens >= 0 && hidden > 0, "input_embedding: bad shape");
    if (tokens == 0) return B200_OK;
    B200_DISPATCH_DTYPE(dtype, {
        const bool vec = hidden % Elem<T>::kVec == 0 && aligned16(table) && aligned16(out);
        launch_pdl(embedding_kernel<T>, dim3(tokens), dim3(256), 0, as_stream(stream), true, ids, (const T *)table, (T *)out, hidden, vec);
    });
    return cuda_status("input_embedding launch");
}

}  // extern "C"
