// sampling.cu -- top-k over the vocabulary and top-k sampling for sm_100a.
// Reference: src/kernels/topk.cu:24-140 (+ includes/topk.cuh:9-42), src/kernels/sampling.cu:14-102.
//
// Top-k is two launches like the reference (B200_TOPK_BLOCKS partial lists per row, then a merge) but with
// -inf sentinels and a total order (value descending, ties -> lower id), so ids are bit-exact against a CPU
// partial sort.  A row of 32000 logits is 128 KB: the stage is latency-, not bandwidth-bound.
#include "common.cuh"

#include <curand_kernel.h>

namespace b200 {

constexpr int kTopkThreads = 256;
constexpr int kMaxK = B200_TOPK_MAX_K;

struct Cand {
    float v;
    int id;
};
__device__ __forceinline__ bool better(float v, int id, float v2, int id2) { return v > v2 || (v == v2 && id < id2); }

__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xffffffffu, c.v, o);
        const int id = __shfl_xor_sync(0xffffffffu, c.id, o);
        if (better(v, id, c.v, c.id)) c.v = v, c.id = id;
    }
    return c;
}

// sorted insertion into a thread-local list of k entries (descending)
__device__ __forceinline__ void insert(float (&lv)[kMaxK], int (&li)[kMaxK], int k, float v, int id) {
    if (!better(v, id, lv[k - 1], li[k - 1])) return;
    lv[k - 1] = v, li[k - 1] = id;
#pragma unroll
    for (int i = kMaxK - 1; i > 0; --i) {
        if (i < k && better(lv[i], li[i], lv[i - 1], li[i - 1])) {
            const float tv = lv[i];
            const int ti = li[i];
            lv[i] = lv[i - 1], li[i] = li[i - 1];
            lv[i - 1] = tv, li[i - 1] = ti;
        }
    }
}

// k rounds of block arg-best over the heads of the thread-local lists; the winner pops its head.
// Thread 0 ends up with the block's top-k in out_v / out_i.
__device__ __forceinline__ void block_select(float (&lv)[kMaxK], int (&li)[kMaxK], int k, float *out_v, int *out_i, Cand *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int r = 0; r < k; ++r) {
        Cand c = warp_best(Cand{lv[0], li[0]});
        __syncthreads();
        if (lane == 0) red[warp] = c;
        __syncthreads();
        Cand b = lane < nwarp ? red[lane] : Cand{-INFINITY, INT_MAX};
        b = warp_best(b);  // every warp computes the same winner
        if (threadIdx.x == 0) out_v[r] = b.v, out_i[r] = b.id == INT_MAX ? -1 : b.id;
        if (li[0] == b.id && b.id != INT_MAX) {
#pragma unroll
            for (int i = 0; i < kMaxK - 1; ++i) lv[i] = lv[i + 1], li[i] = li[i + 1];
            lv[kMaxK - 1] = -INFINITY, li[kMaxK - 1] = INT_MAX;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kTopkThreads)
topk_stage1_kernel(const T *__restrict__ logits, int *__restrict__ tmp_ids, T *__restrict__ tmp_vals, int vocab, int k) {
    __shared__ Cand red[32];
    __shared__ float sv[kMaxK];
    __shared__ int si[kMaxK];
    const int row = blockIdx.y, blk = blockIdx.x;
    const int per = (vocab + B200_TOPK_BLOCKS - 1) / B200_TOPK_BLOCKS;
    const int lo = blk * per, hi = min(vocab, lo + per);
    float lv[kMaxK];
    int li[kMaxK];
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) lv[i] = -INFINITY, li[i] = INT_MAX;
    pdl_wait();
    const T *p = logits + (size_t)row * vocab;
    for (int i = lo + threadIdx.x; i < hi; i += kTopkThreads) insert(lv, li, k, Elem<T>::to_f(p[i]), i);
    block_select(lv, li, k, sv, si, red);
    __syncthreads();
    if (threadIdx.x < k) {
        const size_t o = ((size_t)row * B200_TOPK_BLOCKS + blk) * k + threadIdx.x;
        tmp_ids[o] = si[threadIdx.x];
        tmp_vals[o] = Elem<T>::from_f(sv[threadIdx.x]);
    }
}

template <typename T>
__global__ void __launch_bounds__(64)
topk_stage2_kernel(const int *__restrict__ tmp_ids, const T *__restrict__ tmp_vals, int *__restrict__ final_ids,
                   T *__restrict__ final_vals, int k) {
    __shared__ Cand red[32];
    __shared__ float sv[kMaxK];
    __shared__ int si[kMaxK];
    const int row = blockIdx.x;
    float lv[kMaxK];
    int li[kMaxK];
#pragma unroll
    for (int i = 0; i < kMaxK; ++i) lv[i] = -INFINITY, li[i] = INT_MAX;
    pdl_wait();
    for (int i = threadIdx.x; i < B200_TOPK_BLOCKS * k; i += 64) {
        const int id = tmp_ids[(size_t)row * B200_TOPK_BLOCKS * k + i];
        if (id >= 0) insert(lv, li, k, Elem<T>::to_f(tmp_vals[(size_t)row * B200_TOPK_BLOCKS * k + i]), id);
    }
    block_select(lv, li, k, sv, si, red);
    __syncthreads();
    if (threadIdx.x < k) {
        final_ids[(size_t)row * k + threadIdx.x] = si[threadIdx.x];
        final_vals[(size_t)row * k + threadIdx.x] = Elem<T>::from_f(sv[threadIdx.x]);
    }
}

// one thread per batch row; same arithmetic order as the reference kernel
template <typename T>
__global__ void sampling_kernel(const int *__restrict__ topk_id, T *topk_val, int *seq_len, uint8_t *finished, int *output_id,
                                int batch, int k, int step, int end_id, int vocab) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    if (b >= batch) return;
    T *val = topk_val + (size_t)b * k;
    const int *id = topk_id + (size_t)b * k;
    const float mx = Elem<T>::to_f(val[0]);
    for (int i = 0; i < k; ++i) val[i] = Elem<T>::from_f(expf(Elem<T>::to_f(val[i]) - mx));
    float sum = 0.0f;
    for (int i = 0; i < k; ++i) sum += Elem<T>::to_f(val[i]);
    curandState_t state;
    curand_init((unsigned long long)step, (unsigned long long)b, 0ull, &state);
    float thr = curand_uniform(&state) * sum;
    int chosen = id[0];
    for (int i = 0; i < k; ++i) {
        thr -= Elem<T>::to_f(val[i]);
        if (thr < 0.0f) {
            chosen = id[i];
            break;
        }
    }
    // an empty top-k slot (id -1 / INT_MAX: every candidate NaN or -inf, or vocab < k) must never reach the embedding gather of the next
    // step: the row ends instead (end_id when it is a valid token, else token 0)
    if (chosen < 0 || chosen == INT_MAX) chosen = (end_id >= 0 && end_id < vocab) ? end_id : 0;
    else chosen %= vocab;
    output_id[b] = chosen;
    if (!finished[b]) ++seq_len[b];
    finished[b] = (uint8_t)(chosen == end_id);
}

__global__ void xorwow_uniform_kernel(float *out, int n, unsigned long long seed) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    curandState_t state;
    curand_init(seed, (unsigned long long)b, 0ull, &state);
    out[b] = curand_uniform(&state);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_topk(const void *logits, int *tmp_ids, void *tmp_vals, int *final_ids, void *final_vals, int rows, int vocab, int k,
              int dtype, b200_stream_t stream) {
    B200_REQUIRE(logits && tmp_ids && tmp_vals && final_ids && final_vals, "topk: null pointer");
    B200_REQUIRE(rows >= 0 && vocab > 0, "topk: bad shape");
    B200_REQUIRE(k >= 1 && k <= B200_TOPK_MAX_K, "topk: k=%d outside [1, %d]", k, B200_TOPK_MAX_K);
    B200_REQUIRE(rows <= 65535, "topk: rows %d > 65535", rows);
    if (rows == 0) return B200_OK;
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DTYPE(dtype, {
        launch_pdl(topk_stage1_kernel<T>, dim3(B200_TOPK_BLOCKS, rows), dim3(kTopkThreads), 0, st, true, (const T *)logits, tmp_ids,
                   (T *)tmp_vals, vocab, k);
        launch_pdl(topk_stage2_kernel<T>, dim3(rows), dim3(64), 0, st, true, (const int *)tmp_ids, (const T *)tmp_vals, final_ids,
                   (T *)final_vals, k);
    });
    return cuda_status("topk launch");
}

int b200_sampling(const int *topk_id, void *topk_val, int *seq_len, uint8_t *finished, int *output_id, int batch, int k, int step,
                  int end_id, int vocab, int dtype, b200_stream_t stream) {
    B200_REQUIRE(topk_id && topk_val && seq_len && finished && output_id, "sampling: null pointer");
    B200_REQUIRE(batch >= 0 && k >= 1 && vocab > 0, "sampling: bad shape");
    if (batch == 0) return B200_OK;
    B200_DISPATCH_DTYPE(dtype, launch_pdl(sampling_kernel<T>, dim3((batch + 31) / 32), dim3(32), 0, as_stream(stream), true, topk_id,
                                          (T *)topk_val, seq_len, finished, output_id, batch, k, step, end_id, vocab));
    return cuda_status("sampling launch");
}

int b200_xorwow_uniform(float *out, int n, unsigned long long seed, b200_stream_t stream) {
    B200_REQUIRE(out && n >= 0, "xorwow_uniform: bad argument");
    if (n == 0) return B200_OK;
    xorwow_uniform_kernel<<<(n + 31) / 32, 32, 0, as_stream(stream)>>>(out, n, seed);
    return cuda_status("xorwow_uniform launch");
}

}  // extern "C"
