// prefill.cu -- the context ("prefill") side kernels: padding offsets, causal mask, QKV split + RoPE, KV-cache
// append / GQA gather, scale+mask+softmax, un-padding transpose, and the fused flash-style context attention.
// Reference: src/kernels/{cal_padding_offset,build_causal_mask,qkv_bias_and_rope,concat_past_kv,repeat_kv,
// scale_and_mask_and_softmax,transpose_and_remove_padding}.cu and src/layers/context_attention.cpp:221-289.
// The index kernels are bit-exact by construction (pure copies / integer arithmetic).
#include "common.cuh"

#include <stdlib.h>

#include <float.h>

namespace b200 {

int launch_context_attention_tc(const void *q, const void *k_layer, const void *v_layer, void *out, const int *seq_off, const int *input_len,
                                const int *context_len, int batch, int head_num, int kv_head_num, int max_q_len, int max_seq_len, int head_size,
                                float scale, int dtype, cudaStream_t st, const int *block_table, int max_pages, int num_pages);

__device__ __forceinline__ float rope_denominator_p(float base, int zid, int rot_dim) {
    const float e = (float)zid / (float)rot_dim;
    return (float)pow((double)base, (double)e);  // correctly rounded powf, see attention_decode.cu
}

// ---------------------------------------------------------------- padding offsets (cal_padding_offset.cu:17-43)
// one CTA: thread 0 does the (tiny) serial prefix over the batch into smem, then every thread writes tokens.
__global__ void padding_offset_kernel(int *padding_offset, int *cum_seqlens, const int *input_lengths, int batch, int max_q_len) {
    extern __shared__ int cum[];  // [batch + 1]
    pdl_wait();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int b = 0; b < batch; ++b) {
            cum[b] = total;
            total += input_lengths[b];
        }
        cum[batch] = total;
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= batch; b += blockDim.x) cum_seqlens[b] = cum[b];
    for (int i = threadIdx.x; i < batch * max_q_len; i += blockDim.x) {
        const int b = i / max_q_len, t = i % max_q_len;
        if (t < cum[b + 1] - cum[b]) padding_offset[cum[b] + t] = b * max_q_len - cum[b];
    }
}

// ---------------------------------------------------------------- causal mask (build_causal_mask.cu:4-23)
template <typename T>
__global__ void causal_mask_kernel(T *mask, const int *q_lens, const int *k_lens, int max_q_len, int max_k_len, size_t total) {
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % max_k_len);
        const int q = (int)((i / max_k_len) % max_q_len);
        const int b = (int)(i / ((size_t)max_k_len * max_q_len));
        const int ql = q_lens[b], kl = k_lens[b];
        const bool ok = q < ql && k < kl && k <= q + (kl - ql);
        mask[i] = Elem<T>::from_f(ok ? 1.0f : 0.0f);
    }
}

// ---------------------------------------------------------------- QKV split + transpose + re-pad + RoPE (qkv_bias_and_rope.cu:5-79)
// grid (token, head over H + 2*Hkv), block head_size/2 (rounded up to a warp)
template <typename T>
__global__ void qkv_rope_kernel(T *q, T *k, T *v, const T *__restrict__ qkv, const int *__restrict__ padding_offset,
                                const int *__restrict__ history_len, int seq_len, int head_num, int kv_head_num, int head_size,
                                int rot_dim, float base) {
    const int t = blockIdx.x, head = blockIdx.y;
    pdl_wait();
    const int dst = t + padding_offset[t];
    const int b = dst / seq_len, s = dst % seq_len;
    const T *src = qkv + ((size_t)t * (head_num + 2 * kv_head_num) + head) * head_size;
    const int half = head_size / 2;
    if (head >= head_num + kv_head_num) {  // v: pure re-layout
        T *d = v + (((size_t)b * kv_head_num + (head - head_num - kv_head_num)) * seq_len + s) * head_size;
        for (int i = threadIdx.x; i < head_size; i += blockDim.x) d[i] = src[i];
        return;
    }
    T *d = head < head_num ? q + (((size_t)b * head_num + head) * seq_len + s) * head_size
                           : k + (((size_t)b * kv_head_num + (head - head_num)) * seq_len + s) * head_size;
    const float pos = (float)(history_len[b] + s);
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float x0 = Elem<T>::to_f(src[i]), x1 = Elem<T>::to_f(src[i + half]);
        if (i < rot_dim / 2) {
            const float th = pos / rope_denominator_p(base, 2 * i, rot_dim);
            const float c = cosf(th), sn = sinf(th);
            const float a = x0, bb = x1;
            x0 = a * c - bb * sn;
            x1 = bb * c + a * sn;
        }
        d[i] = Elem<T>::from_f(x0);
        d[i + half] = Elem<T>::from_f(x1);
    }
    if ((head_size & 1) && threadIdx.x == 0) d[head_size - 1] = src[head_size - 1];
}

// ---------------------------------------------------------------- engine prefill: QKV split + RoPE + KV append in ONE pass
// launchFusedQKVAddBiasAndTransposeAndRope + launchConcatKVCache fused: one CTA per token, 128-bit vectors; q goes to the padded
// [B,H,max_q,d] buffer the attention reads, k and v go STRAIGHT into the cache at history_len[b] + s (the padded k/v buffers and
// the second pass over them disappear).  cos/sin of the token's position are evaluated once per CTA (same expressions as above).
template <typename T>
__global__ void __launch_bounds__(256)
prefill_qkv_rope_cache_kernel(T *q, T *k_cache, T *v_cache, const T *__restrict__ qkv, const T *__restrict__ bias,
                              const int *__restrict__ padding_offset, const int *__restrict__ history_len, int seq_len, int head_num,
                              int kv_head_num, int head_size, int max_seq_len, int rot_dim, float base, const int *__restrict__ block_table,
                              int max_pages) {
    constexpr int V = Elem<T>::kVec;
    extern __shared__ float2 cs[];  // [head_size / 2] (cos, sin); identity past rot_dim / 2
    const int t = blockIdx.x;
    const int half = head_size / 2, vph = half / V;  // vector pairs per head
    pdl_wait();
    const int dst = t + padding_offset[t];
    const int b = dst / seq_len, s = dst % seq_len;
    const int pos_i = history_len[b] + s;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float2 v = make_float2(1.0f, 0.0f);
        if (i < rot_dim / 2) {
            const float th = (float)pos_i / rope_denominator_p(base, 2 * i, rot_dim);
            v = make_float2(cosf(th), sinf(th));
        }
        cs[i] = v;
    }
    __syncthreads();
    const int heads = head_num + 2 * kv_head_num;
    const T *src_tok = qkv + (size_t)t * heads * head_size;
    for (int w = threadIdx.x; w < heads * vph; w += blockDim.x) {
        const int head = w / vph, i0 = (w % vph) * V;  // dims [i0, i0 + V) and [i0 + half, ...)
        const T *src = src_tok + (size_t)head * head_size;
        float lo[V], hi[V];
        unpack16<T>(ld_stream_v4(src + i0), lo);
        unpack16<T>(ld_stream_v4(src + i0 + half), hi);
        T *d;
        if (head < head_num) {
            d = q + (((size_t)b * head_num + head) * seq_len + s) * head_size;
        } else {
            if (pos_i >= max_seq_len) continue;  // a prompt longer than the cache: never write past the layer's slab / the block table's reach
            const bool is_k = head < head_num + kv_head_num;
            const int kvh = is_k ? head - head_num : head - head_num - kv_head_num;
            // contiguous [B, Hkv, S, d], or a page pool [num_pages, Hkv, 64, d] addressed through the block table (max_seq_len = max_pages * 64)
            const size_t row = block_table ? ((size_t)block_table[(size_t)b * max_pages + pos_i / 64] * kv_head_num + kvh) * 64 + pos_i % 64
                                           : ((size_t)b * kv_head_num + kvh) * max_seq_len + pos_i;
            d = (is_k ? k_cache : v_cache) + row * head_size;
        }
        if (head < head_num + kv_head_num) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float2 c = cs[i0 + e];
                const float a = lo[e], bb = hi[e];
                lo[e] = a * c.x - bb * c.y;
                hi[e] = bb * c.x + a * c.y;
            }
        }
        if (bias) {
            // same convention as the decode step (decoder_self_attention.cu:93-127): RoPE result stored in T, then + bias in T -- the
            // K / V rows a prompt leaves in the cache are the ones the decode kernel would have appended token by token.  (The
            // reference's prefill launcher drops the bias, qkv_bias_and_rope.cu:28-78; b200_qkv_bias_transpose_rope keeps that.)
            float blo[V], bhi[V];
            unpack16<T>(ld_v4(bias + (size_t)head * head_size + i0), blo);
            unpack16<T>(ld_v4(bias + (size_t)head * head_size + i0 + half), bhi);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                lo[e] = round_to<T>(round_to<T>(lo[e]) + blo[e]);
                hi[e] = round_to<T>(round_to<T>(hi[e]) + bhi[e]);
            }
        }
        st_v4(d + i0, pack16<T>(lo));
        st_v4(d + i0 + half, pack16<T>(hi));
    }
}

// q [B,H,max_q,d]; k_cache / v_cache: LAYER base [B,Hkv,S,d].  B200_ERR_UNSUPPORTED when the shape cannot be vectorised.
int launch_prefill_qkv_rope_cache(void *q, void *k_layer, void *v_layer, const void *qkv, const void *bias, const int *padding_offset,
                                  const int *history_len, int seq_len, int num_tokens, int head_num, int kv_head_num, int head_size,
                                  int max_seq_len, int rot_dim, float base, int dtype, cudaStream_t st, const int *block_table, int max_pages) {
    const int vec = dtype == B200_F32 ? 4 : 8;
    if (head_size % (2 * vec) != 0 || !aligned16(q) || !aligned16(k_layer) || !aligned16(v_layer) || !aligned16(qkv) || (bias && !aligned16(bias)))
        return B200_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float2) * (size_t)(head_size / 2);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(prefill_qkv_rope_cache_kernel<T>, dim3(num_tokens), dim3(256), smem, st, true, (T *)q, (T *)k_layer,
                                          (T *)v_layer, (const T *)qkv, (const T *)bias, padding_offset, history_len, seq_len, head_num, kv_head_num,
                                          head_size, block_table ? max_pages * 64 : max_seq_len, rot_dim, base, block_table, max_pages));
    return cuda_status("prefill_qkv_rope_cache launch");
}

// ---------------------------------------------------------------- KV append (concat_past_kv.cu:10-42), K and V in one launch
// grid (max_q_len, Hkv, 2*B)
template <typename T>
__global__ void concat_kv_kernel(const T *__restrict__ k_src, const T *__restrict__ v_src, T *k_cache, T *v_cache,
                                 const int *__restrict__ cur_len, const int *__restrict__ history_len, int kv_head_num,
                                 int max_q_len, int max_seq_len, int head_size) {
    const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z >> 1;
    const bool is_v = blockIdx.z & 1;
    pdl_wait();
    if (t >= cur_len[b] || history_len[b] + t >= max_seq_len) return;  // never write past the layer's cache slab
    const T *src = (is_v ? v_src : k_src) + (((size_t)b * kv_head_num + h) * max_q_len + t) * head_size;
    T *dst = (is_v ? v_cache : k_cache) + (((size_t)b * kv_head_num + h) * max_seq_len + history_len[b] + t) * head_size;
    for (int i = threadIdx.x; i < head_size; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------- GQA gather (repeat_kv.cu:13-49, intended semantics)
// grid (max_k_len, H, 2*B)
template <typename T>
__global__ void repeat_kv_kernel(const T *__restrict__ k_cache, const T *__restrict__ v_cache, T *k_dst, T *v_dst,
                                 const int *__restrict__ context_len, int head_num, int kv_head_num, int max_k_len,
                                 int max_seq_len, int head_size) {
    const int s = blockIdx.x, h = blockIdx.y, b = blockIdx.z >> 1;
    const bool is_v = blockIdx.z & 1;
    pdl_wait();
    if (s >= context_len[b]) return;
    const int kvh = h / (head_num / kv_head_num);
    const T *src = (is_v ? v_cache : k_cache) + (((size_t)b * kv_head_num + kvh) * max_seq_len + s) * head_size;
    T *dst = (is_v ? v_dst : k_dst) + (((size_t)b * head_num + h) * max_k_len + s) * head_size;
    for (int i = threadIdx.x; i < head_size; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------- scale + mask + softmax (scale_and_mask_and_softmax.cu:64-127)
// one CTA per (q row, head, batch) row; three passes over an L1/L2-resident row; in-place capable
template <typename T>
__global__ void __launch_bounds__(256)
scale_mask_softmax_kernel(const T *qk, const T *__restrict__ mask, T *out, float scale, int head_num, int q_len, int k_len) {
    __shared__ float red[33];
    const int q = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const size_t off = (((size_t)b * head_num + h) * q_len + q) * k_len;
    const T *mk = mask + ((size_t)b * q_len + q) * k_len;
    pdl_wait();
    float m = FLT_MIN;
    for (int k = threadIdx.x; k < k_len; k += blockDim.x)
        m = fmaxf(m, scale * Elem<T>::to_f(qk[off + k]) + (1.0f - Elem<T>::to_f(mk[k])) * (-10000.0f));
    m = block_max(m, red);
    float sum = 0.0f;
    for (int k = threadIdx.x; k < k_len; k += blockDim.x)
        sum += expf(scale * Elem<T>::to_f(qk[off + k]) + (1.0f - Elem<T>::to_f(mk[k])) * (-10000.0f) - m);
    sum = block_sum(sum, red);
    const float inv = 1.0f / (sum + 1e-6f);
    for (int k = threadIdx.x; k < k_len; k += blockDim.x) {
        const float e = expf(scale * Elem<T>::to_f(qk[off + k]) + (1.0f - Elem<T>::to_f(mk[k])) * (-10000.0f) - m);
        out[off + k] = Elem<T>::from_f(e * inv);
    }
}

// ---------------------------------------------------------------- transpose + remove padding (transpose_and_remove_padding.cu:15-43)
template <typename T>
__global__ void transpose_remove_padding_kernel(const T *__restrict__ src, const int *__restrict__ padding_offset, T *dst,
                                                int seq_len, int head_num, int head_size) {
    const int t = blockIdx.x;
    pdl_wait();
    const int d = t + padding_offset[t];
    const int b = d / seq_len, s = d % seq_len;
    for (int i = threadIdx.x; i < head_num * head_size; i += blockDim.x) {
        const int h = i / head_size, e = i % head_size;
        dst[(size_t)t * head_num * head_size + i] = src[(((size_t)b * head_num + h) * seq_len + s) * head_size + e];
    }
}

// ---------------------------------------------------------------- fused context attention (SIMT flash tiles)
// CTA = (q tile of kCaRows rows, head, batch), 4 warps x kCaRows/4 rows each; K/V tiles of kCaKeys positions staged in smem
// as fp32; online softmax whose running max starts at FLT_MIN (the reference's thread_max initial value), masked keys
// are skipped: their weight expf(-10000 + ...) is exactly 0 in fp32.  Output goes straight to the un-padded [T,H,d].
constexpr int kCaRows = 16, kCaKeys = 64, kCaThreads = 128;
template <typename T>
__global__ void __launch_bounds__(kCaThreads)
context_attention_kernel(const T *__restrict__ q, const T *__restrict__ k_cache, const T *__restrict__ v_cache, T *out,
                         const int *__restrict__ cum_seqlens_or_null, const int *__restrict__ padding_offset,
                         const int *__restrict__ input_len, const int *__restrict__ context_len, int head_num, int kv_head_num,
                         int max_q_len, int max_seq_len, int head_size, float scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = head_size, DP = D + 1;
    float *qs = reinterpret_cast<float *>(smem_raw);  // [kCaRows][D]
    float *ks = qs + kCaRows * D;                     // [kCaKeys][DP]
    float *vs = ks + kCaKeys * DP;                    // [kCaKeys][D]
    float *ps = vs + kCaKeys * D;                     // [kCaRows][kCaKeys]
    const int q0 = blockIdx.x * kCaRows, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_wait();
    const int qlen = input_len[b], klen = context_len[b];
    if (q0 >= qlen) return;
    const int kvh = h / (head_num / kv_head_num);
    const T *kc = k_cache + ((size_t)b * kv_head_num + kvh) * max_seq_len * D;
    const T *vc = v_cache + ((size_t)b * kv_head_num + kvh) * max_seq_len * D;
    const T *qb = q + (((size_t)b * head_num + h) * max_q_len) * D;
    for (int i = threadIdx.x; i < kCaRows * D; i += kCaThreads) {
        const int r = i / D, e = i % D;
        qs[i] = (q0 + r < qlen) ? Elem<T>::to_f(qb[(size_t)(q0 + r) * D + e]) : 0.0f;
    }
    constexpr int RPW = kCaRows / 4;  // rows per warp
    constexpr int DPL = 8;            // output dims per lane (head_size <= 256)
    float m_run[RPW], l_run[RPW], o[RPW][DPL];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        m_run[r] = FLT_MIN, l_run[r] = 0.0f;
#pragma unroll
        for (int j = 0; j < DPL; ++j) o[r][j] = 0.0f;
    }
    // last key any row of this tile may see
    const int q_hi = min(q0 + kCaRows, qlen) - 1;
    const int k_hi = min(klen, q_hi + (klen - qlen) + 1);
    for (int k0 = 0; k0 < k_hi; k0 += kCaKeys) {
        __syncthreads();
        for (int i = threadIdx.x; i < kCaKeys * D; i += kCaThreads) {
            const int r = i / D, e = i % D;
            const bool ok = k0 + r < k_hi;
            ks[r * DP + e] = ok ? Elem<T>::to_f(kc[(size_t)(k0 + r) * D + e]) : 0.0f;
            vs[r * D + e] = ok ? Elem<T>::to_f(vc[(size_t)(k0 + r) * D + e]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int row = warp * RPW + r, qi = q0 + row;
            if (qi >= qlen) continue;  // warp-uniform
            const int lim = qi + (klen - qlen);  // keys <= lim are visible
            float s[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int kk = lane + c * 32, kg = k0 + kk;
                float acc = 0.0f;
                for (int e = 0; e < D; ++e) acc = fmaf(qs[row * D + e], ks[kk * DP + e], acc);
                s[c] = (kg < klen && kg <= lim) ? scale * acc : -INFINITY;
            }
            float mx = warp_max(fmaxf(s[0], s[1]));
            const float m_new = fmaxf(m_run[r], mx);
            const float corr = expf(m_run[r] - m_new);
            float psum = 0.0f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float p = s[c] == -INFINITY ? 0.0f : expf(s[c] - m_new);
                ps[row * kCaKeys + lane + c * 32] = p;
                psum += p;
            }
            psum = warp_sum(psum);
            l_run[r] = l_run[r] * corr + psum;
            m_run[r] = m_new;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < DPL; ++j) o[r][j] *= corr;
            for (int kk = 0; kk < kCaKeys; ++kk) {
                const float p = ps[row * kCaKeys + kk];
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const int e = lane + 32 * j;
                    if (e < D) o[r][j] = fmaf(p, vs[kk * D + e], o[r][j]);
                }
            }
            __syncwarp();
        }
    }
    (void)cum_seqlens_or_null;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int qi = q0 + warp * RPW + r;
        if (qi >= qlen) continue;
        // un-padded token index: padded slot = b*max_q_len + qi = t + padding_offset[t]; with left-packed sequences
        // t = padded - (pad slots before) and padding_offset is constant inside a sequence, so search is not needed:
        // the caller passes the per-sequence offset through padding_offset of the sequence's first token.
        const int t = b * max_q_len + qi - padding_offset[b];
        const float inv = 1.0f / (l_run[r] + 1e-6f);
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            const int e = lane + 32 * j;
            if (e < D) out[((size_t)t * head_num + h) * D + e] = Elem<T>::from_f(o[r][j] * inv);
        }
    }
}

// per-sequence pad offset table for the fused attention: seq_off[b] = padding_offset[cum[b]] = b*max_q - cum[b]
__global__ void seq_offset_kernel(int *seq_off, const int *input_len, int batch, int max_q_len) {
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int total = 0;
        for (int b = 0; b < batch; ++b) {
            seq_off[b] = b * max_q_len - total;
            total += input_len[b];
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_cal_padding_offset(int *padding_offset, int *cum_seqlens, const int *input_lengths, int batch, int max_q_len,
                            b200_stream_t stream) {
    B200_REQUIRE(padding_offset && cum_seqlens && input_lengths, "cal_padding_offset: null pointer");
    B200_REQUIRE(batch >= 0 && max_q_len >= 0, "cal_padding_offset: bad shape");
    B200_REQUIRE((size_t)(batch + 1) * sizeof(int) <= 48 * 1024, "cal_padding_offset: batch %d too large", batch);
    launch_pdl(padding_offset_kernel, dim3(1), dim3(1024), (size_t)(batch + 1) * sizeof(int), as_stream(stream), true, padding_offset,
               cum_seqlens, input_lengths, batch, max_q_len);
    return cuda_status("cal_padding_offset launch");
}

int b200_build_causal_masks(void *mask, const int *q_lens, const int *k_lens, int batch, int max_q_len, int max_k_len, int dtype,
                            b200_stream_t stream) {
    B200_REQUIRE(mask && q_lens && k_lens, "build_causal_masks: null pointer");
    B200_REQUIRE(batch >= 0 && max_q_len >= 0 && max_k_len >= 0, "build_causal_masks: bad shape");
    const size_t total = (size_t)batch * max_q_len * max_k_len;
    if (total == 0) return B200_OK;
    size_t g = (total + 255) / 256, cap = (size_t)sm_count() * 16;
    B200_DISPATCH_DTYPE(dtype, launch_pdl(causal_mask_kernel<T>, dim3((unsigned)(g < cap ? g : cap)), dim3(256), 0, as_stream(stream),
                                          true, (T *)mask, q_lens, k_lens, max_q_len, max_k_len, total));
    return cuda_status("build_causal_masks launch");
}

int b200_qkv_bias_transpose_rope(void *q, void *k, void *v, const void *qkv, const void *qkv_bias, const int *padding_offset,
                                 const int *history_len, const int *input_len, int batch, int seq_len, int num_tokens,
                                 int head_num, int kv_head_num, int head_size, int rotary_dim, float rotary_base, int dtype,
                                 b200_stream_t stream) {
    (void)qkv_bias;   // accepted and ignored, like the reference kernel
    (void)input_len;  // unused by the reference kernel as well
    (void)batch;
    B200_REQUIRE(q && k && v && qkv && padding_offset && history_len, "qkv_bias_transpose_rope: null pointer");
    B200_REQUIRE(num_tokens >= 0 && seq_len > 0 && head_num > 0 && kv_head_num > 0 && head_size > 0, "qkv_bias_transpose_rope: bad shape");
    B200_REQUIRE(rotary_dim >= 0 && rotary_dim <= head_size && rotary_dim % 2 == 0, "qkv_bias_transpose_rope: bad rotary_dim");
    B200_REQUIRE(head_num + 2 * kv_head_num <= 65535, "qkv_bias_transpose_rope: too many heads");
    if (num_tokens == 0) return B200_OK;
    const int threads = ((head_size / 2 + 31) / 32) * 32;
    dim3 grid(num_tokens, head_num + 2 * kv_head_num);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(qkv_rope_kernel<T>, grid, dim3(threads < 32 ? 32 : (threads > 256 ? 256 : threads)), 0,
                                          as_stream(stream), true, (T *)q, (T *)k, (T *)v, (const T *)qkv, padding_offset, history_len,
                                          seq_len, head_num, kv_head_num, head_size, rotary_dim, rotary_base));
    return cuda_status("qkv_bias_transpose_rope launch");
}

int b200_concat_kv_cache(const void *k_src, const void *v_src, void *k_cache, void *v_cache, const int *cur_query_len,
                         const int *history_len, int layer, int batch, int kv_head_num, int max_q_len, int max_seq_len, int head_size,
                         int dtype, b200_stream_t stream) {
    B200_REQUIRE(k_src && v_src && k_cache && v_cache && cur_query_len && history_len, "concat_kv_cache: null pointer");
    B200_REQUIRE(batch >= 0 && kv_head_num > 0 && max_q_len >= 0 && max_seq_len > 0 && head_size > 0 && layer >= 0, "concat_kv_cache: bad shape");
    if (batch == 0 || max_q_len == 0) return B200_OK;
    B200_REQUIRE(kv_head_num <= 65535 && 2 * batch <= 65535, "concat_kv_cache: grid too large");
    const size_t loff = (size_t)layer * batch * kv_head_num * max_seq_len * head_size;
    dim3 grid(max_q_len, kv_head_num, 2 * batch);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(concat_kv_kernel<T>, grid, dim3(head_size < 128 ? 32 : 128), 0, as_stream(stream), true,
                                          (const T *)k_src, (const T *)v_src, (T *)k_cache + loff, (T *)v_cache + loff, cur_query_len,
                                          history_len, kv_head_num, max_q_len, max_seq_len, head_size));
    return cuda_status("concat_kv_cache launch");
}

int b200_repeat_kv_cache(const void *k_cache, const void *v_cache, void *k_dst, void *v_dst, const int *context_len, int layer,
                         int batch, int head_num, int kv_head_num, int max_k_len, int max_seq_len, int head_size, int dtype,
                         b200_stream_t stream) {
    B200_REQUIRE(k_cache && v_cache && k_dst && v_dst && context_len, "repeat_kv_cache: null pointer");
    B200_REQUIRE(batch >= 0 && head_num > 0 && kv_head_num > 0 && head_num % kv_head_num == 0 && max_k_len >= 0 && max_seq_len > 0 &&
                     head_size > 0 && layer >= 0,
                 "repeat_kv_cache: bad shape");
    if (batch == 0 || max_k_len == 0) return B200_OK;
    B200_REQUIRE(head_num <= 65535 && 2 * batch <= 65535, "repeat_kv_cache: grid too large");
    const size_t loff = (size_t)layer * batch * kv_head_num * max_seq_len * head_size;
    dim3 grid(max_k_len, head_num, 2 * batch);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(repeat_kv_kernel<T>, grid, dim3(head_size < 128 ? 32 : 128), 0, as_stream(stream), true,
                                          (const T *)k_cache + loff, (const T *)v_cache + loff, (T *)k_dst, (T *)v_dst, context_len,
                                          head_num, kv_head_num, max_k_len, max_seq_len, head_size));
    return cuda_status("repeat_kv_cache launch");
}

int b200_scale_mask_softmax(const void *qk, const void *mask, void *out, float scale, int batch, int head_num, int q_len, int k_len,
                            int dtype, b200_stream_t stream) {
    B200_REQUIRE(qk && mask && out, "scale_mask_softmax: null pointer");
    B200_REQUIRE(batch >= 0 && head_num > 0 && q_len >= 0 && k_len > 0, "scale_mask_softmax: bad shape");
    if (batch == 0 || q_len == 0) return B200_OK;
    B200_REQUIRE(head_num <= 65535 && batch <= 65535, "scale_mask_softmax: grid too large");
    dim3 grid(q_len, head_num, batch);
    B200_DISPATCH_DTYPE(dtype, launch_pdl(scale_mask_softmax_kernel<T>, grid, dim3(k_len <= 128 ? 64 : 256), 0, as_stream(stream), true,
                                          (const T *)qk, (const T *)mask, (T *)out, scale, head_num, q_len, k_len));
    return cuda_status("scale_mask_softmax launch");
}

int b200_transpose_remove_padding(const void *src, const int *padding_offset, void *dst, int num_tokens, int batch, int seq_len,
                                  int head_num, int head_size, int dtype, b200_stream_t stream) {
    (void)batch;
    B200_REQUIRE(src && padding_offset && dst, "transpose_remove_padding: null pointer");
    B200_REQUIRE(num_tokens >= 0 && seq_len > 0 && head_num > 0 && head_size > 0, "transpose_remove_padding: bad shape");
    if (num_tokens == 0) return B200_OK;
    B200_DISPATCH_DTYPE(dtype, launch_pdl(transpose_remove_padding_kernel<T>, dim3(num_tokens), dim3(256), 0, as_stream(stream), true,
                                          (const T *)src, padding_offset, (T *)dst, seq_len, head_num, head_size));
    return cuda_status("transpose_remove_padding launch");
}

static int context_attention_impl(const void *q, const void *k_cache, const void *v_cache, void *out, const int *input_len,
                                  const int *context_len, int layer, int batch, int head_num, int kv_head_num, int max_q_len, int max_seq_len,
                                  int head_size, float scale, int dtype, b200_stream_t stream, const int *block_table, int max_pages,
                                  int num_pages) {
    B200_REQUIRE(q && k_cache && v_cache && out && input_len && context_len, "context_attention: null pointer");
    B200_REQUIRE(batch >= 0 && head_num > 0 && kv_head_num > 0 && head_num % kv_head_num == 0 && max_q_len >= 0 && max_seq_len > 0 &&
                     head_size > 0 && head_size <= 256 && layer >= 0,
                 "context_attention: bad shape (head_size <= 256)");
    if (batch == 0 || max_q_len == 0) return B200_OK;
    B200_REQUIRE(head_num <= 65535 && batch <= 65535, "context_attention: grid too large");
    Workspace ws;
    if (!get_workspace(&ws)) return B200_ERR_WORKSPACE;
    B200_REQUIRE((size_t)batch * sizeof(int) <= ws.scratch_bytes, "context_attention: workspace too small");
    int *seq_off = reinterpret_cast<int *>(ws.scratch);
    cudaStream_t st = as_stream(stream);
    launch_pdl(seq_offset_kernel, dim3(1), dim3(32), 0, st, true, seq_off, input_len, batch, max_q_len);
    const size_t smem = sizeof(float) * ((size_t)kCaRows * head_size + (size_t)kCaKeys * (head_size + 1) + (size_t)kCaKeys * head_size +
                                         (size_t)kCaRows * kCaKeys);
    const size_t eb = dtype == B200_F32 ? 4 : 2;
    // contiguous cache [L, B, Hkv, S, d]; paged pool [L, num_pages, Hkv, 64, d]
    const size_t loff = block_table ? (size_t)layer * num_pages * kv_head_num * 64 * head_size * eb
                                    : (size_t)layer * batch * kv_head_num * max_seq_len * head_size * eb;
    // head size 128, 16-bit: tcgen05 / TMEM flash attention (context_attn_tc.cu); everything else: the SIMT tiles below
    {
        const int rc = launch_context_attention_tc(q, (const char *)k_cache + loff, (const char *)v_cache + loff, out, seq_off, input_len,
                                                   context_len, batch, head_num, kv_head_num, max_q_len, max_seq_len, head_size, scale, dtype, st,
                                                   block_table, max_pages, num_pages);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    if (block_table) {
        set_error("context_attention_paged: served by the tensor-core kernel only (head size 128, 16-bit); got head size %d, dtype %d", head_size, dtype);
        return B200_ERR_UNSUPPORTED;
    }
    dim3 grid((max_q_len + kCaRows - 1) / kCaRows, head_num, batch);
    B200_DISPATCH_DTYPE(dtype, {
        cudaFuncSetAttribute(context_attention_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_pdl(context_attention_kernel<T>, grid, dim3(kCaThreads), smem, st, true, (const T *)q,
                   (const T *)((const char *)k_cache + loff), (const T *)((const char *)v_cache + loff), (T *)out, (const int *)nullptr,
                   (const int *)seq_off, input_len, context_len, head_num, kv_head_num, max_q_len, max_seq_len, head_size, scale);
    });
    return cuda_status("context_attention launch");
}

int b200_context_attention(const void *q, const void *k_cache, const void *v_cache, void *out, const int *padding_offset,
                           const int *input_len, const int *context_len, int layer, int batch, int head_num, int kv_head_num,
                           int max_q_len, int max_seq_len, int head_size, int num_tokens, float scale, int dtype,
                           b200_stream_t stream) {
    (void)padding_offset;
    (void)num_tokens;
    return context_attention_impl(q, k_cache, v_cache, out, input_len, context_len, layer, batch, head_num, kv_head_num, max_q_len, max_seq_len,
                                  head_size, scale, dtype, stream, nullptr, 0, 0);
}

int b200_context_attention_paged(const void *q, const void *k_pool, const void *v_pool, void *out, const int *block_table, const int *input_len,
                                 const int *context_len, int layer, int batch, int head_num, int kv_head_num, int max_q_len, int num_pages,
                                 int max_pages_per_seq, int head_size, float scale, int dtype, b200_stream_t stream) {
    B200_REQUIRE(block_table, "context_attention_paged: null block table");
    B200_REQUIRE(num_pages >= 1 && max_pages_per_seq >= 1, "context_attention_paged: bad pool (num_pages %d, max_pages_per_seq %d)", num_pages, max_pages_per_seq);
    return context_attention_impl(q, k_pool, v_pool, out, input_len, context_len, layer, batch, head_num, kv_head_num, max_q_len,
                                  max_pages_per_seq * 64, head_size, scale, dtype, stream, block_table, max_pages_per_seq, num_pages);
}

}  // extern "C"
