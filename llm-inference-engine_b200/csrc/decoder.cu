// decoder.cu -- the fused decode engine: LlamaSelfDecoder<T>::forward (reference src/layers/self_decoder.cpp:24-122,
// self_attention.cpp:63-151, ffn.cpp:76-144) and the sampling tail (src/models/llama/llama.cpp:247-311) as a
// stream-ordered, allocation-free, sync-free, CUDA-graph-capturable launch sequence.
//
// Per layer (batch <= 4: five launches, all chained with programmatic dependent launch):
//   1. gemv  Wqkv   prologue: fold the pending FFN output into the residual stream + RMSNorm(gamma1)
//   2. attn         RoPE + qkv bias + KV append + split-KV attention + merge
//   3. gemv  Wo
//   4. gemv  Wgu    prologue: residual += attn out; (+ o bias); RMSNorm(gamma2);  epilogue: SwiGLU
//   5. gemv  Wdown
// For batch > 4 the same dataflow runs un-fused: norm kernel -> b200_linear (tensor-core GEMM) -> ...
// The residual stream rotates over engine-owned buffers so that no kernel writes a tensor another CTA of the same kernel
// still reads.
#include "attention_decode.cuh"
#include "gemv.cuh"

#include <math.h>
#include <stdlib.h>
#include <new>
#include <vector>

struct b200_decoder {
    b200_decoder_config_t cfg;
    std::vector<b200_layer_weights_t> layers;
    std::vector<char> layer_set;
    // scratch carve-up
    char *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *res[3] = {nullptr, nullptr, nullptr};  // the residual stream rotates over three buffers (see next_res)
    void *xn = nullptr;      // normalised activations (un-fused path)
    void *qkv = nullptr, *attn = nullptr, *y_attn = nullptr, *gu = nullptr, *act = nullptr, *y_ffn = nullptr;
    float *partials = nullptr;
    unsigned int *tickets = nullptr;
    float2 *rope_cs = nullptr;  // (cos, sin) per (position, rotary pair), filled once by set_scratch
    int max_splits = 0;
    int cur = 0;  // which res[] holds the residual stream
    const int *steps_dev = nullptr;  // per-row steps of a ragged batch (device int[batch]); set only inside b200_decoder_step_ragged / _paged
    // paged cache (set only inside b200_decoder_step_paged): the caches are page pools [L, num_pages, Hkv, 64, d] addressed through a block table
    const int *block_table = nullptr;
    int max_pages = 0, num_pages = 0;
    // fused tensor-parallel exchange (b200_decoder_tp_attach): every rank's exchange buffer as mapped in this process
    char *tp_base[b200::kTpMaxWorld] = {};
    bool tp_attached = false;
};

namespace b200 {

int launch_prefill_qkv_rope_cache(void *q, void *k_layer, void *v_layer, const void *qkv, const void *bias, const int *padding_offset,
                                  const int *history_len, int seq_len, int num_tokens, int head_num, int kv_head_num, int head_size,
                                  int max_seq_len, int rot_dim, float base, int dtype, cudaStream_t st, const int *block_table, int max_pages);

int launch_dequant_vec(const void *w, const void *scales, const void *zeros, void *dst, int N, int K, int w_format, int group, int dtype,
                       cudaStream_t st);  // linear.cu
int launch_gemm_tc(const void *x, const void *w, void *y, int M, int N, int K, int dtype, cudaStream_t st);  // gemm_tc.cu
int launch_gemm_tc_swiglu(const void *x, const void *w, void *act, int M, int inter, int K, int dtype, cudaStream_t st);  // gemm_tc.cu

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
// A kernel never writes the residual buffer other CTAs of the same launch still read: outputs go to the next buffer of the rotation.
static int next_res(int cur) { return (cur + 1) % 3; }
static size_t esize(int dtype) { return dtype == B200_F32 ? 4 : 2; }

struct Carve {
    size_t res, xn, qkv, attn, y, gu, act, partials, tickets, rope, total;
};
static Carve carve(const b200_decoder_config_t &c, int *max_splits) {
    Carve k;
    const size_t e = esize(c.dtype), B = c.max_batch;
    k.res = align_up(B * c.hidden * e);
    k.xn = align_up(B * (size_t)(c.hidden > c.inter_size ? c.hidden : c.inter_size) * e);
    k.qkv = align_up(B * (size_t)(c.head_num + 2 * c.kv_head_num) * c.head_size * e);
    k.attn = align_up(B * (size_t)c.head_num * c.head_size * e);
    k.y = align_up(B * c.hidden * e);
    k.gu = align_up(B * (size_t)2 * c.inter_size * e);
    k.act = align_up(B * (size_t)c.inter_size * e);
    // worst case number of KV splits (decode_attn_plan: chunks of >= 64 positions, at most kAttnMaxSplits splits)
    *max_splits = (c.max_seq_len + 63) / 64 < kAttnMaxSplits ? (c.max_seq_len + 63) / 64 : kAttnMaxSplits;
    // split-KV partials as flagged 8-byte words (attention_decode.cuh: ll_merge): twice the floats, zero between launches
    k.partials = align_up(2 * decode_attn_partials_floats(c.max_batch, c.head_num, c.kv_head_num, c.head_size, *max_splits) * sizeof(float));
    k.tickets = align_up((size_t)c.max_batch * c.kv_head_num * 2 * sizeof(unsigned int));  // x 2: half-group CTAs (GQA group of 8)
    k.rope = c.rotary_dim > 0 ? align_up((size_t)c.max_seq_len * (c.rotary_dim / 2) * sizeof(float2)) : 0;
    k.total = 3 * k.res + k.xn + k.qkv + k.attn + 2 * k.y + k.gu + k.act + k.partials + k.tickets + k.rope;
    return k;
}

// ---- exchange buffer of the fused tensor-parallel path (one per rank, peer-mapped everywhere):
//      [256, 512)  this rank's step counter (epoch) and error word
//      [512, ...)  partial sums as LL words {payload, flag} (common.cuh): [2 slots][world source ranks][max_batch * hidden elements]
constexpr size_t kTpEpochOff = 256, kTpErrorOff = 260, kTpDataOff = 512;
static size_t tp_slot_bytes(const b200_decoder_config_t &c) { return align_up(2 * (size_t)c.max_batch * c.hidden * esize(c.dtype)); }
// where the partial of block `seq` produced by rank `src` lives inside rank `owner`'s buffer
static void *tp_slot(const b200_decoder *d, int owner, int seq, int src) {
    return d->tp_base[owner] + kTpDataOff + ((size_t)(seq & 1) * d->cfg.tp_world + src) * tp_slot_bytes(d->cfg);
}
// descriptor for the consumer of block `seq` (seq >= 1): the P partials, all in this rank's own buffer
static TpExchange tp_consume(const b200_decoder *d, int seq) {
    TpExchange t = {};
    const b200_decoder_config_t &c = d->cfg;
    t.world = c.tp_world, t.rank = c.tp_rank, t.seq = seq;
    for (int r = 0; r < c.tp_world; ++r) t.peer_x[r] = tp_slot(d, c.tp_rank, seq, r);
    t.epoch = reinterpret_cast<const unsigned int *>(d->tp_base[c.tp_rank] + kTpEpochOff);
    t.error = reinterpret_cast<unsigned int *>(d->tp_base[c.tp_rank] + kTpErrorOff);
    return t;
}
// descriptor for the producer of block `seq`: this rank's slot in every rank's buffer
static TpPush tp_produce(const b200_decoder *d, int seq) {
    TpPush p = {};
    const b200_decoder_config_t &c = d->cfg;
    p.n = c.tp_world, p.seq = seq;
    for (int r = 0; r < c.tp_world; ++r) p.dst[r] = tp_slot(d, r, seq, c.tp_rank);
    p.epoch = reinterpret_cast<const unsigned int *>(d->tp_base[c.tp_rank] + kTpEpochOff);
    return p;
}
// batched (M > 4) path: the tensor-core GEMM wrote this rank's partial as a plain tensor; re-emit it as LL words into every rank's buffer
template <typename T>
__global__ void tp_push_kernel(const T *src, TpPush push, size_t n_vec) {
    pdl_wait();
    const unsigned int flag = tp_flag(push.epoch, push.seq);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x)
        tp_push_vec(push, flag, i * 4, ld_v4(reinterpret_cast<const char *>(src) + i * 16));
}
__global__ void tp_begin_step_kernel(unsigned int *epoch) {
    pdl_wait();
    if (threadIdx.x == 0) *epoch += 1;
}

// rows one fused GEMV launch can take: 16 for 16-bit models (tensor-core GEMV, gemv_mma.cuh), 4 for fp32 (SIMT GEMV)
static int gemv_max_rows(const b200_decoder_config_t &c) { return c.dtype != B200_F32 ? 16 : 4; }

// prologue (add residual / bias / RMSNorm) + linear (+ SwiGLU): fused GEMV for M <= 4 (8 quantised), un-fused otherwise.
// tp (optional): x is the fused all-reduce of every rank's partial (TpExchange) instead of a local tensor.
static int norm_linear(b200_decoder *d, const void *x, const void *res_in, void *res_out, const void *bias, const void *gamma,
                       const b200_linear_weight_t &w, int K, int N, bool swiglu, void *y, int M, cudaStream_t st, const TpExchange *tp = nullptr) {
    const b200_decoder_config_t &c = d->cfg;
    const bool b16 = c.dtype != B200_F32;
    // one token (any dtype), fp32 up to 4, quantised weights up to 4: the prologue runs inside the GEMV (SIMT kernels / gemv_q.cuh)
    if (M == 1 || (!b16 && M <= 4) || (c.w_format != B200_W_DENSE && M <= 4)) {
        GemvArgs a = {};
        a.w = w.w, a.scales = w.scales, a.zeros = w.zeros;
        a.x = x, a.y = y;
        a.res_in = res_in, a.res_out = res_out, a.bias = bias, a.gamma = gamma, a.eps = c.rmsnorm_eps, a.norm = 1;
        a.M = M, a.K = K, a.N = N, a.group = c.group, a.inter = swiglu ? N / 2 : 0;
        if (tp) a.tp = *tp;
        const int rc = launch_gemv_nk(a, c.dtype, c.w_format, swiglu, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    // more tokens: the prologue (tensor-parallel reduce, + residual, + bias, RMSNorm) runs ONCE in a small kernel; fused into the GEMV every
    // CTA would redo it for every token (gemv_mma.cuh)
    int rc = launch_norm_tp(c.dtype, x, d->xn, res_in, res_out, bias, gamma, c.rmsnorm_eps, M, K, tp, st);
    if (rc != B200_OK) return rc;
    if (b16 && M <= gemv_max_rows(c)) {  // tensor-core GEMV, SwiGLU in its epilogue
        GemvArgs a = {};
        a.w = w.w, a.scales = w.scales, a.zeros = w.zeros;
        a.x = d->xn, a.y = y;
        a.M = M, a.K = K, a.N = N, a.group = c.group, a.inter = swiglu ? N / 2 : 0;
        rc = launch_gemv_nk(a, c.dtype, c.w_format, swiglu, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    if (!swiglu) return b200_linear(d->xn, w.w, w.scales, w.zeros, y, M, K, N, c.dtype, c.w_format, B200_LAYOUT_NK, c.group, st);
    rc = b200_linear(d->xn, w.w, w.scales, w.zeros, d->gu, M, K, N, c.dtype, c.w_format, B200_LAYOUT_NK, c.group, st);
    if (rc != B200_OK) return rc;
    return b200_silu_and_mul(d->gu, y, M, N / 2, c.dtype, st);
}

// push_seq > 0 (fused tensor-parallel path): y is the partial of block push_seq and must reach every rank's exchange buffer
static int plain_linear(b200_decoder *d, const void *x, const b200_linear_weight_t &w, int K, int N, void *y, int M, cudaStream_t st,
                        int push_seq = 0) {
    const b200_decoder_config_t &c = d->cfg;
    if (M <= gemv_max_rows(c)) {
        GemvArgs a = {};
        a.w = w.w, a.scales = w.scales, a.zeros = w.zeros;
        a.x = x, a.y = y;
        a.M = M, a.K = K, a.N = N, a.group = c.group;
        if (push_seq > 0) a.push = tp_produce(d, push_seq);
        const int rc = launch_gemv_nk(a, c.dtype, c.w_format, false, st);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    int rc = b200_linear(x, w.w, w.scales, w.zeros, y, M, K, N, c.dtype, c.w_format, B200_LAYOUT_NK, c.group, st);
    if (rc != B200_OK || push_seq <= 0) return rc;
    const TpPush push = tp_produce(d, push_seq);
    const size_t n_vec = (size_t)M * N * esize(c.dtype) / 16;
    const int grid = (int)((n_vec + 255) / 256 < 64 ? (n_vec + 255) / 256 : 64);
    B200_DISPATCH_DTYPE(c.dtype, launch_pdl(tp_push_kernel<T>, dim3(grid), dim3(256), 0, st, true, (const T *)y, push, n_vec));
    return cuda_status("tp_push launch");
}

// hidden <- residual + last FFN output (n_vec 16-byte vectors); under tensor parallelism the FFN output is the fused all-reduce of
// every rank's partial (LL words, rank order, rounded to T)
template <typename T>
__global__ void fold_kernel(T *out, const T *a, const T *b, size_t n_vec, size_t n_tail, const TpExchange tp) {
    constexpr int V = Elem<T>::kVec;
    pdl_wait();
    const unsigned int want = tp.world > 1 ? tp_flag(tp.epoch, tp.seq) : 0u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        float add[V], res[V];
        if (tp.world > 1) {
            tp_reduce_vec<T>(tp, want, i * 4, add);
        } else if (b) {
            unpack16<T>(ld_v4(b + i * V), add);
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) add[j] = 0.0f;
        }
        unpack16<T>(ld_v4(a + i * V), res);
#pragma unroll
        for (int j = 0; j < V; ++j) res[j] += add[j];
        st_v4(out + i * V, pack16<T>(res));
    }
    // elements past the last whole vector (hidden sizes that are not a multiple of the vector length: single-GPU only)
    for (size_t i = n_vec * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec * V + n_tail; i += (size_t)gridDim.x * blockDim.x)
        out[i] = Elem<T>::from_f(Elem<T>::to_f(a[i]) + (b ? Elem<T>::to_f(b[i]) : 0.0f));
}

static int check_ready(const b200_decoder *d, int batch) {
    if (!d) {
        set_error("decoder: null handle");
        return B200_ERR_INVALID_ARG;
    }
    if (!d->scratch) {
        set_error("decoder: scratch not set (b200_decoder_set_scratch)");
        return B200_ERR_STATE;
    }
    if (batch < 1 || batch > d->cfg.max_batch) {
        set_error("decoder: batch %d outside [1, %d]", batch, d->cfg.max_batch);
        return B200_ERR_INVALID_ARG;
    }
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

b200_decoder_t *b200_decoder_create(const b200_decoder_config_t *cfg) {
    if (!cfg) {
        set_error("decoder_create: null config");
        return nullptr;
    }
    const b200_decoder_config_t &c = *cfg;
    if (c.hidden <= 0 || c.head_num <= 0 || c.kv_head_num <= 0 || c.head_size <= 0 || c.inter_size <= 0 || c.num_layers <= 0 ||
        c.max_seq_len <= 0 || c.max_batch <= 0 || c.head_num % c.kv_head_num != 0) {
        set_error("decoder_create: bad model shape");
        return nullptr;
    }
    if (c.rotary_dim < 0 || c.rotary_dim > c.head_size || c.rotary_dim % 2 != 0) {
        set_error("decoder_create: rotary_dim %d must be even and within [0, head_size]", c.rotary_dim);
        return nullptr;
    }
    if (c.dtype != B200_F32 && c.dtype != B200_F16 && c.dtype != B200_BF16) {
        set_error("decoder_create: unknown dtype %d", c.dtype);
        return nullptr;
    }
    if (c.w_format < B200_W_DENSE || c.w_format > B200_W_INT4 || (c.w_format != B200_W_DENSE && c.dtype == B200_F32)) {
        set_error("decoder_create: weight format %d not available for dtype %d", c.w_format, c.dtype);
        return nullptr;
    }
    if (c.w_format == B200_W_INT4 && (c.group < 32 || c.group % 32 != 0 || c.hidden % c.group != 0 || c.inter_size % c.group != 0 ||
                                     (c.head_num * c.head_size) % c.group != 0)) {
        set_error("decoder_create: INT4 group %d must divide every K dimension", c.group);
        return nullptr;
    }
    b200_decoder *d = new (std::nothrow) b200_decoder();
    if (!d) return nullptr;
    d->cfg = c;
    d->layers.resize(c.num_layers);
    d->layer_set.assign(c.num_layers, 0);
    return d;
}

void b200_decoder_destroy(b200_decoder_t *dec) { delete dec; }

int b200_decoder_get_config(const b200_decoder_t *dec, b200_decoder_config_t *cfg) {
    B200_REQUIRE(dec && cfg, "decoder_get_config: null argument");
    *cfg = dec->cfg;
    return B200_OK;
}

int b200_decoder_set_layer(b200_decoder_t *dec, int layer, const b200_layer_weights_t *w) {
    B200_REQUIRE(dec && w, "decoder_set_layer: null argument");
    B200_REQUIRE(layer >= 0 && layer < dec->cfg.num_layers, "decoder_set_layer: layer %d out of range", layer);
    B200_REQUIRE(w->attn_norm_gamma && w->ffn_norm_gamma && w->qkv.w && w->o.w && w->gate_up.w && w->down.w,
                 "decoder_set_layer: missing weight pointer");
    if (dec->cfg.w_format != B200_W_DENSE)
        B200_REQUIRE(w->qkv.scales && w->o.scales && w->gate_up.scales && w->down.scales, "decoder_set_layer: missing scales");
    if (dec->cfg.w_format == B200_W_INT4)
        B200_REQUIRE(w->qkv.zeros && w->o.zeros && w->gate_up.zeros && w->down.zeros, "decoder_set_layer: missing zero points");
    dec->layers[layer] = *w;
    dec->layer_set[layer] = 1;
    return B200_OK;
}

size_t b200_decoder_scratch_bytes(const b200_decoder_t *dec) {
    if (!dec) return 0;
    int ms;
    return carve(dec->cfg, &ms).total;
}

int b200_decoder_set_scratch(b200_decoder_t *dec, void *ptr, size_t bytes) {
    B200_REQUIRE(dec && ptr, "decoder_set_scratch: null argument");
    B200_REQUIRE(aligned16(ptr) && ((uintptr_t)ptr & 255) == 0, "decoder_set_scratch: pointer must be 256-byte aligned");
    int ms;
    const Carve k = carve(dec->cfg, &ms);
    B200_REQUIRE(bytes >= k.total, "decoder_set_scratch: need %zu bytes, got %zu", k.total, bytes);
    char *p = (char *)ptr;
    dec->scratch = p, dec->scratch_bytes = bytes, dec->max_splits = ms;
    dec->res[0] = p, p += k.res;
    dec->res[1] = p, p += k.res;
    dec->res[2] = p, p += k.res;
    dec->xn = p, p += k.xn;
    dec->qkv = p, p += k.qkv;
    dec->attn = p, p += k.attn;
    dec->y_attn = p, p += k.y;
    dec->y_ffn = p, p += k.y;
    dec->gu = p, p += k.gu;
    dec->act = p, p += k.act;
    dec->partials = (float *)p, p += k.partials;
    dec->tickets = (unsigned int *)p, p += k.tickets;
    if (cudaMemset(dec->tickets, 0, k.tickets) != cudaSuccess || cudaMemset(dec->partials, 0, k.partials) != cudaSuccess)
        return cuda_status("decoder_set_scratch memset");
    dec->rope_cs = nullptr;
    if (k.rope) {
        dec->rope_cs = (float2 *)p;
        const int rc = launch_rope_table(dec->rope_cs, dec->cfg.max_seq_len, dec->cfg.rotary_dim, dec->cfg.rotary_base, nullptr);
        if (rc != B200_OK) return rc;
        if (cudaStreamSynchronize(nullptr) != cudaSuccess) return cuda_status("decoder_set_scratch rope table");
    }
    dec->cur = 0;
    return B200_OK;
}

// RoPE + qkv bias + KV append + split-KV attention + merge for one layer: dec->qkv -> dec->attn
static int launch_layer_attention(b200_decoder_t *dec, int layer, void *k_cache, void *v_cache, int batch, int step, cudaStream_t st) {
    const b200_decoder_config_t &c = dec->cfg;
    const b200_layer_weights_t &w = dec->layers[layer];
    DecodeAttnArgs a = {};
    const size_t layer_off = dec->block_table ? (size_t)layer * dec->num_pages * c.kv_head_num * kAttnPageSize * c.head_size * esize(c.dtype)
                                              : (size_t)layer * c.max_batch * c.kv_head_num * c.max_seq_len * c.head_size * esize(c.dtype);
    a.qkv = dec->qkv, a.bias = w.qkv_bias;
    a.k_cache = (char *)k_cache + layer_off, a.v_cache = (char *)v_cache + layer_off;
    a.out = dec->attn;
    a.batch = batch, a.head_num = c.head_num, a.kv_head_num = c.kv_head_num, a.head_size = c.head_size;
    a.max_seq_len = dec->block_table ? dec->max_pages * kAttnPageSize : c.max_seq_len, a.step = step, a.steps = dec->steps_dev;
    a.block_table = dec->block_table, a.max_pages = dec->max_pages;
    a.apply_rope = c.rotary_dim > 0, a.rot_dim = c.rotary_dim, a.rot_base = c.rotary_base;
    a.nsplit = decode_attn_plan(batch, c.kv_head_num, step, &a.chunk);
    a.partials = dec->partials, a.tickets = dec->tickets;
    a.ll_merge = 1;  // the engine's own zero-initialised region: flagged words, no fence / ticket (the stand-alone launcher keeps them)
    a.rope_cs = dec->rope_cs;
    a.prefetch = 1;  // the kernel in front of this one is the QKV linear: it does not touch the cache
    return launch_decode_attn(a, c.dtype, st);
}

static int attn_block_impl(b200_decoder_t *dec, int layer, void *hidden, const void *pending, void *k_cache, void *v_cache, void *partial,
                           int batch, int step, b200_stream_t stream, const TpExchange *tp, int push_seq = 0) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    const b200_decoder_config_t &c = dec->cfg;
    B200_REQUIRE(layer >= 0 && layer < c.num_layers && dec->layer_set[layer], "decoder: layer %d not set", layer);
    B200_REQUIRE(k_cache && v_cache && partial, "decoder_attn_block: null pointer");
    B200_REQUIRE(pending || hidden, "decoder_attn_block: need `hidden` (first block) or `pending`");
    B200_REQUIRE(step >= 1 && step <= c.max_seq_len, "decoder: step %d outside [1, %d]", step, c.max_seq_len);
    cudaStream_t st = as_stream(stream);
    const b200_layer_weights_t &w = dec->layers[layer];
    const int qkv_n = (c.head_num + 2 * c.kv_head_num) * c.head_size;
    // 1. residual fold + RMSNorm + QKV
    void *res_out = dec->res[next_res(dec->cur)];
    rc = norm_linear(dec, pending ? pending : hidden, pending ? dec->res[dec->cur] : nullptr, res_out, nullptr, w.attn_norm_gamma, w.qkv,
                     c.hidden, qkv_n, false, dec->qkv, batch, st, pending ? tp : nullptr);
    if (rc != B200_OK) return rc;
    dec->cur = next_res(dec->cur);
    // 2. attention
    rc = launch_layer_attention(dec, layer, k_cache, v_cache, batch, step, st);
    if (rc != B200_OK) return rc;
    // 3. O projection (row-sharded under TP: `partial` is this rank's partial sum)
    return plain_linear(dec, dec->attn, w.o, c.head_num * c.head_size, c.hidden, partial, batch, st, push_seq);
}

int b200_decoder_attn_block(b200_decoder_t *dec, int layer, void *hidden, const void *pending, void *k_cache, void *v_cache,
                            void *partial, int batch, int step, b200_stream_t stream) {
    return attn_block_impl(dec, layer, hidden, pending, k_cache, v_cache, partial, batch, step, stream, nullptr);
}

static int ffn_block_impl(b200_decoder_t *dec, int layer, const void *pending, void *partial, int batch, b200_stream_t stream,
                          const TpExchange *tp, int push_seq = 0) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    const b200_decoder_config_t &c = dec->cfg;
    B200_REQUIRE(layer >= 0 && layer < c.num_layers && dec->layer_set[layer], "decoder: layer %d not set", layer);
    B200_REQUIRE(pending && partial, "decoder_ffn_block: null pointer");
    cudaStream_t st = as_stream(stream);
    const b200_layer_weights_t &w = dec->layers[layer];
    // 4. residual += attention output; + o bias (tp rank 0 semantics: bias is replicated, added after the reduce);
    //    RMSNorm; gate/up; SwiGLU
    rc = norm_linear(dec, pending, dec->res[dec->cur], dec->res[next_res(dec->cur)], w.o_bias, w.ffn_norm_gamma, w.gate_up, c.hidden,
                     2 * c.inter_size, true, dec->act, batch, st, tp);
    if (rc != B200_OK) return rc;
    dec->cur = next_res(dec->cur);
    // 5. down projection
    return plain_linear(dec, dec->act, w.down, c.inter_size, c.hidden, partial, batch, st, push_seq);
}

int b200_decoder_ffn_block(b200_decoder_t *dec, int layer, void *hidden, const void *pending, void *partial, int batch,
                           b200_stream_t stream) {
    (void)hidden;
    return ffn_block_impl(dec, layer, pending, partial, batch, stream, nullptr);
}

static int fold_impl(b200_decoder_t *dec, void *hidden, const void *pending, int batch, b200_stream_t stream, const TpExchange *tpx) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(hidden, "decoder_fold: null pointer");
    const size_t n = (size_t)batch * dec->cfg.hidden, vec = 16 / esize(dec->cfg.dtype);
    // whole rows must be vectors for the vector path (row starts stay 16-byte aligned); otherwise everything goes through the scalar tail
    const bool vec_ok = dec->cfg.hidden % vec == 0 && aligned16(hidden) && (!pending || aligned16(pending));
    B200_REQUIRE(vec_ok || !tpx, "decoder_fold: tensor-parallel fold needs 16-byte rows");
    const size_t n_vec = vec_ok ? n / vec : 0, n_tail = n - n_vec * vec;
    const size_t work = n_vec + n_tail;
    const int grid = (int)((work + 255) / 256 < 1024 ? (work + 255) / 256 : 1024);
    TpExchange tp = {};
    if (tpx) tp = *tpx;
    B200_DISPATCH_DTYPE(dec->cfg.dtype, launch_pdl(fold_kernel<T>, dim3(grid ? grid : 1), dim3(256), 0, as_stream(stream), true, (T *)hidden,
                                                   (const T *)dec->res[dec->cur], (const T *)pending, n_vec, n_tail, tp));
    return cuda_status("decoder_fold launch");
}

int b200_decoder_fold(b200_decoder_t *dec, void *hidden, const void *pending, int batch, b200_stream_t stream) {
    return fold_impl(dec, hidden, pending, batch, stream, nullptr);
}

// ---------------------------------------------------------------- fused tensor-parallel step (no NCCL on the path)
size_t b200_decoder_tp_buffer_bytes(const b200_decoder_t *dec) {
    if (!dec) return 0;
    return kTpDataOff + 2 * (size_t)dec->cfg.tp_world * tp_slot_bytes(dec->cfg);
}

int b200_tp_alloc_exported(size_t bytes, void **ptr, void *handle64) {
    B200_REQUIRE(ptr && handle64 && bytes > 0, "tp_alloc_exported: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return cuda_status("tp_alloc_exported cudaMalloc");
    if (cudaMemset(p, 0, bytes) != cudaSuccess) return cuda_status("tp_alloc_exported memset");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        const int rc = cuda_status("cudaIpcGetMemHandle");
        cudaFree(p);
        return rc;
    }
    memcpy(handle64, &h, 64);
    if (cudaDeviceSynchronize() != cudaSuccess) return cuda_status("tp_alloc_exported sync");
    *ptr = p;
    return B200_OK;
}

int b200_tp_open(const void *handle64, void **ptr) {
    B200_REQUIRE(ptr && handle64, "tp_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return cuda_status("cudaIpcOpenMemHandle");
    *ptr = p;
    return B200_OK;
}

int b200_decoder_tp_attach(b200_decoder_t *dec, int world, int rank, void *const *bases) {
    B200_REQUIRE(dec && bases, "decoder_tp_attach: null argument");
    B200_REQUIRE(world == dec->cfg.tp_world && rank == dec->cfg.tp_rank, "decoder_tp_attach: world/rank differ from the decoder config");
    B200_REQUIRE(world >= 2 && world <= kTpMaxWorld, "decoder_tp_attach: world %d outside [2, %d]", world, kTpMaxWorld);
    B200_REQUIRE((dec->cfg.hidden * esize(dec->cfg.dtype)) % 16 == 0, "decoder_tp_attach: hidden rows must be 16-byte multiples");
    for (int r = 0; r < world; ++r) {
        B200_REQUIRE(bases[r] != nullptr && ((uintptr_t)bases[r] & 255) == 0, "decoder_tp_attach: buffer of rank %d is null or unaligned", r);
        dec->tp_base[r] = (char *)bases[r];
    }
    dec->tp_attached = true;
    return B200_OK;
}

int b200_decoder_tp_error(const b200_decoder_t *dec) {
    if (!dec || !dec->tp_attached) return 0;
    unsigned int e = 0;
    if (cudaMemcpy(&e, dec->tp_base[dec->cfg.tp_rank] + kTpErrorOff, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)e;
}

int b200_decoder_step_tp(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch, int step, b200_stream_t stream) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(dec->tp_attached, "decoder_step_tp: call b200_decoder_tp_attach first");
    B200_REQUIRE(hidden && k_cache && v_cache, "decoder_step_tp: null pointer");
    const b200_decoder_config_t &c = dec->cfg;
    B200_REQUIRE(2 * c.num_layers + 1 < 4096, "decoder_step_tp: too many layers for the flag encoding");
    NvtxRange range("b200 decode step (tensor parallel)");
    cudaStream_t st = as_stream(stream);
    launch_pdl(tp_begin_step_kernel, dim3(1), dim3(32), 0, st, true, reinterpret_cast<unsigned int *>(dec->tp_base[c.tp_rank] + kTpEpochOff));
    if ((rc = cuda_status("tp_begin_step launch")) != B200_OK) return rc;
    int seq = 0;  // sequence number of the last partial produced in this step
    // `partial` / `pending` below are plain staging tensors: the batched (tensor-core) linears write their partial there before it is
    // re-emitted as LL words; the fused GEMVs push straight from their epilogue and never touch them
    for (int l = 0; l < c.num_layers; ++l) {
        // attention block: consumes the previous layer's FFN partials (seq), pushes its O-projection partial as block seq + 1
        TpExchange tin = seq ? tp_consume(dec, seq) : TpExchange{};
        rc = attn_block_impl(dec, l, hidden, seq ? dec->y_ffn : nullptr, k_cache, v_cache, dec->y_attn, batch, step, stream, seq ? &tin : nullptr, seq + 1);
        if (rc != B200_OK) return rc;
        ++seq;
        TpExchange tmid = tp_consume(dec, seq);
        rc = ffn_block_impl(dec, l, dec->y_attn, dec->y_ffn, batch, stream, &tmid, seq + 1);
        if (rc != B200_OK) return rc;
        ++seq;
    }
    TpExchange tlast = tp_consume(dec, seq);
    return fold_impl(dec, hidden, dec->y_ffn, batch, stream, &tlast);
}

int b200_decoder_step(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch, int step, int layer_begin,
                      int layer_end, b200_stream_t stream) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(dec->cfg.tp_world <= 1, "decoder_step: tensor-parallel engines drive attn_block / ffn_block themselves");
    B200_REQUIRE(hidden && k_cache && v_cache, "decoder_step: null pointer");
    B200_REQUIRE(layer_begin >= 0 && layer_end <= dec->cfg.num_layers && layer_begin < layer_end, "decoder_step: bad layer range");
    NvtxRange range(dec->block_table ? "b200 decode step (paged)" : dec->steps_dev ? "b200 decode step (ragged)" : "b200 decode step");
    const void *pending = nullptr;
    for (int l = layer_begin; l < layer_end; ++l) {
        NvtxRange layer_range("layer");
        rc = b200_decoder_attn_block(dec, l, hidden, pending, k_cache, v_cache, dec->y_attn, batch, step, stream);
        if (rc != B200_OK) return rc;
        rc = b200_decoder_ffn_block(dec, l, hidden, dec->y_attn, dec->y_ffn, batch, stream);
        if (rc != B200_OK) return rc;
        pending = dec->y_ffn;
    }
    return b200_decoder_fold(dec, hidden, pending, batch, stream);
}

// Ragged batch: row b sits at its own 1-based position steps[b] (device memory; rows of different prompt lengths decoding together).
// Only the attention kernel looks at positions (RoPE angle, cache row appended, number of cached rows read); its split plan is made for
// max_step and rows that end earlier leave empty partials.  Everything else in the step is position-free.
int b200_decoder_step_ragged(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, int batch, const int *steps, int max_step,
                             int layer_begin, int layer_end, b200_stream_t stream) {
    B200_REQUIRE(dec && steps, "decoder_step_ragged: null argument");
    dec->steps_dev = steps;
    const int rc = b200_decoder_step(dec, hidden, k_cache, v_cache, batch, max_step, layer_begin, layer_end, stream);
    dec->steps_dev = nullptr;
    return rc;
}

// The same step over a PAGED cache: k_pool / v_pool [L, num_pages, Hkv, 64, d]; row b's position p lives in page
// block_table[b * max_pages_per_seq + p / 64].  Rows are ragged by construction (steps[b]); batch <= max_batch: the pool has no batch
// dimension, so sequences join and leave the batch between steps without moving a byte of cache.
int b200_decoder_step_paged(b200_decoder_t *dec, void *hidden, void *k_pool, void *v_pool, const int *block_table, const int *steps, int batch,
                            int max_step, int num_pages, int max_pages_per_seq, int layer_begin, int layer_end, b200_stream_t stream) {
    B200_REQUIRE(dec && block_table && steps, "decoder_step_paged: null argument");
    B200_REQUIRE(num_pages >= 1 && max_pages_per_seq >= 1, "decoder_step_paged: bad pool (num_pages %d, max_pages_per_seq %d)", num_pages, max_pages_per_seq);
    B200_REQUIRE(max_step <= max_pages_per_seq * kAttnPageSize, "decoder_step_paged: max_step %d exceeds the block table's reach (%d pages of %d)",
                 max_step, max_pages_per_seq, kAttnPageSize);
    B200_REQUIRE(dec->cfg.head_size == 128, "decoder_step_paged: the paged attention kernel serves head size 128 (got %d)", dec->cfg.head_size);
    dec->steps_dev = steps, dec->block_table = block_table, dec->max_pages = max_pages_per_seq, dec->num_pages = num_pages;
    const int rc = b200_decoder_step(dec, hidden, k_pool, v_pool, batch, max_step, layer_begin, layer_end, stream);
    dec->steps_dev = nullptr, dec->block_table = nullptr, dec->max_pages = dec->num_pages = 0;
    return rc;
}

// Diagnostic for roofline measurements: exactly the weight-streaming launches of b200_decoder_step (four GEMVs per layer) without the
// attention kernels and the final fold (a tensor-parallel engine runs them on its shard, without the exchange).  The activations are
// whatever the scratch buffers hold: call it after at least one real step;
// results are meaningless, timing is not.
int b200_decoder_linears_only(b200_decoder_t *dec, int batch, int *n_launches, b200_stream_t stream) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    const b200_decoder_config_t &c = dec->cfg;
    cudaStream_t st = as_stream(stream);
    const int L = c.num_layers, qkv_n = (c.head_num + 2 * c.kv_head_num) * c.head_size;
    for (int l = 0; l < L; ++l) B200_REQUIRE(dec->layer_set[l], "decoder_linears_only: layer %d not set", l);
    int launches = 0;
    for (int l = 0; l < L; ++l) {
        const b200_layer_weights_t &w = dec->layers[l];
        rc = norm_linear(dec, dec->y_ffn, dec->res[dec->cur], dec->res[next_res(dec->cur)], nullptr, w.attn_norm_gamma, w.qkv, c.hidden, qkv_n, false,
                         dec->qkv, batch, st);
        if (rc != B200_OK) return rc;
        dec->cur = next_res(dec->cur);
        if ((rc = plain_linear(dec, dec->attn, w.o, c.head_num * c.head_size, c.hidden, dec->y_attn, batch, st)) != B200_OK) return rc;
        rc = norm_linear(dec, dec->y_attn, dec->res[dec->cur], dec->res[next_res(dec->cur)], w.o_bias, w.ffn_norm_gamma, w.gate_up, c.hidden,
                         2 * c.inter_size, true, dec->act, batch, st);
        if (rc != B200_OK) return rc;
        dec->cur = next_res(dec->cur);
        if ((rc = plain_linear(dec, dec->act, w.down, c.inter_size, c.hidden, dec->y_ffn, batch, st)) != B200_OK) return rc;
        // 4 linears; for 2+ tokens of a 16-bit model (5+ in fp32) each of the two prologues is its own small kernel (norm_linear)
        launches += (batch == 1 || ((c.dtype == B200_F32 || c.w_format != B200_W_DENSE) && batch <= 4)) ? 4 : 6;
    }
    if (n_launches) *n_launches = launches;
    return B200_OK;
}

static size_t prefill_carve(const b200_decoder_config_t &c, int batch, int mq, int T, size_t *off /*[12]*/) {
    const size_t e = esize(c.dtype);
    const size_t qkv_heads = (size_t)c.head_num + 2 * c.kv_head_num;
    // quantised weights: one dequantised linear at a time (the largest: gate_up) for the tensor-core GEMM
    size_t wq = 0;
    if (c.w_format != B200_W_DENSE && c.dtype != B200_F32) {
        const size_t h = c.hidden, qn = qkv_heads * c.head_size, on = (size_t)c.head_num * c.head_size, in = c.inter_size;
        wq = h * qn;
        if (on * h > wq) wq = on * h;
        if (2 * in * h > wq) wq = 2 * in * h;
    }
    const size_t sizes[12] = {
        align_up((size_t)T * c.hidden * e),                                // 0 res
        align_up((size_t)T * c.hidden * e),                                // 1 xn
        align_up((size_t)T * qkv_heads * c.head_size * e),                 // 2 qkv
        align_up((size_t)batch * c.head_num * mq * c.head_size * e),       // 3 q padded
        align_up((size_t)batch * c.kv_head_num * mq * c.head_size * e),    // 4 k padded
        align_up((size_t)batch * c.kv_head_num * mq * c.head_size * e),    // 5 v padded
        align_up((size_t)T * c.head_num * c.head_size * e),                // 6 attention out [T, H, d]
        align_up((size_t)T * c.hidden * e),                                // 7 y (O / down output)
        align_up((size_t)T * 2 * c.inter_size * e),                        // 8 gate_up
        align_up(((size_t)batch * mq + batch + 1) * sizeof(int)),          // 9 padding_offset + cum_seqlens
        align_up((size_t)T * c.inter_size * e),                            // 10 SwiGLU activation
        align_up(wq * e),                                                   // 11 dequantised weights of the linear in flight
    };
    size_t total = 0;
    for (int i = 0; i < 12; ++i) {
        off[i] = total;
        total += sizes[i];
    }
    return total;
}

size_t b200_decoder_prefill_scratch_bytes(const b200_decoder_t *dec, int batch, int max_q_len, int num_tokens) {
    if (!dec || batch < 1 || max_q_len < 1 || num_tokens < 1) return 0;
    size_t off[12];
    return prefill_carve(dec->cfg, batch, max_q_len, num_tokens, off);
}

// paging (block_table != NULL): k_cache / v_cache are page pools [L, num_pages, Hkv, 64, d]
static int prefill_impl(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, const int *input_len, const int *history_len,
                        const int *context_len, int batch, int max_q_len, int num_tokens, void *scratch, size_t scratch_bytes,
                        int layer_begin, int layer_end, b200_stream_t stream, const int *block_table, int max_pages, int num_pages,
                        b200_allreduce_fn reduce = nullptr, void *reduce_user = nullptr) {
    B200_REQUIRE(dec, "decoder_prefill: null handle");
    const b200_decoder_config_t &c = dec->cfg;
    B200_REQUIRE(c.tp_world <= 1 || reduce, "decoder_prefill: a tensor-parallel engine prefills through b200_decoder_prefill_tp (all-reduce callback)");
    B200_REQUIRE(hidden && k_cache && v_cache && input_len && history_len && context_len && scratch, "decoder_prefill: null pointer");
    B200_REQUIRE(batch >= 1 && batch <= c.max_batch && max_q_len >= 1 && num_tokens >= 1 && num_tokens <= batch * max_q_len,
                 "decoder_prefill: bad shape (batch %d, max_q_len %d, num_tokens %d)", batch, max_q_len, num_tokens);
    if (block_table) {
        B200_REQUIRE(c.head_size == 128 && c.dtype != B200_F32, "decoder_prefill_paged: head size 128 and a 16-bit dtype (the tensor-core context attention)");
        B200_REQUIRE(max_q_len <= max_pages * kAttnPageSize, "decoder_prefill_paged: max_q_len %d exceeds the block table's reach", max_q_len);
    } else {
        B200_REQUIRE(batch == c.max_batch, "decoder_prefill: batch %d must equal the cache's batch dimension (max_batch %d)", batch, c.max_batch);
    }
    B200_REQUIRE(max_q_len <= c.max_seq_len, "decoder_prefill: max_q_len %d exceeds the cache length %d", max_q_len, c.max_seq_len);
    B200_REQUIRE(layer_begin >= 0 && layer_end <= c.num_layers && layer_begin < layer_end, "decoder_prefill: bad layer range");
    B200_REQUIRE(((uintptr_t)scratch & 255) == 0, "decoder_prefill: scratch must be 256-byte aligned");
    size_t off[12];
    const size_t need = prefill_carve(c, batch, max_q_len, num_tokens, off);
    B200_REQUIRE(scratch_bytes >= need, "decoder_prefill: need %zu bytes of scratch, got %zu", need, scratch_bytes);
    char *base = (char *)scratch;
    void *res = base + off[0], *xn = base + off[1], *qkv = base + off[2], *qp = base + off[3], *kp = base + off[4], *vp = base + off[5];
    void *attn = base + off[6], *y = base + off[7], *gu = base + off[8], *act = base + off[10];
    int *padding_offset = (int *)(base + off[9]), *cum = padding_offset + (size_t)batch * max_q_len;
    cudaStream_t st = as_stream(stream);
    const int T = num_tokens, h = c.hidden, qkv_n = (c.head_num + 2 * c.kv_head_num) * c.head_size, qh = c.head_num * c.head_size;
    const float scale = 1.0f / sqrtf((float)c.head_size);
    int rc = b200_cal_padding_offset(padding_offset, cum, input_len, batch, max_q_len, stream);
    if (rc != B200_OK) return rc;
    void *wdq = base + off[11];
    auto linear = [&](const void *x, const b200_linear_weight_t &w, void *out, int K, int N) {
        // FP8 / INT4 at prefill sizes: dequantise the linear into scratch (one vectorised pass), then the tcgen05 GEMM -- the tensor-core
        // path north_star asks for; round 1 fell to a SIMT kernel here
        if (c.w_format != B200_W_DENSE && c.dtype != B200_F32 && T > 128) {
            int r = launch_dequant_vec(w.w, w.scales, w.zeros, wdq, N, K, c.w_format, c.group, c.dtype, st);
            if (r == B200_OK) r = launch_gemm_tc(x, wdq, out, T, N, K, c.dtype, st);
            if (r != B200_ERR_UNSUPPORTED) return r;
        }
        return b200_linear(x, w.w, w.scales, w.zeros, out, T, K, N, c.dtype, c.w_format, B200_LAYOUT_NK, c.group, stream);
    };
    NvtxRange range(block_table ? "b200 prefill (paged)" : "b200 prefill");
    const void *pending = nullptr;  // output of the previous layer's FFN, folded into the residual stream by the next norm
    for (int l = layer_begin; l < layer_end; ++l) {
        NvtxRange layer_range("layer");
        B200_REQUIRE(dec->layer_set[l], "decoder_prefill: layer %d not set", l);
        const b200_layer_weights_t &w = dec->layers[l];
        // residual <- hidden (+ pending); xn = RMSNorm(residual)
        rc = launch_norm_any(c.dtype, pending ? pending : hidden, xn, pending ? res : nullptr, res, nullptr, w.attn_norm_gamma, c.rmsnorm_eps, T, h, st);
        if (rc != B200_OK) return rc;
        if ((rc = linear(xn, w.qkv, qkv, h, qkv_n)) != B200_OK) return rc;
        // split + RoPE + KV append: one fused pass (k / v go straight into the cache); un-vectorisable shapes take the two launchers
        const size_t layer_off = block_table ? (size_t)l * num_pages * c.kv_head_num * kAttnPageSize * c.head_size * esize(c.dtype)
                                             : (size_t)l * c.max_batch * c.kv_head_num * c.max_seq_len * c.head_size * esize(c.dtype);
        rc = launch_prefill_qkv_rope_cache(qp, (char *)k_cache + layer_off, (char *)v_cache + layer_off, qkv, w.qkv_bias, padding_offset, history_len,
                                           max_q_len, T, c.head_num, c.kv_head_num, c.head_size, c.max_seq_len, c.rotary_dim, c.rotary_base, c.dtype, st,
                                           block_table, max_pages);
        if (rc == B200_ERR_UNSUPPORTED && !block_table) {
            // un-vectorisable head sizes take the reference's two launchers, whose prefill kernel has no bias term
            B200_REQUIRE(!w.qkv_bias, "decoder_prefill: qkv bias needs a head size the fused prefill kernel supports (multiple of %d)",
                         c.dtype == B200_F32 ? 8 : 16);
            rc = b200_qkv_bias_transpose_rope(qp, kp, vp, qkv, w.qkv_bias, padding_offset, history_len, input_len, batch, max_q_len, T, c.head_num,
                                              c.kv_head_num, c.head_size, c.rotary_dim, c.rotary_base, c.dtype, stream);
            if (rc != B200_OK) return rc;
            rc = b200_concat_kv_cache(kp, vp, k_cache, v_cache, input_len, history_len, l, batch, c.kv_head_num, max_q_len, c.max_seq_len,
                                      c.head_size, c.dtype, stream);
        }
        if (rc != B200_OK) return rc;
        rc = block_table ? b200_context_attention_paged(qp, k_cache, v_cache, attn, block_table, input_len, context_len, l, batch, c.head_num,
                                                        c.kv_head_num, max_q_len, num_pages, max_pages, c.head_size, scale, c.dtype, stream)
                         : b200_context_attention(qp, k_cache, v_cache, attn, padding_offset, input_len, context_len, l, batch, c.head_num,
                                                  c.kv_head_num, max_q_len, c.max_seq_len, c.head_size, T, scale, c.dtype, stream);
        if (rc != B200_OK) return rc;
        if ((rc = linear(attn, w.o, y, qh, h)) != B200_OK) return rc;
        // tensor parallel: y is this rank's partial sum of the row-sharded O projection -> the caller's all-reduce (one per block)
        if (reduce && reduce(y, (size_t)T * h, c.dtype, reduce_user, stream) != 0) {
            set_error("decoder_prefill_tp: the all-reduce callback failed (layer %d, attention block)", l);
            return B200_ERR_CUDA;
        }
        // residual += attention output; (+ o bias); xn = RMSNorm
        rc = launch_norm_any(c.dtype, y, xn, res, res, w.o_bias, w.ffn_norm_gamma, c.rmsnorm_eps, T, h, st);
        if (rc != B200_OK) return rc;
        // gate_up + SwiGLU: one tensor-core GEMM whose epilogue applies the activation (16-bit, T > 128; quantised weights: dequantised
        // into scratch first); otherwise the two launchers
        rc = B200_ERR_UNSUPPORTED;
        if (c.dtype != B200_F32 && T > 128) {
            const void *wgu = w.gate_up.w;
            if (c.w_format != B200_W_DENSE) {
                rc = launch_dequant_vec(w.gate_up.w, w.gate_up.scales, w.gate_up.zeros, wdq, 2 * c.inter_size, h, c.w_format, c.group, c.dtype, st);
                if (rc != B200_OK && rc != B200_ERR_UNSUPPORTED) return rc;
                wgu = rc == B200_OK ? wdq : nullptr;
                rc = B200_ERR_UNSUPPORTED;
            }
            if (wgu) rc = launch_gemm_tc_swiglu(xn, wgu, act, T, c.inter_size, h, c.dtype, st);
        }
        if (rc == B200_ERR_UNSUPPORTED) {
            if ((rc = linear(xn, w.gate_up, gu, h, 2 * c.inter_size)) != B200_OK) return rc;
            rc = b200_silu_and_mul(gu, act, T, c.inter_size, c.dtype, stream);
        }
        if (rc != B200_OK) return rc;
        if ((rc = linear(act, w.down, y, c.inter_size, h)) != B200_OK) return rc;
        if (reduce && reduce(y, (size_t)T * h, c.dtype, reduce_user, stream) != 0) {
            set_error("decoder_prefill_tp: the all-reduce callback failed (layer %d, FFN block)", l);
            return B200_ERR_CUDA;
        }
        pending = y;
    }
    // hidden <- residual + last FFN output
    if (cudaMemcpyAsync(hidden, res, (size_t)T * h * esize(c.dtype), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return cuda_status("decoder_prefill copy");
    return b200_add_residual(pending, hidden, T, h, c.dtype, stream);
}

int b200_decoder_prefill(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, const int *input_len, const int *history_len,
                         const int *context_len, int batch, int max_q_len, int num_tokens, void *scratch, size_t scratch_bytes,
                         int layer_begin, int layer_end, b200_stream_t stream) {
    return prefill_impl(dec, hidden, k_cache, v_cache, input_len, history_len, context_len, batch, max_q_len, num_tokens, scratch, scratch_bytes,
                        layer_begin, layer_end, stream, nullptr, 0, 0);
}

// Tensor-parallel prefill: the same pass on this rank's shard (column-sharded QKV / gate_up, row-sharded O / down, head-sharded cache);
// `reduce` all-reduces (sum) the partial [num_tokens, hidden] tensor in place on `stream` after the O projection and after the down
// projection of every layer -- north_star's one all-reduce per attention and per MLP block; the library itself never links NCCL.
int b200_decoder_prefill_tp(b200_decoder_t *dec, void *hidden, void *k_cache, void *v_cache, const int *input_len, const int *history_len,
                            const int *context_len, int batch, int max_q_len, int num_tokens, void *scratch, size_t scratch_bytes,
                            int layer_begin, int layer_end, b200_allreduce_fn reduce, void *user, b200_stream_t stream) {
    B200_REQUIRE(reduce, "decoder_prefill_tp: null all-reduce callback");
    return prefill_impl(dec, hidden, k_cache, v_cache, input_len, history_len, context_len, batch, max_q_len, num_tokens, scratch, scratch_bytes,
                        layer_begin, layer_end, stream, nullptr, 0, 0, reduce, user);
}

// The same pass with the K / V rows written into, and read back from, a PAGE POOL [L, num_pages, Hkv, 64, d] through block_table
// [batch, max_pages_per_seq] (device): the prompt's keys [64 i, 64 i + 64) of row b go to page block_table[b * max_pages_per_seq + i].
// batch <= max_batch (the pool has no batch dimension).  16-bit dtypes, head size 128.
int b200_decoder_prefill_paged(b200_decoder_t *dec, void *hidden, void *k_pool, void *v_pool, const int *block_table, const int *input_len,
                               const int *history_len, const int *context_len, int batch, int max_q_len, int num_tokens, int num_pages,
                               int max_pages_per_seq, void *scratch, size_t scratch_bytes, int layer_begin, int layer_end, b200_stream_t stream) {
    B200_REQUIRE(block_table, "decoder_prefill_paged: null block table");
    B200_REQUIRE(num_pages >= 1 && max_pages_per_seq >= 1, "decoder_prefill_paged: bad pool (num_pages %d, max_pages_per_seq %d)", num_pages, max_pages_per_seq);
    return prefill_impl(dec, hidden, k_pool, v_pool, input_len, history_len, context_len, batch, max_q_len, num_tokens, scratch, scratch_bytes,
                        layer_begin, layer_end, stream, block_table, max_pages_per_seq, num_pages);
}

int b200_lm_head_topk_sample(b200_decoder_t *dec, const void *hidden, const void *final_gamma, const void *lm_head, int vocab,
                             float *logits, int *tmp_ids, float *tmp_vals, int *topk_ids, float *topk_vals, int *seq_len,
                             uint8_t *finished, int *output_id, int batch, int k, int step, int end_id, b200_stream_t stream) {
    int rc = check_ready(dec, batch);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(hidden && final_gamma && lm_head && logits, "lm_head_topk_sample: null pointer");
    B200_REQUIRE(vocab > 0, "lm_head_topk_sample: bad vocab");
    NvtxRange range("b200 lm head + top-k + sampling");
    const b200_decoder_config_t &c = dec->cfg;
    cudaStream_t st = as_stream(stream);
    // final RMSNorm (reference llama.cpp:247-253) fused into the LM-head GEMV; logits in fp32
    const bool b16 = c.dtype != B200_F32;
    const int pass = b16 ? 16 : 4;  // tokens per read of the LM head (tensor-core GEMV / SIMT GEMV)
    const void *xh = hidden;
    bool normed = false;
    if (b16 && batch >= 2) {  // final RMSNorm once, in front (see norm_linear)
        rc = launch_norm_tp(c.dtype, hidden, dec->xn, nullptr, nullptr, nullptr, final_gamma, c.rmsnorm_eps, batch, c.hidden, nullptr, st);
        if (rc != B200_OK) return rc;
        xh = dec->xn, normed = true;
    }
    for (int m0 = 0; m0 < batch; m0 += pass) {
        GemvArgs a = {};
        a.w = lm_head;
        a.x = (const char *)xh + (size_t)m0 * c.hidden * esize(c.dtype);
        a.y = logits + (size_t)m0 * vocab;
        a.y_f32 = 1;
        if (!normed) a.gamma = final_gamma, a.eps = c.rmsnorm_eps, a.norm = 1;
        a.M = batch - m0 < pass ? batch - m0 : pass, a.K = c.hidden, a.N = vocab;
        rc = launch_gemv_nk(a, c.dtype, WF_DENSE, false, st);
        if (rc == B200_ERR_UNSUPPORTED && pass > 4) {  // shapes the tensor-core GEMV cannot split: the SIMT GEMV, 4 tokens at a time
            for (int m1 = m0; m1 < m0 + a.M; m1 += 4) {
                GemvArgs b = a;
                b.x = (const char *)xh + (size_t)m1 * c.hidden * esize(c.dtype);
                b.y = logits + (size_t)m1 * vocab;
                b.M = m0 + a.M - m1 < 4 ? m0 + a.M - m1 : 4;
                if ((rc = launch_gemv_nk(b, c.dtype, WF_DENSE, false, st)) != B200_OK) break;
            }
        }
        if (rc == B200_ERR_UNSUPPORTED) set_error("lm_head_topk_sample: hidden size %d not supported by the GEMV", c.hidden);
        if (rc != B200_OK) return rc;
    }
    if (!topk_ids) return B200_OK;  // logits only
    B200_REQUIRE(tmp_ids && tmp_vals && topk_vals, "lm_head_topk_sample: null top-k buffer");
    rc = b200_topk(logits, tmp_ids, tmp_vals, topk_ids, topk_vals, batch, vocab, k, B200_F32, stream);
    if (rc != B200_OK || !output_id) return rc;
    B200_REQUIRE(seq_len && finished, "lm_head_topk_sample: null sampling buffer");
    return b200_sampling(topk_ids, topk_vals, seq_len, finished, output_id, batch, k, step, end_id, vocab, B200_F32, stream);
}

}  // extern "C"
