// ref_gpu_shim.cu -- TEST INFRASTRUCTURE.  C entry points onto the UNMODIFIED reference kernels.
//
// Compiled by oracle/Makefile together with the reference's own src/kernels/*.cu (from where they lie under
// $(REF), nothing is copied) into oracle/_ref/libref.so.  tests/ uses it on the GPU box to compare this repo's
// CUDA kernels against the reference's fp32 CUDA kernels where those are not defective.
// All pointers are device pointers; fp32 only (the reference's fp16 paths are stubs or broken, D1).
// The reference's TensorWrapper destructor frees the data pointer (D10), so wrappers are leaked on purpose
// via View<>: the tensors belong to the caller.
#include <vector>
#include "src/kernels/includes/rmsnorm.cuh"
#include "src/kernels/includes/add_residual.cuh"
#include "src/kernels/includes/add_residual_and_rmsnorm.cuh"
#include "src/kernels/includes/linear.cuh"
#include "src/kernels/includes/rope.cuh"
#include "src/kernels/includes/decoder_self_attention.cuh"
#include "src/kernels/includes/qkv_bias_and_rope.cuh"
#include "src/kernels/includes/concat_past_kv.cuh"
#include "src/kernels/includes/scale_and_mask_and_softmax.cuh"
#include "src/kernels/includes/build_causal_mask.cuh"
#include "src/kernels/includes/cal_padding_offset.cuh"
#include "src/kernels/includes/transpose_and_remove_padding.cuh"
#include "src/kernels/includes/silu_and_mul.cuh"
#include "src/kernels/includes/input_embedding.cuh"
#include "src/kernels/includes/topk.cuh"
#include "src/kernels/includes/sampling.cuh"

namespace {
template <typename T> TensorWrapper<T> *view(Device dev, std::vector<int> shape, T *data) {
    return new TensorWrapper<T>(dev, getTensorType<T>(), shape, data);  // leaked: see header comment
}
template <typename T> TensorWrapper<T> *gpu(std::vector<int> shape, const T *data) {
    return view<T>(Device::GPU, shape, const_cast<T *>(data));
}
TensorWrapper<int> *host_int(int v) { return view<int>(Device::CPU, {1}, new int(v)); }
int sync_status() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();  // launch-configuration failures do not show up in the sync
    return (int)e;
}
}  // namespace

#define REF_TRY(...)                        \
    try {                                   \
        __VA_ARGS__;                        \
    } catch (const std::exception &e) {     \
        fprintf(stderr, "%s\n", e.what());  \
        return -1;                          \
    }                                       \
    return sync_status();

extern "C" {

int ref_rmsnorm(float *x, float *residual, float *gamma, float eps, int tokens, int hidden) {
    LayerNormWeight<float> w;
    w.gamma = gamma;
    REF_TRY(launchRMSNorm<float>(gpu<float>({tokens, hidden}, x), gpu<float>({tokens, hidden}, residual), &w, eps, false))
}

int ref_fused_add_bias_residual_rmsnorm(float *residual, float *out, float *bias, float *gamma, float eps, int tokens,
                                        int hidden) {
    BaseWeight<float> norm;
    norm.bias = bias;
    norm.data = nullptr;
    REF_TRY(launchFusedAddBiasResidualAndRMSNorm<float>(gpu<float>({tokens, hidden}, residual),
                                                        gpu<float>({tokens, hidden}, out), &norm, gamma, eps))
}

int ref_add_residual(float *residual, float *out, int tokens, int hidden) {
    REF_TRY(launchAddResidual<float>(gpu<float>({tokens, hidden}, residual), gpu<float>({tokens, hidden}, out), false))
}

// y[M,N] = x[M,K] * Wmem; declared weight shape {K,N}, trans_b = false (what src/layers uses).
int ref_linear(float *x, float *w, float *y, int M, int K, int N) {
    static CublasWrapper *cw = nullptr;
    if (!cw) {
        cublasHandle_t h;
        cublasLtHandle_t lt;
        if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS || cublasLtCreate(&lt) != CUBLAS_STATUS_SUCCESS) return -2;
        cw = new CublasWrapper(h, lt);
        cw->setFP32GemmConfig();
    }
    BaseWeight<float> bw;
    bw.type = WeightType::FP32_W;
    bw.shape = {K, N};
    bw.data = w;
    bw.bias = nullptr;
    bw.is_transposed = false;
    REF_TRY(launchLinearGemm<float>(gpu<float>({M, K}, x), &bw, gpu<float>({M, N}, y), cw, false, false))
}

int ref_rope_decode(float *qkv, int batch, int head_num, int kv_head_num, int head_size, int step, int rot_dim,
                    float base) {
    LlamaAttentionStaticParams p;
    p.rotary_embedding_dim = rot_dim;
    p.rotary_embedding_base = base;
    p.max_position_embeddings = 4096;
    p.use_dynamic_ntk = false;
    p.head_size = head_size;
    p.head_num = head_num;
    p.kv_head_num = kv_head_num;
    REF_TRY(launchRope<float>(gpu<float>({batch, head_num + 2 * kv_head_num, head_size}, qkv), host_int(step), &p))
}

int ref_decode_mha(float *qkv, float *bias, float *k_cache, float *v_cache, float *out, int layers, int batch,
                   int head_num, int kv_head_num, int head_size, int max_seq_len, int step, int layer) {
    LlamaAttentionStaticParams p;
    p.rotary_embedding_dim = head_size;
    p.rotary_embedding_base = 10000.0f;
    p.max_position_embeddings = 4096;
    p.use_dynamic_ntk = false;
    p.head_size = head_size;
    p.head_num = head_num;
    p.kv_head_num = kv_head_num;
    BaseWeight<float> bw;
    bw.bias = bias;
    bw.data = nullptr;
    std::vector<int> cshape = {layers, batch, kv_head_num, max_seq_len, head_size};
    bool *fin = nullptr;
    cudaMalloc(&fin, batch);
    cudaMemset(fin, 0, batch);
    REF_TRY(launchDecoderMaskedMultiHeadAttention<float>(
        gpu<float>({batch, head_num + 2 * kv_head_num, head_size}, qkv), &bw, host_int(layer), gpu<float>(cshape, k_cache),
        gpu<float>(cshape, v_cache), gpu<bool>({batch}, fin), host_int(step), gpu<float>({batch, head_num, head_size}, out), &p))
}

int ref_qkv_bias_transpose_rope(float *q, float *k, float *v, float *qkv, int *padding_offset, int *history_len,
                                int *input_len, int batch, int seq_len, int num_tokens, int head_num, int kv_head_num,
                                int head_size, int rot_dim, float base) {
    LlamaAttentionStaticParams p;
    p.rotary_embedding_dim = rot_dim;
    p.rotary_embedding_base = base;
    p.max_position_embeddings = 4096;
    p.use_dynamic_ntk = false;
    p.head_size = head_size;
    p.head_num = head_num;
    p.kv_head_num = kv_head_num;
    BaseWeight<float> bw;
    bw.bias = nullptr;
    bw.data = nullptr;
    REF_TRY(launchFusedQKVAddBiasAndTransposeAndRope<float>(
        gpu<float>({batch, head_num, seq_len, head_size}, q), gpu<float>({batch, kv_head_num, seq_len, head_size}, k),
        gpu<float>({batch, kv_head_num, seq_len, head_size}, v),
        gpu<float>({num_tokens, head_num + 2 * kv_head_num, head_size}, qkv), &bw, gpu<int>({batch, seq_len}, padding_offset),
        gpu<int>({batch}, history_len), gpu<int>({batch}, input_len), &p))
}

int ref_concat_kv_cache(float *k_src, float *v_src, float *k_cache, float *v_cache, int *cur_len, int *history_len,
                        int layers, int layer, int batch, int kv_head_num, int max_q_len, int max_seq_len, int head_size) {
    std::vector<int> cshape = {layers, batch, kv_head_num, max_seq_len, head_size};
    REF_TRY(launchConcatKVCache<float>(gpu<float>({batch, kv_head_num, max_q_len, head_size}, k_src),
                                       gpu<float>({batch, kv_head_num, max_q_len, head_size}, v_src), host_int(layer),
                                       gpu<int>({batch}, cur_len), gpu<int>({batch}, history_len), gpu<float>(cshape, k_cache),
                                       gpu<float>(cshape, v_cache)))
}

int ref_scale_mask_softmax(float *qk, float *mask, float *out, float scale, int batch, int head_num, int q_len,
                           int k_len) {
    REF_TRY(launchFusedScaleMaskAndSoftmax<float>(gpu<float>({batch, head_num, q_len, k_len}, qk),
                                                  gpu<float>({batch, q_len, k_len}, mask),
                                                  gpu<float>({batch, head_num, q_len, k_len}, out), scale))
}

int ref_build_causal_masks(float *mask, int *q_lens, int *k_lens, int batch, int max_q_len, int max_k_len) {
    REF_TRY(launchBuildCausalMasks<float>(gpu<float>({batch, max_q_len, max_k_len}, mask), gpu<int>({batch}, q_lens),
                                          gpu<int>({batch}, k_lens)))
}

int ref_cal_padding_offset(int *padding_offset, int *cum_seqlens, int *input_lengths, int batch, int max_q_len) {
    REF_TRY(launchCalPaddingOffset(gpu<int>({batch, max_q_len}, padding_offset), gpu<int>({batch + 1}, cum_seqlens),
                                   gpu<int>({batch}, input_lengths)))
}

int ref_transpose_remove_padding(float *src, int *padding_offset, float *dst, int num_tokens, int batch, int seq_len,
                                 int head_num, int head_size) {
    REF_TRY(launchFusedTransposeAndRemovePadding<float>(gpu<float>({batch, head_num, seq_len, head_size}, src),
                                                        gpu<int>({num_tokens}, padding_offset),
                                                        gpu<float>({num_tokens, head_num, head_size}, dst)))
}

int ref_silu_and_mul(float *in, float *out, int tokens, int inter) {
    REF_TRY(launchSiluAndMul<float>(gpu<float>({tokens, 2, inter}, in), gpu<float>({tokens, inter}, out)))
}

int ref_input_embedding(int *ids, float *table, float *out, int tokens, int hidden, int vocab) {
    EmbeddingWeight<float> w;
    w.shape = {vocab, hidden};
    w.data = table;
    w.bias = nullptr;
    REF_TRY(launchInputEmbedding<float>(gpu<int>({tokens}, ids), gpu<float>({tokens, hidden}, out), &w))
}

// K = 5, 8 blocks per row are hard-coded by the reference (src/kernels/topk.cu:116-118).
int ref_topk(float *probs, int *tmp_ids, float *tmp_vals, int *final_ids, float *final_vals, int rows, int vocab) {
    REF_TRY(launchTopKForBeamSearch<float>(gpu<float>({rows, vocab}, probs), gpu<int>({rows, 8, 5}, tmp_ids),
                                           gpu<float>({rows, 8, 5}, tmp_vals), gpu<int>({rows, 5}, final_ids),
                                           gpu<float>({rows, 5}, final_vals)))
}

int ref_sampling(int *topk_id, float *topk_val, int *seq_len, bool *finished, int *output_id, int batch, int k,
                 int step, int end_id, int vocab) {
    MapStringToInt params{{"vocab_size", vocab}, {"step", step}, {"end_id", end_id}};
    REF_TRY(launchSampling<float>(gpu<int>({batch, k}, topk_id), gpu<float>({batch, k}, topk_val), gpu<int>({batch}, seq_len),
                                  gpu<bool>({batch}, finished), gpu<int>({batch}, output_id), &params))
}

}  // extern "C"
