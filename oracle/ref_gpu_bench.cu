// ref_gpu_bench.cu -- times the reference's OWN CUDA decode path (LlamaSelfDecoder<float>::forward, src/layers/self_decoder.cpp:24-122, its
// kernels from src/kernels/*.cu, cuBLAS SGEMM for the linears) on the GPU it is run on: SURVEY.md 8(d) "GPU reference baseline (bonus)".
// TEST / MEASUREMENT INFRASTRUCTURE, never linked into libb200llm.so.  Built by `make -C oracle ref_gpu_bench` from the reference sources
// where they lie (needs /root/reference: the build container), linked against oracle/_ref/libref_layers.so + libref.so; the binary travels
// to the GPU box with the snapshot.
//
// The domain is the one on which the reference's decode kernels are valid: fp32, batch 1, step <= 128, H == Hkv.
// All `layers` entries of the weight vector point at ONE LlamaLayerWeight (809 MB of fp32 at the 7B shape, far above the 126 MB L2, so
// every layer still streams its weights from HBM): 32 distinct fp32 layers would only add 25 GB of allocation and fill time.
// The reference's forward() deletes the Tensor objects of the maps it is given when its local TensorMaps go out of scope,
// so fresh wrapper objects are built for every call and never touched again; device buffers are owned here and never freed.
//
// usage: ref_gpu_bench [layers=32] [step=128] [iters=10] [warmup=2] [head_num=32] [head_size=128] [inter=11008] [max_seq=256]
// prints one JSON line: {"impl": "reference-cuda", "ms_per_step": ..., "tokens_per_s": ..., ...}
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "src/layers/includes/self_decoder.h"
#include "src/utils/macro.h"

static float *device_fill(size_t n, float scale, uint64_t seed) {
    std::vector<float> h(n);
    uint64_t s = seed * 6364136223846793005ull + 1442695040888963407ull;
    for (size_t i = 0; i < n; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = scale * ((float)((s >> 40) & 0xffff) / 32768.0f - 1.0f);  // uniform in [-scale, scale)
    }
    float *d = nullptr;
    if (cudaMalloc(&d, n * sizeof(float)) != cudaSuccess) {
        fprintf(stderr, "cudaMalloc of %zu floats failed\n", n);
        _exit(2);
    }
    cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    return d;
}
static void device_refill(float *d, size_t n, float scale, uint64_t seed) {
    float *t = device_fill(n, scale, seed);
    cudaMemcpy(d, t, n * sizeof(float), cudaMemcpyDeviceToDevice);
    cudaFree(t);
}

int main(int argc, char **argv) {
    auto arg = [&](int i, int dflt) { return argc > i ? atoi(argv[i]) : dflt; };
    const int layers = arg(1, 32), step_arg = arg(2, 128), iters = arg(3, 10), warmup = arg(4, 2);
    const int head_num = arg(5, 32), head_size = arg(6, 128), inter = arg(7, 11008), max_seq = arg(8, 256);
    const int kv_head_num = head_num, hidden = head_num * head_size, batch = 1;
    const int qkv_cols = (head_num + 2 * kv_head_num) * head_size;

    // one layer's weights, shapes and flags as the reference's own dummy loader sets them (src/weights/layer_weights.cpp:83-152)
    auto *lw = new LlamaLayerWeight<float>(head_num, kv_head_num, head_size, inter, getWeightType<float>(), true);
    device_refill(lw->attention_norm_weight.gamma, hidden, 1.0f, 1);
    device_refill(lw->ffn_norm_weight.gamma, hidden, 1.0f, 2);
    device_refill(lw->self_attention_weight.qkv.data, (size_t)hidden * qkv_cols, 0.02f, 3);
    device_refill(lw->self_attention_weight.output.data, (size_t)hidden * hidden, 0.02f, 4);
    device_refill(lw->ffn_weight.gate_and_up.data, (size_t)hidden * 2 * inter, 0.02f, 5);
    device_refill(lw->ffn_weight.down.data, (size_t)hidden * inter, 0.02f, 6);
    cudaMemset(lw->self_attention_weight.qkv.bias, 0, sizeof(float) * qkv_cols);
    cudaMemset(lw->self_attention_weight.output.bias, 0, sizeof(float) * hidden);
    lw->ffn_weight.down.bias = device_fill(hidden, 0.0f, 7);
    lw->self_attention_weight.qkv.is_transposed = true;
    lw->self_attention_weight.output.is_transposed = false;
    lw->ffn_weight.gate_and_up.is_transposed = true;
    lw->ffn_weight.down.is_transposed = true;
    auto *layer_weights = new std::vector<LlamaLayerWeight<float> *>((size_t)layers, lw);

    const size_t cache_elems = (size_t)layers * batch * kv_head_num * max_seq * head_size;
    float *d_in = device_fill((size_t)batch * hidden, 1.0f, 8), *d_out = device_fill((size_t)batch * hidden, 0.0f, 9);
    float *d_k = device_fill(cache_elems, 0.5f, 10), *d_v = device_fill(cache_elems, 0.5f, 11);
    float *d_gamma = device_fill(hidden, 1.0f, 12);
    bool *d_finished = nullptr;
    cudaMalloc(&d_finished, batch);
    cudaMemset(d_finished, 0, batch);

    LlamaAttentionStaticParams sp{};
    sp.rotary_embedding_dim = head_size, sp.rotary_embedding_base = 10000, sp.max_position_embeddings = 4096, sp.use_dynamic_ntk = false;
    LlamaAttentionDynamicParams dp{};
    dp.batch_size = batch;
    cublasHandle_t cublas;
    cublasLtHandle_t cublaslt = nullptr;
    cublasCreate(&cublas);
    cublasSetMathMode(cublas, CUBLAS_DEFAULT_MATH);
    auto *wrapper = new CublasWrapper(cublas, cublaslt);
    wrapper->setFP32GemmConfig();
    auto *allocator = new CudaAllocator();
    cudaStream_t stream = nullptr;
    auto *decoder = new LlamaSelfDecoder<float>(head_num, kv_head_num, head_size, inter, layers, sp, 1e-6f, stream, wrapper, allocator);

    int *h_step = new int(step_arg), *h_layer = new int(0);
    const DataType f = getTensorType<float>(), i32 = getTensorType<int>(), b8 = getTensorType<bool>();
    auto forward_once = [&]() {
        // heap maps and wrappers, abandoned after the call (see the header comment)
        auto *in = new TensorMap{{"decoder_input", new TensorWrapper<float>(Device::GPU, f, {batch, hidden}, d_in)},
                                 {"step", new TensorWrapper<int>(Device::CPU, i32, {1}, h_step)},
                                 {"finished", new TensorWrapper<bool>(Device::GPU, b8, {batch}, d_finished)},
                                 {"layer_id", new TensorWrapper<int>(Device::CPU, i32, {1}, h_layer)},
                                 {"output_norm_weight", new TensorWrapper<float>(Device::GPU, f, {hidden}, d_gamma)}};
        auto *out = new TensorMap{{"decoder_output", new TensorWrapper<float>(Device::GPU, f, {batch, hidden}, d_out)},
                                  {"all_k_cache", new TensorWrapper<float>(Device::GPU, f, {layers, batch, kv_head_num, max_seq, head_size}, d_k)},
                                  {"all_v_cache", new TensorWrapper<float>(Device::GPU, f, {layers, batch, kv_head_num, max_seq, head_size}, d_v)}};
        decoder->forward(in, layer_weights, out, &dp);
    };

    // The same work without the layer classes: the reference's launch functions called in LlamaSelfDecoder::forward's order
    // (self_decoder.cpp:69-119, self_attention.cpp:79-139, ffn.cpp:105-140) on buffers allocated once.  `syncs`: keep the
    // cudaDeviceSynchronize the reference does after every launcher (DeviceSyncAndCheckCudaError) -- its actual behaviour -- or not
    // (its kernels + cuBLAS at their best).  Used when forward() itself does not survive on this GPU / driver (recorded in the JSON).
    float *d_res = device_fill((size_t)batch * hidden, 0.0f, 13), *d_qkv = device_fill((size_t)batch * qkv_cols, 0.0f, 14);
    float *d_mha = device_fill((size_t)batch * hidden, 0.0f, 15), *d_gu = device_fill((size_t)batch * 2 * inter, 0.0f, 16);
    float *d_act = device_fill((size_t)batch * inter, 0.0f, 17);
    auto launchers_once = [&](bool syncs) {
        auto sync = [&]() { if (syncs) cudaDeviceSynchronize(); };
        auto *x = new TensorWrapper<float>(Device::GPU, f, {batch, hidden}, d_out);
        auto *res = new TensorWrapper<float>(Device::GPU, f, {batch, hidden}, d_res);
        auto *qkv = new TensorWrapper<float>(Device::GPU, f, {batch, head_num + 2 * kv_head_num, head_size}, d_qkv);
        auto *mha = new TensorWrapper<float>(Device::GPU, f, {batch, hidden}, d_mha);
        auto *gu = new TensorWrapper<float>(Device::GPU, f, {batch, 2, inter}, d_gu);
        auto *act = new TensorWrapper<float>(Device::GPU, f, {batch, inter}, d_act);
        auto *kc = new TensorWrapper<float>(Device::GPU, f, {layers, batch, kv_head_num, max_seq, head_size}, d_k);
        auto *vc = new TensorWrapper<float>(Device::GPU, f, {layers, batch, kv_head_num, max_seq, head_size}, d_v);
        auto *fin = new TensorWrapper<bool>(Device::GPU, b8, {batch}, d_finished);
        auto *stp = new TensorWrapper<int>(Device::CPU, i32, {1}, h_step);
        cudaMemcpyAsync(d_out, d_in, sizeof(float) * batch * hidden, cudaMemcpyDeviceToDevice, 0);
        for (int l = 0; l < layers; ++l) {
            int *h_l = new int(l);
            auto *lid = new TensorWrapper<int>(Device::CPU, i32, {1}, h_l);
            LlamaLayerWeight<float> *w = layer_weights->at(l);
            launchRMSNorm(x, res, &w->attention_norm_weight, 1e-6f);
            sync();
            launchLinearGemm(x, &w->self_attention_weight.qkv, qkv, wrapper, false, w->self_attention_weight.qkv.is_transposed);
            sync();
            launchRope(qkv, stp, &sp);
            sync();
            launchDecoderMaskedMultiHeadAttention<float>(qkv, &w->self_attention_weight.qkv, lid, kc, vc, fin, stp, mha, &sp);
            sync();
            launchLinearGemm(mha, &w->self_attention_weight.output, x, wrapper, false, w->self_attention_weight.output.is_transposed);
            sync();
            launchFusedAddBiasResidualAndRMSNorm(res, x, &w->self_attention_weight.output, w->ffn_norm_weight.gamma, 1e-6f);
            sync();
            launchLinearGemm(x, &w->ffn_weight.gate_and_up, gu, wrapper, false, w->ffn_weight.gate_and_up.is_transposed);
            sync();
            launchSiluAndMul(gu, act);
            sync();
            launchLinearGemm(act, &w->ffn_weight.down, x, wrapper, false, w->ffn_weight.down.is_transposed);
            sync();
            launchAddResidual(res, x, false);
            sync();
        }
    };

    // the reference prints from inside its launchers: keep stdout for the JSON line only
    fflush(stdout);
    const int saved = dup(1);
    FILE *devnull = fopen("/dev/null", "w");
    dup2(fileno(devnull), 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    std::string forward_error;
    float ms = -1.f;
    try {
        for (int i = 0; i < warmup; ++i) forward_once();
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters; ++i) forward_once();
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    } catch (const std::exception &ex) {
        forward_error = ex.what();
        ms = -1.f;
        cudaDeviceSynchronize();
        cudaGetLastError();
    }
    float ms_l[2] = {-1.f, -1.f};  // launcher sequence: [0] with the reference's per-launcher device syncs, [1] without
    std::string launcher_error;
    for (int v = 0; v < 2; ++v) {
        try {
            for (int i = 0; i < warmup; ++i) launchers_once(v == 0);
            cudaDeviceSynchronize();
            cudaEventRecord(e0, 0);
            for (int i = 0; i < iters; ++i) launchers_once(v == 0);
            cudaEventRecord(e1, 0);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_l[v], e0, e1);
        } catch (const std::exception &ex) {
            launcher_error = ex.what();
            cudaDeviceSynchronize();
            cudaGetLastError();
        }
    }
    const cudaError_t err = cudaGetLastError();
    fflush(stdout);
    dup2(saved, 1);
    auto clean = [](std::string t) {
        for (char &ch : t)
            if (ch == '"' || ch == '\\' || ch == '\n') ch = ' ';
        return t;
    };

    const double weight_bytes = (double)layers * 4.0 * ((double)hidden * qkv_cols + (double)hidden * hidden + 3.0 * hidden * (double)inter);
    auto per = [&](float t) { return t > 0 ? (double)t / iters : -1.0; };
    auto tps = [&](float t) { return t > 0 ? 1000.0 / ((double)t / iters) : 0.0; };
    auto gbs = [&](float t) { return t > 0 ? weight_bytes / ((double)t / iters * 1e-3) / 1e9 : 0.0; };
    printf("{\"impl\": \"reference-cuda\", \"what\": \"the reference's kernels + cuBLAS SGEMM, fp32, batch 1\", "
           "\"layers\": %d, \"step\": %d, \"hidden\": %d, \"inter\": %d, \"iters\": %d, \"warmup\": %d, "
           "\"forward\": {\"what\": \"LlamaSelfDecoder<float>::forward\", \"ms_per_step\": %.4f, \"tokens_per_s\": %.3f, \"weight_gb_per_s\": %.1f, \"error\": \"%s\"}, "
           "\"launchers_with_syncs\": {\"what\": \"the reference's launch functions in forward()'s order, device sync after each as the reference does\", "
           "\"ms_per_step\": %.4f, \"tokens_per_s\": %.3f, \"weight_gb_per_s\": %.1f}, "
           "\"launchers_no_syncs\": {\"what\": \"same without the syncs\", \"ms_per_step\": %.4f, \"tokens_per_s\": %.3f, \"weight_gb_per_s\": %.1f}, "
           "\"launcher_error\": \"%s\", \"weight_bytes_per_step\": %.0f, \"cuda_status\": \"%s\", "
           "\"note\": \"decoder layers only (no LM head / sampling: dead code in the reference); one layer's weights shared by all layers\"}\n",
           layers, step_arg, hidden, inter, iters, warmup, per(ms), tps(ms), gbs(ms), clean(forward_error).c_str(), per(ms_l[0]), tps(ms_l[0]),
           gbs(ms_l[0]), per(ms_l[1]), tps(ms_l[1]), gbs(ms_l[1]), clean(launcher_error).c_str(), weight_bytes, cudaGetErrorString(err));
    fflush(stdout);
    _exit((ms_l[0] > 0 || ms > 0) ? 0 : 1);  // no destructors: the reference's wrappers would free buffers they do not own
}
