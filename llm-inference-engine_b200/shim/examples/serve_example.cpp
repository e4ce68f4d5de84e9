// Continuous batching over a paged KV cache, driven from C++ through the C ABI alone (include/b200llm.h) -- the host-side loop a serving
// front end would run.  Builds a small random bf16 Llama-shaped model, then
//   (1) generates 4 prompts of different lengths with b200_generate_ragged (static batch, contiguous cache),
//   (2) serves the same 4 requests with b200_batcher_* (paged cache; all admitted in the first iteration): ids must be IDENTICAL,
//   (3) serves a stream of 9 requests through 3 batch slots and a pool small enough to force waiting (and possibly preemption),
//       and checks that every request finishes with the number of tokens it asked for and that every page returns to the pool.
// usage: serve_example            exit code 0 = passed
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "b200llm.h"

#define CHECK_B200(x)                                                                      \
    do {                                                                                   \
        const int rc_ = (x);                                                               \
        if (rc_ != B200_OK) {                                                              \
            fprintf(stderr, "%s:%d: status %d: %s\n", __FILE__, __LINE__, rc_, b200_last_error_string()); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)
#define CHECK_CUDA(x)                                                                      \
    do {                                                                                   \
        const cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

static uint64_t g_state = 88172645463325252ull;
static float rnd() {  // xorshift: uniform in (-1, 1)
    g_state ^= g_state << 13, g_state ^= g_state >> 7, g_state ^= g_state << 17;
    return (float)((double)(g_state >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}
static void *upload_bf16(size_t n, float scale, float offset = 0.0f) {
    std::vector<__nv_bfloat16> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(offset + scale * rnd());
    void *d = nullptr;
    CHECK_CUDA(cudaMalloc(&d, n * sizeof(__nv_bfloat16)));
    CHECK_CUDA(cudaMemcpy(d, h.data(), n * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
    return d;
}
static void *dev_alloc(size_t bytes) {
    void *d = nullptr;
    CHECK_CUDA(cudaMalloc(&d, bytes + 256));
    CHECK_CUDA(cudaMemset(d, 0, bytes + 256));
    return d;  // cudaMalloc is 256-byte aligned
}

int main() {
    const int hidden = 256, H = 2, Hkv = 2, d = 128, inter = 384, L = 2, vocab = 1000, S = 256, max_batch = 4;
    b200_decoder_config_t cfg = {};
    cfg.hidden = hidden, cfg.head_num = H, cfg.kv_head_num = Hkv, cfg.head_size = d, cfg.inter_size = inter, cfg.num_layers = L;
    cfg.max_seq_len = S, cfg.max_batch = max_batch, cfg.dtype = B200_BF16, cfg.w_format = B200_W_DENSE, cfg.group = 128;
    cfg.rmsnorm_eps = 1e-6f, cfg.rotary_dim = d, cfg.rotary_base = 10000.0f, cfg.tp_world = 1, cfg.tp_rank = 0;
    CHECK_B200(b200_workspace_ensure(0));
    b200_decoder_t *dec = b200_decoder_create(&cfg);
    if (!dec) return fprintf(stderr, "decoder_create: %s\n", b200_last_error_string()), 1;
    const int qkv_n = (H + 2 * Hkv) * d;
    for (int l = 0; l < L; ++l) {
        b200_layer_weights_t w = {};
        w.attn_norm_gamma = upload_bf16(hidden, 0.1f, 1.0f);
        w.qkv.w = upload_bf16((size_t)qkv_n * hidden, 0.06f);
        w.o.w = upload_bf16((size_t)hidden * H * d, 0.06f);
        w.ffn_norm_gamma = upload_bf16(hidden, 0.1f, 1.0f);
        w.gate_up.w = upload_bf16((size_t)2 * inter * hidden, 0.06f);
        w.down.w = upload_bf16((size_t)hidden * inter, 0.06f);
        CHECK_B200(b200_decoder_set_layer(dec, l, &w));
    }
    const size_t scratch_bytes = b200_decoder_scratch_bytes(dec);
    CHECK_B200(b200_decoder_set_scratch(dec, dev_alloc(scratch_bytes), scratch_bytes));

    b200_generate_params_t gp = {};
    gp.embedding = upload_bf16((size_t)vocab * hidden, 1.0f);
    gp.final_gamma = upload_bf16(hidden, 0.1f, 1.0f);
    gp.lm_head = upload_bf16((size_t)vocab * hidden, 0.06f);
    gp.vocab = vocab, gp.top_k = 1, gp.end_id = -1 /* never sampled: every request runs to its token limit */, gp.check_every = 0;

    // ---------------------------------------------------------------- (1) static batch, contiguous cache
    const int lens[4] = {7, 3, 70, 1}, N = 9, max_len = 70;
    std::vector<int> prompts((size_t)4 * max_len, 0);
    for (int b = 0; b < 4; ++b)
        for (int t = 0; t < lens[b]; ++t) prompts[(size_t)b * max_len + t] = 3 + (int)((rnd() * 0.5f + 0.5f) * (vocab - 4));
    gp.max_new_tokens = N;
    const size_t cache_bytes = (size_t)L * max_batch * Hkv * S * d * sizeof(__nv_bfloat16);
    void *kc = dev_alloc(cache_bytes), *vc = dev_alloc(cache_bytes);
    const size_t gen_ws = b200_generate_workspace_bytes(dec, &gp, 4, max_len);
    if (!gen_ws) return fprintf(stderr, "generate workspace: %s\n", b200_last_error_string()), 1;
    std::vector<int> ids_static((size_t)4 * N), ngen(4);
    CHECK_B200(b200_generate_ragged(dec, &gp, prompts.data(), lens, 4, max_len, kc, vc, dev_alloc(gen_ws), gen_ws, ids_static.data(), ngen.data(), nullptr));

    // ---------------------------------------------------------------- (2) the same requests through the batcher, paged cache
    const int num_pages = 12, max_pages_per_seq = S / B200_KV_PAGE_SIZE;
    const size_t pool_bytes = (size_t)L * num_pages * Hkv * B200_KV_PAGE_SIZE * d * sizeof(__nv_bfloat16);
    void *kp = dev_alloc(pool_bytes), *vp = dev_alloc(pool_bytes);
    b200_batcher_config_t bc = {max_batch, num_pages, max_pages_per_seq, 160};
    b200_batcher_t *bat = b200_batcher_create(&bc);
    if (!bat) return fprintf(stderr, "batcher_create: %s\n", b200_last_error_string()), 1;
    const size_t bat_ws = b200_batcher_workspace_bytes(bat, dec, &gp);
    if (!bat_ws) return fprintf(stderr, "batcher workspace: %s\n", b200_last_error_string()), 1;
    void *ws = dev_alloc(bat_ws);
    int rid[4];
    for (int b = 0; b < 4; ++b) rid[b] = b200_batcher_submit(bat, &prompts[(size_t)b * max_len], lens[b], N);
    int iterations = 0, finished = 0;
    while (b200_batcher_pending(bat) > 0) {
        CHECK_B200(b200_batcher_step(bat, dec, &gp, kp, vp, ws, bat_ws, &finished, nullptr));
        ++iterations;
    }
    bool ok = iterations == N;
    for (int b = 0; b < 4; ++b) {
        int out[64], n = 0, state = 0;
        CHECK_B200(b200_batcher_result(bat, rid[b], out, 64, &n, &state));
        const bool same = n == N && state == B200_REQ_FINISHED && memcmp(out, &ids_static[(size_t)b * N], N * sizeof(int)) == 0;
        printf("request %d (prompt %2d tokens): %s\n", b, lens[b], same ? "identical to the static batch" : "DIFFERS");
        ok = ok && same;
    }
    ok = ok && b200_batcher_free_pages(bat) == num_pages;
    b200_batcher_destroy(bat);

    // ---------------------------------------------------------------- (3) a stream: 9 requests, 3 slots, 5 pages
    b200_batcher_config_t bc2 = {3, 5, max_pages_per_seq, 128};
    bat = b200_batcher_create(&bc2);
    const size_t ws2_bytes = b200_batcher_workspace_bytes(bat, dec, &gp);
    void *ws2 = dev_alloc(ws2_bytes);
    const int plen[9] = {60, 5, 62, 17, 63, 1, 40, 100, 9}, want[9] = {12, 20, 10, 6, 9, 15, 8, 5, 30};
    int rids[9];
    for (int r = 0; r < 9; ++r) {
        std::vector<int> p(plen[r]);
        for (int t = 0; t < plen[r]; ++t) p[t] = 3 + (int)((rnd() * 0.5f + 0.5f) * (vocab - 4));
        rids[r] = b200_batcher_submit(bat, p.data(), plen[r], want[r]);
        if (rids[r] < 0) return fprintf(stderr, "submit: %s\n", b200_last_error_string()), 1;
    }
    iterations = 0;
    int total = 0, preempted = 0;
    while (b200_batcher_pending(bat) > 0 && iterations < 1000) {
        CHECK_B200(b200_batcher_step(bat, dec, &gp, kp, vp, ws2, ws2_bytes, &finished, nullptr));
        ++iterations;
    }
    for (int r = 0; r < 9; ++r) {
        int n = 0, state = 0;
        CHECK_B200(b200_batcher_result(bat, rids[r], nullptr, 0, &n, &state));
        ok = ok && n == want[r] && state == B200_REQ_FINISHED;
        total += n, preempted += b200_batcher_preemptions(bat, rids[r]);
    }
    ok = ok && b200_batcher_free_pages(bat) == 5;
    printf("stream: 9 requests, %d tokens in %d iterations through 3 slots and 5 pages (%d preemptions)\n", total, iterations, preempted);
    b200_batcher_destroy(bat);
    b200_decoder_destroy(dec);
    printf(ok ? "serve_example passed\n" : "serve_example FAILED\n");
    return ok ? 0 : 1;
}
