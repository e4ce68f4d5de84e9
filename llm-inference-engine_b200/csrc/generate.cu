// generate.cu -- the generation loop around the decoder-layer path: what the reference's LlamaModel<T>::response /
// generateFirstToken / generateNextToken / LMHeadAndTopKSample intend (src/models/llama/llama.cpp:165-398 -- dead code there: the class
// does not compile and hard-codes a 13-token prompt), as one stream-ordered C-ABI call:
//
//   prompt ids -> launchInputEmbedding -> context decoder over the whole prompt (b200_decoder_prefill, KV cache filled at history 0)
//              -> final RMSNorm + LM head on the LAST prompt token of every sequence -> top-k -> sampling        (llama.cpp:166-217, 259-318)
//   then per new token:  embedding of the sampled id -> self decoder step over the cache -> final RMSNorm + LM head -> top-k -> sampling
//                                                                                                                 (llama.cpp:219-257)
//
// Differences from the reference's loop, all on the host side: no per-kernel device synchronisation, no allocation inside the loop,
// the sampled id never visits the host between steps (the next embedding reads it from device memory), and the recorded ids are
// scanned for the end id every `check_every` steps instead of every step.  `step` follows the reference: the sampling seed of the first token is the
// prompt length, then it is incremented once per token (llama.cpp:353,372) and is also the self decoder's 1-based position count.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace b200 {

static size_t g_align(size_t v) { return (v + 255) & ~(size_t)255; }

struct GenCarve {
    size_t ids, lens, hidden_prompt, prefill, hidden, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len, finished, output_id, out_ids, steps, total;
};

static GenCarve gen_carve(const b200_decoder_t *dec, const b200_generate_params_t *p, int batch, int prompt_len) {
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    const size_t e = c.dtype == B200_F32 ? 4 : 2, T = (size_t)batch * prompt_len;
    GenCarve k = {};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += g_align(bytes);
        return at;
    };
    k.ids = take(T * sizeof(int));
    k.lens = take((size_t)4 * batch * sizeof(int));  // input / history / context lengths + the packed row of every sequence's last prompt token
    k.hidden_prompt = take(T * c.hidden * e);
    k.prefill = take(b200_decoder_prefill_scratch_bytes(dec, batch, prompt_len, (int)T));
    k.hidden = take((size_t)batch * c.hidden * e);
    k.logits = take((size_t)batch * p->vocab * sizeof(float));
    k.tmp_ids = take((size_t)batch * B200_TOPK_BLOCKS * p->top_k * sizeof(int));
    k.tmp_vals = take((size_t)batch * B200_TOPK_BLOCKS * p->top_k * sizeof(float));
    k.topk_ids = take((size_t)batch * p->top_k * sizeof(int));
    k.topk_vals = take((size_t)batch * p->top_k * sizeof(float));
    k.seq_len = take((size_t)batch * sizeof(int));
    k.finished = take((size_t)batch);
    k.output_id = take((size_t)batch * sizeof(int));
    k.out_ids = take((size_t)batch * p->max_new_tokens * sizeof(int));
    k.steps = take((size_t)batch * p->max_new_tokens * sizeof(int));  // ragged prompts: per-row positions of every decode step
    k.total = o;
    return k;
}

static int gen_check(const b200_decoder_t *dec, const b200_generate_params_t *p, int batch, int prompt_len) {
    B200_REQUIRE(dec && p, "generate: null argument");
    B200_REQUIRE(p->embedding && p->final_gamma && p->lm_head, "generate: missing embedding / final norm / LM head");
    B200_REQUIRE(p->vocab > 0 && p->top_k >= 1 && p->top_k <= B200_TOPK_MAX_K, "generate: bad vocab %d or top_k %d", p->vocab, p->top_k);
    B200_REQUIRE(p->max_new_tokens >= 1, "generate: max_new_tokens %d < 1", p->max_new_tokens);
    B200_REQUIRE(batch >= 1 && prompt_len >= 1, "generate: bad batch %d / prompt_len %d", batch, prompt_len);
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    B200_REQUIRE(c.tp_world <= 1, "generate: single-GPU engines only");
    B200_REQUIRE(batch == c.max_batch, "generate: batch %d must equal the engine's (and the cache's) batch dimension %d", batch, c.max_batch);
    B200_REQUIRE(prompt_len + p->max_new_tokens - 1 <= c.max_seq_len, "generate: prompt %d + %d new tokens exceed max_seq_len %d", prompt_len,
                 p->max_new_tokens, c.max_seq_len);
    return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_generate_workspace_bytes(const b200_decoder_t *dec, const b200_generate_params_t *p, int batch, int prompt_len) {
    if (gen_check(dec, p, batch, prompt_len) != B200_OK) return 0;
    return gen_carve(dec, p, batch, prompt_len).total;
}

int b200_generate_ragged(b200_decoder_t *dec, const b200_generate_params_t *p, const int *prompt_ids, const int *prompt_lens, int batch,
                         int prompt_len, void *k_cache, void *v_cache, void *workspace, size_t workspace_bytes, int *out_ids, int *n_generated,
                         b200_stream_t stream) {
    int rc = gen_check(dec, p, batch, prompt_len);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(prompt_ids && k_cache && v_cache && workspace && out_ids, "generate: null pointer");
    B200_REQUIRE(((uintptr_t)workspace & 255) == 0, "generate: workspace must be 256-byte aligned");
    const GenCarve k = gen_carve(dec, p, batch, prompt_len);
    B200_REQUIRE(workspace_bytes >= k.total, "generate: need %zu bytes of workspace, got %zu", k.total, workspace_bytes);
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    const int V = p->vocab, K = p->top_k, N = p->max_new_tokens;
    // ---- the prompts, packed back to back (the context decoder's un-padded token layout); prompt_lens == NULL: all of length prompt_len
    std::vector<int> len(batch, prompt_len), packed;
    bool ragged = false;
    int max_len = 0;
    for (int b = 0; b < batch; ++b) {
        if (prompt_lens) len[b] = prompt_lens[b];
        B200_REQUIRE(len[b] >= 1 && len[b] <= prompt_len, "generate: prompt length %d of sequence %d outside [1, %d]", len[b], b, prompt_len);
        ragged = ragged || len[b] != len[0];
        max_len = len[b] > max_len ? len[b] : max_len;
        for (int t = 0; t < len[b]; ++t) {
            const int id = prompt_ids[(size_t)b * prompt_len + t];
            B200_REQUIRE(id >= 0 && id < V, "generate: prompt id %d at (%d, %d) outside the vocabulary", id, b, t);
            packed.push_back(id);
        }
    }
    const int T = (int)packed.size();
    char *w = (char *)workspace;
    int *ids = (int *)(w + k.ids), *lens = (int *)(w + k.lens);
    void *hidden_prompt = w + k.hidden_prompt, *hidden = w + k.hidden;
    float *logits = (float *)(w + k.logits), *tmp_vals = (float *)(w + k.tmp_vals), *topk_vals = (float *)(w + k.topk_vals);
    int *tmp_ids = (int *)(w + k.tmp_ids), *topk_ids = (int *)(w + k.topk_ids), *seq_len = (int *)(w + k.seq_len);
    uint8_t *finished = (uint8_t *)(w + k.finished);
    int *output_id = (int *)(w + k.output_id), *out_dev = (int *)(w + k.out_ids), *steps_dev = (int *)(w + k.steps);
    cudaStream_t st = as_stream(stream);

    // ---- host -> device: the prompt, the three length vectors of the context decoder (input = context = prompt length, history 0), the
    //      packed row of every sequence's last prompt token, and (ragged only) the position of every row at every decode step
    std::vector<int> hl((size_t)4 * batch), steps_host;
    for (int b = 0, cum = 0; b < batch; ++b) {
        hl[b] = len[b], hl[batch + b] = 0, hl[2 * batch + b] = len[b];
        cum += len[b];
        hl[3 * batch + b] = cum - 1;
    }
    if (ragged) {
        steps_host.resize((size_t)N * batch);
        for (int i = 0; i < N; ++i)
            for (int b = 0; b < batch; ++b) steps_host[(size_t)i * batch + b] = len[b] + i;
    }
    if (cudaMemcpyAsync(ids, packed.data(), (size_t)T * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(lens, hl.data(), hl.size() * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(seq_len, len.data(), (size_t)batch * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||  // llama.cpp: sequence_lengths
        (ragged && cudaMemcpyAsync(steps_dev, steps_host.data(), steps_host.size() * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) ||
        cudaMemsetAsync(finished, 0, (size_t)batch, st) != cudaSuccess)
        return cuda_status("generate H2D");
    if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_status("generate H2D sync");  // pageable host vectors: done with them here

    NvtxRange range("b200 generate");
    // ---- first token: embedding -> context decoder -> last prompt token of every sequence (a row gather) -> sampling tail
    if ((rc = b200_input_embedding(ids, p->embedding, hidden_prompt, T, c.hidden, c.dtype, stream)) != B200_OK) return rc;
    rc = b200_decoder_prefill(dec, hidden_prompt, k_cache, v_cache, lens, lens + batch, lens + 2 * batch, batch, max_len, T, w + k.prefill,
                              b200_decoder_prefill_scratch_bytes(dec, batch, max_len, T), 0, c.num_layers, stream);
    if (rc != B200_OK) return rc;
    if ((rc = b200_input_embedding(lens + 3 * batch, hidden_prompt, hidden, batch, c.hidden, c.dtype, stream)) != B200_OK) return rc;
    int step = max_len;  // llama.cpp:353: step->data = &context_length (ragged: the longest prompt -- the sampling seed is shared by the batch)
    rc = b200_lm_head_topk_sample(dec, hidden, p->final_gamma, p->lm_head, V, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len, finished,
                                  output_id, batch, K, step, p->end_id, stream);
    if (rc != B200_OK) return rc;

    // ---- the decode loop: the sampled ids stay on the device
    std::vector<int> poll((size_t)batch * N);
    int produced = 1;
    for (int i = 1; i <= N; ++i) {
        // record token i-1 (column i-1 of out_dev[batch, N])
        if (cudaMemcpy2DAsync(out_dev + (i - 1), (size_t)N * sizeof(int), output_id, sizeof(int), sizeof(int), batch, cudaMemcpyDeviceToDevice, st) !=
            cudaSuccess)
            return cuda_status("generate record");
        if (i == N) break;
        if (p->check_every > 0 && i % p->check_every == 0) {
            // stop early once every sequence HAS produced end_id: decided from the recorded ids, not from the `finished` flags -- the
            // sampling kernel keeps the reference's flag semantics (finished = "the token sampled just now is end_id"), which is not sticky
            if (cudaMemcpyAsync(poll.data(), out_dev, poll.size() * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess)
                return cuda_status("generate poll");
            bool all = true;
            for (int b = 0; b < batch && all; ++b) {
                bool hit = false;
                for (int t = 0; t < i && !hit; ++t) hit = poll[(size_t)b * N + t] == p->end_id;
                all = hit;
            }
            if (all) break;
        }
        ++step;  // llama.cpp:372
        if ((rc = b200_input_embedding(output_id, p->embedding, hidden, batch, c.hidden, c.dtype, stream)) != B200_OK) return rc;
        rc = ragged ? b200_decoder_step_ragged(dec, hidden, k_cache, v_cache, batch, steps_dev + (size_t)i * batch, step, 0, c.num_layers, stream)
                    : b200_decoder_step(dec, hidden, k_cache, v_cache, batch, step, 0, c.num_layers, stream);
        if (rc != B200_OK) return rc;
        rc = b200_lm_head_topk_sample(dec, hidden, p->final_gamma, p->lm_head, V, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len, finished,
                                      output_id, batch, K, step, p->end_id, stream);
        if (rc != B200_OK) return rc;
        ++produced;
    }

    // ---- device -> host; everything after a sequence's first end_id is end_id, n_generated excludes it (the reference stops before
    //      emitting eos, llama.cpp:366-369)
    std::vector<int> host((size_t)batch * N);
    if (cudaMemcpyAsync(host.data(), out_dev, host.size() * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return cuda_status("generate D2H");
    for (int b = 0; b < batch; ++b) {
        int n = 0;
        while (n < produced && host[(size_t)b * N + n] != p->end_id) ++n;
        for (int i = 0; i < N; ++i) out_ids[(size_t)b * N + i] = i < n ? host[(size_t)b * N + i] : p->end_id;
        if (n_generated) n_generated[b] = n;
    }
    return B200_OK;
}

int b200_generate(b200_decoder_t *dec, const b200_generate_params_t *p, const int *prompt_ids, int batch, int prompt_len, void *k_cache,
                  void *v_cache, void *workspace, size_t workspace_bytes, int *out_ids, int *n_generated, b200_stream_t stream) {
    return b200_generate_ragged(dec, p, prompt_ids, nullptr, batch, prompt_len, k_cache, v_cache, workspace, workspace_bytes, out_ids, n_generated,
                                stream);
}

}  // extern "C"
