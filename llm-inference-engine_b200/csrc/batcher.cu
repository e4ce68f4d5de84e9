// batcher.cu -- continuous batching over a paged KV cache ("batching / paged KV"; the reference keeps one static
// [L,B,Hkv,S,d] cache and one hard-coded prompt, src/models/llama/llama.cpp:47-48,327-341).
//
// Two halves:
//   * the SCHEDULER (host only, no CUDA call: b200_batcher_submit / _plan / _commit): a FCFS queue, a page allocator (free list of
//     B200_KV_PAGE_SIZE-position pages shared by all sequences) and the per-iteration plan -- which waiting requests are admitted and
//     prefilled, which running sequences take a decode step, which page every position lives in.  A running sequence that needs a page
//     when none is free preempts the most recently admitted one (its pages are freed, it re-queues at the FRONT with prompt + generated
//     tokens and is recomputed by a later prefill).  Pure bookkeeping: tested on the CPU with a fake sampler.
//   * the ITERATION (b200_batcher_step): plan -> [ids -> embedding -> b200_decoder_prefill_paged -> last rows -> LM head / top-k /
//     sampling] for the admitted requests -> [last ids -> embedding -> b200_decoder_step_paged -> tail] for the running ones -> sampled
//     ids back to the host -> commit.  Sequences join and leave between iterations without moving cache bytes: a row of the batch is
//     just a block-table row.
#include "attention_decode.cuh"

#include <algorithm>
#include <deque>
#include <new>
#include <vector>

namespace b200 {

struct Seq {
    int id = -1;
    std::vector<int> tokens;  // prompt + generated so far
    int prompt_len = 0, max_new = 0;
    int cached = 0;           // positions whose K / V rows are in the pool (0 while waiting)
    std::vector<int> pages;
    int state = B200_REQ_WAITING;
    int admitted_at = -1;     // iteration of the (last) admission: the preemption victim is the youngest
    int preemptions = 0;
    bool vocab_checked = false;
};

}  // namespace b200

struct b200_batcher {
    b200_batcher_config_t cfg;
    std::vector<b200::Seq> seqs;   // by request id
    std::deque<int> waiting;       // request ids, FCFS (preempted ones re-enter at the front)
    std::vector<int> running;      // request ids in admission order
    std::vector<int> free_pages;   // stack
    long long iteration = 0;
    // the current plan
    b200_batch_plan_t plan = {};
    std::vector<int> p_ids, p_lens, p_req, p_bt, d_tok, d_steps, d_req, d_bt, p_last;
    bool planned = false;
};

using namespace b200;

static int pages_for(int positions) { return (positions + B200_KV_PAGE_SIZE - 1) / B200_KV_PAGE_SIZE; }

static void release_pages(b200_batcher *b, Seq &s) {
    for (int p : s.pages) b->free_pages.push_back(p);
    s.pages.clear();
    s.cached = 0;
}

static void fill_bt_row(const b200_batcher *b, const Seq &s, std::vector<int> &bt) {
    const size_t at = bt.size();
    bt.resize(at + b->cfg.max_pages_per_seq, -1);
    for (size_t i = 0; i < s.pages.size(); ++i) bt[at + i] = s.pages[i];
}

extern "C" {

b200_batcher_t *b200_batcher_create(const b200_batcher_config_t *cfg) {
    if (!cfg || cfg->max_batch < 1 || cfg->num_pages < 1 || cfg->max_pages_per_seq < 1 || cfg->max_prefill_tokens < 1) {
        set_error("batcher_create: bad configuration");
        return nullptr;
    }
    b200_batcher *b = new (std::nothrow) b200_batcher();
    if (!b) return nullptr;
    b->cfg = *cfg;
    for (int p = cfg->num_pages - 1; p >= 0; --p) b->free_pages.push_back(p);  // page 0 is handed out first
    return b;
}

void b200_batcher_destroy(b200_batcher_t *b) { delete b; }

int b200_batcher_submit(b200_batcher_t *b, const int *prompt_ids, int prompt_len, int max_new_tokens) {
    B200_REQUIRE(b && prompt_ids, "batcher_submit: null argument");
    B200_REQUIRE(prompt_len >= 1 && max_new_tokens >= 1, "batcher_submit: bad prompt_len %d / max_new_tokens %d", prompt_len, max_new_tokens);
    // the longest prefill this request can ever need is prompt + all but its last token (after a preemption); its last position is
    // prompt_len + max_new_tokens - 2 (the final sampled token is never fed back)
    const int reach = prompt_len + max_new_tokens - 1;
    B200_REQUIRE(reach <= b->cfg.max_pages_per_seq * B200_KV_PAGE_SIZE, "batcher_submit: %d positions exceed a block-table row (%d pages of %d)",
                 reach, b->cfg.max_pages_per_seq, B200_KV_PAGE_SIZE);
    B200_REQUIRE(reach <= b->cfg.max_prefill_tokens, "batcher_submit: %d tokens exceed max_prefill_tokens %d (a preempted request is recomputed in one prefill)",
                 reach, b->cfg.max_prefill_tokens);
    B200_REQUIRE(pages_for(reach) <= b->cfg.num_pages, "batcher_submit: the request alone needs %d pages, the pool has %d", pages_for(reach),
                 b->cfg.num_pages);
    Seq s;
    s.id = (int)b->seqs.size();
    s.tokens.assign(prompt_ids, prompt_ids + prompt_len);
    s.prompt_len = prompt_len, s.max_new = max_new_tokens;
    b->seqs.push_back(std::move(s));
    b->waiting.push_back(b->seqs.back().id);
    return b->seqs.back().id;
}

int b200_batcher_plan(b200_batcher_t *b, b200_batch_plan_t *out) {
    B200_REQUIRE(b && out, "batcher_plan: null argument");
    B200_REQUIRE(!b->planned, "batcher_plan: the previous plan has not been committed");
    const b200_batcher_config_t &c = b->cfg;
    b->plan = {};
    b->p_ids.clear(), b->p_lens.clear(), b->p_req.clear(), b->p_bt.clear(), b->p_last.clear();
    b->d_tok.clear(), b->d_steps.clear(), b->d_req.clear(), b->d_bt.clear();

    // ---- 1. running sequences: the decode step appends position `cached`; a page boundary needs a page.  No page free: preempt the
    //         youngest running sequence (possibly the one asking) and try again.
    for (size_t i = 0; i < b->running.size();) {
        Seq &s = b->seqs[b->running[i]];
        if (pages_for(s.cached + 1) <= (int)s.pages.size()) {
            ++i;
            continue;
        }
        if (!b->free_pages.empty()) {
            s.pages.push_back(b->free_pages.back());
            b->free_pages.pop_back();
            ++i;
            continue;
        }
        size_t victim = 0;
        for (size_t k = 1; k < b->running.size(); ++k)
            if (b->seqs[b->running[k]].admitted_at >= b->seqs[b->running[victim]].admitted_at) victim = k;
        Seq &v = b->seqs[b->running[victim]];
        release_pages(b, v);
        v.state = B200_REQ_WAITING;
        ++v.preemptions;
        b->waiting.push_front(v.id);
        b->running.erase(b->running.begin() + victim);
        ++b->plan.n_preempted;
        if (victim < i) --i;  // the list shifted; re-examine the same sequence (unless it was the victim itself)
    }
    for (int id : b->running) {
        const Seq &s = b->seqs[id];
        b->d_req.push_back(id);
        b->d_tok.push_back(s.tokens.back());
        b->d_steps.push_back(s.cached + 1);
        fill_bt_row(b, s, b->d_bt);
        b->plan.decode_max_step = std::max(b->plan.decode_max_step, s.cached + 1);
    }
    b->plan.n_decode = (int)b->running.size();

    // ---- 2. admission, FCFS: a batch slot, the prompt's pages plus one spare page per admitted sequence (so that its first decode steps
    //         cannot preempt anybody), the token budget of one prefill pass and its padded-q budget (n * longest <= 2 * budget)
    int slots = c.max_batch - (int)b->running.size();
    int tokens = 0, longest = 0, n = 0;
    std::vector<int> admitted;
    while (slots > 0 && !b->waiting.empty() && b->plan.n_preempted == 0) {
        Seq &s = b->seqs[b->waiting.front()];
        const int len = (int)s.tokens.size();
        // the prompt's pages, and room for the first decode step's row unless this prefill already produces the request's last token
        const int need = pages_for(std::min(len + 1, s.prompt_len + s.max_new - 1));
        const int new_longest = std::max(longest, len);
        if (need + 1 > (int)b->free_pages.size() && !(b->running.empty() && admitted.empty() && need <= (int)b->free_pages.size())) break;
        if (tokens + len > c.max_prefill_tokens) break;
        if ((long long)(n + 1) * new_longest > 2LL * c.max_prefill_tokens) break;
        for (int i = 0; i < need; ++i) {
            s.pages.push_back(b->free_pages.back());
            b->free_pages.pop_back();
        }
        tokens += len, longest = new_longest, ++n, --slots;
        s.state = B200_REQ_RUNNING;
        s.admitted_at = (int)b->iteration;
        admitted.push_back(s.id);
        b->waiting.pop_front();
    }
    int cum = 0;
    for (int id : admitted) {
        const Seq &s = b->seqs[id];
        b->p_req.push_back(id);
        b->p_lens.push_back((int)s.tokens.size());
        b->p_ids.insert(b->p_ids.end(), s.tokens.begin(), s.tokens.end());
        cum += (int)s.tokens.size();
        b->p_last.push_back(cum - 1);
        fill_bt_row(b, s, b->p_bt);
    }
    b->plan.n_prefill = n, b->plan.prefill_tokens = tokens, b->plan.prefill_max_len = longest;
    b->plan.free_pages = (int)b->free_pages.size();
    b->plan.n_waiting = (int)b->waiting.size();
    b->planned = true;
    *out = b->plan;
    return B200_OK;
}

const int *b200_batcher_plan_array(const b200_batcher_t *b, int which) {
    if (!b || !b->planned) return nullptr;
    switch (which) {
        case B200_PLAN_PREFILL_IDS: return b->p_ids.data();
        case B200_PLAN_PREFILL_LENS: return b->p_lens.data();
        case B200_PLAN_PREFILL_REQUESTS: return b->p_req.data();
        case B200_PLAN_PREFILL_BLOCK_TABLE: return b->p_bt.data();
        case B200_PLAN_PREFILL_LAST_ROWS: return b->p_last.data();
        case B200_PLAN_DECODE_TOKENS: return b->d_tok.data();
        case B200_PLAN_DECODE_STEPS: return b->d_steps.data();
        case B200_PLAN_DECODE_REQUESTS: return b->d_req.data();
        case B200_PLAN_DECODE_BLOCK_TABLE: return b->d_bt.data();
        default: return nullptr;
    }
}

// Drop the current plan without results (a launch failed): the admitted requests give their pages back and return to the FRONT of the
// queue in their original order; running sequences keep the pages they were granted (they will need them at the next attempt);
// preempted ones stay re-queued.  The scheduler is then in a state from which b200_batcher_plan can be called again.
int b200_batcher_abort(b200_batcher_t *b) {
    B200_REQUIRE(b, "batcher_abort: null handle");
    if (!b->planned) return B200_OK;
    for (int i = b->plan.n_prefill - 1; i >= 0; --i) {
        Seq &s = b->seqs[b->p_req[i]];
        release_pages(b, s);
        s.state = B200_REQ_WAITING;
        b->waiting.push_front(s.id);
    }
    b->planned = false;
    return B200_OK;
}

int b200_batcher_commit(b200_batcher_t *b, const int *prefill_sampled, const int *decode_sampled, int end_id) {
    B200_REQUIRE(b, "batcher_commit: null handle");
    B200_REQUIRE(b->planned, "batcher_commit: nothing planned");
    B200_REQUIRE((b->plan.n_prefill == 0 || prefill_sampled) && (b->plan.n_decode == 0 || decode_sampled), "batcher_commit: missing sampled ids");
    int finished = 0;
    auto accept = [&](Seq &s, int tok, int now_cached) {
        s.cached = now_cached;
        s.tokens.push_back(tok);
        const int generated = (int)s.tokens.size() - s.prompt_len;
        if (tok == end_id || generated >= s.max_new) {
            s.state = B200_REQ_FINISHED;
            release_pages(b, s);
            ++finished;
            return true;
        }
        return false;
    };
    // the decode rows first (they were planned from `running` in order), then the admitted ones join `running`
    std::vector<int> still;
    for (int i = 0; i < b->plan.n_decode; ++i) {
        Seq &s = b->seqs[b->d_req[i]];
        if (!accept(s, decode_sampled[i], b->d_steps[i])) still.push_back(s.id);
    }
    for (int i = 0; i < b->plan.n_prefill; ++i) {
        Seq &s = b->seqs[b->p_req[i]];
        if (!accept(s, prefill_sampled[i], b->p_lens[i])) still.push_back(s.id);
    }
    b->running.swap(still);
    b->planned = false;
    ++b->iteration;
    return finished;
}

int b200_batcher_result(const b200_batcher_t *b, int request, int *out_ids, int capacity, int *n_generated, int *state) {
    B200_REQUIRE(b, "batcher_result: null handle");
    B200_REQUIRE(request >= 0 && request < (int)b->seqs.size(), "batcher_result: unknown request %d", request);
    const Seq &s = b->seqs[request];
    const int n = (int)s.tokens.size() - s.prompt_len;
    if (n_generated) *n_generated = n;
    if (state) *state = s.state;
    if (out_ids)
        for (int i = 0; i < n && i < capacity; ++i) out_ids[i] = s.tokens[s.prompt_len + i];
    return B200_OK;
}

int b200_batcher_pending(const b200_batcher_t *b) { return b ? (int)(b->waiting.size() + b->running.size()) : 0; }
int b200_batcher_free_pages(const b200_batcher_t *b) { return b ? (int)b->free_pages.size() : 0; }
int b200_batcher_preemptions(const b200_batcher_t *b, int request) {
    return b && request >= 0 && request < (int)b->seqs.size() ? b->seqs[request].preemptions : -1;
}

// ---------------------------------------------------------------- the GPU iteration
namespace {
struct Carve {
    size_t ids, ints, hidden_prompt, prefill, prefill_bytes, hidden, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len, finished, output_id, total;
};
size_t al(size_t v) { return (v + 255) & ~(size_t)255; }
Carve carve(const b200_batcher *b, const b200_decoder_t *dec, const b200_generate_params_t *p) {
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    const b200_batcher_config_t &k = b->cfg;
    const size_t e = c.dtype == B200_F32 ? 4 : 2, B = k.max_batch, T = k.max_prefill_tokens;
    Carve v = {};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += al(bytes);
        return at;
    };
    v.ids = take(T * sizeof(int));
    v.ints = take((4 * B + 2 * B * (size_t)k.max_pages_per_seq + 2 * B) * sizeof(int));  // lens x3, last rows, two block tables, steps, tokens
    v.hidden_prompt = take(T * c.hidden * e);
    // the prefill scratch of the worst pass the admission policy lets through: n sequences, n * longest <= 2 T, sum <= T
    size_t worst = 0;
    for (int n = 1; n <= k.max_batch; ++n) {
        const long long mq = std::min<long long>(T, 2LL * T / n);
        if (mq < 1) break;
        worst = std::max(worst, b200_decoder_prefill_scratch_bytes(dec, n, (int)mq, (int)std::min<long long>(T, n * mq)));
    }
    v.prefill = take(worst);
    v.prefill_bytes = al(worst);
    v.hidden = take(B * c.hidden * e);
    v.logits = take(B * (size_t)p->vocab * sizeof(float));
    v.tmp_ids = take(B * B200_TOPK_BLOCKS * (size_t)p->top_k * sizeof(int));
    v.tmp_vals = take(B * B200_TOPK_BLOCKS * (size_t)p->top_k * sizeof(float));
    v.topk_ids = take(B * (size_t)p->top_k * sizeof(int));
    v.topk_vals = take(B * (size_t)p->top_k * sizeof(float));
    v.seq_len = take(B * sizeof(int));
    v.finished = take(B);
    v.output_id = take(B * sizeof(int));
    v.total = o;
    return v;
}
int check(const b200_batcher *b, const b200_decoder_t *dec, const b200_generate_params_t *p) {
    B200_REQUIRE(b && dec && p, "batcher: null argument");
    B200_REQUIRE(p->embedding && p->final_gamma && p->lm_head, "batcher: missing embedding / final norm / LM head");
    B200_REQUIRE(p->vocab > 0 && p->top_k >= 1 && p->top_k <= B200_TOPK_MAX_K, "batcher: bad vocab %d or top_k %d", p->vocab, p->top_k);
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    B200_REQUIRE(c.tp_world <= 1, "batcher: single-GPU engines only");
    B200_REQUIRE(b->cfg.max_batch <= c.max_batch, "batcher: max_batch %d exceeds the engine's %d", b->cfg.max_batch, c.max_batch);
    B200_REQUIRE(b->cfg.max_pages_per_seq * B200_KV_PAGE_SIZE <= c.max_seq_len, "batcher: a block-table row reaches %d positions, the engine %d",
                 b->cfg.max_pages_per_seq * B200_KV_PAGE_SIZE, c.max_seq_len);
    B200_REQUIRE(c.head_size == 128 && c.dtype != B200_F32, "batcher: the paged kernels serve head size 128 and 16-bit dtypes");
    return B200_OK;
}
}  // namespace

size_t b200_batcher_workspace_bytes(const b200_batcher_t *b, const b200_decoder_t *dec, const b200_generate_params_t *p) {
    if (check(b, dec, p) != B200_OK) return 0;
    return carve(b, dec, p).total;
}

int b200_batcher_step(b200_batcher_t *b, b200_decoder_t *dec, const b200_generate_params_t *p, void *k_pool, void *v_pool, void *workspace,
                      size_t workspace_bytes, int *n_finished, b200_stream_t stream) {
    int rc = check(b, dec, p);
    if (rc != B200_OK) return rc;
    B200_REQUIRE(k_pool && v_pool && workspace, "batcher_step: null pointer");
    B200_REQUIRE(((uintptr_t)workspace & 255) == 0, "batcher_step: workspace must be 256-byte aligned");
    const Carve k = carve(b, dec, p);
    B200_REQUIRE(workspace_bytes >= k.total, "batcher_step: need %zu bytes of workspace, got %zu", k.total, workspace_bytes);
    b200_decoder_config_t c;
    b200_decoder_get_config(dec, &c);
    NvtxRange range("b200 batcher iteration");
    // a queued prompt with an id outside the vocabulary would read outside the embedding table: reject the REQUEST (state
    // B200_REQ_REJECTED, out of the queue), not the iteration
    for (auto it = b->waiting.begin(); it != b->waiting.end();) {
        Seq &s = b->seqs[*it];
        bool ok = true;
        if (!s.vocab_checked) {
            for (int t = 0; t < s.prompt_len && ok; ++t) ok = s.tokens[t] >= 0 && s.tokens[t] < p->vocab;
            s.vocab_checked = true;
        }
        if (!ok) {
            s.state = B200_REQ_REJECTED;
            it = b->waiting.erase(it);
        } else {
            ++it;
        }
    }
    b200_batch_plan_t plan;
    if ((rc = b200_batcher_plan(b, &plan)) != B200_OK) return rc;
    char *w = (char *)workspace;
    const int B = b->cfg.max_batch, MP = b->cfg.max_pages_per_seq;
    int *ids = (int *)(w + k.ids), *ints = (int *)(w + k.ints);
    int *lens = ints, *last = ints + 3 * B, *bt_p = ints + 4 * B, *bt_d = bt_p + (size_t)B * MP, *steps = bt_d + (size_t)B * MP, *toks = steps + B;
    void *hidden_prompt = w + k.hidden_prompt, *hidden = w + k.hidden;
    float *logits = (float *)(w + k.logits), *tmp_vals = (float *)(w + k.tmp_vals), *topk_vals = (float *)(w + k.topk_vals);
    int *tmp_ids = (int *)(w + k.tmp_ids), *topk_ids = (int *)(w + k.topk_ids), *seq_len = (int *)(w + k.seq_len);
    uint8_t *finished = (uint8_t *)(w + k.finished);
    int *output_id = (int *)(w + k.output_id);
    cudaStream_t st = as_stream(stream);
    std::vector<int> sampled_p(plan.n_prefill), sampled_d(plan.n_decode);
    auto fail = [&](int code) {
        b200_batcher_abort(b);  // the admitted requests return to the front of the queue with their pages freed: the caller may retry
        return code;
    };
    auto up = [&](void *dst, const void *src, size_t bytes) { return bytes == 0 || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess; };

    // ---- admitted requests: one packed prefill pass, first token from the last prompt position of each
    if (plan.n_prefill > 0) {
        NvtxRange phase("admitted: prefill");
        const int n = plan.n_prefill, T = plan.prefill_tokens;
        std::vector<int> hl((size_t)3 * n);
        for (int i = 0; i < n; ++i) hl[i] = b->p_lens[i], hl[n + i] = 0, hl[2 * n + i] = b->p_lens[i];
        if (!up(ids, b->p_ids.data(), (size_t)T * sizeof(int)) || !up(lens, hl.data(), hl.size() * sizeof(int)) ||
            !up(last, b->p_last.data(), (size_t)n * sizeof(int)) || !up(bt_p, b->p_bt.data(), b->p_bt.size() * sizeof(int)) ||
            !up(seq_len, b->p_lens.data(), (size_t)n * sizeof(int)) || cudaMemsetAsync(finished, 0, (size_t)n, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)  // `hl` is pageable host memory about to go out of scope
            return fail(cuda_status("batcher_step H2D (prefill)"));
        if ((rc = b200_input_embedding(ids, p->embedding, hidden_prompt, T, c.hidden, c.dtype, stream)) != B200_OK) return fail(rc);
        const size_t need = b200_decoder_prefill_scratch_bytes(dec, n, plan.prefill_max_len, T);
        if (need > k.prefill_bytes || need == 0) {
            set_error("batcher_step: prefill scratch %zu exceeds the reservation", need);
            return fail(B200_ERR_WORKSPACE);
        }
        rc = b200_decoder_prefill_paged(dec, hidden_prompt, k_pool, v_pool, bt_p, lens, lens + n, lens + 2 * n, n, plan.prefill_max_len, T,
                                        b->cfg.num_pages, MP, w + k.prefill, need, 0, c.num_layers, stream);
        if (rc != B200_OK) return fail(rc);
        if ((rc = b200_input_embedding(last, hidden_prompt, hidden, n, c.hidden, c.dtype, stream)) != B200_OK) return fail(rc);
        rc = b200_lm_head_topk_sample(dec, hidden, p->final_gamma, p->lm_head, p->vocab, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len,
                                      finished, output_id, n, p->top_k, plan.prefill_max_len, p->end_id, stream);
        if (rc != B200_OK) return fail(rc);
        if (cudaMemcpyAsync(sampled_p.data(), output_id, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return fail(cuda_status("batcher_step D2H (prefill)"));
    }
    // ---- running sequences: one decode step, every row at its own position, through the block table
    if (plan.n_decode > 0) {
        NvtxRange phase("running: decode step");
        const int n = plan.n_decode;
        if (!up(toks, b->d_tok.data(), (size_t)n * sizeof(int)) || !up(steps, b->d_steps.data(), (size_t)n * sizeof(int)) ||
            !up(bt_d, b->d_bt.data(), b->d_bt.size() * sizeof(int)) || !up(seq_len, b->d_steps.data(), (size_t)n * sizeof(int)) ||
            cudaMemsetAsync(finished, 0, (size_t)n, st) != cudaSuccess)
            return fail(cuda_status("batcher_step H2D (decode)"));
        if ((rc = b200_input_embedding(toks, p->embedding, hidden, n, c.hidden, c.dtype, stream)) != B200_OK) return fail(rc);
        rc = b200_decoder_step_paged(dec, hidden, k_pool, v_pool, bt_d, steps, n, plan.decode_max_step, b->cfg.num_pages, MP, 0, c.num_layers, stream);
        if (rc != B200_OK) return fail(rc);
        rc = b200_lm_head_topk_sample(dec, hidden, p->final_gamma, p->lm_head, p->vocab, logits, tmp_ids, tmp_vals, topk_ids, topk_vals, seq_len,
                                      finished, output_id, n, p->top_k, plan.decode_max_step, p->end_id, stream);
        if (rc != B200_OK) return fail(rc);
        if (cudaMemcpyAsync(sampled_d.data(), output_id, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return fail(cuda_status("batcher_step D2H (decode)"));
    }
    const int fin = b200_batcher_commit(b, sampled_p.data(), sampled_d.data(), p->end_id);
    if (fin < 0) return fin;
    if (n_finished) *n_finished = fin;
    return B200_OK;
}

}  // extern "C"
