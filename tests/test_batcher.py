"""Continuous batching over a paged KV cache (absent from the reference: one static [L,B,Hkv,S,d] cache,
src/models/llama/llama.cpp:47-48, one prompt at a time, :327-398).

CPU part (no GPU): the scheduler half of b200_batcher_* -- submit / plan / commit are pure host bookkeeping -- driven through the C ABI
with a FAKE sampler whose next token is a function of the whole sequence so far: whatever the batch composition, admission order or
preemptions, every request must end up with exactly the tokens the fake model produces for it alone, and the page bookkeeping must hold
at every iteration (no page held twice, pages conserved, block tables = the pages held, positions covered).

GPU part: the paged kernels are bit-identical to the contiguous ones on the same rows (decode attention, context attention, prefill,
engine step), and the iteration loop reproduces b200_generate_ragged / the oracle."""
import numpy as np
import pytest

from util import b200

PAGE = 64
END = 7


def fake_next(tokens, vocab=1000):
    """A deterministic 'model': the next token depends on every token so far (so a wrong recompute after a preemption shows)."""
    h = 1469598103934665603
    for t in tokens:
        h = ((h ^ (int(t) + 1)) * 1099511628211) % (1 << 64)
    return int(h % vocab)


def alone(prompt, max_new):
    toks, out = list(prompt), []
    for _ in range(max_new):
        t = fake_next(toks)
        out.append(t)
        toks.append(t)
        if t == END:
            break
    return out


def drive(bat, requests, mod, check_every_iteration=True, max_iterations=10000):
    """Run plan / commit until nothing is pending; returns per-iteration plans.  requests: id -> (prompt, max_new)."""
    cfg = bat.cfg
    tokens_of = {rid: list(p) for rid, (p, _) in requests.items()}  # the scheduler's view, mirrored here
    history = []
    it = 0
    while bat.pending() > 0:
        it += 1
        assert it < max_iterations, "scheduler does not make progress"
        plan, v = bat.plan()
        history.append(plan)
        assert plan.n_prefill + plan.n_decode > 0 or plan.n_preempted > 0, "an iteration that does nothing"
        assert plan.n_prefill + plan.n_decode <= cfg.max_batch
        # ---- prefill rows: the packed ids are exactly each admitted sequence's tokens so far
        cum = 0
        for i, rid in enumerate(v["prefill_requests"]):
            n = int(v["prefill_lens"][i])
            assert list(v["prefill_ids"][cum:cum + n]) == tokens_of[int(rid)]
            cum += n
            assert int(v["prefill_last_rows"][i]) == cum - 1
        assert cum == plan.prefill_tokens <= cfg.max_prefill_tokens
        if plan.n_prefill:
            assert plan.prefill_max_len == int(v["prefill_lens"].max())
            assert plan.n_prefill * plan.prefill_max_len <= 2 * cfg.max_prefill_tokens
        # ---- decode rows: feed the last token at step = len(tokens)
        for i, rid in enumerate(v["decode_requests"]):
            assert int(v["decode_tokens"][i]) == tokens_of[int(rid)][-1]
            assert int(v["decode_steps"][i]) == len(tokens_of[int(rid)])
        if plan.n_decode:
            assert plan.decode_max_step == int(v["decode_steps"].max())
        if check_every_iteration:
            # ---- pages: every position a row touches this iteration has a page; no page is held by two rows; pages are conserved
            held = []
            for bt, need in list(zip(v["prefill_block_table"], [int(n) for n in v["prefill_lens"]])) + \
                    list(zip(v["decode_block_table"], [int(s) for s in v["decode_steps"]])):
                pages = [int(p) for p in bt if p >= 0]
                assert len(pages) >= (need + PAGE - 1) // PAGE, f"positions [0, {need}) not covered by {pages}"
                assert all(0 <= p < cfg.num_pages for p in pages)
                assert list(bt[:len(pages)]) == pages and all(p == -1 for p in bt[len(pages):]), "pages must be a prefix of the row"
                held += pages
            assert len(held) == len(set(held)), "a page is held by two sequences"
            assert len(held) + plan.free_pages == cfg.num_pages, "pages leaked or invented"
        sp = [fake_next(tokens_of[int(r)]) for r in v["prefill_requests"]]
        sd = [fake_next(tokens_of[int(r)]) for r in v["decode_requests"]]
        for r, t in list(zip(v["prefill_requests"], sp)) + list(zip(v["decode_requests"], sd)):
            tokens_of[int(r)].append(t)
        bat.commit(sp, sd, END)
    assert bat.free_pages() == cfg.num_pages, "pages not returned at the end"
    return history


def test_scheduler_every_request_gets_its_own_tokens_whatever_the_batch():
    mod = b200()
    rng = np.random.default_rng(3)
    bat = mod.Batcher(max_batch=4, num_pages=40, max_pages_per_seq=6, max_prefill_tokens=400)
    requests = {}
    for _ in range(17):
        prompt = rng.integers(8, 1000, size=int(rng.integers(1, 200))).tolist()
        max_new = int(rng.integers(1, 150))
        requests[bat.submit(prompt, max_new)] = (prompt, max_new)
    history = drive(bat, requests, mod)
    for rid, (prompt, max_new) in requests.items():
        got, state = bat.result(rid)
        assert state == mod.REQ_FINISHED
        assert list(got) == alone(prompt, max_new), f"request {rid}"
    assert max(p.n_decode for p in history) > 1, "the test never batched anything"


def test_scheduler_admission_is_first_come_first_served_and_respects_the_budgets():
    mod = b200()
    bat = mod.Batcher(max_batch=3, num_pages=100, max_pages_per_seq=4, max_prefill_tokens=200)
    reqs = {}
    for n in (150, 40, 100, 5):
        p = list(range(10, 10 + n))
        reqs[bat.submit(p, 3)] = (p, 3)
    plan, v = bat.plan()
    # 150 + 40 fit the 200-token pass; 100 does not (and nothing overtakes it: FCFS)
    assert list(v["prefill_requests"]) == [0, 1] and plan.n_waiting == 2
    bat.commit([fake_next(reqs[0][0]), fake_next(reqs[1][0])], [], END)
    plan, v = bat.plan()
    assert plan.n_decode == 2 and list(v["prefill_requests"]) == [2], "one slot left of max_batch 3"
    bat.commit([11], [12, 13], END)
    with pytest.raises(mod.B200Error):
        bat.submit(list(range(300)), 2)  # longer than one prefill pass
    with pytest.raises(mod.B200Error):
        bat.submit([1, 2, 3], 4 * PAGE)  # would outgrow a block-table row


def test_scheduler_preempts_the_youngest_and_recomputes_it():
    """A pool too small for everybody: sequences grow across page boundaries until no page is free; the most recently admitted running
    sequence is pushed back to the FRONT of the queue with prompt + generated tokens and later recomputed -- same final tokens."""
    mod = b200()
    bat = mod.Batcher(max_batch=4, num_pages=7, max_pages_per_seq=4, max_prefill_tokens=256)
    requests = {}
    for i in range(4):
        prompt = [20 + i] * 60  # one page each, crossing into a second page after 4 generated tokens
        requests[bat.submit(prompt, 150)] = (prompt, 150)
    history = drive(bat, requests, mod)
    assert sum(p.n_preempted for p in history) > 0, "the scenario was meant to run out of pages"
    pre = [bat.preemptions(r) for r in requests]
    assert pre[0] == 0, "the oldest sequence is never the victim while younger ones run"
    for rid, (prompt, max_new) in requests.items():
        got, state = bat.result(rid)
        assert state == mod.REQ_FINISHED and list(got) == alone(prompt, max_new), f"request {rid} (preempted {pre[rid]} times)"


def test_scheduler_a_request_as_large_as_the_pool_runs_alone():
    mod = b200()
    bat = mod.Batcher(max_batch=2, num_pages=2, max_pages_per_seq=2, max_prefill_tokens=128)
    p0, p1 = list(range(10, 110)), list(range(200, 230))
    r0, r1 = bat.submit(p0, 29), bat.submit(p1, 5)  # r0 reaches 128 positions = both pages
    drive(bat, {r0: (p0, 29), r1: (p1, 5)}, mod)
    assert list(bat.result(r0)[0]) == alone(p0, 29) and list(bat.result(r1)[0]) == alone(p1, 5)


@pytest.mark.parametrize("seed", range(12))
def test_scheduler_randomised_stress(seed):
    """Random pools (often too small), batch sizes, budgets and request mixes, requests submitted WHILE the loop runs: the invariants of
    drive() must hold at every iteration, nothing may starve, and every request ends with its own tokens."""
    mod = b200()
    rng = np.random.default_rng(1000 + seed)
    max_pages = int(rng.integers(2, 6))
    cfg = dict(max_batch=int(rng.integers(1, 7)), num_pages=int(rng.integers(max_pages, 4 * max_pages + 1)), max_pages_per_seq=max_pages,
               max_prefill_tokens=max_pages * PAGE)
    bat = mod.Batcher(**cfg)
    requests, pending_submissions = {}, []
    for _ in range(int(rng.integers(5, 25))):
        reach = int(rng.integers(1, max_pages * PAGE + 1))        # prompt + max_new - 1
        plen = int(rng.integers(1, reach + 1))
        pending_submissions.append((rng.integers(8, 1000, size=plen).tolist(), reach - plen + 1))
    # a third of the requests are queued up front, the rest trickle in between iterations
    tokens_of, it = {}, 0
    def submit_some(n):
        for _ in range(min(n, len(pending_submissions))):
            p, m = pending_submissions.pop()
            rid = bat.submit(p, m)
            requests[rid] = (p, m)
            tokens_of[rid] = list(p)
    submit_some(max(1, len(pending_submissions) // 3))
    while bat.pending() > 0 or pending_submissions:
        it += 1
        assert it < 20000, "scheduler does not make progress"
        if pending_submissions and (bat.pending() == 0 or rng.random() < 0.3):
            submit_some(int(rng.integers(1, 4)))
        plan, v = bat.plan()
        assert plan.n_prefill + plan.n_decode > 0 or plan.n_preempted > 0
        held = []
        for bt, need in list(zip(v["prefill_block_table"], [int(n) for n in v["prefill_lens"]])) + \
                list(zip(v["decode_block_table"], [int(s_) for s_ in v["decode_steps"]])):
            pages = [int(p) for p in bt if p >= 0]
            assert len(pages) >= (need + PAGE - 1) // PAGE
            held += pages
        assert len(held) == len(set(held)) and len(held) + plan.free_pages == cfg["num_pages"]
        sp = [fake_next(tokens_of[int(r)]) for r in v["prefill_requests"]]
        sd = [fake_next(tokens_of[int(r)]) for r in v["decode_requests"]]
        for r, t in list(zip(v["prefill_requests"], sp)) + list(zip(v["decode_requests"], sd)):
            tokens_of[int(r)].append(t)
        bat.commit(sp, sd, END)
    assert bat.free_pages() == cfg["num_pages"]
    for rid, (p, m) in requests.items():
        got, state = bat.result(rid)
        assert state == mod.REQ_FINISHED and list(got) == alone(p, m), f"request {rid} (preempted {bat.preemptions(rid)} times) under {cfg}"


def test_scheduler_abort_returns_the_admitted_requests_to_the_front_of_the_queue():
    """A failed launch must not leak pages or lose requests: after b200_batcher_abort the same plan comes out again."""
    mod = b200()
    bat = mod.Batcher(max_batch=3, num_pages=20, max_pages_per_seq=4, max_prefill_tokens=256)
    reqs = {}
    for n in (70, 10, 100):
        p = list(range(10, 10 + n))
        reqs[bat.submit(p, 5)] = (p, 5)
    plan1, v1 = bat.plan()
    assert plan1.n_prefill == 3 and bat.free_pages() < 20
    bat.abort()
    assert bat.free_pages() == 20 and bat.pending() == 3
    plan2, v2 = bat.plan()
    assert list(v2["prefill_requests"]) == list(v1["prefill_requests"]) and np.array_equal(v2["prefill_ids"], v1["prefill_ids"])
    bat.abort()
    drive(bat, reqs, mod)
    for rid, (p, m) in reqs.items():
        assert list(bat.result(rid)[0]) == alone(p, m)


def test_batcher_argument_errors():
    import ctypes as C

    mod = b200()
    bad = mod.BatcherConfig(0, 1, 1, 1)
    assert not mod.lib().b200_batcher_create(C.byref(bad))
    bat = mod.Batcher(2, 4, 2, 64)
    with pytest.raises(mod.B200Error):
        bat.commit([], [], END)  # nothing planned
    bat.submit([1, 2, 3], 2)
    bat.plan()
    with pytest.raises(mod.B200Error):
        bat.plan()  # the previous plan has not been committed


# ------------------------------------------------------------------------------------------------------------------ GPU
def scatter_to_pages(cache, num_pages, lens, rng, max_pages):
    """cache [L, B, Hkv, S, d] (numpy) -> (pool [L, num_pages, Hkv, 64, d], block_table [B, max_pages]) with the rows [0, lens[b]) of every
    batch row spread over randomly chosen, non-contiguous pages; the rest of the pool is NaN."""
    L, B, Hkv, S, d = cache.shape
    pool = np.full((L, num_pages, Hkv, PAGE, d), np.nan, np.float32)
    free = list(rng.permutation(num_pages))
    bt = np.full((B, max_pages), -1, np.int32)
    for b in range(B):
        for i in range((int(lens[b]) + PAGE - 1) // PAGE):
            pg = int(free.pop())
            bt[b, i] = pg
            n = min(PAGE, S - i * PAGE)
            pool[:, pg, :, :n] = cache[:, b, :, i * PAGE:i * PAGE + n]
    return pool, bt


def gather_from_pages(pool, bt, lens, S):
    L, P, Hkv, _, d = pool.shape
    B = bt.shape[0]
    out = np.zeros((L, B, Hkv, S, d), np.float32)
    for b in range(B):
        for i in range((int(lens[b]) + PAGE - 1) // PAGE):
            n = min(PAGE, S - i * PAGE)
            out[:, b, :, i * PAGE:i * PAGE + n] = pool[:, bt[b, i], :, :n]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("H,Hkv,steps", [(32, 32, [1, 64, 65, 1024]), (8, 1, [513, 2, 640, 129, 128]), (8, 2, [300, 300, 7])])
def test_decode_mha_paged_is_the_ragged_kernel_bit_for_bit(H, Hkv, steps, dtype):
    import torch

    from test_ops_gpu import _mha_case
    from util import rounded, to_dev, to_np

    mod = b200()
    d, L, layer = 128, 2, 1
    B, S = len(steps), (max(steps) + PAGE - 1) // PAGE * PAGE
    qkv, bias, kc, vc = _mha_case(B, H, Hkv, d, S, L, max(steps), layer, dtype, seed=51)
    rng = np.random.default_rng(52)
    num_pages, mp = 3 * B * (S // PAGE), S // PAGE + 2
    kpool, bt = scatter_to_pages(kc, num_pages, steps, rng, mp)
    vpool = np.full_like(kpool, np.nan)
    for b in range(B):
        for i in range((steps[b] + PAGE - 1) // PAGE):
            vpool[:, bt[b, i]] = vc[:, b, :, i * PAGE:(i + 1) * PAGE]
    sd = torch.tensor(steps, dtype=torch.int32, device="cuda")
    kcd, vcd = to_dev(kc, dtype), to_dev(vc, dtype)
    ref = mod.decode_mha(to_dev(qkv, dtype), to_dev(bias, dtype), kcd, vcd, H, Hkv, max(steps), layer, apply_rope=True, rot_dim=d, steps=sd)
    kpd, vpd = to_dev(kpool, dtype), to_dev(vpool, dtype)
    got = mod.decode_mha_paged(to_dev(qkv, dtype), to_dev(bias, dtype), kpd, vpd, to_dev(bt), sd, H, Hkv, layer, apply_rope=True, rot_dim=d)
    assert np.array_equal(to_np(got), to_np(ref)), "paged decode attention differs from the contiguous kernel"
    # the appended rows landed in the right pages, nothing else moved
    gk, gv = gather_from_pages(to_np(kpd), bt, steps, S), gather_from_pages(to_np(vpd), bt, steps, S)
    rk, rv = to_np(kcd), to_np(vcd)
    for b, s in enumerate(steps):
        assert np.array_equal(gk[:, b, :, :s], rk[:, b, :, :s]) and np.array_equal(gv[:, b, :, :s], rv[:, b, :, :s]), f"cache rows of batch row {b}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("B,H,Hkv,input_len,hist", [(2, 8, 2, [300, 129], [290, 0]), (1, 4, 4, [256], [0]), (3, 2, 1, [1, 128, 2], [0, 2, 127]),
                                                      (1, 2, 2, [1024], [60])])
def test_context_attention_paged_is_the_contiguous_kernel_bit_for_bit(B, H, Hkv, input_len, hist, dtype):
    from oracle import oracle
    from util import rounded, to_dev, to_np

    mod = b200()
    r = np.random.default_rng(77)
    d, L, layer = 128, 2, 1
    input_len, hist = np.array(input_len, np.int32), np.array(hist, np.int32)
    ctx = input_len + hist
    mq = int(input_len.max())
    S = (int(ctx.max()) + PAGE - 1) // PAGE * PAGE
    po, cum = oracle.cal_padding_offset(input_len, mq)
    T = int(cum[-1])
    q = np.zeros((B, H, mq, d), np.float32)
    for b in range(B):
        q[b, :, :input_len[b]] = r.standard_normal((H, input_len[b], d))
    q = rounded(q, dtype)
    kc = rounded(0.5 * r.standard_normal((L, B, Hkv, S, d)), dtype)
    vc = rounded(0.5 * r.standard_normal((L, B, Hkv, S, d)), dtype)
    for b in range(B):  # what lies beyond the context is never initialised by anybody
        kc[:, b, :, ctx[b]:] = np.nan
        vc[:, b, :, ctx[b]:] = np.nan
    kpool, bt = scatter_to_pages(kc, 4 * B * (S // PAGE), ctx, np.random.default_rng(5), S // PAGE + 1)
    vpool = np.full_like(kpool, np.nan)
    for b in range(B):
        for i in range((ctx[b] + PAGE - 1) // PAGE):
            vpool[:, bt[b, i]] = vc[:, b, :, i * PAGE:(i + 1) * PAGE]
    scale = 1.0 / np.sqrt(d)
    ref = mod.context_attention(to_dev(q, dtype), to_dev(kc, dtype), to_dev(vc, dtype), to_dev(po.reshape(-1)), to_dev(input_len), to_dev(ctx), layer,
                                T, scale)
    got = mod.context_attention_paged(to_dev(q, dtype), to_dev(kpool, dtype), to_dev(vpool, dtype), to_dev(bt), to_dev(input_len), to_dev(ctx),
                                      layer, T, scale)
    g = to_np(got)
    assert np.isfinite(g).all(), "NaN rows of the pool leaked"
    assert np.array_equal(g, to_np(ref).reshape(g.shape)), "paged context attention differs from the contiguous kernel"


def _paged_model(dtype, seed=5, bias=True):
    from test_decoder_engine import make_model
    from test_generate import CFG, tail_weights

    cfg = dict(CFG, max_seq=256)
    return cfg, make_model(cfg, seed=seed, bias=bias), tail_weights(6, dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_engine_prefill_and_step_paged_equal_the_contiguous_engine_bit_for_bit(dtype):
    """b200_decoder_prefill_paged + b200_decoder_step_paged against b200_decoder_prefill + b200_decoder_step_ragged on the same ragged
    batch: same kernels, same batch composition -> identical hidden states; and the pool holds exactly the contiguous cache's rows."""
    import torch

    from test_decoder_engine import build_decoder
    from util import rounded, to_dev, to_np

    cfg, model, _ = _paged_model(dtype)
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16}[dtype]
    lens = np.array([150, 37, 129], np.int32)
    B, T, S, L, Hkv, d = len(lens), int(lens.sum()), cfg["max_seq"], cfg["layers"], cfg["kv_head_num"], cfg["head_size"]
    rng = np.random.default_rng(8)
    x = rounded(rng.standard_normal((T, cfg["hidden"])), dtype)
    zero = torch.zeros(B, dtype=torch.int32, device="cuda")
    # contiguous
    dec = build_decoder(model, cfg, dtype, B)
    xd = to_dev(x, dtype)
    kc = torch.zeros((L, B, Hkv, S, d), dtype=tdt, device="cuda")
    vc = torch.zeros_like(kc)
    dec.prefill(xd, kc, vc, to_dev(lens), zero, to_dev(lens), int(lens.max()))
    # paged: pages handed out in a scrambled order
    num_pages, mp = 16, S // PAGE
    perm = rng.permutation(num_pages)
    bt = np.full((B, mp), -1, np.int32)
    k = 0
    for b in range(B):
        for i in range((int(lens[b]) + 1 + PAGE - 1) // PAGE):
            bt[b, i] = perm[k]
            k += 1
    dec2 = build_decoder(model, cfg, dtype, B)
    xp = to_dev(x, dtype)
    kp = torch.full((L, num_pages, Hkv, PAGE, d), float("nan"), dtype=tdt, device="cuda")
    vp = torch.full_like(kp, float("nan"))
    btd = to_dev(bt)
    dec2.prefill_paged(xp, kp, vp, btd, to_dev(lens), zero, to_dev(lens), int(lens.max()))
    assert np.array_equal(to_np(xp), to_np(xd)), "prefill output"
    gk = gather_from_pages(to_np(kp), bt, lens, S)
    for b in range(B):
        assert np.array_equal(gk[:, b, :, :lens[b]], to_np(kc)[:, b, :, :lens[b]]), f"K rows of sequence {b}"
    # two decode steps on top
    h1 = to_dev(rounded(rng.standard_normal((B, cfg["hidden"])), dtype), dtype)
    h2 = h1.clone()
    for i in range(2):
        steps = to_dev((lens + 1 + i).astype(np.int32))
        dec.step_ragged(h1, kc, vc, steps, int(lens.max()) + 1 + i)
        dec2.step_paged(h2, kp, vp, btd, steps, int(lens.max()) + 1 + i)
        assert np.array_equal(to_np(h1), to_np(h2)), f"decode step {i}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_batcher_with_everybody_admitted_at_once_is_generate_ragged_bit_for_bit(dtype):
    """All requests fit the first iteration and run the same number of tokens: the batcher's iterations are then exactly the launches of
    b200_generate_ragged (packed prefill, ragged decode steps of the same batch) on a paged instead of a contiguous cache."""
    import torch

    from test_decoder_engine import build_decoder
    from test_generate import V
    from util import to_dev

    mod = b200()
    cfg, model, (emb, gamma, lm) = _paged_model(dtype)
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16}[dtype]
    lens, N = [7, 3, 70, 1], 9
    B, L, Hkv, d, S = len(lens), cfg["layers"], cfg["kv_head_num"], cfg["head_size"], cfg["max_seq"]
    rng = np.random.default_rng(21)
    prompts = [rng.integers(3, V, size=n).astype(np.int32) for n in lens]
    embd, gd, lmd = to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype)
    dec = build_decoder(model, cfg, dtype, B)
    padded = np.zeros((B, max(lens)), np.int32)
    for b, p in enumerate(prompts):
        padded[b, :len(p)] = p
    kc = torch.zeros((L, B, Hkv, S, d), dtype=tdt, device="cuda")
    ids, _ = dec.generate(padded, embd, gd, lmd, kc, torch.zeros_like(kc), N, top_k=1, end_id=-1, prompt_lens=lens)

    dec2 = build_decoder(model, cfg, dtype, B)
    bat = mod.Batcher(max_batch=B, num_pages=12, max_pages_per_seq=S // PAGE, max_prefill_tokens=160)  # 4 x 70 padded rows <= 2 x 160
    rids = [bat.submit(p, N) for p in prompts]
    kp = torch.full((L, 12, Hkv, PAGE, d), float("nan"), dtype=tdt, device="cuda")
    vp = torch.full_like(kp, float("nan"))
    it = 0
    while bat.pending():
        bat.step(dec2, embd, gd, lmd, kp, vp, top_k=1, end_id=-1)
        it += 1
        assert it < 100
    assert it == N, "one prefill iteration + N - 1 decode iterations"
    for b, rid in enumerate(rids):
        got, state = bat.result(rid)
        assert state == mod.REQ_FINISHED and np.array_equal(got, ids[b]), f"request {rid}: {got} vs {ids[b]}"
    assert bat.free_pages() == 12


@pytest.mark.gpu
def test_batcher_rejects_a_prompt_outside_the_vocabulary_and_serves_the_rest():
    import torch

    from test_decoder_engine import build_decoder
    from test_generate import V
    from util import to_dev

    mod = b200()
    dtype = "bf16"
    cfg, model, (emb, gamma, lm) = _paged_model(dtype)
    L, Hkv, d, S = cfg["layers"], cfg["kv_head_num"], cfg["head_size"], cfg["max_seq"]
    dec = build_decoder(model, cfg, dtype, 2)
    bat = mod.Batcher(max_batch=2, num_pages=6, max_pages_per_seq=S // PAGE, max_prefill_tokens=128)
    good = bat.submit([5, 6, 7, 8], 4)
    bad = bat.submit([5, V + 3, 7], 4)  # would read past the embedding table
    good2 = bat.submit([9, 10], 3)
    kp = torch.full((L, 6, Hkv, PAGE, d), float("nan"), dtype=torch.bfloat16, device="cuda")
    vp = torch.full_like(kp, float("nan"))
    it = 0
    while bat.pending():
        bat.step(dec, to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype), kp, vp, top_k=1, end_id=-1)
        it += 1
        assert it < 50
    assert bat.result(bad)[1] == mod.REQ_REJECTED and len(bat.result(bad)[0]) == 0
    assert bat.result(good)[1] == mod.REQ_FINISHED and len(bat.result(good)[0]) == 4
    assert bat.result(good2)[1] == mod.REQ_FINISHED and len(bat.result(good2)[0]) == 3
    assert bat.free_pages() == 6


@pytest.mark.gpu
def test_batcher_stream_of_requests_matches_the_oracle_up_to_near_ties():
    """More requests than batch slots, a pool small enough to force preemptions: requests join and leave between iterations.  Every
    request's greedy ids must equal the fp32 oracle's for the same prompt generated alone -- until the first step where the oracle's own
    top two logits are closer than the 16-bit engine can resolve (there both continuations are valid and the comparison stops)."""
    import torch

    from test_decoder_engine import build_decoder
    from test_generate import V, oracle_generate
    from util import to_dev

    mod = b200()
    dtype = "f16"
    cfg, model, (emb, gamma, lm) = _paged_model(dtype, bias=False)
    import test_generate

    old_cfg = test_generate.CFG
    test_generate.CFG = cfg  # oracle_generate reads the module's configuration (cache length 256 here)
    try:
        L, Hkv, d, S = cfg["layers"], cfg["kv_head_num"], cfg["head_size"], cfg["max_seq"]
        rng = np.random.default_rng(33)
        reqs = [(rng.integers(3, V, size=int(n)).astype(np.int32), int(m)) for n, m in [(60, 12), (5, 20), (62, 10), (17, 6), (63, 9), (1, 15), (40, 8)]]
        dec = build_decoder(model, cfg, dtype, 3)
        bat = mod.Batcher(max_batch=3, num_pages=4, max_pages_per_seq=S // PAGE, max_prefill_tokens=128)
        rids = [bat.submit(p, m) for p, m in reqs]
        kp = torch.full((L, 4, Hkv, PAGE, d), float("nan"), dtype=torch.float16, device="cuda")
        vp = torch.full_like(kp, float("nan"))
        embd, gd, lmd = to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype)
        it = 0
        while bat.pending():
            bat.step(dec, embd, gd, lmd, kp, vp, top_k=1, end_id=2)
            it += 1
            assert it < 500
        compared = 0
        for rid, (p, m) in zip(rids, reqs):
            got, state = bat.result(rid)
            assert state == mod.REQ_FINISHED
            ref, logits_all = oracle_generate(model, emb, gamma, lm, p.reshape(1, -1), m)
            ref = ref[0]
            for i in range(len(got)):
                srt = np.sort(logits_all[i][0])
                near_tie = (srt[-1] - srt[-2]) < 2e-2 * max(1.0, abs(srt[-1]))
                if got[i] != ref[i]:
                    assert near_tie, f"request {rid} token {i}: {got[i]} vs oracle {ref[i]} with a clear margin {srt[-1] - srt[-2]:.3f}"
                    break
                compared += 1
                if ref[i] == 2:
                    assert i == len(got) - 1, "generation continued past end_id"
                    break
            else:
                assert len(got) == m or got[-1] == 2
        assert compared >= 30, f"only {compared} tokens compared: pick another seed"
        assert sum(bat.preemptions(r) for r in rids) > 0, "the pool was meant to be too small: no preemption happened"
        assert bat.free_pages() == 4
    finally:
        test_generate.CFG = old_cfg
