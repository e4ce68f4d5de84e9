"""Property tests (hypothesis) of the CPU checker and of the host-side sharding / packing algebra: independent numpy statements of
what each op means, on random shapes -- the second line of defence behind the reference's own golden vectors (tests/test_oracle_golden.py)."""
import importlib

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import oracle

tp = importlib.import_module("llm-inference-engine_b200.tp")
W = importlib.import_module("llm-inference-engine_b200.weights")
SET = settings(max_examples=25, deadline=None, derandomize=True, database=None)  # same examples on every run


@SET
@given(st.lists(st.integers(1, 9), min_size=1, max_size=6))
def test_padding_offset_counts_the_padding_in_front_of_each_token(lens):
    """cal_padding_offset.cuh:9-15: one entry per UN-PADDED token, sequences back to back; the offset of a token = number of padding
    slots in front of it in the padded [batch, max_q_len] grid (lens [4,3,5] -> [0 x 4, 1 x 3, 3 x 5])."""
    lens = np.asarray(lens, np.int32)
    mq = int(lens.max())
    off, cum = oracle.cal_padding_offset(lens, mq)
    assert np.array_equal(cum, np.concatenate([[0], np.cumsum(lens)]))
    flat, pad, t = off.reshape(-1), 0, 0
    for n in lens:
        assert (flat[t:t + n] == pad).all()
        t += int(n)
        pad += mq - int(n)


@SET
@given(st.integers(1, 4), st.integers(6, 300), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_topk_is_the_descending_prefix_with_ties_to_the_lower_id(rows, vocab, k, seed):
    """topk.cuh:28-41: value descending, equal values -> lower id first (what the engine's integer output is compared with)."""
    k = min(k, vocab)
    rng = np.random.default_rng(seed)
    logits = rng.integers(-5, 6, size=(rows, vocab)).astype(np.float32)  # many ties
    ids, vals = oracle.topk(logits, k)
    for r in range(rows):
        order = sorted(range(vocab), key=lambda i: (-logits[r, i], i))[:k]
        assert list(ids[r]) == order
        assert np.array_equal(vals[r], logits[r, order])


@SET
@given(st.integers(1, 5), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_sampling_is_the_inverse_cdf_over_the_k_candidates(batch, k, seed):
    """sampling.cu:31-69: w_i = exp(v_i - v_0); thr = u * sum(w); first i with (thr -= w_i) < 0, else candidate 0; id %= vocab."""
    rng = np.random.default_rng(seed)
    vocab, end_id = 50, 2
    ids = rng.integers(0, 200, size=(batch, k)).astype(np.int32)
    vals = -np.sort(-rng.standard_normal((batch, k)).astype(np.float32), axis=1)
    u = rng.random(batch).astype(np.float32)
    seq, fin = np.full(batch, 7, np.int32), rng.integers(0, 2, batch).astype(bool)
    seq0, fin0 = seq.copy(), fin.copy()
    out = oracle.sampling(ids, vals.copy(), seq, fin, u, end_id, vocab)
    for b in range(batch):
        w = np.exp(vals[b] - vals[b, 0]).astype(np.float32)
        thr = np.float32(u[b]) * np.float32(w.sum(dtype=np.float32))
        pick = ids[b, 0] % vocab
        for i in range(k):
            thr = np.float32(thr - w[i])
            if thr < 0:
                pick = ids[b, i] % vocab
                break
        # the float32 running sum may differ in the last bit from the C loop: accept the neighbour only when thr is within an ulp of 0
        assert out[b] == pick or abs(float(thr)) < 1e-6
        assert seq[b] == seq0[b] + (0 if fin0[b] else 1)
        assert fin[b] == (out[b] == end_id)


@SET
@given(st.sampled_from([1, 2, 4]), st.integers(1, 3), st.integers(1, 3), st.integers(0, 2 ** 31 - 1))
def test_tensor_parallel_shards_tile_the_layer(world, hmul, imul, seed):
    """tp.py: the ranks' QKV rows / gate-up rows partition the full tensors, their O / down columns tile K, and the sharded computation of
    a row-sharded linear sums to the un-sharded one."""
    rng = np.random.default_rng(seed)
    d, H, Hkv, I, h = 8, 4 * hmul, 4, 8 * imul, 16
    cfg = dict(head_num=H, kv_head_num=Hkv, head_size=d, inter=I)
    w = dict(g1=rng.standard_normal(h), g2=rng.standard_normal(h), wqkv=rng.standard_normal(((H + 2 * Hkv) * d, h)), bqkv=rng.standard_normal((H + 2 * Hkv) * d),
             wo=rng.standard_normal((h, H * d)), bo=rng.standard_normal(h), wgu=rng.standard_normal((2 * I, h)), wd=rng.standard_normal((h, I)))
    shards = [tp.shard_layer(w, cfg, r, world) for r in range(world)]
    rows = np.concatenate([s["wqkv"] for s in shards])
    assert sorted(map(tuple, rows)) == sorted(map(tuple, w["wqkv"]))  # every row exactly once
    assert np.array_equal(np.concatenate([s["wo"] for s in shards], axis=1), w["wo"])
    assert np.array_equal(np.concatenate([s["wd"] for s in shards], axis=1), w["wd"])
    x = rng.standard_normal((3, H * d))
    parts = [x[:, r * (H * d // world):(r + 1) * (H * d // world)] @ shards[r]["wo"].T for r in range(world)]
    assert np.allclose(sum(parts), x @ w["wo"].T)
    # gate / up stay paired per rank: SwiGLU of the shard = the shard of SwiGLU
    a = rng.standard_normal((2, h))
    full = a @ w["wgu"].T
    act = full[:, :I] / (1 + np.exp(-full[:, :I])) * full[:, I:]
    Il = I // world
    for r in range(world):
        loc = a @ shards[r]["wgu"].T
        assert np.allclose(loc[:, :Il] / (1 + np.exp(-loc[:, :Il])) * loc[:, Il:], act[:, r * Il:(r + 1) * Il])


@SET
@given(st.integers(1, 6), st.integers(1, 4), st.integers(0, 2 ** 31 - 1))
def test_quantisers_round_trip_within_their_step(n, kblocks, seed):
    """FP8-e4m3 per-row scale: relative error <= 2^-4 (3 mantissa bits, round to nearest).  INT4 group-128 scale + zero point: within one
    step everywhere (the rounded zero point shifts the grid by up to half a step, so the two ends may clamp) and within half a step for
    values at least one step inside the group's range."""
    rng = np.random.default_rng(seed)
    k = 128 * kblocks
    w = rng.standard_normal((n, k)).astype(np.float32)
    q, sc = oracle.quantize_fp8(w)
    back = oracle.dequantize_fp8(q, sc)
    assert np.all(np.abs(back - w) <= np.abs(w) * 2.0 ** -4 * 1.001 + sc[:, None] * 2.0 ** -9 + 1e-7)  # denormal step 2^-9 * scale
    q4, s4, z4 = oracle.quantize_int4(w, 128)
    back4 = oracle.dequantize_int4(q4, s4, z4, 128)
    step = np.repeat(s4.astype(np.float32), 128, axis=1)
    err = np.abs(back4 - w)
    assert np.all(err <= step * 1.01 + 1e-6)
    grp = w.reshape(n, kblocks, 128)
    lo, hi = np.repeat(grp.min(axis=2), 128, axis=1), np.repeat(grp.max(axis=2), 128, axis=1)
    inside = (w >= lo + step) & (w <= hi - step)
    assert np.all(err[inside] <= 0.5 * step[inside] * 1.01 + 1e-6)


@SET
@given(st.sampled_from([1, 2, 4, 8]), st.integers(1, 3), st.integers(1, 6), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_vocab_sharded_topk_merge_equals_the_global_topk(world, rows, k, per_rank, seed):
    """tp.merge_topk (vocab-sharded LM head): local top-k per vocabulary shard + merge == top-k of the whole row, ids and
    values bit for bit, with heavily tied logits (the tie rule 'lower id first' must survive the cut into shards)."""
    k = min(k, per_rank)
    vocab = world * per_rank
    rng = np.random.default_rng(seed)
    logits = rng.integers(-3, 4, size=(rows, vocab)).astype(np.float32)
    want_ids, want_vals = oracle.topk(logits, k)
    vals, ids = [], []
    for r in range(world):
        lo, hi = tp.vocab_range(vocab, r, world)
        i, v = oracle.topk(np.ascontiguousarray(logits[:, lo:hi]), k)
        ids.append(i.astype(np.int64) + lo)
        vals.append(v)
    got_ids, got_vals = tp.merge_topk(np.stack(vals), np.stack(ids), k)
    assert np.array_equal(got_ids, want_ids) and np.array_equal(got_vals, want_vals)
