#!/usr/bin/env python
"""Which kernel, how many passes over the weights and which ring geometry each linear of a decode step gets -- a Python MODEL of the host-side
dispatch (no GPU): decoder.cu norm_linear / plain_linear (fused GEMV for M <= gemv_max_rows, un-fused otherwise), linear.cu:300-330
(tcgen05 GEMM for dense M > 4, else GEMV passes of 8 / 4 rows), gemv_inst.cuh:32-60 (gemv_nk geometry) and gemv_q.cuh launch_gemv_q_inst.
It restates those few formulas; when they change, this file must follow (tests/test_dispatch_plan.py pins the cases measured on the B200).
usage: python scripts/dispatch_plan.py > profiles/<tag>_dispatch_plan.txt"""
BUDGET = 224 * 1024
GROUPS, GW, MAX_STAGES, ROWS, PIECE = 2, 8, 8, 2, 8192
Q_ROWS, Q_TOK = 16, 8
MMA_WARPS, MMA_ROWS = 16, 8
WARPS = GROUPS * GW


def gemv_nk_plan(M, K, fmt, ebytes=2):
    """gemv_inst.cuh launch_gemv_inst: None when unsupported, else dict(pieces, piece_bytes, stages)."""
    if M > 4:
        return None
    MB = 1 if M <= 1 else (2 if M <= 2 else 4)
    row_bytes = {"dense": K * ebytes, "fp8": K, "int4": K // 2}[fmt]
    pieces = -(-row_bytes // PIECE)
    piece_bytes = row_bytes if pieces == 1 else (-(-row_bytes // pieces) + 511) // 512 * 512
    pieces = -(-row_bytes // piece_bytes)
    stage_bytes = ROWS * ((piece_bytes + 127) // 128 * 128)
    block = {"dense": 1, "fp8": 512, "int4": 1024}[fmt]
    xs_bytes = ebytes if fmt == "dense" else 4
    Kp = -(-K // block) * block
    fixed = (MB * Kp * xs_bytes + 127) // 128 * 128 + GROUPS * (2 * MAX_STAGES + 4) * 8 + GROUPS * GW * 2 * ROWS * MB * 32 * 4
    per_stage = GROUPS * stage_bytes
    if fixed + 3 * per_stage > BUDGET:
        return None
    return dict(kernel="gemv_nk", pieces=pieces, piece_bytes=piece_bytes, stages=min((BUDGET - fixed) // per_stage, MAX_STAGES))


def gemv_q_plan(M, K, fmt):
    """gemv_q.cuh launch_gemv_q_t / launch_gemv_q_inst (16-bit activations): None when unsupported."""
    if fmt not in ("fp8", "int4") or M < 1 or M > Q_TOK or K % 128:
        return None
    xs_stride = K + (16 if fmt == "fp8" else 32)
    fixed = (M * xs_stride * 2 + 127) // 128 * 128 + GROUPS * (2 * MAX_STAGES + 4) * 8 + WARPS * 2 * Q_ROWS * Q_TOK * 4
    for piece in (2048, 1024):
        per_stage = GROUPS * Q_ROWS * (piece + 16)
        if fixed + 3 * per_stage <= BUDGET:
            k_per_piece = piece * (1 if fmt == "fp8" else 2)
            return dict(kernel="gemv_q", pieces=-(-K // k_per_piece), piece_bytes=piece, stages=min((BUDGET - fixed) // per_stage, MAX_STAGES))
    return None


def gemv_mma_plan(M, K, fmt, max_units=8):
    """gemv_mma.cuh gemv_mma_geometry (16-bit activations, 1..16 tokens): None when unsupported."""
    if M < 1 or M > 16 or K % 128:
        return None
    dense = fmt == "dense"
    tok = 8 if M <= 8 else 16
    xs_rows = tok if M == tok else M + 1
    k_round = -(-K // 512) * 512
    best, best_score = None, -1
    piece_bytes0 = 2048 if fmt == "int4" else 4096
    while piece_bytes0 >= 1024:
        piece_bytes = piece_bytes0
        piece_bytes0 //= 2
        piece_k = piece_bytes // 2 if dense else (piece_bytes if fmt == "fp8" else piece_bytes * 2)
        if piece_k > k_round:
            piece_k = k_round
            piece_bytes = piece_k * 2 if dense else (piece_k if fmt == "fp8" else piece_k // 2)
        if fmt == "int4" and (piece_k // MMA_WARPS) % 128:
            continue
        if fmt == "int4" and piece_k // MMA_WARPS > 256:
            continue
        rg = 2 if fmt == "fp8" else 1  # row groups of 8 weight rows per unit
        row_stride = piece_bytes + (64 if dense else (32 if fmt == "fp8" else 16))
        stage_bytes = MMA_ROWS * rg * row_stride
        pieces_total = -(-K // piece_k)
        for parts in range(1, pieces_total + 1):
            ppp = -(-pieces_total // parts)
            if -(-pieces_total // ppp) != parts:
                continue
            part_k = ppp * piece_k
            tile = 16 * MMA_ROWS * rg * 4
            fixed = (xs_rows * (part_k + 32) * 2 + 127) // 128 * 128 + (2 * MAX_STAGES + 6) * 8 + 2 * MMA_WARPS * tile + (max_units * tile if parts > 1 else 0)
            if fixed + 3 * stage_bytes > 226 * 1024:
                continue
            stages = min((226 * 1024 - fixed) // stage_bytes, MAX_STAGES)
            score = min(stages * stage_bytes, 196608) + (49152 if piece_bytes >= 4096 else (0 if piece_bytes >= 2048 else -49152)) - (49152 if M <= 8 else 8192) * (parts - 1)
            if best is None or score > best_score:
                best_score = score
                best = dict(kernel="gemv_mma", pieces=pieces_total, piece_bytes=piece_bytes, stages=stages, parts=parts, part_k=part_k)
    return best


def gemv_any(M, K, fmt):
    """gemv_f32.cu launch_gemv_nk for a 16-bit model: the tensor-core GEMV for 2..16 dense tokens and 5..16 quantised ones, the round-1
    quantised kernel for 1..4 quantised tokens (fused prologue), the SIMT kernel for one dense token."""
    if 2 <= M <= 16 and (fmt == "dense" or M > 4):
        g = gemv_mma_plan(M, K, fmt)
        if g:
            return g
    if fmt != "dense" and M <= 8:
        g = gemv_q_plan(M, K, fmt)
        if g:
            return g
    return gemv_nk_plan(M, K, fmt)


def linear_plan(M, K, fmt, fused_rows):
    """One linear of the decode step: (path, passes over the weights, geometry)."""
    if M <= fused_rows:
        g = gemv_any(M, K, fmt)
        if g:  # gemv_nk / gemv_q run the norm / residual / TP-reduce prologue themselves; gemv_mma takes plain activations (norm_kernel in front)
            return (("norm_kernel + " if g["kernel"] == "gemv_mma" else "fused ") + g["kernel"], 1, g)
    if M <= 16:  # linear.cu: one GEMV pass for decode batches
        g = gemv_any(M, K, fmt)
        if g:
            return ("un-fused " + g["kernel"], 1, g)
    if fmt == "dense" and M > 4:
        return ("un-fused gemm_tc (tcgen05, swap-AB stream-K)" if M <= 128 else "un-fused gemm_tc", 1, None)
    for step in (16, 8, 4):  # linear.cu: quantised weights up to 64 tokens in passes
        if M > 4 * step or M > 64:
            continue
        plans = [gemv_any(min(step, M - m0), K, fmt) for m0 in range(0, M, step)]
        if all(plans):
            return (f"un-fused {plans[0]['kernel']} x {len(plans)} passes of <= {step} tokens", len(plans), plans[0])
    return ("un-fused generic SIMT fallback (launch_simt, linear.cu)", 1, None)


def step_plan(name, hidden, heads, kv_heads, d, inter, tp, batch, fmt):
    fused_rows = 16  # decoder.cu gemv_max_rows (16-bit models)
    shapes = [("qkv", hidden, (heads + 2 * kv_heads) * d // tp), ("o", heads * d // tp, hidden), ("gate_up", hidden, 2 * inter // tp),
              ("down", inter // tp, hidden)]
    lines = [f"{name}, batch {batch}, weights {fmt}" + (f", TP-{tp} (one rank)" if tp > 1 else "")]
    for lin, K, N in shapes:
        path, passes, g = linear_plan(batch, K, fmt, fused_rows)
        geo = (f"pieces {g['pieces']} x {g['piece_bytes']} B, {g['stages']} stages" + (f", {g['parts']} activation parts of {g['part_k']} k" if g.get("parts") else "")) if g else "-"
        wbytes = K * N * {"dense": 2, "fp8": 1, "int4": 0.5}[fmt]
        lines.append(f"  {lin:8s} K={K:6d} N={N:6d}  {path:58s} weights read {passes} x {wbytes / 1e6:7.1f} MB   {geo}")
    return lines


CASES = [("Llama-2-7B", 4096, 32, 32, 128, 11008, 1, b, f) for f in ("dense", "fp8", "int4") for b in (1, 4, 8, 16, 32)] + \
        [("Llama-2-70B-shaped", 8192, 64, 8, 128, 28672, 8, b, "dense") for b in (1, 8)]

if __name__ == "__main__":
    print(__doc__.split("usage:")[0].strip() + "\n")
    for c in CASES:
        print("\n".join(step_plan(*c)) + "\n")
