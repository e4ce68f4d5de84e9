"""scripts/dispatch_plan.py restates the host-side kernel selection (decoder.cu / linear.cu / gemv_inst.cuh / gemv_q.cuh) so that the plan of a
decode step can be read without a GPU.  These cases pin it to geometries that were observed on the B200 (six 16 KB stages per
group for the bf16 GEMV at batch 1; the 3-stage ring of the quantised kernel; 8 tokens of a K = 11008 row do not fit) so the model cannot
drift silently from the numbers the design document quotes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import dispatch_plan as dp  # noqa: E402


def test_dense_batch1_geometry_matches_the_design_document():
    g = dp.gemv_nk_plan(1, 4096, "dense")
    assert g == dict(kernel="gemv_nk", pieces=1, piece_bytes=8192, stages=6)  # 6 stages x 16 KB x 2 groups = 192 KB in flight per SM
    g = dp.gemv_nk_plan(1, 11008, "dense")
    assert g["pieces"] == 3 and g["piece_bytes"] % 512 == 0 and g["pieces"] * g["piece_bytes"] >= 11008 * 2


def test_quantised_round1_kernel_geometry_is_still_modelled():
    assert dp.gemv_q_plan(1, 4096, "int4") == dict(kernel="gemv_q", pieces=1, piece_bytes=2048, stages=3)
    assert dp.gemv_q_plan(8, 11008, "fp8") is None and dp.gemv_q_plan(8, 11008, "int4") is None  # 176 KB of activations: no ring fits


def test_tensor_core_gemv_reads_the_weights_once_up_to_16_tokens():
    """gemv_mma.cuh: activations staged in K parts of 64 KiB, so K = 8192 and K = 11008 keep a >= 4-stage ring at 8 and at 16 tokens; every
    decode batch up to 16 is ONE pass over the weights, in every format (round 1: 2 passes at batch 16, 4 for the K = 11008 down projection)."""
    for m, k in ((2, 4096), (8, 4096), (16, 4096), (8, 8192), (8, 11008), (16, 11008)):  # incl. the 70B-shaped hidden size and the 7B down projection
        g = dp.gemv_mma_plan(m, k, "dense")
        assert g and g["stages"] >= 3 and g["parts"] * g["part_k"] >= k, (m, k, g)
        assert (m + 1) * (g["part_k"] + 32) * 2 + g["stages"] * 8 * (g["piece_bytes"] + 64) <= 226 * 1024  # the staged part and the ring share the SM
    assert dp.gemv_mma_plan(8, 4096, "dense")["parts"] == 1 and dp.gemv_mma_plan(16, 4096, "dense")["parts"] == 2
    assert dp.gemv_mma_plan(8, 8192, "dense")["parts"] == 2  # 8 tokens x 8192 k do not fit beside a ring: two parts of 4096
    for fmt in ("dense", "fp8", "int4"):
        for m in (2, 8, 16):
            for k in (4096, 11008):
                path, passes, geo = dp.linear_plan(m, k, fmt, fused_rows=16)
                want = "fused gemv_q" if (fmt != "dense" and m <= 4) else "norm_kernel + gemv_mma"
                assert passes == 1 and path == want, (fmt, m, k, path)
    assert dp.linear_plan(1, 4096, "dense", fused_rows=16)[0] == "fused gemv_nk"  # a single dense token stays on the SIMT kernel
    assert dp.linear_plan(1, 4096, "int4", fused_rows=16)[0] == "fused gemv_q"  # up to 4 quantised tokens: the kernel with the fused prologue
    path, passes, _ = dp.linear_plan(32, 4096, "fp8", fused_rows=16)
    assert passes == 2 and "16 tokens" in path


def test_dense_batches_beyond_16_go_to_the_tensor_core_gemm():
    for m in (17, 32, 128):
        assert dp.linear_plan(m, 4096, "dense", fused_rows=16)[0].startswith("un-fused gemm_tc")
