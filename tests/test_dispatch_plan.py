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


def test_quantised_geometry_and_the_down_projection_limit():
    assert dp.gemv_q_plan(1, 4096, "int4") == dict(kernel="gemv_q", pieces=1, piece_bytes=2048, stages=3)
    assert dp.gemv_q_plan(8, 4096, "fp8")["piece_bytes"] == 1024            # eight staged tokens leave room for 1-KiB pieces only
    assert dp.gemv_q_plan(8, 11008, "fp8") is None and dp.gemv_q_plan(8, 11008, "int4") is None  # 176 KB of activations: no ring fits
    assert dp.gemv_q_plan(4, 11008, "fp8")["stages"] >= 3
    path, passes, _ = dp.linear_plan(16, 11008, "fp8", fused_rows=8)
    assert passes == 4 and "4 tokens" in path                               # batch 16: the down weights are streamed four times
    path, passes, _ = dp.linear_plan(16, 4096, "fp8", fused_rows=8)
    assert passes == 2


def test_dense_batched_decode_goes_to_the_tensor_core_gemm():
    for m in (5, 8, 32, 128):
        assert dp.linear_plan(m, 4096, "dense", fused_rows=4)[0].startswith("un-fused gemm_tc")
    assert dp.linear_plan(4, 4096, "dense", fused_rows=4)[0] == "fused gemv_nk"
