"""The shim's data carriers (llm-inference-engine_b200/shim/src/utils/{tensor,string_utils,macro}.h) against the reference's
own headers (src/utils/tensor.h:112-295, string_utils.h:10-36): the scripted driver tests/tools/tensor_driver.cpp is compiled against each
and what it prints -- sizes, host reads, refusals and their messages, TensorMap insert / overwrite / validation rules, vec2str's doubled last
element -- must be identical line for line.  Fixture tests/golden/tensor_golden.txt is the reference build's output (oracle/Makefile ref_tensor)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "llm-inference-engine_b200", "shim", "_own_programs", "tensor_shim")
REF = os.path.join(ROOT, "oracle", "_ref", "tensor_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden", "tensor_golden.txt")


def run(exe):
    return subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True, timeout=60).stdout.decode().splitlines()


def test_shim_carriers_reproduce_the_reference_golden():
    if not os.path.exists(SHIM):
        pytest.skip("shim/_own_programs/tensor_shim not built (run __graft_entry__.build())")
    want = open(GOLDEN).read().splitlines()
    got = run(SHIM)
    assert len(want) >= 35
    assert got == want


def test_golden_is_what_the_reference_headers_print():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/tensor_ref only exists where /root/reference does")
    assert run(REF) == open(GOLDEN).read().splitlines()


DEBUG = os.path.join(ROOT, "llm-inference-engine_b200", "shim", "_own_programs", "debug_utils_shim")


def test_save_tensor_and_print_tensor(tmp_path):
    """saveTensor / print_tensor of the shim (reference: src/utils/debug_utils.h:17-119, output_utils.h:9-33) on host tensors: file names,
    the first-three-layers rule, byte counts per rank, and the two printed lines per tensor."""
    import numpy as np

    if not os.path.exists(DEBUG):
        pytest.skip("shim/_own_programs/debug_utils_shim not built (run __graft_entry__.build())")
    env = dict(os.environ, LLM_SAVE_TENSOR_DIR=str(tmp_path))
    out = subprocess.run([DEBUG], stdout=subprocess.PIPE, check=True, timeout=60, env=env).stdout.decode().splitlines()
    assert sorted(os.listdir(tmp_path)) == ["0_rank3.bin", "1_rank4.bin", "rank1.bin", "rank2.bin"]
    want = (0.25 * np.arange(24)).astype(np.float32)
    for name in ("rank2.bin", "0_rank3.bin", "1_rank4.bin"):
        assert np.array_equal(np.fromfile(tmp_path / name, dtype=np.float32), want)
    assert os.path.getsize(tmp_path / "rank1.bin") == 0
    assert out == ["Saving intermediate tensor in rank2.bin", "Saving intermediate tensor in rank3.bin", "Saving intermediate tensor in rank4.bin",
                   "Saving intermediate tensor in rank1.bin",
                   "number of dimensions: 3", "2 3 4 ", "number of dimensions: 4", "1 2 3 4 ", "number of dimensions: 2", "11008 4096 "]
