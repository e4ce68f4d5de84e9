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
