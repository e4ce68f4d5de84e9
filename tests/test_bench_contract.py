"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs on host cores only and prints one JSON line with the
keys the driver reads; the algorithmic-byte model of the roofline reproduces SURVEY.md 8(d)'s figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode tokens/s" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3 * d["config"]["batch"]) < 1e-3 * d["ms_per_step"] * d["value"]
    assert d["config"]["workload"].startswith("Llama-2-7B 32-layer bf16 decode, batch 1, 1024-token context")
    cpu = d["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] == d["value"] and cpu["sample"]
    if cpu["kind"] == "reference":  # the reference's own unit-test loops were loadable: the all-cores port is reported beside them
        assert cpu["port_all_cores"]["kind"] == "port" and cpu["port_all_cores"]["value"] > 0
        assert cpu["single_thread"]["cores"] == 1 and cpu["single_thread"]["kind"] == "reference"
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_algorithmic_bytes_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench

    step_bytes, weight_bytes = bench.algorithmic_bytes(bench.CONFIGS["7b"], 1, 1024, "bf16")
    assert int(weight_bytes) == 13_214_154_752 and int(step_bytes) == 13_751_025_664  # weights + 536,870,912 B of KV per step
    # 70B-shaped, tensor-parallel over 8 ranks, batch 8 (configs[4]): 17.113 GB of layer weights per rank + the replicated bf16 LM head + KV
    step70, w70 = bench.algorithmic_bytes(bench.CONFIGS["70b"], 8, 1024, "bf16", tp=8)
    assert int(w70) == 80 * (855_638_016 // 8) * 2 + 32000 * 8192 * 2
    assert int(step70 - w70) == 80 * 8 * 2 * 1 * 128 * 1024 * 2
