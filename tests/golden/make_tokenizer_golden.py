"""Generate tests/golden/tokenizer_golden.json: a synthetic vocabulary (in the file format the reference reads, src/models/tokenizer.h:138-166),
a list of texts / id lists, and what the REFERENCE'S OWN tokenizer header answers (oracle/_ref/tokenizer_ref = tests/tools/tokenizer_driver.cpp
compiled against /root/reference/src/models/tokenizer.h by `make -C oracle ref_tokenizer`).  Build container only.

    python tests/golden/make_tokenizer_golden.py
"""
import json
import os
import struct
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
from test_tokenizer import CASES, synthetic_vocabulary, write_vocabulary  # noqa: E402


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "tokenizer_ref")
    assert os.path.exists(exe), "build it first: make -C oracle ref_tokenizer (needs /root/reference)"
    vocab = synthetic_vocabulary()
    path = "/tmp/tokenizer_golden_vocab.bin"
    write_vocabulary(path, vocab)
    out = subprocess.run([exe, path], input="\n".join(CASES).encode() + b"\n", stdout=subprocess.PIPE, check=True).stdout.decode()
    json.dump({"vocabulary": [[t.hex(), i, s] for t, i, s in vocab], "cases": CASES, "expected": out.splitlines()},
              open(os.path.join(HERE, "tokenizer_golden.json"), "w"), indent=0)
    print(out)


if __name__ == "__main__":
    main()
