"""The shim's CPU tokenizer (llm-inference-engine_b200/shim/src/models/tokenizer.h) against golden vectors from the
reference's own header (src/models/tokenizer.h:57-347, compiled where it lies into oracle/_ref/tokenizer_ref; fixture
tests/golden/tokenizer_golden.json, script tests/golden/make_tokenizer_golden.py).  Ids and decoded bytes must be identical, quirks included
(prefixes that are no vocabulary entry merge with score 0 and read id 0; "<FLM_FIX_TOKEN_n>" literals; "<0xNN>" byte fallback; "<|blank_n|>")."""
import json
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "llm-inference-engine_b200", "shim", "_own_programs", "tokenizer_shim")
REF = os.path.join(ROOT, "oracle", "_ref", "tokenizer_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json")
MARK = "▁".encode()


def synthetic_vocabulary():
    """[(bytes, id, score)]: byte-fallback entries, single characters, a few merges with SentencePiece-like negative scores (longer = better),
    ties, a positive score, entries reachable only through non-entry prefixes, and the special decode strings."""
    v = [(b"<unk>", 0, 0.0), (b"<s>", 1, 0.0), (b"</s>", 2, 0.0)]
    nid = 3
    for b in range(256):
        v.append((("<0x%02X>" % b).encode(), nid, 0.0))
        nid += 1
    for ch in "abcdefghijklmnoprstuwy.,!?'-:>0123456789":
        v.append((ch.encode(), nid, -20.0 - (nid % 7)))
        nid += 1
    v.append((MARK, nid, -5.0)); nid += 1
    for b in (0x96, 0x81, 0xA9):  # raw continuation bytes as entries of their own (byte-level vocabularies have them): without them the
        v.append((bytes([b]), nid, -30.0)); nid += 1  # mark's 2nd and 3rd byte start no entry and it can never be merged
    for piece, score in [("he", -9.0), ("ll", -9.0), ("lo", -9.0), ("hell", -7.0), ("hello", -4.0), ("wor", -8.0), ("world", -3.5), ("ld", -9.5),
                         ("or", -9.0), ("th", -6.0), ("the", -2.0), ("er", -6.0), ("re", -6.0), ("there", -1.0), ("an", -6.5), ("and", -2.5),
                         ("in", -6.5), ("ing", -3.0), ("abc", 1.5), ("abcd", -1.0), ("é", -10.0), ("日本", -4.0), ("日", -12.0), ("本", -12.0),
                         ("can't", -2.0), ("<n>", -1.0), ("<|tab|>", -1.0), ("<|blank_4|>", -1.0), ("xyzzy", -3.0)]:
        v.append((piece.encode(), nid, score)); nid += 1
    for piece, score in [("hello", -4.0), ("the", -2.0), ("there", -1.0), ("and", -2.5), ("world", -3.5), ("abc", 1.5), ("日本", -4.0)]:
        v.append((MARK + piece.encode(), nid, score + 0.5)); nid += 1
    v.append((b"dup", nid, -3.0)); nid += 1
    v.append((b"dup", nid, -2.0)); nid += 1  # inserted twice: the later id / score win
    return v


CASES = [
    "T hello world", "T   hello   world  ", "T the there and thing", "T hellohello worldwide", "T abc abcd abcde", "T can't handle héllo, 日本!",
    "T xyz xyzzy q", "T <FLM_FIX_TOKEN_42>abc", "T a<FLM_FIX_TOKEN_7>b<FLM_FIX_TOKEN_123456>", "T <FLM_FIX_TOKEN_9", "T <FLM and <FLM_FIX_TOKEN_x>", "T dup dupe",
    "T 1234567890 -> 42!", "T ~^`|{}", "T in an inn innings", "T \t tab and \x7f del", "T a", "T  ", "T", "T ther there's therein",
    "I 1 2 0", "I 268 300 301", "I 3 4 100 200 258", "I 99999 5", "I", "I 13 35 68 104",
]


def write_vocabulary(path, vocab, version=1):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", version))
        if version >= 1:
            pairs = [(b"tokenizer_use_score", b"1"), (b"model_type", b"llama")]
            f.write(struct.pack("<i", len(pairs)))
            for k, val in pairs:
                f.write(struct.pack("<i", len(k)) + k + struct.pack("<i", len(val)) + val)
        f.write(struct.pack("<i", len(vocab)))
        for text, tid, score in vocab:
            f.write(struct.pack("<i", len(text)))
            for b in text:
                f.write(struct.pack("<i", b if b < 128 else b - 256))  # one int32 per byte, as a (signed) char
            f.write(struct.pack("<if", tid, score))


def replay(exe, vocab_path, cases):
    return subprocess.run([exe, vocab_path], input="\n".join(cases).encode() + b"\n", stdout=subprocess.PIPE, check=True, timeout=120).stdout.decode().splitlines()


def ids_of_special(vocab, text):
    return [i for t, i, _ in vocab if t == text][-1]


def test_shim_tokenizer_reproduces_the_reference_golden(tmp_path):
    if not os.path.exists(SHIM):
        pytest.skip("shim/_own_programs/tokenizer_shim not built (run __graft_entry__.build())")
    g = json.load(open(GOLDEN))
    vocab = [(bytes.fromhex(t), i, s) for t, i, s in g["vocabulary"]]
    assert [[t.hex(), i, s] for t, i, s in synthetic_vocabulary()] == g["vocabulary"] and g["cases"] == CASES, "fixture is stale: regenerate it"
    path = str(tmp_path / "vocab.bin")
    write_vocabulary(path, vocab)
    got = replay(SHIM, path, g["cases"])
    assert len(got) == len(g["expected"])
    for i, (a, b) in enumerate(zip(got, g["expected"])):
        assert a == b, f"line {i}: shim {a!r} != reference {b!r}"
    # sanity of the fixture itself: a plain word round-trips, the longest-scoring merge wins, the byte fallback is used
    exp = g["expected"]
    hello = ids_of_special(vocab, MARK + b"hello")
    assert exp[0].split()[1:2] == [str(hello)]
    assert bytes.fromhex(exp[1][2:]) == b" hello world"
    assert str(ids_of_special(vocab, b"<0x7E>")) in exp[2 * CASES.index("T ~^`|{}")].split()


def test_shim_tokenizer_matches_the_reference_binary_on_more_text(tmp_path):
    """Where the reference header is available (build container), compare live on pseudo-random text over the vocabulary's alphabet."""
    if not (os.path.exists(REF) and os.path.exists(SHIM)):
        pytest.skip("needs oracle/_ref/tokenizer_ref (built from /root/reference) and the shim driver")
    import random

    rnd = random.Random(5)
    alphabet = "abcdehlnorstw  .,'日本é<>_FLMIXTOKEN0123456789xyz~"
    cases = ["T " + "".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 40))) for _ in range(300)]
    cases += ["T " + " ".join(rnd.choice(["hello", "world", "the", "there", "and", "abc", "abcd", "can't", "dup", "xyzzy", "日本", "in", "ing"]) for _ in range(rnd.randint(1, 9)))
              for _ in range(200)]
    cases += ["I " + " ".join(str(rnd.randint(0, 340)) for _ in range(rnd.randint(0, 12))) for _ in range(100)]
    path = str(tmp_path / "vocab.bin")
    write_vocabulary(path, synthetic_vocabulary())
    a, b = replay(SHIM, path, cases), replay(REF, path, cases)
    assert len(a) == len(b)
    bad = [(c, x, y) for c, x, y in zip([c for c in cases for _ in range(2 if c[0] == "T" else 1)], a, b) if x != y]
    assert not bad, bad[:3]
    # version-0 files (no key-value table) load the same
    write_vocabulary(path, synthetic_vocabulary(), version=0)
    assert replay(SHIM, path, cases[:20]) == replay(REF, path, cases[:20])
