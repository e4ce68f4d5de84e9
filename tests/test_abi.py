"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly what include/b200llm.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

from util import b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200llm.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        out[m.group(1)] = n
    return out


def test_header_declares_functions():
    d = declared()
    assert len(d) >= 35 and "b200_linear" in d and "b200_decoder_step" in d


def test_library_exports_every_declared_symbol():
    mod = b200()
    if not os.path.exists(mod.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(mod.LIB_PATH)
    for name in declared():
        assert hasattr(lib, name), f"{name} declared in include/b200llm.h but not exported"


def test_python_binding_matches_header():
    mod = b200()
    d = declared()
    assert set(d) == set(mod.SIGNATURES), (set(d) ^ set(mod.SIGNATURES))
    for name, n in d.items():
        assert len(mod.SIGNATURES[name]) == n, f"{name}: header has {n} args, binding {len(mod.SIGNATURES[name])}"


def test_no_cpu_fallback_and_no_oracle_in_product():
    """The product path must not reference the oracle, and must fail loudly without the CUDA library."""
    pkg = os.path.join(ROOT, "llm-inference-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "__init__.py" and "oracle" not in text, f"{f} mentions the oracle"
    mod = b200()
    saved = mod.LIB_PATH
    try:
        mod.LIB_PATH = "/nonexistent/libb200llm.so"
        mod._lib = None
        with pytest.raises(mod.B200Error):
            mod.lib()
    finally:
        mod.LIB_PATH = saved
        mod._lib = None


def test_abi_version_and_error_string_without_gpu():
    lib = b200().lib()
    assert lib.b200_abi_version() == 1
    # argument validation happens before any CUDA call
    rc = lib.b200_rmsnorm(None, None, None, 1e-6, 1, 8, 0, None)
    assert rc == -1 and b"non-null" in lib.b200_last_error_string()
    rc = lib.b200_topk(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 10, 99, 0, None)
    assert rc == -1 and b"k=99" in lib.b200_last_error_string()
