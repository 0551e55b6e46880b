"""CPU: the C-ABI library builds, loads, exports every symbol include/cql_b200.h declares,
and refuses (loudly) to work without a GPU -- no compute calls here."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "cql_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(cql_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_functions():
    fns = declared_functions()
    assert "cql_create" in fns and "cql_score_topk" in fns and len(fns) >= 20


def test_library_exports_every_declared_symbol():
    from replay_cql_b200 import build, _lib
    lib_path = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_path)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(cql_[a-z_0-9]+)", out))
    missing = [f for f in declared_functions() if f not in exported]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == declared_functions()   # the ctypes binding covers the whole header
    lib = _lib.load()
    assert lib.cql_abi_version() == 2


def test_config_struct_matches_header_size():
    from replay_cql_b200 import _lib
    # 9 int32 + 13 float + padding + uint64
    assert ctypes.sizeof(_lib.CqlConfig) == 96


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from replay_cql_b200 import _lib
    from replay_cql_b200.engine import CqlEngine
    with pytest.raises(_lib.CqlLibraryError, match="no CUDA device|no CPU fallback"):
        CqlEngine()


def test_product_never_imports_oracle():
    for f in (ROOT / "replay_cql_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
