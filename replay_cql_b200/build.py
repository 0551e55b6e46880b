"""Builds the C-ABI CUDA library in-tree:  python -m replay_cql_b200.build

Two translation units (csrc/capi.cu includes the kernel headers; csrc/mdp_gpu.cu the CUB sorts), compiled for
sm_100a only.  The .so lands next to this file so it travels with the repo
snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libcql_b200.so"
STAMP = HERE / ".libcql_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-ldl",
]


def _source_digest() -> str:
    hsh = hashlib.sha256()
    files = sorted(CSRC.glob("*")) + [HERE.parent / "include" / "cql_b200.h"]
    for f in files:
        if f.is_file():
            hsh.update(f.name.encode())
            hsh.update(f.read_bytes())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _source_digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("CQL_EXTRA_NVCC", "").split()          # debugging only (e.g. -DTKS_TIMING)
    digest = digest + "|" + " ".join(extra) if extra else digest
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", str(LIB), str(CSRC / "capi.cu"), str(CSRC / "mdp_gpu.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
