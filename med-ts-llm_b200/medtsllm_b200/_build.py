"""Builds libmtsb200.so (sm_100a only) in-tree with nvcc.

The shared library is the product's only native artefact: hand-written CUDA kernels behind the C ABI
declared in include/mts_b200.h.  It is built IN-TREE (med-ts-llm_b200/medtsllm_b200/lib/) so that it
travels with the repo snapshot to the GPU box; object files go to build/ (git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR.parent / "csrc"
REPO = PKG_DIR.parent.parent
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libmtsb200.so"
OBJ_DIR = REPO / "build" / "mtsb200"

SOURCES = ["common.cu", "gemm_tcgen05.cu", "gemm_tcgen05_2cta.cu", "rowops.cu", "frontend.cu", "attention.cu", "attention_tc.cu", "backward.cu", "precise.cu", "stats.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libmtsb200.so")


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libmtsb200.so.  Idempotent."""
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [REPO / "include" / "mts_b200.h"]
    stamp = LIB_DIR / ".build_stamp"
    digest = _digest(deps)
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(REPO / "include"), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    if verbose:
        print(f"[mtsb200] built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
