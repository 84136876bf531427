"""In-tree build of libmw_b200.so (nvcc, sm_100a only).  `python -m manual_whisper_b200.build`.

The built library sits next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
REPO = PKG.parent
OBJ = REPO / "build" / "obj"
LIB = PKG / "libmw_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(REPO / "include"),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libmw_b200.so cannot be built (there is no CPU fallback)")


def _newest_header() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((REPO / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hs)


def build(force: bool = False, verbose: bool = False) -> Path:
    """MW_STORAGE_BF16=1 in the environment builds the bf16-storage variant next to the default fp16 one
    (libmw_b200_bf16.so, objects in build/obj_bf16; manual_whisper_b200/_lib.py loads it under the same variable): A/B runs."""
    global OBJ, LIB
    nvcc = _nvcc()
    bf16 = os.environ.get("MW_STORAGE_BF16") == "1"
    OBJ = REPO / "build" / ("obj_bf16" if bf16 else "obj")
    LIB = PKG / ("libmw_b200_bf16.so" if bf16 else "libmw_b200.so")
    OBJ.mkdir(parents=True, exist_ok=True)
    extra = (["-DMW_STORAGE_BF16"] if bf16 else []) + os.environ.get("MW_EXTRA_NVCC", "").split()
    srcs = sorted(CSRC.glob("*.cu"))
    hdr = max(_newest_header(), Path(__file__).stat().st_mtime)
    jobs = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, hdr):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(s), "-o", str(o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (s.stem + ".ptxas.log")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stderr[-6000:]}")
        if verbose:
            sys.stderr.write(r.stderr)
        return s.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                print(f"[build] compiled {name}", file=sys.stderr)
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if jobs or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
        print(f"[build] linked {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
