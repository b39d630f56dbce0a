"""Build ``libgpjax_b200.so`` (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m gpjax_b200.build            # incremental
    python -m gpjax_b200.build --force

No CPU fallback exists: if this library is missing, importing any compute path of
``gpjax_b200`` raises.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB_PATH = os.path.join(OUT_DIR, "libgpjax_b200.so")

CU_SOURCES = ["gemm_f64.cu", "gram.cu", "potrf_leaf.cu", "level2.cu", "sgpr_kernels.cu", "svgp_kernels.cu", "ozaki_i8.cu", "profile.cu"]
CPP_SOURCES = ["algorithms.cpp", "sgpr.cpp", "abi.cpp", "collective.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
if os.environ.get("GPB_NB"):  # experiment hook: block size of the blocked algorithms (default 1024, see algorithms.h)
    COMMON += [f"-DGPB_NB={int(os.environ['GPB_NB'])}"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: gpjax_b200 needs the CUDA toolkit to build its kernels")
    return nvcc


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".cu", ".cpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _local_deps(path: str, seen=None) -> set:
    """`path` plus every quoted #include reachable from it (so one header edit rebuilds only its users)."""
    import re

    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
        _local_deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [s for s in CU_SOURCES + CPP_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _deps_mtime():
        return LIB_PATH
    nvcc = _nvcc()
    flags_tag = " ".join(COMMON)

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        tag = obj + ".flags"
        newest = max(os.path.getmtime(d) for d in _local_deps(os.path.join(CSRC, src)))
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest and os.path.exists(tag)
                and open(tag).read() == flags_tag):
            return obj
        cmd = [nvcc, *ARCH, *COMMON, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        open(tag, "w").write(flags_tag)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
