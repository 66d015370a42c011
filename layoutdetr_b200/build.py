"""Build liblayoutdetr_sm100.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m layoutdetr_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot.  The library has no dependency on torch: cudart is linked statically and the one
driver entry point (cuTensorMapEncodeTiled) is resolved at run time.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "liblayoutdetr_sm100.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build liblayoutdetr_sm100.so")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile_one(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return src


def build(force=False, verbose=True):
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    hdr_m = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        obj = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(obj)
        src_m = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_m)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_m:
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print("[layoutdetr_b200.build] nvcc sm_100a: %s" % ", ".join(s for s, _ in jobs), flush=True)
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _compile_one(nvcc, *j), jobs))
    need_link = bool(jobs) or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[layoutdetr_b200.build] linked %s" % LIB, flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
