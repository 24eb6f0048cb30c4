"""Build libscp_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m speechclip_plus_b200.build [--force]

The shared object is written next to this file so that it travels with the source tree; it is the ONLY
compute backend of the package (there is no CPU or eager-PyTorch fallback).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libscp_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")

SOURCES = ["scp_runtime.cu", "scp_wsum.cu", "scp_kwbn.cu", "scp_splice.cu", "scp_cif.cu", "scp_vq.cu", "scp_nce.cu", "scp_optim.cu", "scp_p2p.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]
if os.environ.get("SCP_BUILD_ABLATION") == "1":  # timing-ablation switches (results invalid), see csrc/scp_common.cuh
    NVCC_FLAGS += ["-DSCP_ABLATION"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "scp_b200.h"), __file__]
    return max(os.path.getmtime(f) for f in files)


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _deps_mtime()


def build(force: bool = False, verbose: bool = False) -> str:
    """Serialised across processes (torchrun starts one per GPU and every rank may find the library stale at once):
    an exclusive file lock is held for the whole build, the late ranks re-check under the lock and find the library
    fresh; objects and the linked library are written under per-process names and renamed into place."""
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    tag = f".{os.getpid()}"

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(INCLUDE, "scp_b200.h"), __file__]
    hdr_mtime = max(os.path.getmtime(f) for f in headers)

    def compile_one(src: str) -> str:
        final = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if os.path.exists(final) and os.path.getmtime(final) > max(hdr_mtime, os.path.getmtime(os.path.join(CSRC, src))):
            return final  # this unit and every header it can include are older than its object
        obj = os.path.join(OBJ_DIR, src.replace(".cu", tag + ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp = LIB_PATH + tag + ".tmp"
    # cudart is linked statically (nvcc default): the library has no runtime dependency besides libcuda (resolved
    # lazily through cudaGetDriverEntryPoint) and libstdc++
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    for obj in objs:  # canonical object names: the next build recompiles only the units that changed
        if obj.endswith(tag + ".o"):
            os.replace(obj, obj.replace(tag + ".o", ".o"))
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose=True)
    print("built", path)
