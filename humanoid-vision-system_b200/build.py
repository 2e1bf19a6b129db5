"""Build libhvs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libhvs_b200.so")
STAMP = os.path.join(PKG_DIR, "build", "sources.sha256")

SOURCES = ["api_common.cu", "mhc_stream_fwd.cu", "mhc_stream_generic.cu", "mhc_stream_generic_bwd.cu", "mhc_stream_bwd.cu", "mhc_stream_bwd_fused.cu", "sinkhorn.cu", "yolo_decode.cu", "nms.cu", "norm.cu", "k2_gemm.cu", "k2_coeffs.cu", "k2_chain.cu", "train_ops.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc() -> str:
    """Path of nvcc, or RuntimeError when there is none."""
    return _nvcc()


def _fingerprint() -> str:
    h = hashlib.sha256()
    root = os.path.dirname(PKG_DIR)
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(root, "include", "hvs_b200.h"))
    files.append(os.path.abspath(__file__))
    for f in files:
        if os.path.isfile(f):
            h.update(os.path.basename(f).encode())     # content + name, not the checkout path: the stamp travels with the tree
            with open(f, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def is_fresh() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library.  Returns its path.
    HVS_FUSED_TRACE=1 in the environment adds the cycle-counter instrumentation of the fused backward
    (development aid, tools/time_fused.py)."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    obj_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    # one builder at a time (several ranks may import the package at once); the others find a fresh library
    import fcntl
    lock = open(os.path.join(obj_dir, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and is_fresh():
            return LIB_PATH
        return _build_locked(nvcc, obj_dir, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(nvcc: str, obj_dir: str, verbose: bool) -> str:
    procs = []
    objs = []
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        if not os.path.exists(spath):
            raise RuntimeError(f"missing source {spath}")
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", spath, "-o", obj]
        if os.environ.get("HVS_FUSED_TRACE"):
            cmd.insert(1, "-DHVS_FUSED_TRACE")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[hvs_b200 build] {src} failed:\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)                          # atomic: a concurrent loader sees the old or the new library, never half of one
    with open(STAMP, "w") as f:
        f.write(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
