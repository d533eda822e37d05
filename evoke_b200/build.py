"""Build libevoke_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the
library is a plain C ABI, see include/evoke_b200.h).

    python -m evoke_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libevoke_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, "csrc", ".build_stamp")

SOURCES = ["evk_api.cu", "k1_l2norm.cu", "k2_posmask.cu", "k_small.cu", "k_stats.cu", "k_pos.cu", "k_peer.cu", "k_wstrip.cu", "k_local.cu", "k_topk.cu", "tc_engine.cu"]
HEADERS = ["evk_common.cuh", "tc_ptx.cuh", "peer_sync.cuh", os.path.join(ROOT, "include", "evoke_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _source_hash() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link libevoke_b200.so next to the package."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = os.path.join(PKG_DIR, "csrc", "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif verbose and out:
            print(f"--- {src} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", *objs, "-o", LIB_PATH]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout)
    with open(STAMP_PATH, "w") as f:
        f.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
