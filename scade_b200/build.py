"""Build libscade_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch C++ ABI).

    python -m scade_b200.build [--force] [--verbose]

The .so lands in scade_b200/_lib/ so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libscade_b200.so")
STAMP = os.path.join(LIBDIR, "build.stamp")
TRACE = os.environ.get("SCADE_TC_TRACE", "0") == "1"      # timeline-tracing variant of the library (tools/tc_trace.py)
if TRACE:
    LIB = os.path.join(LIBDIR, "libscade_b200_trace.so")
    STAMP = os.path.join(LIBDIR, "build_trace.stamp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
if TRACE:
    FLAGS.append("-DSCADE_TC_TRACE=1")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB          # GPU box without a toolkit: use the prebuilt library that travelled with the repo
        raise RuntimeError(f"nvcc not found at {NVCC} and no prebuilt {LIB}")
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ("_trace.o" if TRACE else ".o"))
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart", "-lcuda"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
