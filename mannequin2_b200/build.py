"""Build libmq_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m mannequin2_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to mannequin2_b200/csrc/_build/,
the library to mannequin2_b200/libmq_b200.so (git-ignored, shipped by gpurun).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libmq_b200.so")
BUILD = os.path.join(CSRC, "_build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
if os.environ.get("MQ_TC_PROFILE"):          # developer aid: per-phase cycle counters in the tcgen05 kernel (tools/tc_phases.py)
    FLAGS.append("-DMQ_TC_PROFILE")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s: cannot build libmq_b200.so" % NVCC)

    def compile_one(src):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        for src, obj, r in ex.map(compile_one, sources()):
            log = os.path.join(BUILD, src[:-3] + ".ptxas.log")
            open(log, "w").write(r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(r.stderr)
            objs.append(obj)
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    open(stamp, "w").write(dig)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
