"""Build libcardelino_b200.so in-tree with nvcc for sm_100a (no CPU fallback exists)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcardelino_b200.so")
SOURCES = ["api.cu", "formats.cu", "gibbs.cu", "cells_sell.cu", "element.cu", "small.cu", "post.cu", "synth.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"] + os.environ.get("CDL_EXTRA_NVCC_FLAGS", "").split() + [
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "cardelino_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        log = o.replace(".o", ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, r.stdout + r.stderr))
        return s

    with ThreadPoolExecutor(max_workers=8) as ex:
        for s in ex.map(compile_one, jobs):
            if verbose:
                print("compiled", s)
    if jobs or force or _stale(OUT, objs):
        r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-ldl", "-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
