"""Builds swne_b200/libswne_nmf.so (sm_100a only) with nvcc: one object per KP instantiation, in parallel.

    python -m swne_b200.build [--force]

The .so is built in-tree (git-ignored, shipped to the GPU box by gpurun).  No torch involved: the
library links only libcudart (static) and libdl: NCCL is dlopen'ed and the driver entry point for TMA
descriptors is fetched with cudaGetDriverEntryPoint, so the .so also loads on a box without a GPU driver.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libswne_nmf.so")
KPS = (8, 16, 24, 32, 40, 48, 56, 64)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "swne_nmf.h"))
    return sorted(out)


def _digest():
    # file names relative to the package: the tree is copied to another absolute path on the GPU box and the
    # shipped .so must still count as current there
    h = hashlib.sha256()
    for p in _sources():
        h.update(os.path.relpath(p, HERE).replace(os.sep, "/").encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _current(stamp, dig):
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig


def build(force=False, verbose=False):
    stamp = os.path.join(OBJ, "digest.txt")
    dig = _digest()
    if not force and _current(stamp, dig):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # several ranks of one job may get here at once: one builds, the others wait and then find the stamp current
    import fcntl
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _current(stamp, dig):
                return LIB
            return _build_locked(stamp, dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


KP_DEPS = ("kernels_kp.cu", "kernels_simt.cuh", "common.cuh", "kp_ops.h")


def _files_digest(paths, extra=""):
    h = hashlib.sha256()
    for p in paths:
        h.update(os.path.relpath(p, HERE).replace(os.sep, "/").encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update((" ".join(FLAGS) + extra).encode())
    return h.hexdigest()


def _build_locked(stamp, dig, verbose):
    # per-object stamps: the eight per-KP translation units only depend on KP_DEPS, so an edit of the tcgen05
    # engine or the host loop recompiles one file
    jobs, objs, stamps = [], [], []

    def add(obj, cmd, deps, extra=""):
        objs.append(obj)
        d = _files_digest(deps, extra)
        st = obj + ".digest"
        if os.path.exists(obj) and os.path.exists(st) and open(st).read() == d:
            return
        jobs.append(cmd)
        stamps.append((st, d))

    main_o = os.path.join(OBJ, "swne_nmf.o")
    add(main_o, [NVCC] + FLAGS + ["-c", os.path.join(CSRC, "swne_nmf.cu"), "-o", main_o], _sources())
    kp_deps = [os.path.join(CSRC, f) for f in KP_DEPS] + [os.path.join(HERE, "..", "include", "swne_nmf.h")]
    for kp in KPS:
        o = os.path.join(OBJ, "kernels_kp%d.o" % kp)
        add(o, [NVCC] + FLAGS + ["-DKP_VALUE=%d" % kp, "-c", os.path.join(CSRC, "kernels_kp.cu"), "-o", o], kp_deps, " kp%d" % kp)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            outs = list(ex.map(_run, jobs))
        if verbose:
            for o in outs:
                if o.strip():
                    print(o)
        for st, d in stamps:
            with open(st, "w") as f:
                f.write(d)
    _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
