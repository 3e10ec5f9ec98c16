"""Builds chimeracl_b200/libchimera_b200.so (sm_100a) in-tree with nvcc.

    python -m chimeracl_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels
to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libchimera_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include")]
# per-file extra flags (none at present; integer-deciding arithmetic uses explicit
# round-to-nearest intrinsics instead of a global -fmad=false)
SOURCES = {
    "capi.cu": [],
    "particles.cu": [],
    "deposit.cu": [],
    "gather.cu": [],
    "spectral.cu": [],
    "dht.cu": [],
    "fft.cu": [],
    "peer.cu": [],
}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps(src):
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "chimera_b200.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return deps


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    procs = []
    for name, extra in SOURCES.items():
        src = os.path.join(CSRC, name)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ_DIR, name.replace(".cu", ".o"))
        objs.append(obj)
        newest = max(os.path.getmtime(d) for d in _deps(src))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            continue
        cmd = [nvcc] + ARCH + COMMON + extra + os.environ.get("CHB_NVCC_FLAGS", "").split() + \
              (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", src, "-o", obj]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for name, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (name, out))
        elif verbose or out.strip():
            sys.stderr.write("[%s]\n%s\n" % (name, out))
    if failed:
        raise RuntimeError("chimera_b200: CUDA build failed")
    need_link = force or procs or not os.path.exists(LIB) or \
        any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
