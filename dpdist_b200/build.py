"""Builds libdpdist_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dpdist_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
BUILD_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libdpdist_b200.so")
STAMP_PATH = LIB_PATH + ".stamp"     # next to the .so: travels with it (build/ directories may be dropped by snapshots)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libdpdist_b200.so cannot be built")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p):
                h.update(f.encode())
                with open(p, "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    if not os.path.exists(STAMP_PATH):
        # a prebuilt library without its stamp: rebuild if a compiler is here, otherwise trust the library
        try:
            _nvcc()
        except RuntimeError:
            return False
        return True
    with open(STAMP_PATH) as fh:
        return fh.read().strip() != _stamp()


def _compile_one(nvcc, src):
    obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(obj + ".ptxas.txt", "w") as fh:
        fh.write(r.stderr)
    return obj


def build_library(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libdpdist_b200.so.  Returns the .so path.
    Safe under concurrent callers (one rank per GPU all importing the package): the whole build runs under an exclusive
    file lock, whoever gets it second finds the library up to date, and the .so appears by an atomic rename so that no
    process can dlopen a half-written file."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(BUILD_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(BUILD_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it while we waited
                return LIB_PATH
            srcs = _sources()
            with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
                objs = list(ex.map(lambda s: _compile_one(nvcc, s), srcs))
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + ".tmp", "w") as fh:
                fh.write(_stamp())
            os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    if verbose:
        print("built", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
