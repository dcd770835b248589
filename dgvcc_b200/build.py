"""Build libdgvcc_b200.so in-tree with nvcc for sm_100a (no JIT cache, so the .so travels with the repo).

    python -m dgvcc_b200.build [--force] [-v]

Every .cu under csrc/ is compiled to its own object (in parallel, only the stale ones) and the objects
are linked into one shared library.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libdgvcc_b200.so")
STAMP_PATH = LIB_PATH + ".sha256"
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    # no --use_fast_math / -ftz: the rounding sequence of the reference is part of the contract
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")


def _newer(path, than):
    return not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in than)


def _digest():
    """Content hash of everything the library is built from (mtimes do not survive a copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sorted(sources() + _headers()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    try:
        with open(STAMP_PATH) as f:
            return not os.path.exists(LIB_PATH) or f.read().strip() != _digest()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _newer(_obj(s), [s] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE, "-c", src, "-o", _obj(src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, proc

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as pool:
        for src, proc in pool.map(compile_one, todo):
            if proc.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
            if verbose:
                print(src, proc.stderr, flush=True)
    objs = [_obj(s) for s in sources()]
    for stray in set(glob.glob(os.path.join(OBJ_DIR, "*.o"))) - set(objs):
        os.remove(stray)
    tmp = LIB_PATH + ".tmp"
    proc = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs,
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    with open(STAMP_PATH, "w") as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
