"""Build libdgvcc_b200.so in-tree with nvcc for sm_100a (no JIT cache, so the .so travels with the repo).

    python -m dgvcc_b200.build [--force] [-v] [--checked]

Every .cu under csrc/ is compiled to its own object (in parallel, only the stale ones) and the objects
are linked into one shared library.

``--checked`` (or ``DGVCC_BOUNDS_CHECK=1`` in the environment) builds the CHECKED variant beside the product library:
``lib/libdgvcc_b200_chk.so``, compiled with ``-DDGVCC_BOUNDS_CHECK`` (device-side index checks that trap, see
csrc/common.cuh).  ``dgvcc_b200._native`` loads it instead of the product library when ``DGVCC_BOUNDS_CHECK=1`` is set
at import, so ``DGVCC_BOUNDS_CHECK=1 python -m pytest tests -m gpu`` runs the whole GPU suite through the checks.
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
CHK_OBJ_DIR = os.path.join(LIB_DIR, "obj_chk")
CHK_LIB_PATH = os.path.join(LIB_DIR, "libdgvcc_b200_chk.so")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    # no --use_fast_math / -ftz: the rounding sequence of the reference is part of the contract
]
CHK_FLAGS = ["-DDGVCC_BOUNDS_CHECK"]


def checked_from_env():
    return os.environ.get("DGVCC_BOUNDS_CHECK", "0") not in ("", "0")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def _obj(src, obj_dir=OBJ_DIR):
    return os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")


def _newer(path, than):
    return not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in than)


def _digest(flags=NVCC_FLAGS):
    """Content hash of everything the library is built from (mtimes do not survive a copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for path in sorted(sources() + _headers()):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(lib_path=LIB_PATH, flags=NVCC_FLAGS):
    try:
        with open(lib_path + ".sha256") as f:
            return not os.path.exists(lib_path) or f.read().strip() != _digest(flags)
    except OSError:
        return True


def build(force=False, verbose=False, checked=None):
    """Compile every .cu under csrc/ into one shared library; returns its path.  ``checked``: the variant with the
    device-side index checks (default: what DGVCC_BOUNDS_CHECK says)."""
    if checked is None:
        checked = checked_from_env()
    flags = NVCC_FLAGS + (CHK_FLAGS if checked else [])
    lib_path, obj_dir = (CHK_LIB_PATH, CHK_OBJ_DIR) if checked else (LIB_PATH, OBJ_DIR)
    if not force and not _stale(lib_path, flags):
        return lib_path
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _newer(_obj(s, obj_dir), [s] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + flags + (["-Xptxas=-v"] if verbose else []) + ["-I", INCLUDE, "-c", src, "-o", _obj(src, obj_dir)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, proc

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as pool:
        for src, proc in pool.map(compile_one, todo):
            if proc.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
            if verbose:
                print(src, proc.stderr, flush=True)
    objs = [_obj(s, obj_dir) for s in sources()]
    for stray in set(glob.glob(os.path.join(obj_dir, "*.o"))) - set(objs):
        os.remove(stray)
    tmp = f"{lib_path}.{os.getpid()}.tmp"   # several ranks may find the library stale at the same moment
    proc = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs,
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, lib_path)
    with open(lib_path + ".sha256", "w") as f:
        f.write(_digest(flags))
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, checked=True if "--checked" in sys.argv else None))
