"""Loader of the UNMODIFIED reference files (TEST / BENCH INFRASTRUCTURE ONLY).

``oracle/make_ref.sh`` copies losses/bl.py, utils/dmap_gen.py and models/ISW/instance_whitening.py byte for byte into
the git-ignored ``oracle/_ref/``; this module imports them BY FILE PATH (the reference's packages pull in modules
that are not installed -- models/ISW/__init__.py imports kmeans1d) under private module names, so
nothing of the reference's package layout leaks into ``sys.modules``.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOTS = [os.path.join(HERE, "_ref"), "/root/reference"]
FILES = {"bl": "losses/bl.py", "dmap_gen": "utils/dmap_gen.py", "instance_whitening": "models/ISW/instance_whitening.py"}


def path_of(name):
    for root in ROOTS:
        p = os.path.join(root, FILES[name])
        if os.path.exists(p):
            return p
    return None


def available(name="bl"):
    return path_of(name) is not None


def load(name):
    """The unmodified reference module ``name`` ('bl', 'dmap_gen', 'instance_whitening')."""
    key = f"_dgvcc_reference_{name}"
    if key in sys.modules:
        return sys.modules[key]
    path = path_of(name)
    if path is None:
        raise FileNotFoundError(f"reference file {FILES[name]} not found under {ROOTS}; run oracle/make_ref.sh where "
                                "/root/reference exists")
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True
    try:
        spec = importlib.util.spec_from_file_location(key, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[key] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = dont
    return mod
