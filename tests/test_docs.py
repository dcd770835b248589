"""The documents do not rot: every `from dgvcc_b200... import ...` INTEGRATION.md shows resolves, and every
`file.py:line[-line]` citation into the reference (header, DESIGN, INTEGRATION, docstrings, kernels, oracle, tests)
names an existing reference file and lines it has.  CPU only; the citation check needs /root/reference."""
import glob
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_integration_md_imports_resolve():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    found = set(re.findall(r"from (dgvcc_b200[\w\.]*) import ([\w, ]+)", text))
    assert len(found) >= 10
    for module, names in sorted(found):
        mod = importlib.import_module(module)
        for name in (n.strip() for n in names.split(",")):
            assert hasattr(mod, name), f"INTEGRATION.md imports {name} from {module}, which has no such name"


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
def test_reference_citations_point_at_existing_lines():
    files = [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "bench.py")]
    for pattern in ("include/*.h", "dgvcc_b200/**/*.py", "dgvcc_b200/csrc/*", "oracle/*.py", "tests/*.py"):
        files += glob.glob(os.path.join(ROOT, pattern), recursive=True)
    cite = re.compile(r"(?<![\w/])((?:\w+/)*\w+\.py):(\d+)(?:-(\d+))?")
    lengths = {}

    def ref_lines(rel):
        if rel not in lengths:
            path = os.path.join(REF, rel)
            hits = [path] if os.path.exists(path) else glob.glob(os.path.join(REF, "**", rel), recursive=True)
            lengths[rel] = sum(1 for _ in open(hits[0], errors="ignore")) if hits else None
        return lengths[rel]

    checked, bad = 0, []
    for f in files:
        if not os.path.isfile(f):
            continue
        for m in cite.finditer(open(f, errors="ignore").read()):
            rel, first, last = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            own = any(os.path.exists(os.path.join(ROOT, sub, rel)) for sub in ("", "dgvcc_b200", "tests"))
            if own and not os.path.exists(os.path.join(REF, rel)):
                continue                      # a citation into this repository, not into the reference
            checked += 1
            n = ref_lines(rel)
            if n is None or first > last or last > n:
                bad.append((os.path.relpath(f, ROOT), m.group(0), "no such reference file" if n is None else f"file has {n} lines"))
    assert checked >= 300, checked
    assert not bad, bad[:20]


def test_integration_md_ctypes_stub_matches_the_signature_table():
    """The raw ctypes stub of INTEGRATION.md section 3 (what a maintainer would write without this package) declares the
    argument types dgvcc_b200._native declares for the same entry point."""
    import ctypes
    from dgvcc_b200 import _native
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"lib\.(dgvcc_\w+)\.argtypes = (.*?)\n\n", text, flags=re.S)
    assert m, "stub not found"
    name, expr = m.group(1), m.group(2).replace("\\\n", " ")
    stub = eval(expr, {"ctypes": ctypes})
    res, args = _native.SIGNATURES[name]
    assert list(stub) == list(args), name
    assert re.search(rf"lib\.{name}\.restype = ctypes\.c_int\b", text) and res is ctypes.c_int
