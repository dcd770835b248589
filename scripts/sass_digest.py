"""sha256 of the SASS text (`cuobjdump -sass`, comment lines dropped) of every shipped object: two builds with the same
digests run the same device code, whatever happened to comments, line numbers (-lineinfo) or host code in between.

    python -m dgvcc_b200.build && python scripts/sass_digest.py [> profiles/<tag>_sass_digest.txt]
"""
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "dgvcc_b200", "lib", "obj")
for name in sorted(f for f in os.listdir(OBJ) if f.endswith(".o")):
    text = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", os.path.join(OBJ, name)], capture_output=True, text=True).stdout
    # function headers and instruction lines only: the fatbin header carries the absolute source path
    body = "\n".join(line for line in text.splitlines() if line.lstrip().startswith(("/*", "Function :")))
    print(hashlib.sha256(body.encode()).hexdigest(), name)
