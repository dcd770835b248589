#!/bin/sh
# TEST / BENCH INFRASTRUCTURE ONLY -- never imported by dgvcc_b200/.
#
# Puts the reference's own files for the hot path into the git-ignored oracle/_ref/ (it travels to the GPU box with the
# snapshot, like the built .so; /root/reference does not exist there), so that `bench.py --impl reference`, bench.py's
# cpu_baseline / gpu_eager legs and scripts/ can time the UNMODIFIED reference instead of the oracle port.
# The reference is pure Python: nothing is compiled, nothing is edited; the three files are byte copies.
#   losses/bl.py                         Post_Prob / Bay_Loss / BL          
#   utils/dmap_gen.py                    gaussian_filter_density{,_fixed}   (rows a5-a7)
#   models/ISW/instance_whitening.py     InstanceWhitening, Gram loss       (rows a8-a10)
set -e
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "make_ref: $REF not present (GPU box?) -- keeping whatever is in $OUT"; exit 0; }
rm -rf "$OUT"
mkdir -p "$OUT/losses" "$OUT/utils" "$OUT/models/ISW"
cp "$REF/losses/bl.py" "$OUT/losses/bl.py"
cp "$REF/utils/dmap_gen.py" "$OUT/utils/dmap_gen.py"
cp "$REF/models/ISW/instance_whitening.py" "$OUT/models/ISW/instance_whitening.py"
( cd "$OUT" && sha256sum losses/bl.py utils/dmap_gen.py models/ISW/instance_whitening.py > SHA256SUMS )
echo "make_ref: unmodified reference files in $OUT"
