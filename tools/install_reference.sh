#!/bin/sh
# Copy the files of the UNMODIFIED reference that its sampling / proximal path needs to baseline/_ref/ (git-ignored,
# shipped to the GPU box with the snapshot): the `src` package, the three model yamls and the two PDB fixtures.
# bench.py --impl reference, bench.py's cpu_baseline and tests/test_dropin_reference.py run it from there under
# tools/ref_shims.py (PACKPPI_REFERENCE=baseline/_ref).  Nothing of it is imported by the product path.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
DST="$HERE/baseline/_ref"
[ -d "$REF/src" ] || { echo "install_reference: $REF/src not found (nothing to do on the GPU box)"; exit 0; }
rm -rf "$DST"
mkdir -p "$DST/configs" "$DST/data"
cp -r "$REF/src" "$DST/src"
cp -r "$REF/configs/model" "$DST/configs/model"
cp "$REF/data/1BRS.pdb" "$REF/data/T1124_lig.pdb" "$DST/data/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
rm -rf "$DST/src/models/components/cache"
echo "installed reference into $DST ($(find "$DST" -type f | wc -l) files)"
