#!/usr/bin/env bash
# Copy the reference's importable sources (src/, read-only at /root/reference in the build container) into the
# git-ignored oracle/_ref/ so that they travel to the GPU box with the gpurun snapshot.  Nothing is edited: the
# reference arm of bench.py (--impl reference), the CPU/eager-GPU baselines of the bench line and the suite
# integration test import them from there, unmodified.  oracle/_ref/ never enters git history (.gitignore) and is
# test / baseline infrastructure only -- nothing under nerf_dbr_b200/ imports it.
set -euo pipefail
REF="${NERF_DBR_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")/.." && pwd)"
DST="$HERE/oracle/_ref"
if [ ! -d "$REF/src" ]; then
    echo "vendor_reference: $REF/src not found (only the build container has the reference)" >&2
    exit 3
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$REF/src" "$DST/src"
chmod -R u+w "$DST"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
# (matplotlib, which the reference imports but this image lacks, is stubbed in-process by oracle/refload.py)
( cd "$REF" && find src -name "*.py" -type f | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "vendored $(find "$DST/src" -name '*.py' | wc -l) files from $REF/src into $DST"
