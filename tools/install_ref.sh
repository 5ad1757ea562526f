#!/bin/bash
# Copy the UNMODIFIED reference package (pure Python, numpy + scipy only) from /root/reference
# into oracle/_ref/ so that it travels to the GPU box with the gpurun snapshot.  oracle/_ref/ is
# git-ignored (the reference's sources never enter this repository's history) but not
# gpurun-ignored.  Used only as a checker and as the CPU arm of bench.py (--impl reference,
# cpu_baseline.kind = "reference"); the product never imports it.
set -e
here="$(cd "$(dirname "$0")/.." && pwd)"
src="${1:-/root/reference}"
if [ ! -d "$src/skfem" ]; then
  echo "install_ref: $src/skfem not found (nothing to do on a box without the reference)" >&2
  exit 0
fi
mkdir -p "$here/oracle/_ref"
rm -rf "$here/oracle/_ref/skfem"
cp -r "$src/skfem" "$here/oracle/_ref/skfem"
find "$here/oracle/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "install_ref: $(du -sh "$here/oracle/_ref" | cut -f1) in oracle/_ref"
