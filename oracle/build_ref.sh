#!/bin/sh
# ORACLE — TEST INFRASTRUCTURE ONLY.
# Makes the UNMODIFIED reference importable on the GPU box: /root/reference exists only in the build container, so its
# Python packages for the sampling path (MuseDiffusion/ and the commu/ preprocessor it imports) are copied verbatim
# into the git-ignored directory oracle/_ref/, which travels with the gpurun snapshot like the built .so files do.
# Nothing here is compiled and nothing under oracle/_ref/ is ever committed or imported by the product package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${MUSEDIFFUSION_REFERENCE_SRC:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/MuseDiffusion" ]; then
    echo "build_ref.sh: no reference tree at $SRC (fine on the GPU box: oracle/_ref/ is prebuilt)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/MuseDiffusion" "$DST/MuseDiffusion"
cp -r "$SRC/commu" "$DST/commu"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
( cd "$SRC" && find MuseDiffusion commu -type f -name "*.py" | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
echo "build_ref.sh: copied $(find "$DST" -name '*.py' | wc -l) python files to $DST"
