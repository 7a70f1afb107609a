#!/bin/bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference `genmap` binary from the read-only sources
# under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
#
# The reference's own build system (cmake + FindSeqAn) is NOT used: the program is one translation
# unit (src/genmap.cpp) over the vendored header-only SeqAn, so we compile that file where it lies.
# Flags mirror /root/reference/src/CMakeLists.txt:28-38,81-85 with -march=native replaced by the
# portable "-msse4.2 -mpopcnt" (README.rst:60-62) so the binary also runs on the GPU box's host CPU.
# Takes ~15 min and ~4 GB RSS (320 template instantiations of computeMappability); never rebuilt when
# the binary already exists.  Nothing from /root/reference is copied into the repository.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GENMAP_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
BIN="$OUT/genmap_ref"
mkdir -p "$OUT"
if [ -x "$BIN" ] && [ "${1:-}" != "--force" ]; then
    echo "oracle/_ref/genmap_ref already built"; exit 0
fi
# A binary produced by exactly this recipe during the survey may be stashed under baseline/_ref/.
if [ "${1:-}" != "--force" ] && [ -x "$HERE/../baseline/_ref/genmap_ref_sse4" ]; then
    cp "$HERE/../baseline/_ref/genmap_ref_sse4" "$BIN"
    echo "oracle/_ref/genmap_ref taken from baseline/_ref/genmap_ref_sse4 (same recipe)"; exit 0
fi
if [ ! -f "$REF/src/genmap.cpp" ]; then
    echo "reference sources not present at $REF; cannot build oracle/_ref" >&2; exit 3
fi
g++ -std=c++14 -O3 -DNDEBUG -msse4.2 -mpopcnt -fopenmp \
    -DSEQAN_APP_VERSION='"1.3.0"' -DCMAKE_BUILD_TYPE='"Release"' -DSEQAN_HAS_OPENMP=1 \
    -DSEQAN_DISABLE_VERSION_CHECK=YES -D_FILE_OFFSET_BITS=64 -D_LARGEFILE_SOURCE -w \
    -I"$REF/include/seqan/include" "$REF/src/genmap.cpp" -o "$BIN.tmp" -lpthread -lrt
mv "$BIN.tmp" "$BIN"
echo "built $BIN"
