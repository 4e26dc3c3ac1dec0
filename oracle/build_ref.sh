#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *real* reference BuildGraph (OpenMP) as a parity checker / CPU baseline.
#
# Compiles /root/reference/src/BuildGraph/src/*.cpp (sources stay where they lie; nothing is copied into the repo)
# in a throw-away scratch directory with the two patches SURVEY.md section 8(c) documents:
#   1. compile fix: Common.h:68  SSTR() dynamic_cast on an rvalue stream is rejected by libstdc++ >= 11
#   2. canonical numbering: sort reads by fileIndex before IDs are assigned (before Dataset.cpp:133),
#      which is what the MPI siblings do (their loader has no OpenMP) and makes the output thread-count independent.
# Output: oracle/_ref/buildG (git-ignored, travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${DISCO_REFERENCE:-/root/reference}/src/BuildGraph/src"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT/buildG" >&2
  [ -x "$OUT/buildG" ] && exit 0 || exit 3
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/disco_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/*.cpp "$REF"/*.h "$TMP"/
chmod -R u+w "$TMP"
sed -i 's|^#define SSTR( x ).*|#define SSTR( x ) (static_cast< std::ostringstream \&\& >( std::ostringstream() << std::dec << x ).str())|' "$TMP/Common.h"
# insert the sort right before the "Assing ID's to the reads" loop
python3 - "$TMP/Dataset.cpp" <<'PY'
import sys
p = sys.argv[1]
src = open(p).read()
needle = "\tfor(UINT64 i = 0 ; i < reads->size(); i++) \t\t// Assing ID's to the reads."
assert src.count(needle) == 1, "Dataset.cpp numbering loop not found"
patch = "\tstd::sort(reads->begin(), reads->end(), [](Read*a, Read*b){return a->getFileIndex()<b->getFileIndex();});\n"
open(p, "w").write(src.replace(needle, patch + needle))
PY
GZ=""
if echo '#include <zlib.h>' | g++ -x c++ -fsyntax-only - 2>/dev/null; then GZ="-DINCLUDE_READGZ"; GZL="-lz"; else GZL=""; fi
g++ $GZ -Wno-sign-compare -fopenmp -std=c++11 -O3 -w -o "$OUT/buildG" "$TMP"/*.cpp $GZL
echo "build_ref: built $OUT/buildG"

# The first consumer of the hot path's files: parsimplify (src/SimplifyGraph), used only to check that our
# files are accepted and lead to the same contracted graph, and as the oracle of the GPU contraction (simplify.cu).  Same
# treatment: scratch copy, compile fixes (SSTR in Config.h:47 and OverlapGraphSimple.h:17, missing <cstdint> in Utils.cpp)
# plus one determinism fix (3. below).
SG="${DISCO_REFERENCE:-/root/reference}/src/SimplifyGraph/src"
TMP2="$(mktemp -d /tmp/disco_ref_build2.XXXXXX)"
trap 'rm -rf "$TMP" "$TMP2"' EXIT
cp -r "$SG"/. "$TMP2"/
chmod -R u+w "$TMP2"
for f in Config.h OverlapGraphSimple.h; do
  sed -i 's|^#define SSTR( x ).*|#define SSTR( x ) (static_cast< std::ostringstream \&\& >( std::ostringstream() << std::dec << x ).str())|' "$TMP2/$f"
done
sed -i '1i #include <cstdint>' "$TMP2/Utils.cpp"
# 3. determinism fix: EdgeSimple::copyEdge (EdgeSimple.cpp:46-61) copies an edge WITHOUT its two read lengths, and the copy
#    constructor initialises nothing else -- every composite edge grown from such a copy (OverlapGraphSimple.cpp:365, :408)
#    carries an uninitialised m_destinationLen, which is printed as part of the edge length (:672) and decides
#    removeParDeadEndNodes' "edge long enough" test (:181).  The unpatched binary's output depends on heap garbage (observed:
#    length = offset + 0 / + 2954, short tips kept at random); with the two members copied it is a function of its input.
python3 - "$TMP2/EdgeSimple.cpp" <<'PY'
import sys
p = sys.argv[1]
src = open(p).read()
needle = "\tm_source = edge.m_source;\n"
assert src.count(needle) == 1, "EdgeSimple::copyEdge not found"
open(p, "w").write(src.replace(needle, needle + "\tm_sourceLen = edge.m_sourceLen;\n\tm_destinationLen = edge.m_destinationLen;\n"))
PY
( cd "$TMP2" && g++ $GZ -Wno-sign-compare -fopenmp -std=c++11 -O3 -w -o "$OUT/parsimplify" Config.cpp DataSet.cpp EdgeSimple.cpp OverlapGraphSimple.cpp Read.cpp Utils.cpp dna.cpp mainParSimplify.cpp $GZL ) \
  && echo "build_ref: built $OUT/parsimplify" || echo "build_ref: parsimplify did not build (acceptance test will be skipped)"
