#!/usr/bin/env bash
# Builds the REFERENCE's Cython module `giggles.core` with src/genotypehmm.{h,cpp} replaced by the shim in this
# directory and linked against giggles_b200/libgiggles_b200.so — exactly what a Giggles maintainer would do
# (INTEGRATION.md).  Runs in a scratch copy of the reference tree; only built artefacts and the package's plain
# Python helpers land in integration/_build/giggles/ (git-ignored, shipped to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${GIGGLES_REFERENCE:-/root/reference}"
OUT="$HERE/_build"
if [[ ! -d "$REF/src" ]]; then
    echo "build_dropin: $REF not present; keeping whatever is already in $OUT" >&2
    exit 0
fi
W="$(mktemp -d /tmp/giggles_dropin.XXXXXX)"
trap 'rm -rf "$W"' EXIT
cp -r "$REF/giggles" "$W/giggles"; cp -r "$REF/src" "$W/src"; chmod -R u+w "$W"
cp "$HERE/genotypehmm.h" "$HERE/genotypehmm.cpp" "$W/src/"          # <- the change to the reference tree: the HMM ...
cp "$HERE/readselect.pyx" "$W/giggles/readselect.pyx"                # <- ... and read selection
echo 'version = "0+b200"' > "$W/giggles/_version.py"
CXXFLAGS="-std=c++11 -include cstdint -O2 -fPIC -UNDEBUG -w"
PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
# the host core the bindings need; the reference's column/iterator/graycode/transition units are no longer linked
UNITS="pedigree entry read readset indexset genotype binomial phredgenotypelikelihoods genotypedistribution genotypehmm"
mkdir -p "$W/obj"
for f in $UNITS; do
    g++ $CXXFLAGS -I"$ROOT/include" -c "$W/src/$f.cpp" -o "$W/obj/$f.o" &
done
wait
(cd "$W" && cython --cplus -3 giggles/core.pyx -o "$W/core.cpp" >/dev/null 2>&1 \
         && cython --cplus -3 giggles/priorityqueue.pyx -o "$W/pq.cpp" >/dev/null 2>&1)
g++ $CXXFLAGS -I"$PYINC" -I"$W/giggles" -I"$ROOT/include" -c "$W/core.cpp" -o "$W/core_mod.o"
rm -rf "$OUT"; mkdir -p "$OUT/giggles"
g++ -shared -o "$OUT/giggles/core$EXT" "$W/core_mod.o" "$W"/obj/*.o \
    -L"$ROOT/giggles_b200" -lgiggles_b200 -Wl,-rpath,'$ORIGIN/../../../giggles_b200'
g++ -std=c++11 -O2 -fPIC -w -I"$PYINC" -I"$W/giggles" -c "$W/pq.cpp" -o "$W/pq.o"
g++ -shared -o "$OUT/giggles/priorityqueue$EXT" "$W/pq.o"
for f in __init__.py _version.py variant.py coverage.py graph.py; do cp "$W/giggles/$f" "$OUT/giggles/"; done
PYTHONPATH="$OUT" python3 -c "from giggles.core import GenotypeHMM, ReadSet, Read, Pedigree, NumericSampleIds, PhredGenotypeLikelihoods, Genotype; print('build_dropin: giggles.core with the B200 GenotypeHMM imports')"
