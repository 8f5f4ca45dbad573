#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference core into oracle/_ref/.
#
# Compiles the reference's own C++ sources where they lie under $REF/src (the
# 16 translation units of $REF/setup.py:23-38) with g++ directly — the
# reference's build system is not run and no reference source is copied into
# this repository.  Outputs (git-ignored, but shipped to the GPU box):
#   oracle/_ref/libgiggles_ref.so   reference core + oracle/ref_driver.cpp (C entry points)
#   oracle/_ref/pyref/giggles/      (optional, --python) the reference's Cython modules giggles.core and
#                                   giggles.align, for generating tests/golden fixtures
# Flags follow SURVEY Appendix C: -std=c++11 (std::bind2nd), -include cstdint
# (GCC 13 needs it for genotype.h:54), asserts live (-UNDEBUG) like setup.py:14.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GIGGLES_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
WITH_PY=0
[[ "${1:-}" == "--python" ]] && WITH_PY=1

if [[ ! -d "$REF/src" ]]; then
    echo "build_ref: $REF/src not present; keeping whatever is already in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT/obj"
CXXFLAGS="-std=c++11 -include cstdint -O2 -fPIC -UNDEBUG -w"
UNITS="pedigree columnindexingiterator column entry graycodes read readset columniterator indexset genotype binomial phredgenotypelikelihoods genotypedistribution genotypehmm backwardcolumniterator transitionprobabilitycomputer"
for f in $UNITS; do
    if [[ ! -f "$OUT/obj/$f.o" || "$REF/src/$f.cpp" -nt "$OUT/obj/$f.o" ]]; then
        g++ $CXXFLAGS -c "$REF/src/$f.cpp" -o "$OUT/obj/$f.o" &
    fi
done
wait
g++ $CXXFLAGS -I"$REF/src" -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
OBJS=""
for f in $UNITS; do OBJS="$OBJS $OUT/obj/$f.o"; done
g++ -shared -o "$OUT/libgiggles_ref.so" "$OUT/obj/ref_driver.o" $OBJS
echo "build_ref: built $OUT/libgiggles_ref.so"

if [[ $WITH_PY == 1 ]]; then
    # The Cython module needs a writable package dir with a hand-written _version.py
    # (setuptools_scm is absent); that scratch copy lives in /tmp, only built
    # artefacts and the package's plain-Python helpers land in oracle/_ref/pyref.
    W="$(mktemp -d /tmp/giggles_ref_py.XXXXXX)"
    cp -r "$REF/giggles" "$W/giggles"; cp -r "$REF/src" "$W/src"; chmod -R u+w "$W"
    echo 'version = "0+ref"' > "$W/giggles/_version.py"
    PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
    EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
    (cd "$W" && cython --cplus -3 giggles/core.pyx -o "$W/core.cpp" >/dev/null 2>&1 \
             && cython --cplus -3 giggles/priorityqueue.pyx -o "$W/pq.cpp" >/dev/null 2>&1)
    g++ $CXXFLAGS -I"$PYINC" -I"$W/giggles" -c "$W/core.cpp" -o "$W/core_mod.o"
    g++ -shared -o "$W/giggles/core$EXT" "$W/core_mod.o" $OBJS
    g++ -std=c++11 -O2 -fPIC -w -I"$PYINC" -I"$W/giggles" -c "$W/pq.cpp" -o "$W/pq.o"
    g++ -shared -o "$W/giggles/priorityqueue$EXT" "$W/pq.o"
    rm -rf "$OUT/pyref"; mkdir -p "$OUT/pyref/giggles"
    # edit_distance (giggles/align.pyx): plain C Cython module, the pin of the batched device edit distance
    (cd "$W" && cython -3 giggles/align.pyx -o "$W/align.c" >/dev/null 2>&1)
    gcc -O2 -fPIC -w -I"$PYINC" -c "$W/align.c" -o "$W/align.o"
    gcc -shared -o "$W/giggles/align$EXT" "$W/align.o"
    cp "$W/giggles/core$EXT" "$W/giggles/priorityqueue$EXT" "$W/giggles/align$EXT" "$OUT/pyref/giggles/"
    for f in __init__.py _version.py variant.py coverage.py graph.py; do cp "$W/giggles/$f" "$OUT/pyref/giggles/"; done
    rm -rf "$W"
    PYTHONPATH="$OUT/pyref" python3 -c "from giggles.core import GenotypeHMM, ReadSet, Read, Pedigree, NumericSampleIds, PhredGenotypeLikelihoods, Genotype; print('build_ref: giggles.core (reference) imports')"
fi
