#!/bin/bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference HMC core (15 TUs, everything except
# HMC.cpp / Options.cpp / Main.cpp, which need boost::program_options) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored).  No reference source is copied into
# the repo: two TUs are streamed through sed on their way into the compiler to fix constructs
# g++ rejects (neither touches arithmetic):
#   PatternTree.cpp:9   static member definition of a template specialisation needs `template<>`
#   HaploFile.cpp:162   `filename` (undeclared) -> m_filename.c_str()
# Mixed-type min/max calls (HaploBuilder.cpp:320-321, PatternManager.cpp:57,328) are satisfied by
# overloads in shim/prelude.h.  Output: oracle/_ref/hmc_ref (driver = oracle/ref_driver.cpp).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${HMC_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then
  if [ -x "$OUT/hmc_ref" ]; then echo "build_ref: $REF absent, keeping prebuilt $OUT/hmc_ref"; exit 0; fi
  echo "build_ref: $REF absent and no prebuilt binary" >&2; exit 1
fi
CXX="${CXX:-g++}"
FLAGS="-std=gnu++14 -O2 -w -I$HERE/shim -I$REF -include $HERE/shim/prelude.h"
TUS="Allele Constant GenoData Genotype HaploBuilder HaploComp HaploData HaploFile HaploModel HaploPair HaploPattern Haplotype PatternManager PatternTree Utils"
OBJS=""
for t in $TUS; do
  src="$REF/$t.cpp"; obj="$OUT/$t.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$0" -nt "$obj" ]; then
    case $t in
      PatternTree) sed 's/^boost::pool<> PatternNode::m_pool/template<> &/' "$src" | $CXX $FLAGS -x c++ -c - -o "$obj" ;;
      HaploFile)   sed 's/Can not open file %s!", filename);/Can not open file %s!", m_filename.c_str());/' "$src" | $CXX $FLAGS -x c++ -c - -o "$obj" ;;
      *)           $CXX $FLAGS -c "$src" -o "$obj" ;;
    esac
  fi
  OBJS="$OBJS $obj"
done
$CXX $FLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/ref_driver.o"
$CXX -O2 -o "$OUT/hmc_ref" "$OUT/ref_driver.o" $OBJS
echo "build_ref: built $OUT/hmc_ref"
