#!/bin/sh
# Build the reference's own assignment solver (vendored lp_solve 5.5.2.5 + lp_transbig_edit)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored).
# TEST INFRASTRUCTURE ONLY.  No reference source is copied into the repo: the C function
# lp_transbig_edit (src/my_lpsolve.cpp:34-122, plain C inside an Rcpp file) is extracted at
# build time into the ignored output directory.  Include flags follow src/Makevars:1.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${BMM_REFERENCE:-/root/reference}
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OUT/obj"
S="$REF/src"
INC="-I$S/ls_source -I$S/ls_source/bfp -I$S/ls_source/bfp/bfp_LUSOL -I$S/ls_source/bfp/bfp_LUSOL/LUSOL -I$S/ls_source/colamd -I$S/ls_source/shared"
{
  echo '#include <stdlib.h>'
  echo '#include "lp_lib.h"'
  echo 'void lp_transbig_edit(int, int, double *, double *);'
  sed -n '34,122p' "$S/my_lpsolve.cpp"
} > "$OUT/lp_transbig_edit_extracted.c"
OBJS=""
for f in "$S"/*.c "$OUT/lp_transbig_edit_extracted.c"; do
  o="$OUT/obj/$(basename "$f" .c).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
    gcc -O2 -fPIC -w $INC -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
gcc -shared -o "$OUT/liblpsolve_ref.so" $OBJS -lm -ldl
echo "build_ref: built $OUT/liblpsolve_ref.so"

# The reference's own samplers: every first-party C++ file of /root/reference/src, UNMODIFIED and compiled
# where it lies, against the header-only Rcpp/Armadillo stand-in in oracle/shim/ (R, Rcpp and Armadillo are
# not in this image).  -O2 is R's default optimisation level.  Linked with the lp_solve objects above.
mkdir -p "$OUT/objcpp"
COBJS=""
for f in full_gibbs stickbreaking collapsed_gibbs collapsed_gibbs_dp stephens utils my_lpsolve RcppExports; do
  o="$OUT/objcpp/$f.o"
  if [ ! -f "$o" ] || [ "$S/$f.cpp" -nt "$o" ] || [ "$HERE/shim/RcppArmadillo.h" -nt "$o" ] || [ "$HERE/rrng.h" -nt "$o" ]; then
    g++ -O2 -std=gnu++17 -fPIC -w -I"$HERE/shim" -I"$S" $INC -c "$S/$f.cpp" -o "$o" &
  fi
  COBJS="$COBJS $o"
done
g++ -O2 -std=gnu++17 -fPIC -Wall -I"$HERE/shim" -I"$S" $INC -c "$HERE/shim/ref_capi.cpp" -o "$OUT/objcpp/ref_capi.o" &
wait
for o in $COBJS "$OUT/objcpp/ref_capi.o"; do
  [ -f "$o" ] || { echo "build_ref: $o did not compile"; exit 1; }
done
LPOBJS=""
for f in "$S"/*.c; do LPOBJS="$LPOBJS $OUT/obj/$(basename "$f" .c).o"; done
g++ -shared -o "$OUT/libbmm_ref.so" $COBJS "$OUT/objcpp/ref_capi.o" $LPOBJS -lm -ldl
echo "build_ref: built $OUT/libbmm_ref.so"
