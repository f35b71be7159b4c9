#!/usr/bin/env bash
# build_ref.sh -- compile the UNMODIFIED FemTech reference (plus its vendored
# ParMETIS/METIS/GKlib and jsoncpp) from the sources where they lie under
# /root/reference into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.
#
# The reference's own build system is not run (it needs MPI, BLAS/LAPACK and
# Boost, none of which exist in this image).  Instead its source files are
# compiled directly with gcc against three small shims kept in this directory:
#   mpi.h + ftmpi.c   process-per-rank MPI over shared memory (+ ftmpirun)
#   blas_shim.c       naive dgemm_/dgemv_ (fixes the summation order)
# No reference source is copied into the repository; only objects/binaries are
# written, and only under oracle/_ref/ (git-ignored, travels with gpurun).
#
# Two flavours of the first-party code are built:
#   exact: -O2 -ffp-contract=off            parity pinning (bit-reproducible,
#                                           matches oracle/femtech_oracle.c)
#   fast : -O3 -DNDEBUG -march=x86-64-v3    CPU baseline timing (the
#          reference's Release flags are -O3 -DNDEBUG -march=native,
#          CMakeLists.txt:44; x86-64-v3 keeps the binary portable to the GPU
#          box's host CPU)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${FEMTECH_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present, nothing to do"; exit 0; fi
mkdir -p "$OUT/obj/tp" "$OUT/obj/exact" "$OUT/obj/fast" "$OUT/gen"
PM="$REF/third-party/parmetis-4.0.3"

# 1. jsoncpp (zip, no unzip binary in the image)
if [ ! -d "$OUT/jsoncpp-1.8.4" ]; then
  python3 -m zipfile -e "$REF/third-party/jsoncpp-1.8.4.zip" "$OUT" >/dev/null
fi
printf '#define GIT_BRANCH "oracle"\n#define GIT_COMMIT_HASH "reference-snapshot"\n' > "$OUT/gen/gitbranch.h"

compile_list() { # stdin: "<src> <obj> <compiler+flags...>"
  xargs -P "$JOBS" -L 1 bash -c 'src="$0"; obj="$1"; shift; if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then "$@" -c "$src" -o "$obj"; fi'
}

# 2. third-party C: GKlib, METIS, ParMETIS (flags: SURVEY.md 8c recipe)
TPFLAGS="gcc -O2 -w -fcommon -fPIC -DLINUX -D_FILE_OFFSET_BITS=64 -DNDEBUG -DNDEBUG2 -DHAVE_EXECINFO_H -DHAVE_GETLINE -std=gnu99 -I$HERE"
{
  for f in "$PM"/metis/GKlib/*.c; do echo "$f $OUT/obj/tp/gk_$(basename "$f" .c).o $TPFLAGS -I$PM/metis/GKlib"; done
  for f in "$PM"/metis/libmetis/*.c; do echo "$f $OUT/obj/tp/metis_$(basename "$f" .c).o $TPFLAGS -I$PM/metis/libmetis -I$PM/metis/GKlib -I$PM/metis/include"; done
  for f in "$PM"/libparmetis/*.c; do
    [ "$(basename "$f")" = frename.c ] && continue   # Fortran bindings need MPI_Comm_f2c
    echo "$f $OUT/obj/tp/pm_$(basename "$f" .c).o $TPFLAGS -I$PM/libparmetis -I$PM/include -I$PM/metis/GKlib -I$PM/metis/include"
  done
  for f in "$OUT"/jsoncpp-1.8.4/src/lib_json/*.cpp; do echo "$f $OUT/obj/tp/json_$(basename "$f" .cpp).o g++ -O2 -w -fPIC -I$OUT/jsoncpp-1.8.4/include"; done
  echo "$HERE/ftmpi.c $OUT/obj/tp/ftmpi.o gcc -O2 -fPIC -w -I$HERE"
  echo "$HERE/blas_shim.c $OUT/obj/tp/blas_shim.o gcc -O2 -fPIC -ffp-contract=off -w"
} | compile_list
rm -f "$OUT/libftref_tp.a"; ar rcs "$OUT/libftref_tp.a" "$OUT"/obj/tp/*.o

# 3. first-party reference sources, two flavours
INC="-I$HERE -I$OUT/gen -I$REF/include -I$OUT/jsoncpp-1.8.4/include -I$PM/include -I$PM/metis/include"
for flav in exact fast; do
  if [ "$flav" = exact ]; then FL="-O2 -ffp-contract=off"; else FL="-O3 -DNDEBUG -march=x86-64-v3 -mtune=generic"; fi
  {
    find "$REF/src" -name '*.cpp' | while read -r f; do
      o="$OUT/obj/$flav/$(echo "${f#$REF/src/}" | tr '/' '_' | sed 's/\.cpp$/.o/')"
      echo "$f $o g++ -std=c++11 -w -fPIC $FL $INC"
    done
    echo "$HERE/ref_dump.cpp $OUT/obj/$flav/ref_dump.o g++ -std=c++11 -w -fPIC $FL $INC"
    # the shipped benchmark driver, unmodified (BASELINE config 1)
    echo "$REF/examples/Benchmarking-Parallel/Benchmarking-Parallel.cpp $OUT/obj/$flav/driver_benchmarking_parallel.o g++ -std=c++11 -w -fPIC $FL $INC"
    echo "$REF/examples/ex9/ex9.cpp $OUT/obj/$flav/driver_ex9.o g++ -std=c++11 -w -fPIC $FL $INC"
  } | compile_list
  LIBOBJ=$(ls "$OUT"/obj/$flav/*.o | grep -v -e ref_dump.o -e driver_)
  rm -f "$OUT/libftref_$flav.a"; ar rcs "$OUT/libftref_$flav.a" $LIBOBJ
  for exe in ref_dump driver_benchmarking_parallel driver_ex9; do
    g++ -o "$OUT/${exe}_$flav" "$OUT/obj/$flav/$exe.o" "$OUT/libftref_$flav.a" "$OUT/libftref_tp.a" -lm
  done
done
gcc -O2 -o "$OUT/ftmpirun" "$HERE/ftmpirun.c"
echo "build_ref: done -> $OUT"
