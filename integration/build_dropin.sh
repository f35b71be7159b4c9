#!/usr/bin/env bash
# build_dropin.sh -- link the reference's UNMODIFIED example drivers against femtech_b200.
#   driver objects : compiled from /root/reference/examples/... by oracle/ref/build_ref.sh (unchanged sources)
#   hot path       : integration/femtech_host.o  ->  libftb200.so (CUDA)
#   everything else: the reference's own objects (readers, PartitionMesh, log, VTU, ParMETIS) from oracle/_ref
# femtech_host.o comes first, so the archive members of the replaced translation units are never pulled.
# Output: oracle/_ref/dropin_benchmarking_parallel, dropin_ex9, dropin_resident (integration/resident_driver.cpp), dropin_ref_dump (the oracle harness driver, which
# also exercises CalculateMaximumPrincipalStrain / the injury loop in legacy mode); git-ignored, travel with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$HERE/.."
REF="${FEMTECH_REFERENCE:-/root/reference}"
OUT="$ROOT/oracle/_ref"
if [ ! -d "$REF/src" ] || [ ! -f "$OUT/libftref_fast.a" ]; then echo "build_dropin: reference build not present, nothing to do"; exit 0; fi
PM="$REF/third-party/parmetis-4.0.3"
INC="-I$ROOT/oracle/ref -I$OUT/gen -I$REF/include -I$OUT/jsoncpp-1.8.4/include -I$PM/include -I$PM/metis/include -I$ROOT/include"
g++ -std=c++11 -O2 -w -fPIC $INC -c "$HERE/femtech_host.cpp" -o "$OUT/obj/femtech_host.o"
for drv in benchmarking_parallel ex9 ref_dump; do
  obj="$OUT/obj/fast/driver_$drv.o"; [ "$drv" = ref_dump ] && obj="$OUT/obj/fast/ref_dump.o"
  g++ -o "$OUT/dropin_$drv" "$obj" "$OUT/obj/femtech_host.o" "$OUT/libftref_fast.a" "$OUT/libftref_tp.a" \
      -L"$ROOT/femtech_b200" -lftb200 -Wl,-rpath,'$ORIGIN/../../femtech_b200' -lm
done
# a driver of our own that uses the resident mode through the reference's API names (ExplicitDynamics), see the file header
g++ -std=c++11 -O2 -w -fPIC $INC -c "$HERE/resident_driver.cpp" -o "$OUT/obj/resident_driver.o"
g++ -o "$OUT/dropin_resident" "$OUT/obj/resident_driver.o" "$OUT/obj/femtech_host.o" "$OUT/libftref_fast.a" "$OUT/libftref_tp.a" \
    -L"$ROOT/femtech_b200" -lftb200 -Wl,-rpath,'$ORIGIN/../../femtech_b200' -lm
echo "build_dropin: done"
