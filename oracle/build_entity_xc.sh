#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE: builds the UNMODIFIED reference (entity.xc) with its own cmake
# build from a scratch copy of /root/reference, once per (backend, deposit) flavour, with the
# dump-wrapper problem generators of oracle/pgens/. Only binaries are kept, under baseline/_ref/
# (git-ignored, travels to the GPU box). Recipe = BASELINE.md §3 / SURVEY.md §8(c).
#
#   oracle/build_entity_xc.sh omp      # Kokkos-OpenMP, zigzag: streaming, reconnection, turbulence (2D), magnetosphere, wald, accretion
#   oracle/build_entity_xc.sh omp3     # Kokkos-OpenMP, esirkepov shape_order=3: turbulence
#   oracle/build_entity_xc.sh cuda     # Kokkos-CUDA sm_100 (Kokkos_ARCH_BLACKWELL100): reconnection
#   oracle/build_entity_xc.sh cuda3    # Kokkos-CUDA sm_100, esirkepov 3: turbulence
#   oracle/build_entity_xc.sh cuda_shim # the cuda flavour with integration/eb200_shim.hpp patched in: the reference's
#                                        # engine calling libentity_b200.so (drop-in demonstration)
set -e
FLAVOUR=${1:-omp}
REPO=$(cd "$(dirname "$0")/.." && pwd)
WORK=${EB_REF_WORK:-/tmp/eb_refbuild}/$FLAVOUR
OUT=$REPO/baseline/_ref/$FLAVOUR
[ -d /root/reference ] || { echo "no /root/reference here"; exit 0; }
mkdir -p "$WORK" "$OUT"
if [ ! -f "$WORK/CMakeLists.txt" ]; then
  (cd /root/reference && tar cf - --exclude=extern/adios2 .) | (cd "$WORK" && tar xf -)
  chmod -R u+w "$WORK"
fi
P=$REPO/oracle/pgens
case $FLAVOUR in
  omp)   PG="$P/dump_streaming;$P/dump_reconnection;$P/dump_turbulence;$P/dump_magnetosphere;$P/dump_wald;$P/dump_accretion"
         EXTRA="-D Kokkos_ENABLE_OPENMP=ON" ;;
  omp3)  PG="$P/dump_turbulence"
         EXTRA="-D Kokkos_ENABLE_OPENMP=ON -D deposit=esirkepov -D shape_order=3" ;;
  cuda)  PG="$P/dump_reconnection;$P/dump_streaming"
         EXTRA="-D Kokkos_ENABLE_CUDA=ON -D Kokkos_ARCH_BLACKWELL100=ON" ;;
  cuda3) PG="$P/dump_turbulence"
         EXTRA="-D Kokkos_ENABLE_CUDA=ON -D Kokkos_ARCH_BLACKWELL100=ON -D deposit=esirkepov -D shape_order=3" ;;
  cuda_shim)
         # the Kokkos-CUDA build with the SRPIC Minkowski dispatchers handed to libentity_b200.so
         # (integration/): starts from the cuda flavour's scratch tree (Kokkos already built)
         SRC="${EB_REF_WORK:-/tmp/eb_refbuild}/cuda"
         if [ ! -d "$WORK/build" ] && [ -d "$SRC/build" ]; then
           rm -rf "$WORK"; cp -a "$SRC" "$WORK"
           # a copied cmake build tree is bound to its path: rebind it (objects keep their timestamps)
           grep -rlI "$SRC" "$WORK/build" | xargs sed -i "s#$SRC\\b#$WORK#g"
         fi
         python "$REPO/integration/apply_shim.py" "$WORK"
         PG="$P/dump_reconnection;$P/dump_streaming"
         EXTRA="-D Kokkos_ENABLE_CUDA=ON -D Kokkos_ARCH_BLACKWELL100=ON"
         SHIM_LD="-L$REPO/entity_b200 -lentity_b200 -Wl,-rpath,\$ORIGIN/../../../entity_b200 -Wl,-rpath,$REPO/entity_b200" ;;
  *) echo "unknown flavour"; exit 1 ;;
esac
cd "$WORK"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -B build -D pgens="$PG" -D output=OFF -D mpi=OFF -D OFFLINE=ON $EXTRA \
   ${SHIM_CXX:+-D CMAKE_CXX_FLAGS="$SHIM_CXX"} ${SHIM_LD:+-D CMAKE_EXE_LINKER_FLAGS="$SHIM_LD"} \
   > "$OUT/cmake_configure.log" 2>&1 || { tail -30 "$OUT/cmake_configure.log"; exit 1; }
cmake --build build -j${EB_REF_JOBS:-8} > "$OUT/cmake_build.log" 2>&1 || { tail -40 "$OUT/cmake_build.log"; exit 1; }
for f in $(find build -name 'entity*.xc'); do n=$(basename "$f"); cp "$f" "$OUT/entity_${n##*_dump_}"; done
ls -la "$OUT"
