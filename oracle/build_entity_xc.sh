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
  *) echo "unknown flavour"; exit 1 ;;
esac
cd "$WORK"
CC=/usr/bin/gcc CXX=/usr/bin/g++ cmake -B build -D pgens="$PG" -D output=OFF -D mpi=OFF -D OFFLINE=ON $EXTRA \
   > "$OUT/cmake_configure.log" 2>&1 || { tail -30 "$OUT/cmake_configure.log"; exit 1; }
cmake --build build -j${EB_REF_JOBS:-8} > "$OUT/cmake_build.log" 2>&1 || { tail -40 "$OUT/cmake_build.log"; exit 1; }
for f in $(find build -name 'entity*.xc'); do n=$(basename "$f"); cp "$f" "$OUT/entity_${n##*_dump_}"; done
ls -la "$OUT"
