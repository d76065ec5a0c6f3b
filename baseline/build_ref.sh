#!/bin/bash
# Build the UNMODIFIED reference (NVIDIA/mpi-acx, /root/reference) into baseline/_ref against
# the bundled mini-MPI (the image has no MPI). Nothing under /root/reference is edited: the
# tree is copied to a scratch dir because its Makefile writes objects next to the sources.
#   - MPI_ACX_MEMOPS_V2=1: the v1 memOps attribute no longer exists in CUDA 12
#   - gencode sm_100a via the reference Makefile's own NVCC_GENCODE knob (it defaults to sm_80); the "a" target is
#     needed because the shared pipeline-GEMM benchmark links a tcgen05 kernel against the library (RDC)
# Then build the shared benchmark/test programs (bench/*.c,*.cu — written against the common
# MPIX_* API) against the reference library, so both arms run the same harness.
set -e
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF_SRC=${REF_SRC:-/root/reference}
OUT=$REPO/baseline/_ref
SCRATCH=$(mktemp -d /tmp/acx_ref_build.XXXXXX)
GEN="-gencode=arch=compute_100a,code=sm_100a"
make -C "$REPO" lib >/dev/null
rm -rf "$OUT"; mkdir -p "$OUT"/{include,lib,bin,mpi/include,mpi/lib}
cp "$REPO/include/mpi.h" "$OUT/mpi/include/mpi.h"
cp "$REPO/lib/libacxmpi.a" "$OUT/mpi/lib/libmpi.a"
cp -r "$REF_SRC"/. "$SCRATCH"/
chmod -R u+w "$SCRATCH"
make -C "$SCRATCH" NVCC=nvcc MPI_HOME="$OUT/mpi" NVCC_GENCODE="$GEN" MPI_ACX_MEMOPS_V2=1 MPI_ACX_PARTITIONED_SUPPORT=1
cp "$SCRATCH/libmpi-acx.a" "$OUT/lib/"
cp "$SCRATCH/include/mpi-acx.h" "$OUT/include/"
# the reference's own tests that still compile with CUDA 12 (graph tests use the removed 5-arg cudaGraphInstantiate)
for t in ring ring-all ring-all-device; do
  nvcc -I"$OUT/include" -I"$OUT/mpi/include" $GEN -o "$OUT/bin/$t" "$SCRATCH/test/src/$t.c" -L"$OUT/lib" -L"$OUT/mpi/lib" -lmpi-acx -lmpi -lcuda -lpthread -ldl -lrt 2>&1 | tail -3 || echo "reference test $t did not build"
done
nvcc -dc -I"$OUT/include" -I"$OUT/mpi/include" $GEN -o "$SCRATCH/rp.o" "$SCRATCH/test/src/ring-partitioned.cu" && \
nvcc -I"$OUT/include" $GEN -o "$OUT/bin/ring-partitioned" "$SCRATCH/rp.o" -L"$OUT/lib" -L"$OUT/mpi/lib" -lmpi-acx -lmpi -lcuda -lpthread -ldl -lrt || echo "reference ring-partitioned did not build"
# shared benchmark programs against the reference library (RDC: its device Pready/Parrived are not inline)
for b in "$REPO"/bench/*.cu; do
  [ -e "$b" ] || continue
  n=$(basename "$b" .cu)
  nvcc -dc -O3 -std=c++17 -DACX_BENCH_REFERENCE=1 -I"$OUT/include" -I"$OUT/mpi/include" -I"$REPO/bench" $GEN -o "$SCRATCH/$n.o" "$b" && \
  nvcc -O3 $GEN -o "$OUT/bin/$n" "$SCRATCH/$n.o" -L"$OUT/lib" -L"$OUT/mpi/lib" -lmpi-acx -lmpi -lcuda -lpthread -ldl -lrt || echo "bench $n did not build against the reference"
done
rm -rf "$SCRATCH"
ls -la "$OUT/lib" "$OUT/bin"
