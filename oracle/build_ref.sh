#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own hot-path sources, where they lie under $HACC_REFERENCE
# (default /root/reference), into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun):
#   libhaccref.so       unmodified sources (VMAX = 16384 as shipped, RCBForceTree.cxx:921)
#   libhaccref_vmax.so  same, but the VMAX #define is rewritten *in a pipe* (no copy of the source is
#                       ever written) so clustered snapshots do not abort at assert(SIZE < VMAX).
# No reference source is copied into the repository.  Does not use the reference's build system.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${HACC_REFERENCE:-/root/reference}"
HF="$REF/src/halo_finder"
OUT="$HERE/_ref"
if [ ! -d "$HF" ]; then
  echo "build_ref.sh: $HF not found; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
# type flags of the reference build (src/halo_finder/include.mk:4); unthreaded tree build as in the
# production BG/Q configuration (src/env/bashrc.mira:12) so node numbering is deterministic (DFS).
FL="-O3 -fopenmp -fPIC -DID_64 -DPOSVEL_32 -DGRID_32 -DLONG_INTEGER -DRCB_UNTHREADED_BUILD -I$HERE/stubs -I$HF -w"
g++ $FL -c "$HF/RCBForceTree.cxx" -o "$OUT/obj/RCBForceTree.o"
sed 's/^#define VMAX 16384/#define VMAX 1048576/' "$HF/RCBForceTree.cxx" | \
  g++ $FL -x c++ -c - -o "$OUT/obj/RCBForceTree_vmax.o"
g++ $FL -c "$HF/ForceLaw.cxx" -o "$OUT/obj/ForceLaw.o"
g++ $FL -c "$HERE/stubs/partition_stub.cxx" -o "$OUT/obj/partition_stub.o"
gcc -O3 -fPIC -std=c99 -w -c "$HF/BGQCM.c" -o "$OUT/obj/BGQCM.o"
gcc -O3 -fPIC -w -c "$HF/bigchunk.c" -o "$OUT/obj/bigchunk.o"
g++ $FL -c "$HERE/ref_harness.cxx" -o "$OUT/obj/ref_harness.o"
COMMON="$OUT/obj/ForceLaw.o $OUT/obj/partition_stub.o $OUT/obj/BGQCM.o $OUT/obj/bigchunk.o $OUT/obj/ref_harness.o"
g++ -shared -fopenmp -o "$OUT/libhaccref.so" "$OUT/obj/RCBForceTree.o" $COMMON -lrt -lpthread
g++ -shared -fopenmp -o "$OUT/libhaccref_vmax.so" "$OUT/obj/RCBForceTree_vmax.o" $COMMON -lrt -lpthread
# libhaccref_cic.so: the reference's own cloud-in-cell loops (src/cpu/Particles.cxx: array_index, cic, inverse_cic).  The file
# as a whole needs MPI; the three function bodies are cut out of it *in a pipe* between our scaffolding (stubs/cic_pre.h,
# stubs/cic_post.cxx) -- again no reference text is written anywhere.
PX="$REF/src/cpu/Particles.cxx"
fn() { awk -v start="$1" '$0 ~ start {f=1} f {print} f && /^}/ {exit}' "$PX"; }
{ cat "$HERE/stubs/cic_pre.h"; fn '^inline int *$'; fn '^void Particles::cic\\(\\) \\{'; fn '^void Particles::inverse_cic\\('; cat "$HERE/stubs/cic_post.cxx"; } | \
  g++ -O3 -fopenmp -fPIC -w -x c++ -shared - -o "$OUT/libhaccref_cic.so"
rm -rf "$OUT/obj"
echo "built $OUT/libhaccref.so $OUT/libhaccref_vmax.so $OUT/libhaccref_cic.so"
