#!/usr/bin/env bash
# Builds the UNMODIFIED reference (xp3i4/linear) from the sources where they lie
# under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).
# No reference source is copied into this repo. Flags: SURVEY.md App. D.
#   - plain g++ (the reference's vendored FindOpenMP.cmake aborts under CMake>=4)
#   - "-include array": include/pmpfinder.h:71 uses std::array without <array>
#   - gap_extra.cpp is dead code with an undefined reference; not built.
set -euo pipefail
R=${LNR_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
B=$HERE/_ref
mkdir -p "$B"
if [ ! -d "$R/src" ]; then echo "reference not present at $R; using prebuilt $B" >&2; exit 0; fi
FLAGS="-include array -O3 -DNDEBUG -std=c++14 -fopenmp -w -fPIC -DSEQAN_HAS_ZLIB=1 -DSEQAN_HAS_OPENMP=1 \
 -DSEQAN_ENABLE_TESTING=0 -D_LARGEFILE_SOURCE -D_FILE_OFFSET_BITS=64 -DSEQAN_HAS_EXECINFO=1 \
 -I$R/include -I$R/external -I$R/seqan/include"
SRCS="base cords shape_extend index_util args_parser cluster_util pmpfinder gap_util gap align_util \
 f_io align_bands align_interface parallel_io mapper linear"
pids=()
for s in $SRCS; do
  if [ ! -f "$B/$s.o" ] || [ "$R/src/$s.cpp" -nt "$B/$s.o" ]; then
    g++ $FLAGS -c "$R/src/$s.cpp" -o "$B/$s.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ -fopenmp $(for s in $SRCS; do echo "$B/$s.o"; done) -lz -lpthread -lrt -o "$B/linear"
# function-level harness (our own code, links the reference objects except linear.o)
if [ -f "$HERE/ref_harness.cpp" ]; then
  g++ $FLAGS -c "$HERE/ref_harness.cpp" -o "$B/ref_harness.o"
  g++ -shared -fopenmp "$B/ref_harness.o" $(for s in $SRCS; do [ $s != linear ] && echo "$B/$s.o"; done) \
      -lz -lpthread -lrt -o "$B/libref_harness.so"
fi
# hybrid program for the drop-in test (tests/test_gpu_hybrid.py): the reference's own main, scheduler, mapGaps and
# writers, with createIndexDynamic + apxMap resolved to integration/lnr_seqan_shim.cpp (our code over liblnr_b200.so).
# The reference objects are used as built above; only copies of two of them get the two symbols weakened, so that the
# shim's strong definitions win at link time. No reference source is modified or copied.
SHIM=$HERE/../integration/lnr_seqan_shim.cpp
LIBDIR=$HERE/../linear_b200/csrc
if [ -f "$SHIM" ] && [ -f "$LIBDIR/liblnr_b200.so" ]; then
  APX=$(nm "$B/pmpfinder.o" | awk '$2=="T" && $3 ~ /^_Z6apxMapR12IndexDynamic/ {print $3}')
  CID=$(nm "$B/index_util.o" | awk '$2=="T" && $3 ~ /^_Z18createIndexDynamic/ {print $3}')
  objcopy --weaken-symbol="$APX" "$B/pmpfinder.o" "$B/pmpfinder_weak.o"
  objcopy --weaken-symbol="$CID" "$B/index_util.o" "$B/index_util_weak.o"
  g++ $FLAGS -I$HERE/../include -c "$SHIM" -o "$B/lnr_seqan_shim.o"
  g++ -fopenmp $(for s in $SRCS; do case $s in pmpfinder) echo "$B/pmpfinder_weak.o";; index_util) echo "$B/index_util_weak.o";; *) echo "$B/$s.o";; esac; done) \
      "$B/lnr_seqan_shim.o" -L"$LIBDIR" -llnr_b200 -Wl,-rpath,'$ORIGIN/../../linear_b200/csrc' -lz -lpthread -lrt -o "$B/linear_hybrid"
  echo "built $B/linear_hybrid"
fi
echo "built $B/linear"
