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
echo "built $B/linear"
