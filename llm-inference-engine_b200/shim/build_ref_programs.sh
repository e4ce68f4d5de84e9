#!/bin/bash
# Compile the REFERENCE's own unit tests (tests/unit_tests/*.cu), layer examples (examples/cpp/*.cpp) and chat entry (user_entry.cpp), unmodified, against
# the shim headers of this directory and libb200llm.so.  This is the drop-in check of the boundary: the reference sources
# include "../../src/<...>" relative to their own location, so a scratch tree of SYMLINKS is laid out in which tests/ and
# examples/ point at the reference files and src/ points at shim/src -- nothing is copied into the repository.
# Output: shim/_ref_programs/<name> (git-ignored; travels to the GPU box with the snapshot).
#   usage: build_ref_programs.sh [REF_DIR]      (default /root/reference)
set -u
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$HERE/_ref_programs
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
LIBDIR=$ROOT/llm-inference-engine_b200/lib
[ -d "$REF/tests/unit_tests" ] || { echo "no reference at $REF: nothing to build"; exit 0; }
TREE=$(mktemp -d /tmp/b200shim_tree.XXXXXX)
trap 'rm -rf "$TREE"' EXIT
mkdir -p "$TREE/tests/unit_tests" "$TREE/examples/cpp" "$OUT"
ln -s "$HERE/src" "$TREE/src"
for f in "$REF"/tests/unit_tests/*.cu; do ln -s "$f" "$TREE/tests/unit_tests/$(basename "$f")"; done
for f in "$REF"/examples/cpp/*.cpp; do ln -s "$f" "$TREE/examples/cpp/$(basename "$f")"; done
[ -f "$REF/user_entry.cpp" ] && ln -s "$REF/user_entry.cpp" "$TREE/user_entry.cpp"   # the chat entry: includes "src/utils/model_utils.h"
fail=0
build_one() {  # $1 = source (in the symlink tree), $2 = output name
    "$NVCC" -std=c++17 -O2 -w -x cu -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$TREE" \
        "$1" -o "$OUT/$2" -L"$LIBDIR" -lb200llm -lcublas -lcublasLt -Xlinker -rpath -Xlinker '$ORIGIN/../../lib' > "$OUT/$2.build.log" 2>&1
    if [ $? -eq 0 ]; then rm -f "$OUT/$2.build.log"; echo "built  $2"; else echo "FAILED $2 (see $OUT/$2.build.log)"; fail=$((fail + 1)); fi
}
pids=()
for f in "$TREE"/tests/unit_tests/*.cu; do build_one "$f" "$(basename "$f" .cu)" & pids+=($!); done
for f in "$TREE"/examples/cpp/*.cpp; do build_one "$f" "$(basename "$f" .cpp)" & pids+=($!); done
[ -f "$TREE/user_entry.cpp" ] && { build_one "$TREE/user_entry.cpp" user_entry & pids+=($!); }
for p in "${pids[@]}"; do wait "$p"; done
ls "$OUT"/*.build.log > /dev/null 2>&1 && { echo "some reference programs did not compile against the shim"; exit 1; }
echo "all reference programs compiled against the shim -> $OUT"
