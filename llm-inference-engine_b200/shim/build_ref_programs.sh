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

# ---- second configuration: the reference's OWN layer sources (src/layers/*.cpp + src/layers/includes/*.h, unmodified) on top of the shim's
#      launchers.  north_star: "keeps the repo's C++ launch-function ... API surface so it drops into src/layers ... unchanged".  In this tree
#      src/layers points at the reference, src/{kernels,utils,weights,memory,models} at the shim; the five layer examples are linked against
#      the resulting library.  Output: shim/_ref_programs/ref_layers.d/{libref_layers_on_b200.so,<example>}
TREE2=$(mktemp -d /tmp/b200shim_tree2.XXXXXX)
trap 'rm -rf "$TREE" "$TREE2"' EXIT
OUT2=$OUT/ref_layers.d
rm -rf "$OUT2"
mkdir -p "$TREE2/src/layers/includes" "$TREE2/examples/cpp" "$TREE2/obj" "$OUT2"
for d in kernels utils weights memory models; do ln -s "$HERE/src/$d" "$TREE2/src/$d"; done
for f in "$REF"/src/layers/*.cpp; do ln -s "$f" "$TREE2/src/layers/$(basename "$f")"; done
for f in "$REF"/src/layers/includes/*.h; do ln -s "$f" "$TREE2/src/layers/includes/$(basename "$f")"; done
for f in "$REF"/examples/cpp/*.cpp; do ln -s "$f" "$TREE2/examples/cpp/$(basename "$f")"; done
pids=()
for f in "$TREE2"/src/layers/*.cpp; do
    n=$(basename "$f" .cpp)
    ( "$NVCC" -std=c++17 -O2 -w -x cu -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$TREE2" -Xcompiler -fPIC -c "$f" -o "$TREE2/obj/$n.o" \
        > "$OUT2/$n.build.log" 2>&1 && rm -f "$OUT2/$n.build.log" ) & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
if ! ls "$OUT2"/*.build.log > /dev/null 2>&1; then
    "$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT2/libref_layers_on_b200.so" "$TREE2"/obj/*.o -L"$LIBDIR" -lb200llm -lcublas -lcublasLt \
        -Xlinker -rpath -Xlinker '$ORIGIN/../../../lib' > "$OUT2/libref_layers_on_b200.build.log" 2>&1 && rm -f "$OUT2/libref_layers_on_b200.build.log"
fi
if ! ls "$OUT2"/*.build.log > /dev/null 2>&1; then
    pids=()
    for f in "$TREE2"/examples/cpp/*.cpp; do
        n=$(basename "$f" .cpp)
        ( "$NVCC" -std=c++17 -O2 -w -x cu -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$TREE2" "$f" -o "$OUT2/$n" -L"$OUT2" -lref_layers_on_b200 \
            -L"$LIBDIR" -lb200llm -lcublas -lcublasLt -Xlinker -rpath -Xlinker '$ORIGIN' -Xlinker -rpath -Xlinker '$ORIGIN/../../../lib' \
            > "$OUT2/$n.build.log" 2>&1 && rm -f "$OUT2/$n.build.log" ) & pids+=($!)
    done
    for p in "${pids[@]}"; do wait "$p"; done
fi
ls "$OUT2"/*.build.log > /dev/null 2>&1 && { echo "the reference's layer sources did not build on the shim's launchers (see $OUT2/*.build.log)"; exit 1; }
echo "the reference's own src/layers/*.cpp + layer examples built on the shim's launchers -> $OUT2"
