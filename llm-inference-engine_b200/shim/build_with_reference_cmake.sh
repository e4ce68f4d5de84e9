#!/bin/bash
# Build the REFERENCE's own tests and examples with the REFERENCE's own build system (root CMakeLists.txt, tests/unit_tests/CMakeLists.txt,
# examples/cpp/CMakeLists.txt -- all unmodified), with `src` replaced by this shim: shim/src/CMakeLists.txt defines every library name the
# reference's targets link (rmsnorm, linear, ..., llama_self_decoder, layer_weights) as a forwarder to libb200llm.so.
# A scratch tree of per-file symlinks is configured (per-file, so that the sources' relative includes "../../src/..." stay inside the tree);
# nothing is copied into the repository.  Output: shim/_ref_programs/cmake.d/<target> (git-ignored; travels to the GPU box).
#   usage: build_with_reference_cmake.sh [REF_DIR]      (default /root/reference)
set -u
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref_programs/cmake.d
[ -f "$REF/CMakeLists.txt" ] || { echo "no reference at $REF: nothing to build"; exit 0; }
command -v cmake > /dev/null || { echo "cmake not installed: skipped"; exit 0; }
TREE=$(mktemp -d /tmp/b200shim_cmake.XXXXXX)
trap 'rm -rf "$TREE"' EXIT
mkdir -p "$TREE/tests/unit_tests" "$TREE/examples/cpp" "$TREE/build"
ln -s "$REF/CMakeLists.txt" "$TREE/CMakeLists.txt"
ln -s "$HERE/src" "$TREE/src"
ln -s "$REF/tests/CMakeLists.txt" "$TREE/tests/CMakeLists.txt"
ln -s "$REF/examples/CMakeLists.txt" "$TREE/examples/CMakeLists.txt"
for f in "$REF"/tests/unit_tests/*; do ln -s "$f" "$TREE/tests/unit_tests/$(basename "$f")"; done
for f in "$REF"/examples/cpp/*; do ln -s "$f" "$TREE/examples/cpp/$(basename "$f")"; done
rm -rf "$OUT"; mkdir -p "$OUT"
( cd "$TREE/build" && cmake -DCMAKE_CUDA_COMPILER=${NVCC:-/usr/local/cuda/bin/nvcc} -DCMAKE_BUILD_TYPE=Release '-DCMAKE_BUILD_RPATH=$ORIGIN/../../../lib' -Wno-dev .. > "$OUT/configure.log" 2>&1 ) \
    || { echo "cmake configure failed (see $OUT/configure.log)"; tail -20 "$OUT/configure.log"; exit 1; }
( cd "$TREE/build" && cmake --build . -j 8 -- -k > "$OUT/build.log" 2>&1 ); rc=$?
n=0
for exe in $(find "$TREE/build/tests" "$TREE/build/examples" -maxdepth 3 -type f -perm -u+x ! -name '*.so' ! -name '*.a' ! -name '*.o' ! -name '*.cmake' ! -name '*.bin' 2>/dev/null | grep -v CMakeFiles); do
    cp "$exe" "$OUT/$(basename "$exe")"; n=$((n + 1))
done
if [ $rc -ne 0 ]; then echo "cmake --build failed for some targets (see $OUT/build.log)"; grep -m 10 -E "error|Error" "$OUT/build.log"; exit 1; fi
rm -f "$OUT/configure.log" "$OUT/build.log"
echo "the reference's own CMake build, src/ replaced by the shim: $n executables -> $OUT"
