#!/bin/bash
# TEST INFRASTRUCTURE ONLY: memory check of the kernel logic without a GPU.  Builds the emulator library with
# -fsanitize=address into a scratch directory, points the loader at it and runs tests/emul/asan_driver.py.
# (compute-sanitizer is not available on the GPU pool; this catches the same class of indexing bugs on the host.)
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../quantum_inferno_b200/csrc"
out="${TMPDIR:-/tmp}/qi_emul_asan"
mkdir -p "$out"
pids=""
for f in "$src"/*.cu; do
  g++ -std=c++17 -O1 -g -fPIC -fsanitize=address -fno-omit-frame-pointer -DQI_EMUL -I"$here" -I"$src" -x c++ -c "$f" \
      -o "$out/$(basename "$f" .cu).o" &
  pids="$pids $!"
done
g++ -std=c++17 -O1 -g -fPIC -fsanitize=address -fno-omit-frame-pointer -DQI_EMUL -I"$here" -c "$here/cuda_emul.cpp" -o "$out/cuda_emul.o" &
pids="$pids $!"
for p in $pids; do wait "$p"; done
g++ -shared -fsanitize=address -o "$out/libqi_emul.so" "$out"/*.o -lpthread
QI_EMUL_LIB="$out/libqi_emul.so" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 \
  LD_PRELOAD="$(gcc -print-file-name=libasan.so)" python "$here/asan_driver.py"
