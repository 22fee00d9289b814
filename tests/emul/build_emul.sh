#!/bin/bash
# TEST INFRASTRUCTURE ONLY: builds tests/emul/libqi_emul.so, a g++ build of the kernel sources
# on top of the fiber emulator in this directory, for debugging kernel logic without a GPU.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../quantum_inferno_b200/csrc"
out="$here/libqi_emul.so"
objs=""
pids=""
mkdir -p "$here/build"
for f in "$src"/*.cu; do
  o="$here/build/$(basename "$f" .cu).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$src" "$here" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ]; then
    g++ -std=c++17 -O2 -g -fPIC -DQI_EMUL -I"$here" -I"$src" -x c++ -c "$f" -o "$o" &
    pids="$pids $!"
  fi
  objs="$objs $o"
done
o="$here/build/cuda_emul.o"
if [ ! -f "$o" ] || [ "$here/cuda_emul.cpp" -nt "$o" ] || [ "$here/cuda_emul.h" -nt "$o" ]; then
  g++ -std=c++17 -O2 -g -fPIC -DQI_EMUL -I"$here" -c "$here/cuda_emul.cpp" -o "$o" &
  pids="$pids $!"
fi
for p in $pids; do wait "$p" || { echo "emulator compile failed" >&2; exit 1; }; done
g++ -shared -o "$out" $objs "$o" -lpthread
echo "built $out"
