#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compiles the product's .cu sources for the HOST with
# the fiber emulation of tests/emu/cuda_emu.h into tests/emu/_build/libreveal_emu.so
# (same C-ABI as libreveal_b200.so) so kernel logic can be checked without a GPU.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$here/_build"
mkdir -p "$out"
CXXFLAGS="$RV_EMU_EXTRA -O1 -g -std=c++17 -fPIC -DRV_EMU -I$here -I$root/reveal_b200/csrc -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable"
for f in rv_api rv_sa rv_lcp rv_sweep rv_split rv_chain rv_tiny; do
  g++ $CXXFLAGS -x c++ -c "$root/reveal_b200/csrc/$f.cu" -o "$out/$f.o" &
done
g++ $CXXFLAGS -c "$here/cuda_emu.cpp" -o "$out/cuda_emu.o" &
wait
g++ -shared -o "$out/libreveal_emu.so" "$out"/*.o
echo "$out/libreveal_emu.so"
