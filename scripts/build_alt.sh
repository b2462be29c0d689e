#!/bin/bash
# A/B builds of the library with alternative compile-time switches (git-ignored, shipped to the GPU box by gpurun):
#   scripts/build_alt.sh <name> "<-D flags>"  ->  storygen_b200/csrc/alt/<name>/libstorygen_b200.so
# Select at run time with STORYGEN_B200_LIB=<path> (Python) or LD_LIBRARY_PATH=<dir> (csrc/selftest).
set -e
cd "$(dirname "$0")/../storygen_b200/csrc"
name=$1; shift
[ -n "$name" ] || { echo "usage: build_alt.sh <name> [flags]"; exit 2; }
flags="$*"
out="alt/$name"
obj="alt/$name/obj"
mkdir -p "$obj"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
for f in $(ls *.cu | grep -v selftest | sed 's/\.cu$//'); do
  $NV $flags -c $f.cu -o "$obj/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libstorygen_b200.so" "$obj"/*.o -cudart static
find "$obj" -name '*.o' -delete
rmdir "$obj"
echo "built $out/libstorygen_b200.so"
