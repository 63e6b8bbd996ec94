#!/bin/bash
# A/B builds of the library with extra -D flags: tools/build_variant.sh <out.so> [-DFLAG=...]...   (use with LPI_LIB_PATH=<out.so>)
set -e
out=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/variant_$$
pids=()
for f in lpi_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$f" -o build/variant_$$/$(basename "$f" .cu).o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out" build/variant_$$/*.o -ldl
rm -rf build/variant_$$
echo "built $out"
