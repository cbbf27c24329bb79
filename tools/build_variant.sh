#!/bin/bash
# Build a kernel-variant copy of libeml_b200.so for A/B measurements:  tools/build_variant.sh NAME "-DEML_...=..."
# -> build/variants/libeml_NAME.so, selected at run time with EML_B200_LIB=build/variants/libeml_NAME.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
TMP=$(mktemp -d)
cp "$ROOT"/effective_model_learning_b200/csrc/*.cu "$ROOT"/effective_model_learning_b200/csrc/*.cuh "$ROOT"/effective_model_learning_b200/csrc/*.hpp "$TMP"/
mkdir -p "$ROOT/build/variants"
cd "$TMP"
pids=()
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I"$ROOT/include" -I. "$@" -c $f -o ${f%.cu}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/build/variants/libeml_$NAME.so" *.o -cudart static
rm -rf "$TMP"
echo "$ROOT/build/variants/libeml_$NAME.so"
