#!/bin/bash
# Build a variant of libfdql.so for A/B runs: profiles/build_variant.sh <name> <file.cu> "<extra nvcc flags>"
# (only <file.cu> is recompiled with the flags; the other objects come from the in-tree build) -> profiles/variants/libfdql_<name>.so
set -e
name=$1; src=$2; extra=$3
cd "$(dirname "$0")/../fastdeepqlearning_b200/csrc"
mkdir -p ../../profiles/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr $extra -c $src -o /tmp/variant_$name.o
objs=""
for f in arena sample tqc hotpath vmap; do
  if [ "$f.cu" == "$src" ]; then objs="$objs /tmp/variant_$name.o"; else objs="$objs $f.o"; fi
done
$NVCC $ARCH -shared -o ../../profiles/variants/libfdql_$name.so $objs -lcudart
echo built profiles/variants/libfdql_$name.so
