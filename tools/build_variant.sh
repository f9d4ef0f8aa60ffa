#!/bin/bash
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS" file.cu [file.cu ...]
# Recompiles the named sources with extra flags and links them with the product's other objects into
# gpurun_variants/lib_NAME.so (a probe library for tools/variant_probe.sh; never the shipped one).
set -e
cd "$(dirname "$0")/../lineslam_b200/csrc"
make -s
name=$1; flags=$2; shift 2
mkdir -p build_var/$name ../../gpurun_variants
objs=""
for o in $(sed -n "s/^SRCS = //p" Makefile | sed "s/\([a-z_0-9]*\)\.cu/build\/\1.o/g"); do
  b=$(basename $o .o); skip=0
  for f in "$@"; do [ "$b.cu" = "$f" ] && skip=1; done
  [ $skip = 0 ] && objs="$objs $o"
done
for f in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -O3 -std=c++17 -Xcompiler -fPIC,-O2,-ffp-contract=off -ccbin /usr/bin/g++ $flags -c $f -o build_var/$name/$(basename $f .cu).o &
done
wait
for f in "$@"; do objs="$objs build_var/$name/$(basename $f .cu).o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/lib_$name.so $objs -ldl
echo built gpurun_variants/lib_$name.so
