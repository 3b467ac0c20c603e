#!/bin/bash
# usage: tools/build_variants.sh name1 "<flags1>" name2 "<flags2>" ...
# Builds build/variants/<name>/libmlvfs_b200.so for each flag set HERE (nvcc cross-compiles): only $SRC (default
# fused; e.g. SRC=amaze for the AMaZE tile program) is recompiled, the other objects are the default build's.
# tools/run_variants.sh measures them on the GPU box.
set -e
cd "$(dirname "$0")/.."
SRC=${SRC:-fused}
make -s -C mlvfs_b200/csrc >/dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=build/variants/$name; mkdir -p $d
  nvcc $ARCH -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude $flags \
       -c mlvfs_b200/csrc/$SRC.cu -o $d/$SRC.o
  objs=$(ls build/obj/*.o | grep -v /$SRC.o)
  nvcc $ARCH -shared -cudart static -Xlinker --version-script=mlvfs_b200/csrc/exports.map -o $d/libmlvfs_b200.so $objs $d/$SRC.o -lpthread
  echo "$flags" > $d/flags.txt
  echo "built $name: $flags"
done
