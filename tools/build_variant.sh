#!/bin/bash
# Build a tuning variant of the library: tools/build_variant.sh NAME "-DFLAG ..." [file.cu ...]
# Recompiles the named .cu files (default fused_fwd_tmem.cu) with the extra flags and links them with the other objects of
# build/csrc/ into armnet_b200/tuning/libNAME.so (select it with ARMNET_B200_LIB).
set -e
cd "$(dirname "$0")/../armnet_b200/csrc"
NAME=$1; FLAGS=$2; shift 2 || true
FILES=${@:-fused_fwd_tmem.cu}
OUT=../../build/variant_$NAME
mkdir -p $OUT ../tuning
OBJS=""
for f in ../../build/csrc/*.o; do
  b=$(basename $f .o)
  skip=0
  for c in $FILES; do [ "$b" == "$(basename $c .cu)" ] && skip=1; done
  [ $skip == 0 ] && OBJS="$OBJS $f"
done
for c in $FILES; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v -DARMNET_MAX_WARPS=12 $FLAGS \
    -c $c -o $OUT/$(basename $c .cu).o 2> $OUT/$(basename $c .cu).ptxas.log || { cat $OUT/$(basename $c .cu).ptxas.log; exit 1; }
  OBJS="$OBJS $OUT/$(basename $c .cu).o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../tuning/lib$NAME.so $OBJS
echo built armnet_b200/tuning/lib$NAME.so
