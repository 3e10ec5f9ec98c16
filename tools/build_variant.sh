#!/bin/bash
# usage: tools/build_variant.sh NAME file.cu "-DFLAG=1 ..."   ->  chimeracl_b200/libchimera_b200_NAME.so
# (A/B experiments: the variant is picked at run time with CHB_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; FLAGS=$3
B=chimeracl_b200/build
mkdir -p $B/var_$NAME
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include $FLAGS -Xptxas -v \
  -c chimeracl_b200/csrc/$SRC -o $B/var_$NAME/${SRC%.cu}.o 2> $B/var_$NAME/ptxas.log
OBJS=""
for o in $B/*.o; do
  if [ "$(basename $o)" = "${SRC%.cu}.o" ]; then OBJS="$OBJS $B/var_$NAME/${SRC%.cu}.o"; else OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o chimeracl_b200/libchimera_b200_$NAME.so $OBJS -lcudart
echo built chimeracl_b200/libchimera_b200_$NAME.so
