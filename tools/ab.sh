#!/bin/bash
# usage (GPU box): tools/ab.sh TAG variant1 variant2 ...   (variants = suffixes of libchimera_b200_*.so, "default" = the product lib)
TAG=$1; shift
for v in "$@"; do
  echo "=== $v" >> gpurun_out/ab_$TAG.txt
  if [ "$v" = default ]; then unset CHB_LIB; else export CHB_LIB=$PWD/chimeracl_b200/libchimera_b200_$v.so; fi
  python tools/profile_step.py --time --steps 10 $AB_ARGS 2>&1 | head -${AB_LINES:-14} >> gpurun_out/ab_$TAG.txt
done
cat gpurun_out/ab_$TAG.txt
