#!/bin/bash
# follow-up to ablate_sep.sh: L2 prefetch distance of the input patches (1024: 2 K-blocks ahead, 4096: 4, 8192: 9)
out=${1:-gpurun_out/ablate_sep2.txt}
export DEEPHAR_B200_LIB=$PWD/deephar_b200/libdeephar_b200_ablate.so
: > $out
for bits in 0 1024 4096 8192 1056 4128 8224 1025 4097; do
  echo -n "dbg=$bits  " >> $out
  DH_DBG=$bits timeout 60 python tools/prof_conv.py sep 256 32 32 576 576 5 3 10 >> $out 2>&1
done
cat $out
