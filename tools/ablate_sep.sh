#!/bin/bash
# Timing ablation of the dominant kernel (conv_sep.cu, 576 -> 576 5x5 on 32x32 maps, 256 frames): which pipeline stage is
# the limiter?  Needs the ablation build next to the shipped library:  make -C deephar_b200/csrc ABLATE=1
# Results are WRONG numerically when a bit is set; only the timings mean something.
#   1 no depthwise math   8 no DSMEM push   16 no weight TMA   32 epilogue without global traffic   64 no MMA issue
out=${1:-gpurun_out/ablate_sep.txt}
export DEEPHAR_B200_LIB=$PWD/deephar_b200/libdeephar_b200_ablate.so
: > $out
for bits in 0 1 32 64 33 65 96 97 16 8; do
  echo -n "dbg=$bits  " >> $out
  DH_DBG=$bits timeout 60 python tools/prof_conv.py sep 256 32 32 576 576 5 3 10 >> $out 2>&1
done
echo "--- 16x16 288->288 (no cluster sharing)" >> $out
for bits in 0 1 32 64 97; do
  echo -n "dbg=$bits  " >> $out
  DH_DBG=$bits timeout 60 python tools/prof_conv.py sep 256 16 16 288 288 5 3 10 >> $out 2>&1
done
cat $out
