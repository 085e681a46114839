#!/bin/bash
# times the hot layers of the headline model (256-frame launches) through the C ABI: dense layers with the
# patch kernel (DH_PATCH=1) and the register producer (0); separable layers; fReMap on both kernels
out=${1:-gpurun_out/layers.txt}
: > $out
for shape in "256 128 128 32 64 3" "256 128 128 32 32 3" "256 64 64 64 96 3" "256 32 32 576 48 1" "256 16 16 288 576 1" \
             "256 16 16 576 288 1" "256 64 64 160 64 1" "256 32 32 384 576 1" "256 32 32 144 288 3" "256 128 128 48 96 3"; do
  for pk in 1 0; do
    echo -n "patch=$pk " >> $out
    DH_PATCH=$pk timeout 60 python tools/prof_conv.py conv $shape 3 10 >> $out 2>&1 || echo "FAILED rc=$?" >> $out
  done
done
for shape in "256 32 32 576 576 5" "256 16 16 288 288 5" "256 16 16 288 576 5" "256 8 8 288 288 5" "256 32 32 384 576 3"; do
  timeout 60 python tools/prof_conv.py sep $shape 3 10 >> $out 2>&1 || echo "FAILED rc=$?" >> $out
done
for pw in 1 0; do
  echo -n "fReMap pw_smallk=$pw " >> $out
  DH_PWSMALLK=$pw DH_RES2=1 timeout 60 python tools/prof_conv.py conv 256 32 32 48 576 1 3 10 >> $out 2>&1 || echo "FAILED rc=$?" >> $out
done
cat $out
