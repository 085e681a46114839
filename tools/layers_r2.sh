#!/bin/bash
# times the dense layers of the headline model with the patch kernel (DH_PATCH=1) and the register producer (0)
out=${1:-gpurun_out/layers.txt}
: > $out
for shape in "256 128 128 32 64 3" "256 128 128 32 32 3" "256 64 64 64 96 3" "256 32 32 576 48 1" "256 16 16 288 576 1" \
             "256 16 16 576 288 1" "256 64 64 160 64 1" "256 32 32 384 576 1" "256 32 32 144 288 3" "256 128 128 48 96 3"; do
  for pk in 1 0; do
    echo -n "patch=$pk " >> $out
    DH_PATCH=$pk timeout 120 python tools/prof_conv.py conv $shape 3 10 >> $out 2>&1 || echo "FAILED rc=$?" >> $out
  done
done
cat $out
