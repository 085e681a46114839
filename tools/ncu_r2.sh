#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 asked for (one launch each, warm), into $1 (default gpurun_out/ncu)
out=${1:-gpurun_out/ncu}; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 170 $NCU -k regex:sep_tma     -s 3 -o $out/r2_sep576_k5      python tools/prof_conv.py sep 128 32 32 576 576 5 3 2   > $out/sep576.log 2>&1
timeout 170 $NCU -k regex:patch_dense -s 3 -o $out/r2_patch_stem3x3  python tools/prof_conv.py conv 128 128 128 32 64 3 3 2 > $out/stem.log 2>&1
DH_NORES=1 timeout 170 $NCU -k regex:patch_dense -s 3 -o $out/r2_patch_regmap python tools/prof_conv.py conv 128 32 32 576 48 1 3 2 > $out/regmap.log 2>&1
DH_RES2=1 timeout 170 $NCU -k regex:pw_smallk -s 3 -o $out/r2_pw_smallk_fremap python tools/prof_conv.py conv 128 32 32 48 576 1 3 2 > $out/fremap.log 2>&1
timeout 170 $NCU -k regex:sam_stream  -s 3 -o $out/r2_softargmax2d_stream python tools/prof_sam.py 1024 2 2d > $out/sam2d.log 2>&1
timeout 170 $NCU -k regex:sam3d_stream -s 3 -o $out/r2_softargmax3d_stream python tools/prof_sam.py 256 2 3d > $out/sam3d.log 2>&1
ls -la $out
