#!/bin/bash
# full single-GPU evidence run of the round: tests, bench, per-model profiles, layer table, ncu of the top kernels
out=${1:-gpurun_out/final}; mkdir -p $out
timeout 500 python -m pytest tests -q -m gpu --timeout 120 --timeout-method thread > $out/test_gpu_1.txt 2>&1; tail -3 $out/test_gpu_1.txt
timeout 400 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; tail -3 $out/bench.err
for w in reception2d reception3d spnet_penn spnet_ntu; do timeout 200 python tools/profile_model.py $w 256 > $out/prof_$w.txt 2>&1; head -3 $out/prof_$w.txt; done
bash tools/layers_r2.sh $out/layers.txt > /dev/null
timeout 100 python tools/prof_sam.py 4096 10 2d >> $out/layers.txt; timeout 100 python tools/prof_sam.py 256 10 3d >> $out/layers.txt; timeout 100 python tools/prof_sam.py 32 10 3d >> $out/layers.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file $out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-graph > $out/bench_under_ncu.log 2>&1
bash tools/ncu_r2.sh $out/ncu > /dev/null 2>&1
ls $out $out/ncu
