timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 3000 gpurun_out/bench_r1.json | head -c 700; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
tail -n 2 gpurun_out/r1_launches_bench.csv | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:sep_tma -s 3 -c 1 -o gpurun_out/r1_septma_final -f python tools/prof_conv.py sep 128 32 32 576 576 5 3 2 > gpurun_out/septma_final.log 2>&1
tail -n 1 gpurun_out/septma_final.log
