timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
tail -n 1 gpurun_out/r1_launches_bench.csv | cut -c1-120
