timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['cpu_baseline']['value'])
for k in d['kernel_profile']: print(k)
PY
