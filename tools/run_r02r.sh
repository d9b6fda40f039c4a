#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest_gpu.log; tail -4 gpurun_out/r02r_pytest_gpu.log | cut -c1-300
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-paths > /dev/null 2> gpurun_out/r02r_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02r_prof.err | tail -7 | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-paths > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
PY
for n in 20000 28284; do timeout 300 python tools/repro_grow.py $n $n 2>&1 | tail -1 | cut -c1-200; done
