#!/bin/bash
# state check after the container was re-created: full GPU suite + the bench exactly as the driver runs it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest_gpu.log; tail -6 gpurun_out/r02u_pytest_gpu.log | cut -c1-300
( time timeout 900 python bench.py ) > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02u_bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[0])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
print('cpu', d.get('cpu_baseline'))
PY
