#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/spattr_err.py > gpurun_out/r02e_spattr_err.txt 2>&1; cat gpurun_out/r02e_spattr_err.txt | tail -8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -8 gpurun_out/r02e_pytest_gpu.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu --no-paths > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02e_bench.err
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02e_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02e_prof.err | tail -7 | cut -c1-330
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['e2e']['last_step_ms'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'], 'node', d['config'].get('host_numa_node'))
PY
