#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -15 gpurun_out/r02d_pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02d_bench.err
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02d_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02d_prof.err | tail -8
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], 'api', d['e2e_api']['ms_per_step'])
print(json.dumps(d.get('other_paths'), indent=1)[:6000])
print(json.dumps(d.get('cpu_baseline'))[:600])
PY
