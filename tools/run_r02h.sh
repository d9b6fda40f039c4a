#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fused_and_multikernel or transport or attr" > gpurun_out/r02h_pytest.log 2>&1; tail -12 gpurun_out/r02h_pytest.log | cut -c1-250
timeout 600 python bench.py --graphs 20000 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -3 | cut -c1-600
GRAKEL_B200_WL_V1=1 timeout 600 python bench.py --graphs 20000 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -2 | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench.err
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02h_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02h_prof.err | tail -7 | cut -c1-330
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['e2e']['last_step_ms'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
c5 = d['other_paths']['config5_spattr']; print('config5', c5['ms_per_step'], c5['stages_ms'], c5['roofline']['frac'])
PY
