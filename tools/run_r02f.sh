#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/spattr_err.py > gpurun_out/r02f_spattr_err.txt 2>&1; tail -7 gpurun_out/r02f_spattr_err.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
tail -8 gpurun_out/r02f_pytest_gpu.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
c5 = d['other_paths']['config5_spattr']; print('config5', c5['ms_per_step'], c5['stages_ms'], c5['roofline']['frac'])
PY
