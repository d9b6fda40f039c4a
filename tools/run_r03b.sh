#!/bin/bash
# asynchronous pass (gk_wl_gram): suite + bench async on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "asynchronous or config2 or wl" ) > gpurun_out/r03b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03b_pytest_gpu.log; tail -25 gpurun_out/r03b_pytest_gpu.log | cut -c1-300
for v in 0 1; do
  if [ $v = 1 ]; then export GRAKEL_B200_NO_ASYNC=1; fi
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths > gpurun_out/r03b_bench_noasync$v.json 2> gpurun_out/r03b_bench$v.err; echo "bench rc=$?"
  tail -3 gpurun_out/r03b_bench$v.err | cut -c1-300
  python - <<PY
import json
d = json.loads(open('gpurun_out/r03b_bench_noasync$v.json').read().strip().splitlines()[0])
print('NO_ASYNC=$v ms/step', d['ms_per_step'], d['stages_ms'], 'launches', d['gpu_launches'])
print('   e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
PY
done
