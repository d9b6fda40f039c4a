#!/bin/bash
# fat col_classify / feat_scatter launches in the asynchronous pass: async parity test + bench (3 repeats)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r03i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03i_pytest_gpu.log; tail -4 gpurun_out/r03i_pytest_gpu.log | cut -c1-300
for r in 1 2 3; do
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r03i_bench$r.json 2> gpurun_out/r03i_bench$r.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r03i_bench$r.json').read().strip().splitlines()[0])
print('ms/step', d['ms_per_step'], d['stages_ms'], d['clocks'])
PY
done
