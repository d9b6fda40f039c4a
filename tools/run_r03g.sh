#!/bin/bash
# batched tile staging in wl_fused2: parity subset + per-level profile at 10000 / 14142 / 28284 graphs + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_wloa.py tests/test_next_rows.py -m gpu -q -x ) > gpurun_out/r03g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03g_pytest_gpu.log; tail -4 gpurun_out/r03g_pytest_gpu.log | cut -c1-300
for n in 10000 14142 28284; do timeout 300 python tools/prof_wl_n.py $n 2>&1 | grep -A8 "wl_fused2 prof\|ms_features" | cut -c1-330 | tee gpurun_out/r03g_wl_prof_$n.txt; done
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r03g_bench.json 2> gpurun_out/r03g_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/r03g_bench.json').read().strip().splitlines()[0])
print('ms/step', d['ms_per_step'], d['stages_ms'])
PY
