#!/bin/bash
# slot payload verification in wl_fused2: suite + per-level profile + bench, payload on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_wloa.py -m gpu -q -x ) > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest_gpu.log; tail -3 gpurun_out/r02x_pytest_gpu.log | cut -c1-300
for v in 1 0; do
  export GRAKEL_B200_WL_PAYLOAD=$v
  echo "== payload $v"
  timeout 300 python tools/prof_wl.py 2>&1 | grep -A8 "wl_fused2 prof\|ms_features" | cut -c1-330 | tee gpurun_out/r02x_wl_prof_payload$v.txt
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r02x_bench_payload$v.json 2> gpurun_out/r02x_bench$v.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02x_bench_payload$v.json').read().strip().splitlines()[0])
print('payload=$v ms/step', d['ms_per_step'], d['stages_ms'])
PY
  for n in 20000 28284; do timeout 300 python tools/repro_grow.py $n $n 2>&1 | tail -1 | cut -c1-200; done
done
