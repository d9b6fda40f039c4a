#!/bin/bash
# merged Gram prologue (one host sync per pass), cached tile lists: suite + A/B bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest_gpu.log; tail -6 gpurun_out/r02v_pytest_gpu.log | cut -c1-300
for v in 0 1; do
  if [ $v = 1 ]; then export GRAKEL_B200_NO_PROLOGUE=1; fi
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r02v_bench_noprologue$v.json 2> gpurun_out/r02v_bench$v.err; echo "bench rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02v_bench_noprologue$v.json').read().strip().splitlines()[0])
print('NO_PROLOGUE=$v ms/step', d['ms_per_step'], d['stages_ms'])
PY
done
