#!/bin/bash
# tail fused into the GEMM epilogue: parity (async test + gpu parity file) + bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r03n_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03n_pytest_gpu.log; tail -12 gpurun_out/r03n_pytest_gpu.log | cut -c1-300
for v in 1; do
export GRAKEL_B200_TAIL_FUSED=$v
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r03n_bench_fused$v.json 2> gpurun_out/r03n_bench$v.err; echo "bench rc=$?"
tail -2 gpurun_out/r03n_bench$v.err | cut -c1-200
python - <<PY
import json
d = json.loads(open('gpurun_out/r03n_bench_fused$v.json').read().strip().splitlines()[0])
print('TAIL_FUSED=$v ms/step', d['ms_per_step'], d['stages_ms'])
PY
done
