#!/bin/bash
# last verification of the round: full GPU suite, the async test with the symmetric tail lists, bench (default as the driver runs it; A/B of the tail variants), memcheck of the small script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/r03z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03z_pytest_gpu.log; tail -6 gpurun_out/r03z_pytest_gpu.log | cut -c1-300
GRAKEL_B200_TB_SYM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k asynchronous 2>&1 | tail -2
( time timeout 600 python bench.py ) > gpurun_out/r03z_bench.json 2> gpurun_out/r03z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r03z_bench.json').read().strip().splitlines()[0])
print('default ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], 'traffic', d['roofline']['traffic'])
PY
for v in "GRAKEL_B200_TB_SYM=1" "GRAKEL_B200_TAIL_FUSED=0"; do
env $v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-paths --no-e2e > gpurun_out/r03z_bench_$v.json 2> /dev/null
python - <<PY
import json
d = json.loads(open('gpurun_out/r03z_bench_$v.json').read().strip().splitlines()[0])
print('$v ms/step', d['ms_per_step'], d['stages_ms'])
PY
done
( time timeout 300 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 python tools/sanitize_small.py ) > gpurun_out/r03_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r03_sanitizer_memcheck.log; tail -6 gpurun_out/r03_sanitizer_memcheck.log | cut -c1-200
