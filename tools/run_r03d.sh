#!/bin/bash
# e2e outliers: 100 steps with pinned vs floating delivery workers; which step is slow?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1 0 1; do
  export GRAKEL_B200_HOST_FLOAT=$v
  timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu --no-paths > gpurun_out/r03d_bench_float$v.json 2> gpurun_out/r03d_bench$v.err; echo "bench rc=$?"
  python - <<PY
import json, numpy as np
d = json.loads(open('gpurun_out/r03d_bench_float$v.json').read().strip().splitlines()[0])
e = d['e2e']; x = np.array(e['ms_each_step'])
print('FLOAT=$v e2e mean', round(e['ms_per_step'],3), 'min/med/max', [round(t,2) for t in e['ms_per_step_min_median_max']], 'slow steps (>8ms):', [(int(i), float(x[i])) for i in np.nonzero(x > 8)[0]], 'value step', round(d['ms_per_step'],4))
PY
done
