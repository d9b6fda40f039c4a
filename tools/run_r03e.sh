#!/bin/bash
# e2e outliers: count-based or time-based?  gc off / longer warm-up
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps 60 --warmup 3 --no-cpu --no-paths > gpurun_out/r03e_$1.json 2> gpurun_out/r03e_$1.err; echo "bench rc=$?"
  python - <<PY
import json, numpy as np
d = json.loads(open('gpurun_out/r03e_$1.json').read().strip().splitlines()[0])
e = d['e2e']; x = np.array(e['ms_each_step'])
print('$1 e2e mean', round(e['ms_per_step'],3), 'min/med/max', [round(t,2) for t in e['ms_per_step_min_median_max']], 'slow steps (>8ms):', [(int(i), float(x[i])) for i in np.nonzero(x > 8)[0]])
PY
}
run base
GRAKEL_B200_BENCH_GC=0 run gc_off
GRAKEL_B200_E2E_WARMUP=40 run warm40
GRAKEL_B200_E2E_WARMUP=40 GRAKEL_B200_BENCH_GC=0 run warm40_gc_off
