#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== v2 30000 fresh"; timeout 300 python tools/repro_grow.py 30000 2>&1 | tail -2
echo "== v1 30000"; GRAKEL_B200_WL_V1=1 timeout 300 python tools/repro_grow.py 30000 2>&1 | tail -2
echo "== v2 25000 28284"; timeout 300 python tools/repro_grow.py 25000 28284 2>&1 | tail -3
echo "== v2 10000 tiles/cta 4, 8"; GRAKEL_B200_WL_TILES_PER_CTA=4 timeout 300 python tools/repro_grow.py 10000 2>&1 | tail -1; GRAKEL_B200_WL_TILES_PER_CTA=8 timeout 300 python tools/repro_grow.py 10000 2>&1 | tail -1
echo "== multikernel 30000"; GRAKEL_B200_WL_FUSED=0 timeout 300 python tools/repro_grow.py 30000 2>&1 | tail -1
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02j_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02j_prof.err | tail -7 | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --no-paths > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['e2e']['last_step_ms'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
PY
