#!/bin/bash
# round 2 evidence: ncu launch list + full capture of one steady-state pass of every workload, then a clean bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/profile.sh r02 > gpurun_out/r02_profile.log 2>&1
tail -5 gpurun_out/r02_launches.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r02_full_raw.csv > gpurun_out/r02_full_summary.md 2>> gpurun_out/r02_profile.log; head -30 gpurun_out/r02_full_summary.md | cut -c1-220
python tools/make_traffic.py gpurun_out/r02_full_raw.csv r02 > gpurun_out/r02_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/r02_traffic.json
gzip -kf gpurun_out/r02_full_raw.csv
rm -f gpurun_out/r02_full.ncu-rep
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'], 'traffic', d['roofline']['traffic'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
print(json.dumps(d['cpu_baseline'])[:400])
PY
echo "== 28284 graphs, V2 vs V1 relabel"
timeout 600 python bench.py --graphs 28284 --steps 5 --warmup 3 --no-cpu --no-e2e --no-paths 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('V2', d['ms_per_step'], d['stages_ms'])"
GRAKEL_B200_WL_V1=1 timeout 600 python bench.py --graphs 28284 --steps 5 --warmup 3 --no-cpu --no-e2e --no-paths 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('V1', d['ms_per_step'], d['stages_ms'])"
