#!/bin/bash
# round-3 evidence in one call: full GPU suite, the bench as the driver runs it, ncu launch list + full capture of one
# steady-state pass, per-level WL profile, compute-sanitizer memcheck + racecheck of the small end-to-end script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r03_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03_pytest_gpu.log; tail -5 gpurun_out/r03_pytest_gpu.log | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/r03_bench.json 2> gpurun_out/r03_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r03_bench.json').read().strip().splitlines()[0])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
print('roofline', d['roofline']['frac'], 'relabel', d['roofline_relabel']['frac'], 'paths', json.dumps(d.get('other_paths'))[:600])
PY
timeout 300 python tools/prof_wl.py 2>&1 | grep -A8 "wl_fused2 prof\|ms_features" | cut -c1-330 > gpurun_out/r03_wl_prof.txt; tail -3 gpurun_out/r03_wl_prof.txt | cut -c1-200
bash tools/profile.sh r03 > gpurun_out/r03_profile.log 2>&1
python tools/ncu_summary.py gpurun_out/r03_full_raw.csv > gpurun_out/r03_full_summary.md 2> gpurun_out/r03_summary.err; head -14 gpurun_out/r03_full_summary.md | cut -c1-200
gzip -f gpurun_out/r03_full_raw.csv
rm -f gpurun_out/r03_full.ncu-rep
for tool in memcheck racecheck; do
  ( time timeout 600 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python tools/sanitize_small.py ) > gpurun_out/r03_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r03_sanitizer_$tool.log
  tail -8 gpurun_out/r03_sanitizer_$tool.log | cut -c1-200
done
