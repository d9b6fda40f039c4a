#!/bin/bash
# round 2, third GPU pass: wl_fused2 after the diagonal fix, delivery with dynamic tasks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -15 gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02c_bench.err
GRAKEL_B200_PROF=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2> gpurun_out/r02c_prof.err; grep -A8 "wl_fused2 prof" gpurun_out/r02c_prof.err | tail -9
for t in 6 8 16 24; do
  GRAKEL_B200_HOST_THREADS=$t timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02c_bench_ht$t.json 2> gpurun_out/r02c_bench_ht$t.err
done
GRAKEL_B200_HOST_NO_PIN=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02c_bench_nopin.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02c_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        a = d.get('e2e_api') or {}
        print(f, 'ms/step %.3f' % d['ms_per_step'], 'e2e %s' % e.get('ms_per_step'), 'last', (e.get('last_step_ms') or {}).get('d2h'), 'e2e_api', a.get('ms_per_step'), a.get('min_ms'), d['stages_ms'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
