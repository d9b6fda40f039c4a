#!/bin/bash
# round 2, first GPU pass: GPU test suite, bench line, host-delivery / packer sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== host"; lscpu | egrep "Model name|Socket|Core|Thread|NUMA|^CPU\(s\)"; cat /sys/kernel/mm/transparent_hugepage/enabled; cat /sys/kernel/mm/transparent_hugepage/defrag; free -g | head -2
nvidia-smi topo -m 2>/dev/null | head -20
} > gpurun_out/r02a_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest_gpu.log
tail -5 gpurun_out/r02a_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
for t in 8 16 32 48; do
  GRAKEL_B200_HOST_THREADS=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench_ht$t.json 2> gpurun_out/r02a_bench_ht$t.err
done
GRAKEL_B200_NO_TRI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench_notri.json 2>/dev/null
GRAKEL_B200_WIDEN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench_nowiden.json 2>/dev/null
GRAKEL_B200_HOST_ANY_NODE=1 GRAKEL_B200_HOST_THREADS=32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench_anynode32.json 2>/dev/null
for t in 1 4 8 16 32; do
  GRAKEL_B200_PACK_THREADS=$t GRAKEL_B200_PACK_DEBUG=1 python - <<PY 2>&1 | tail -12 > gpurun_out/r02a_pack_t$t.txt
import sys, time
sys.path.insert(0, '.')
from bench import gen_list
from grakel_b200 import packing
X = gen_list(10000)
for _ in range(3):
    t = time.perf_counter(); b = packing.pack(X, 'wl', len_ok=lambda n: n >= 2); t1 = time.perf_counter()
    ids, d = packing.label_ids(b.labels, None, True); t2 = time.perf_counter()
    print('pack %.2f ms  label_ids %.2f ms' % ((t1 - t) * 1e3, (t2 - t1) * 1e3))
PY
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02a_bench*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.3f' % d['ms_per_step'], 'e2e %.2f ms' % d['e2e']['ms_per_step'], 'e2e_api', d.get('e2e_api') and '%.1f ms (first %.0f, pack %.1f)' % (d['e2e_api']['ms_per_step'], d['e2e_api']['first_call_ms'], d['e2e_api']['host_ms']['pack(list -> CSR)']), d['stages_ms'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
cat gpurun_out/r02a_pack_t*.txt | grep -v worker
