#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sizes"; timeout 600 python tools/repro_grow.py 14142 20000 25000 30000 10000 2>&1 | tail -5 | cut -c1-200
for k in 2 4 8; do echo "== 10000 knob $k"; GRAKEL_B200_WL_TILES_PER_CTA=$k timeout 300 python tools/repro_grow.py 10000 2>&1 | tail -1 | cut -c1-200; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest_gpu.log; tail -4 gpurun_out/r02m_pytest_gpu.log | cut -c1-300
GRAKEL_B200_CONFIG4_GRAPHS=20000 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02m_bench2.json 2> gpurun_out/r02m_bench2.err; echo "bench2 rc=$?"; grep -v Warning gpurun_out/r02m_bench2.err | grep -i "error\|Traceback\|SIG" -A3 | head -12 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02m_bench2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], d['dist_check'])
    for r in d['stages_ms_per_rank']: print(r)
    print('e2e', d['e2e']['ms_per_step'], d['e2e'].get('ms_per_step_min_median_max'))
    c = d['config4']; print({k: c[k] for k in ('ms_per_step','ms_relabel_replicated','ms_columns_panel_gemm','ms_barrier_tail_allgather','allgather_GBps_in_per_rank','checksum_equal_on_all_ranks','prefix_equals_single_gpu')})
except Exception as e:
    print('N=2 unreadable', e)
PY
