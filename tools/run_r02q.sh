#!/bin/bash
# 8-GPU box: the weak-scaling bench at 8 and 4 ranks (GK_DIST tiles, config 4 at 8 ranks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02q_bench8.json 2> gpurun_out/r02q_bench8.err; echo "bench8 rc=$?"
grep -v "Warning" gpurun_out/r02q_bench8.err | grep -i "error\|Traceback\|Fatal" -A6 | head -30 | cut -c1-250
GRAKEL_B200_CONFIG4_GRAPHS=0 GRAKEL_B200_DIST_TMA=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02q_bench8_tma.json 2> gpurun_out/r02q_bench8_tma.err; echo "bench8 tma rc=$?"
python - <<'PY'
import json
for f in ('r02q_bench8', 'r02q_bench8_tma'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], d['dist_check'])
        for r in d['stages_ms_per_rank'][:3]: print('  ', r)
        if d.get('e2e'): print('   e2e', d['e2e']['ms_per_step'])
        if d.get('config4'):
            c = d['config4']; print('   config4', {k: c[k] for k in ('ms_per_step','pairs_per_s','ms_relabel_replicated','ms_columns_panel_gemm','ms_barrier_tail_allgather','allgather_GBps_in_per_rank','checksum_equal_on_all_ranks','prefix_equals_single_gpu')})
    except Exception as e:
        print(f, 'unreadable', e)
PY
