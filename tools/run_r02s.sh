#!/bin/bash
# 4-GPU box: tile sharing with rotated owner order vs full tiles vs TMA mirrored stores; e2e with per-process pools
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02s_bench4.json 2> gpurun_out/r02s_bench4.err; echo "bench4 rc=$?"
GRAKEL_B200_DIST_SHARE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02s_bench4_full.json 2> gpurun_out/r02s_bench4_full.err; echo "bench4 full tiles rc=$?"
GRAKEL_B200_DIST_TMA=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02s_bench4_tma.json 2> gpurun_out/r02s_bench4_tma.err; echo "bench4 tma rc=$?"
python - <<'PY'
import json
for f in ('r02s_bench4', 'r02s_bench4_full', 'r02s_bench4_tma'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], d['dist_check'])
        for r in d['stages_ms_per_rank'][:2]: print('  ', r)
        if d.get('e2e'): print('   e2e', d['e2e']['ms_per_step'], d['e2e'].get('ms_per_step_min_median_max'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
