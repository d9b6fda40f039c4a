#!/bin/bash
# 8-GPU box: the bench exactly as the driver launches it (rotated tile order, per-process delivery pools, config 4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02t_bench8.json 2> gpurun_out/r02t_bench8.err; echo "bench8 rc=$?"
grep -v "Warning" gpurun_out/r02t_bench8.err | grep -i "error\|Traceback\|Fatal" -A6 | head -20 | cut -c1-250
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02t_bench8.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], d['dist_check'])
for r in d['stages_ms_per_rank'][:3]: print('  ', r)
print('   e2e', d['e2e']['ms_per_step'], d['e2e'].get('ms_per_step_min_median_max'))
c = d['config4']; print('   config4', {k: c[k] for k in ('ms_per_step','pairs_per_s','ms_relabel_replicated','ms_columns_panel_gemm','ms_barrier_tail_allgather','allgather_GBps_in_per_rank','checksum_equal_on_all_ranks','prefix_equals_single_gpu')})
PY
