#!/bin/bash
# 2-GPU box: the bench as the driver launches it (dist path after the tile-cache / prologue changes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r03c_bench2.json 2> gpurun_out/r03c_bench2.err; echo "bench2 rc=$?"
grep -v "Warning" gpurun_out/r03c_bench2.err | grep -i "error\|Traceback\|Fatal" -A6 | head -20 | cut -c1-250
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r03c_bench2.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], d['dist_check'])
for r in d['stages_ms_per_rank'][:2]: print('  ', r)
print('   e2e', d['e2e']['ms_per_step'], d['e2e'].get('ms_per_step_min_median_max'))
c = d.get('config4')
if c: print('   config4', {k: c[k] for k in ('ms_per_step','pairs_per_s','ms_relabel_replicated','ms_columns_panel_gemm','ms_barrier_tail_allgather','allgather_GBps_in_per_rank','checksum_equal_on_all_ranks','prefix_equals_single_gpu')})
PY
