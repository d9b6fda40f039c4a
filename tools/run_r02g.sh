#!/bin/bash
# 2-GPU box: u16 transport + tf32 timing on one GPU, then the tiled Gram (GK_DIST) and config 4 at 2 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "transport or attr or spellings" > gpurun_out/r02g_pytest.log 2>&1; tail -3 gpurun_out/r02g_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r02g_bench1.json 2> gpurun_out/r02g_bench1.err; echo "bench1 rc=$?"; tail -3 gpurun_out/r02g_bench1.err
timeout 600 python tools/spattr_err.py 2>&1 | tail -3
GRAKEL_B200_CONFIG4_GRAPHS=20000 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench2.json 2> gpurun_out/r02g_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/r02g_bench2.err | cut -c1-400
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02g_bench1.json').read().strip().splitlines()[-1])
print('N=1 ms/step', d['ms_per_step'], d['stages_ms'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['ms_per_step_min_median_max'], d['e2e']['last_step_ms'], 'api', d['e2e_api']['ms_per_step'], d['e2e_api']['min_ms'])
c5 = d['other_paths']['config5_spattr']; print('config5', c5['ms_per_step'], c5['stages_ms'], c5['roofline']['frac'])
try:
    d = json.loads(open('gpurun_out/r02g_bench2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], d['stages_ms'], d['dist_check'])
    print('e2e', d['e2e']['ms_per_step'])
    print(json.dumps(d['config4'], indent=1))
except Exception as e:
    print('N=2 unreadable', e)
PY
