#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/repro_grow.py 2>&1 | tail -5
GRAKEL_B200_CONFIG4_GRAPHS=20000 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02i_bench2.json 2> gpurun_out/r02i_bench2.err; echo "bench2 rc=$?"; grep -v Warning gpurun_out/r02i_bench2.err | grep -i "error\|Traceback" -A3 | head -20 | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02i_bench2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], d['stages_ms'], d['dist_check'])
    print('e2e', d['e2e']['ms_per_step'])
    print(json.dumps(d['config4'], indent=1))
except Exception as e:
    print('N=2 unreadable', e)
PY
