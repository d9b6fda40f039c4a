#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
GRAKEL_B200_CONFIG4_GRAPHS=20000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02n_bench2.json 2> gpurun_out/r02n_bench2.err; echo "bench2 (e2e+config4) rc=$?"
grep -v "Warning" gpurun_out/r02n_bench2.err | grep -B2 -A30 "Fatal Python error\|Segmentation" | head -70 | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02n_bench2b.json 2> gpurun_out/r02n_bench2b.err; echo "bench2 (e2e only) rc=$?"
grep -v "Warning" gpurun_out/r02n_bench2b.err | grep -B2 -A30 "Fatal Python error\|Segmentation" | head -40 | cut -c1-200
for f in 2 4; do echo "== table factor $f"; GRAKEL_B200_WL_TABLE_FACTOR=$f timeout 300 python tools/repro_grow.py 10000 10000 28284 28284 2>&1 | tail -4 | cut -c1-200; done
