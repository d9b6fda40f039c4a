#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 2 3 4; do echo "== 10000 knob $k"; GRAKEL_B200_WL_TILES_PER_CTA=$k timeout 300 python tools/repro_grow.py 10000 2>&1 | tail -1; done
for n in 500 2000 5000; do echo "== $n knob 4"; GRAKEL_B200_WL_TILES_PER_CTA=4 timeout 300 python tools/repro_grow.py $n 2>&1 | tail -1; done
echo "== 25000"; timeout 300 python tools/repro_grow.py 25000 2>&1 | tail -1
