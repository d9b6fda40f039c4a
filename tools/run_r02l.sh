#!/bin/bash
cd "$(dirname "$0")/.."
for d in 4 5 6; do echo "== 2000 knob 2 dbg $d"; GRAKEL_B200_WL_DBG=$d GRAKEL_B200_WL_TILES_PER_CTA=2 timeout 300 python tools/repro_grow.py 2000 2>&1 | tail -4 | cut -c1-300; done
echo "== 20000"; GRAKEL_B200_WL_DBG=4 timeout 300 python tools/repro_grow.py 20000 2>&1 | tail -2 | cut -c1-300
