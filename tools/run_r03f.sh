#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 14142 28284; do timeout 300 python tools/prof_wl_n.py $n 2>&1 | grep -A8 "wl_fused2 prof\|ms_features" | cut -c1-330 | tee gpurun_out/r03f_wl_prof_$n.txt; done
