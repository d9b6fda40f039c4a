#!/bin/bash
# ncu launch list + full capture of ONE steady-state WL pass of the final pipeline (asynchronous pass, tail in the GEMM
# epilogue), and racecheck of the small end-to-end script
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=r03z
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --what wl > gpurun_out/${TAG}_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -o gpurun_out/${TAG}_full python tools/profile_step.py --what wl > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>> gpurun_out/${TAG}_full.log
python tools/ncu_summary.py gpurun_out/${TAG}_full_raw.csv > gpurun_out/${TAG}_full_summary.md 2> /dev/null; cat gpurun_out/${TAG}_full_summary.md | cut -c1-200
gzip -f gpurun_out/${TAG}_full_raw.csv; rm -f gpurun_out/${TAG}_full.ncu-rep
( time timeout 200 compute-sanitizer --tool racecheck --target-processes all --print-limit 20 python tools/sanitize_small.py ) > gpurun_out/r03_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r03_sanitizer_racecheck.log; tail -5 gpurun_out/r03_sanitizer_racecheck.log | cut -c1-200
