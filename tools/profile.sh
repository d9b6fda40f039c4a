#!/bin/bash
# ncu evidence for profiles/: run on a B200 box from the repo root (gpurun).
#   tools/profile.sh <tag>        e.g. tools/profile.sh r01b
# Writes gpurun_out/<tag>_launches.csv (per-launch durations of one steady-state pass of
# every workload), gpurun_out/<tag>_full.ncu-rep (ncu --set full of the same pass) and its
# raw page as CSV.  Copy what should be judged into profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py --what wl,dense,sp,spattr,wloa > $OUT/${TAG}_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -o $OUT/${TAG}_full python tools/profile_step.py --what wl,dense,sp,wloa > $OUT/${TAG}_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>> $OUT/${TAG}_full.log
ls -la $OUT | tail -20
