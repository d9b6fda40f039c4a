# WL-OA expansion after the block-aggregated reservations: parity, timing, launch list
O=gpurun_out
(timeout 300 python -m pytest tests/test_wloa.py tests/test_tu_reader.py -m gpu -x -q 2>&1 | tail -8) > $O/r01i_wloa_pytest.log 2>&1
timeout 200 python tools/bench_paths.py > $O/r01i_bench_paths.json 2> $O/r01i_bench_paths.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r01i_launches_wloa.csv python tools/profile_step.py --what wloa > $O/r01i_launches.log 2>&1
tail -3 $O/r01i_wloa_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r01i_bench_paths.json"))["config2_wloa"]; print(d)
PY
grep -E "oa_|wl_fused" $O/r01i_launches_wloa.csv | awk -F'","' '{print $5, $(NF)}' | head
