# singleton shortcut of the fused WL kernel (wl_fused_kernel<true>, GRAKEL_B200_WL_SKIP=1): full GPU suite + A/B bench on one box
O=gpurun_out
(time GRAKEL_B200_WL_SKIP=1 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > $O/r01l_pytest.log 2>&1
GRAKEL_B200_WL_SKIP=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r01l_bench_skip1.json 2> $O/r01l_bench_skip1.err
GRAKEL_B200_WL_SKIP=0 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $O/r01l_bench_skip0.json 2> $O/r01l_bench_skip0.err
tail -4 $O/r01l_pytest.log
python - <<'PY'
import json
for f in ("r01l_bench_skip1.json","r01l_bench_skip0.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, d["ms_per_step"], d["stages_ms"])
    except Exception as e: print(f, "ERR", e)
PY
