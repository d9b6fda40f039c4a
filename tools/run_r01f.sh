set -o pipefail
O=gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/r01f_pytest.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 3 > $O/r01f_bench.json 2> $O/r01f_bench.err
GRAKEL_B200_MIRROR_TMA=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $O/r01f_bench_mirror0.json 2> $O/r01f_bench_mirror0.err
GRAKEL_B200_PROF=1 timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2> $O/r01f_prof_mirror2.err
GRAKEL_B200_PROF=1 GRAKEL_B200_MIRROR_TMA=0 timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2> $O/r01f_prof_mirror0.err
timeout 200 python tools/bench_tu.py > $O/r01f_bench_tu.json 2> $O/r01f_bench_tu.err
tail -4 $O/r01f_pytest.log
python - <<'PY'
import json
for f in ("r01f_bench.json","r01f_bench_mirror0.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, d["ms_per_step"], d["stages_ms"], d["roofline"]["frac"], d.get("e2e") and d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
grep "gram_tc prof" $O/r01f_prof_mirror2.err | tail -2
grep "gram_tc prof" $O/r01f_prof_mirror0.err | tail -2
cat $O/r01f_bench_tu.json
