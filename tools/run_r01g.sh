# round-1 experiment: CTA-pair (cta_group::2) Gram GEMM vs the one-CTA kernel (A/B, same box)
O=gpurun_out
(timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta_pair" 2>&1 | tail -25) > $O/r01g_cta2_pytest.log 2>&1
tail -5 $O/r01g_cta2_pytest.log
if grep -q " passed" $O/r01g_cta2_pytest.log && ! grep -q "failed" $O/r01g_cta2_pytest.log; then
  GRAKEL_B200_CTA2=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $O/r01g_bench_cta2.json 2> $O/r01g_bench_cta2.err
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > $O/r01g_bench_cta1.json 2> $O/r01g_bench_cta1.err
  python - <<'PY'
import json
for f in ("r01g_bench_cta2.json","r01g_bench_cta1.json"):
    try:
        d=json.load(open("gpurun_out/"+f)); print(f, d["ms_per_step"], d["stages_ms"]["gram_gemm"], d["roofline"]["frac"], d["dense_gemm_mode"])
    except Exception as e: print(f, "ERR", e)
PY
fi
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
