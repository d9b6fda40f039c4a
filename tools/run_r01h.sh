# round-1 final evidence run: full GPU suite, smoke, bench line, ncu launch list + full capture (tools/profile.sh)
O=gpurun_out
(time timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/r01h_pytest.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r01h_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > $O/r01h_bench.json 2> $O/r01h_bench.err
timeout 200 python tools/bench_paths.py > $O/r01h_bench_paths.json 2> $O/r01h_bench_paths.err
timeout 600 bash tools/profile.sh r01h > $O/r01h_profile.log 2>&1
python tools/ncu_summary.py $O/r01h_full_raw.csv > $O/r01h_full_summary.md 2>> $O/r01h_profile.log
gzip -f $O/r01h_full_raw.csv
rm -f $O/r01h_full.ncu-rep
tail -4 $O/r01h_pytest.log; tail -1 $O/r01h_smoke.log; head -c 400 $O/r01h_bench.json; echo; ls -la $O | tail -12
