# stdout hygiene of the bench line at N=2 and N=1 (one JSON line, nothing else)
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/r01k_bench_n2.json 2> $O/r01k_bench_n2.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r01k_bench_n1.json 2> $O/r01k_bench_n1.err
wc -l $O/r01k_bench_n2.json $O/r01k_bench_n1.json; head -c 200 $O/r01k_bench_n2.json; echo; head -c 200 $O/r01k_bench_n1.json; echo; grep -c "NCCL version" $O/r01k_bench_n2.err
