#!/usr/bin/env python
"""BASELINE config 4: WL-subtree (h=5) Gram of N synthetic graphs, K row-tiled over the ranks of a
torchrun job and assembled on every rank with ONE NCCL all-gather over NVLink.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29533 tools/bench_config4.py --graphs 50000

Every rank relabels the replicated CSR block (56 MB at 50 000 graphs), computes its row block of K
(fp32, written by the GEMM epilogue straight into a torch tensor through GK_OUT_DEVICE), then
all_gather_into_tensor.  Parity: K[:p, :p] must equal the single-rank Gram of the first p graphs
(an entry depends on its two graphs only), checked on rank 0 against the CPU oracle for a small p
and against a one-GPU run for p = 2000; the checksum of K must agree on all ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import H, pack_workload  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from grakel_b200 import _lib
    from grakel_b200.dist import all_gather_rows, row_block
    eng = _lib.Engine(local)
    n = args.graphs
    gp, rp, ci, lab = pack_workload(n)
    rb, re_ = row_block(n, rank, world)
    per = (n + world - 1) // world
    K_local = torch.empty((per, n), dtype=torch.float32, device="cuda")  # padded to the all-gather block
    eng.pack(gp, rp, ci, lab)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_feat, t_gram, t_gather, t_total = [], [], [], []
    K = None
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.wl_features(H)
        eng.gram(n, dtype=np.float32, row_range=(rb, re_), stats=st, want_diag=False,
                 device_ptr=K_local.data_ptr(), ld=n)  # returns after the engine's stream is synchronised
        t1 = time.perf_counter()
        ev[0].record()
        K = all_gather_rows(K_local, n) if world > 1 else K_local[:n]
        ev[1].record()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if it >= args.warmup:
            t_feat.append(st.ms_features)
            t_gram.append(st.ms_panel + st.ms_gemm + st.ms_tail)
            t_gather.append(ev[0].elapsed_time(ev[1]))
            t_total.append((t2 - t0) * 1e3)
    tt = torch.tensor([float(np.mean(t_total)), float(np.mean(t_gather))], device="cuda")
    cs = torch.tensor([float(K.double().sum().item())], device="cuda", dtype=torch.float64)
    cs_min, cs_max = cs.clone(), cs.clone()
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(cs_max, op=dist.ReduceOp.MAX)
    if rank == 0:
        # parity of the assembled matrix: prefix blocks against a single-device run and the CPU oracle
        p1 = min(2000, n)
        e1 = _lib.Engine(local)
        e1.pack(*pack_workload(p1))
        s1 = e1.wl_features(H)
        K1, _, _ = e1.gram(p1, dtype=np.float32, stats=s1, want_diag=False)
        ok_dev = bool(np.array_equal(K[:p1, :p1].cpu().numpy(), K1))
        from oracle.gk_oracle import WLOracle, gen
        p2 = min(300, n)
        Ko = WLOracle(n_iter=H).fit_transform(gen(p2, 40, 0))
        ok_cpu = bool(np.array_equal(K[:p2, :p2].double().cpu().numpy(), Ko))
        ms = float(tt[0].item())
        print(json.dumps({
            "workload": f"config4: {n} ER graphs (avg 40 nodes, 7 labels, seed 0), WL-subtree h={H}",
            "n_gpus": world, "ms_per_step_wall": ms, "pairs_per_s": n * n / (ms * 1e-3),
            "ms_features(replicated)": float(np.mean(t_feat)), "ms_gram(row block)": float(np.mean(t_gram)),
            "ms_allgather": float(tt[1].item()), "allgather_bytes_per_rank": int(per * n * 4),
            "allgather_GBps_per_rank_in": (world - 1) * per * n * 4 / (float(tt[1].item()) * 1e-3) / 1e9 if world > 1 else None,
            "head_columns": int(st.n_dense_columns), "threshold_T": int(st.threshold),
            "checksum": float(cs.item()), "checksum_equal_on_all_ranks": bool(cs_min.item() == cs_max.item()),
            "prefix_equals_single_gpu": ok_dev, "prefix_equals_cpu_oracle": ok_cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
