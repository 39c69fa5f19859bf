#!/usr/bin/env python
"""Secondary throughput check of the BPRMF / WARP path (BASELINE configs[2], [3]) on ONE B200: positives per second
of whole epochs (sampling + update kernels + optimizer step) with everything resident on the device.

  python benchmarks/sgd_bench.py --algo warp --users 1000000 --items 100000 --nnz 50000000 --dim 64
  python benchmarks/sgd_bench.py --algo bpr  --users 10000000 --items 1000000 --nnz 500000000 --dim 128

Under torchrun (one rank per GPU) the users are sharded by nonzeros over the ranks (parallel/dist.py::ShardedSGD):
gradient-accumulating configurations all-reduce the accumulators once per epoch, plain-SGD BPR exchanges item deltas.

bench.py (ALS, the headline metric) is what the driver runs; this prints one JSON line of the same style.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", choices=["bpr", "warp"], default="warp")
    ap.add_argument("--users", type=int, default=1000000)
    ap.add_argument("--items", type=int, default=100000)
    ap.add_argument("--nnz", type=int, default=50000000)
    ap.add_argument("--dim", type=int, default=64)   # not "--d": torchrun's own parser rejects it as ambiguous
    ap.add_argument("--optimizer", default=None)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench import make_workload
    from buffalo_b200 import backend
    from buffalo_b200.parallel.dist import ShardedSGD
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = make_workload(dict(users=args.users, items=args.items, nnz=args.nnz), dev)
    U, I, nnz, d = wl["U"], wl["I"], wl["nnz"], args.dim
    optimizer = args.optimizer or ("adagrad" if args.algo == "warp" else "sgd")
    opt = dict(d=d, num_workers=1, optimizer=optimizer, use_bias=(args.algo == "bpr"), update_i=True, update_j=True,
               reg_u=0.025, reg_i=0.025, reg_j=0.025, reg_b=0.025, lr=0.05, min_lr=0.0001, beta1=0.9, beta2=0.999,
               per_coordinate_normalize=False, num_negative_samples=1, sampling_power=0.0, verify_neg=True,
               random_seed=7, num_iters=args.epochs + args.warmup, compute_loss_on_training=True, max_trials=50,
               threshold=1.0, score_func="dot")
    g = backend.CuSGD(args.algo)
    assert g.init(opt)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    P = (torch.randn(U, d, device=dev, generator=gen) * 0.1).contiguous()
    Q = (torch.randn(I, d, device=dev, generator=gen) * 0.1).contiguous()
    Qb = torch.zeros(I, 1, device=dev)
    shard = ShardedSGD(None, None, P, Q, Qb, wl["r_indptr"], rank, world, dist if world > 1 else None)
    g.bind_factors(P, Q, Qb, shard.local_positives(wl["r_indptr"]))
    g.bind_csr(wl["r_indptr"], wl["r_keys"])
    g.launch_workers()
    grads = None
    if optimizer != "sgd":
        grads = [g.grad_tensor(0, P.shape), g.grad_tensor(1, Q.shape)] + ([g.grad_tensor(2, (I,))] if args.algo == "bpr" else [])
        grads += [g.count_tensor(0, U), g.count_tensor(1, I)]
    drv = ShardedSGD(g.add_jobs_device, g.update_parameters_device, P, Q, Qb, wl["r_indptr"], rank, world,
                     dist if world > 1 else None, grads=grads)
    times = []
    for ep in range(args.warmup + args.epochs):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        drv.epoch()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # device time, max over ranks
        if ep >= args.warmup:
            times.append(float(t.item()))
    loss, updates = g.read_stats() if args.algo == "warp" else (float("nan"), 0)
    ms = float(np.mean(times))
    finite = bool(torch.isfinite(P).all().item() and torch.isfinite(Q).all().item())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # rows touched per positive: user + positive + negative row, read and updated (3 x 2 x 4d bytes).  NOT DRAM
    # traffic: a warp walks one user's positives back to back, so the user row and popular items hit L1/L2; WARP reads
    # one more item row per extra trial (data dependent, not counted)
    alg = 3 * 2 * 4 * d
    print(json.dumps({"metric": "positives/sec %s d=%d" % (args.algo.upper(), d), "value": nnz / (ms / 1e3), "unit": "nnz/s",
                      "n_gpus": world, "ms_per_epoch": ms, "epochs": args.epochs, "config": dict(users=U, items=I, nnz=nnz,
                      d=d, optimizer=optimizer), "row_bytes_touched_gbs": alg * nnz / (ms / 1e3) / 1e9,
                      "finite": finite, "warp_stats_rank0": {"loss": loss, "updates": updates}}))


if __name__ == "__main__":
    main()
