"""torchrun worker for tests/test_mgpu.py: row-sharded ALS on N GPUs (both exchange modes) must reproduce the
single-GPU factors, and every rank must hold identical replicas."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from buffalo_b200 import backend  # noqa: E402
from buffalo_b200.parallel.dist import ShardedALS  # noqa: E402
from tests.helpers import init_factors, make_csr, transpose_csr  # noqa: E402


def run(world, rank, mode, P0, Q0, rw, cw, dev, d):
    opt = dict(d=d, optimizer="manual_cg", compute_loss_on_training=False, alpha=8.0, reg_u=0.1, reg_i=0.1,
               block_size=32, num_cg_max_iters=3, eps=1e-10, cg_tolerance=1e-10)
    obj = backend.CuALS()
    assert obj.init(opt)
    P, Q = torch.from_numpy(P0.copy()).to(dev), torch.from_numpy(Q0.copy()).to(dev)
    if mode == "p2p":
        from buffalo_b200.parallel.dist import exportable_like
        P, Q = exportable_like(P), exportable_like(Q)
    obj.bind_factors(P, Q)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    obj.bind_csr(0, t(rw[0]), t(rw[1]), t(rw[2]))
    obj.bind_csr(1, t(cw[0]), t(cw[1]), t(cw[2]))
    drv = ShardedALS(obj.precompute_device, obj.update_device, P, Q, rank, world, dist if world > 1 else None,
                     exchange=mode, backend=obj, indptrs=(rw[0], cw[0]))   # p2p: rows split by nonzeros
    for _ in range(2):
        drv.iteration()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return P.cpu().numpy(), Q.cpu().numpy()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    d, U, I = 128, 4096, 2048
    indptr, keys, vals, _ = make_csr(U, I, 200000, seed=3, empty_rows=9)
    cw = transpose_csr(indptr, keys, vals, U, I)
    P0, Q0 = init_factors(U, d, d, 1, 0.1, True), init_factors(I, d, d, 2, 0.1, True)
    ref = run(1, 0, "none", P0, Q0, (indptr, keys, vals), cw, dev, d)
    ok = True
    for mode in ("allgather", "p2p"):
        P, Q = run(world, rank, mode, P0, Q0, (indptr, keys, vals), cw, dev, d)
        err = max(np.abs(P - ref[0]).max() / np.abs(ref[0]).max(), np.abs(Q - ref[1]).max() / np.abs(ref[1]).max())
        good = err < 1e-5      # same kernels, same inputs: only the loss-free row order differs
        print("rank %d mode %s rel err vs single GPU %.2e %s" % (rank, mode, err, "OK" if good else "FAIL"), flush=True)
        ok = ok and good
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
