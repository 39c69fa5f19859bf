"""world_size-2 gloo test of the multi-GPU host logic (row sharding + one all-gather per half-epoch) on CPU.
The per-rank row update is the CPU oracle here (tests may use it); on the GPU box bench.py plugs the CUDA
backend into the same ShardedALS driver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import init_factors, make_csr, rel_err, transpose_csr

OPT = dict(d=16, optimizer="manual_cg", num_workers=1, compute_loss_on_training=False, alpha=8.0, reg_u=0.1, reg_i=0.1)


def _problem():
    U, I = 64, 48
    indptr, keys, vals, _ = make_csr(U, I, 900, seed=1)
    cind, ckeys, cvals = transpose_csr(indptr, keys, vals, U, I)
    return U, I, (indptr, keys, vals), (cind, ckeys, cvals), init_factors(U, 16, 16, 1, 0.1, True), init_factors(I, 16, 16, 2, 0.1, True)


def _driver(P, Q, csr, rank, world, d, sharded_gram=False):
    import oracle
    from buffalo_b200.parallel.dist import ShardedALS
    o = oracle.OracleALS()
    o.init(OPT)
    o.initialize_model(P.numpy(), Q.numpy())     # shares memory with the torch tensors

    def update(axis, lo, hi):
        ind, k, v = csr[axis]
        beg = 0 if lo == 0 else int(ind[lo - 1])
        end = int(ind[hi - 1]) if hi > lo else beg
        o.partial_update(lo, hi, ind, np.ascontiguousarray(k[beg:max(end, beg + 1)]),
                         np.ascontiguousarray(v[beg:max(end, beg + 1)]), axis)
    class RangeGram(object):
        """the two backend calls ShardedALS uses for the sharded Gram (CuALS.precompute_rows_device / gram_tensor)"""

        def precompute_rows_device(self, axis, lo, hi):
            F = (o.Q if axis == 0 else o.P)[lo:hi]
            o.FF[:] = 0.0
            if hi > lo:
                oracle.lib().orc_als_precompute(oracle._f32(np.ascontiguousarray(F)), int(hi - lo), o.o.d, oracle._f32(o.FF),
                                                o.o.num_workers)

        def gram_tensor(self):
            return torch.from_numpy(o.FF)      # shares memory: the all-reduce lands in the oracle's Gram
    return ShardedALS(o.precompute, update, P, Q, rank, world, d, backend=RangeGram() if sharded_gram else None)


def _worker(rank, world, port, out, sharded_gram=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, I, rw, cw, P0, Q0 = _problem()
    P, Q = torch.from_numpy(P0.copy()), torch.from_numpy(Q0.copy())
    drv = _driver(P, Q, (rw, cw), rank, world, dist, sharded_gram)
    assert drv.sharded_gram == sharded_gram
    for _ in range(2):
        drv.iteration()
    out[rank] = (P.numpy().copy(), Q.numpy().copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("sharded_gram", [False, True])
def test_sharded_equals_single_process(sharded_gram):
    """sharded_gram: every rank sums the Gram over its own rows of the opposite factor and the d x d partials are
    all-reduced (what ShardedALS does with the CUDA backend) instead of each rank recomputing the full matrix."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out, sharded_gram), nprocs=2, join=True)
    U, I, rw, cw, P0, Q0 = _problem()
    P, Q = torch.from_numpy(P0.copy()), torch.from_numpy(Q0.copy())
    single = _driver(P, Q, (rw, cw), 0, 1, None)
    for _ in range(2):
        single.iteration()
    for r in (0, 1):
        Pr, Qr = out[r]
        tol = 1e-5 if sharded_gram else 1e-6     # the all-reduced Gram has a different fp32 summation order
        assert rel_err(Pr, P.numpy()) < tol and rel_err(Qr, Q.numpy()) < tol          # replicas match the 1-process run
    assert not np.array_equal(P.numpy(), P0)


def test_row_shard():
    from buffalo_b200.parallel.dist import row_shard
    assert [row_shard(10, r, 4)[:2] for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert row_shard(8, 1, 2) == (4, 8, 4) and row_shard(3, 3, 4)[:2] == (3, 3)


def test_nnz_shard_covers_rows_and_balances_nonzeros():
    """rows are split by a prefix sum of nonzeros (used with the fused p2p exchange)."""
    from buffalo_b200.parallel.dist import nnz_shard
    rng = np.random.default_rng(0)
    deg = np.minimum(rng.zipf(1.3, 5000), 4000)            # heavily skewed row lengths
    deg[:7] = 0
    ind = np.cumsum(deg).astype(np.int64)
    for world in (1, 2, 3, 8):
        parts = [nnz_shard(ind, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == len(ind)
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))        # contiguous, disjoint, complete
        per = [int(ind[hi - 1] - (ind[lo - 1] if lo else 0)) if hi > lo else 0 for lo, hi, _ in parts]
        assert sum(per) == int(ind[-1])
        assert max(per) <= int(ind[-1]) / world + deg.max()                         # within one row of the ideal
        tparts = [nnz_shard(torch.from_numpy(ind), r, world) for r in range(world)]
        assert [p[:2] for p in tparts] == [p[:2] for p in parts]
    assert nnz_shard(np.zeros(0, np.int64), 0, 2)[:2] == (0, 0)
    empty = [nnz_shard(np.zeros(5, np.int64), r, 2)[:2] for r in range(2)]          # all-empty matrix: still a cover
    assert empty[0][0] == 0 and empty[0][1] == empty[1][0] and empty[1][1] == 5


def _sgd_worker(rank, world, port, out, kind, optimizer):
    import oracle
    from buffalo_b200.parallel.dist import ShardedSGD
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, I, d = 300, 120, 16
    indptr, keys, vals, _ = make_csr(U, I, 4000, seed=5)
    opt = dict(d=d, num_workers=1, optimizer=optimizer, use_bias=(kind == "bpr"), update_i=True, update_j=True,
               reg_u=0.02, reg_i=0.02, reg_j=0.02, reg_b=0.02, lr=0.05, min_lr=0.0001, beta1=0.9, beta2=0.999,
               per_coordinate_normalize=(optimizer == "adam"), num_negative_samples=2, sampling_power=0.0,
               verify_neg=True, random_seed=3, num_iters=3, compute_loss_on_training=True, max_trials=30,
               threshold=1.0, score_func="dot")

    def make(n_local):
        o = oracle.OracleSGD(warp=(kind == "warp"), use_lut=False)
        o.init(opt)
        P, Q = init_factors(U, d, d, 1, 0.2, True), init_factors(I, d, d, 2, 0.2, True)
        Qb = np.zeros((I, 1), np.float32)
        o.initialize_model(P, Q, Qb, n_local)
        return o, P, Q, Qb
    # single process
    o1, P1, Q1, Qb1 = make(len(keys))
    for _ in range(3):
        o1.add_jobs(0, U, indptr, keys)
        o1.update_parameters()
    # sharded
    from buffalo_b200.parallel.dist import nnz_shard
    lo, hi, _ = nnz_shard(indptr, rank, world)
    n_local = int(indptr[hi - 1]) - (int(indptr[lo - 1]) if lo else 0)
    o, P, Q, Qb = make(n_local)
    tP, tQ, tQb = torch.from_numpy(P), torch.from_numpy(Q), torch.from_numpy(Qb)     # share memory with the oracle
    grads = None
    if optimizer != "sgd":
        grads = [torch.from_numpy(o.gP), torch.from_numpy(o.gQ)] + ([torch.from_numpy(o.gQb)] if kind == "bpr" else [])
        grads += [torch.from_numpy(o.cP), torch.from_numpy(o.cQ)]

    def accumulate(a, b):
        beg = 0 if a == 0 else int(indptr[a - 1])
        o.add_jobs(a, b, indptr, np.ascontiguousarray(keys[beg:int(indptr[b - 1])]))
    drv = ShardedSGD(accumulate, o.update_parameters, tP, tQ, tQb, indptr, rank, world, dist, grads=grads)
    assert (drv.lo, drv.hi) == (lo, hi)
    drv.begin()
    for _ in range(3):
        drv.epoch()
    drv.finalize()      # sgd mode: the user ranges are gathered once, after the last epoch
    gathered = [torch.zeros_like(tQ) for _ in range(world)]
    dist.all_gather(gathered, tQ)
    out[rank] = dict(same=bool(all(torch.equal(g, gathered[0]) for g in gathered)),
                     errP=rel_err(P, P1), errQ=rel_err(Q, Q1), errB=rel_err(Qb + 1.0, Qb1 + 1.0),
                     lr=(o.lr, o1.lr), moved=rel_err(Q, init_factors(I, d, d, 2, 0.2, True)))
    dist.destroy_process_group()


def _run_sgd(kind, optimizer):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sgd_worker, args=(2, port, out, kind, optimizer), nprocs=2, join=True)
    return [out[r] for r in range(2)]


def test_sharded_warp_and_bpr_adagrad_equal_single_process():
    """Gradient-accumulating epochs: sharding users over 2 ranks + all-reduce of the accumulators reproduces the
    single-process result (same Philox draws, same optimizer step), replicas identical."""
    for kind, optimizer in (("warp", "adagrad"), ("bpr", "adam")):
        for r in _run_sgd(kind, optimizer):
            assert r["same"] and r["moved"] > 1e-3, (kind, r)
            assert r["errP"] < 1e-5 and r["errQ"] < 1e-5 and r["errB"] < 1e-5, (kind, r)
            assert abs(r["lr"][0] - r["lr"][1]) < 1e-9


def test_sharded_bpr_sgd_bounded_staleness():
    """Plain-SGD BPR: item deltas are summed once per epoch (bounded staleness), so the result is close to but not
    equal to the sequential run; replicas must be identical on every rank."""
    for r in _run_sgd("bpr", "sgd"):
        assert r["same"] and r["moved"] > 1e-3
        assert r["errP"] < 0.15 and r["errQ"] < 0.15, r   # small problem, lr 0.05: a few per cent is the staleness effect
