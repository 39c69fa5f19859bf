"""Row-sharded multi-GPU driver for ALS: one process per GPU, contiguous user / item row ranges
per rank, full factor replicas everywhere, ONE exchange step per half-epoch -- an in-place all-gather of the
freshly updated factor shard over NCCL (NVLink 5 / NVSwitch).  The same class drives the gloo CPU tests
(tests/test_dist_cpu.py) with a CPU row-update function, so the sharding / exchange logic is covered without GPUs.
"""


def row_shard(total, rank, world):
    """Equal contiguous pieces (the in-place all-gather needs equal counts); the last pieces may be empty."""
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total), per


class ShardedALS(object):
    """precompute(axis), update(axis, row_begin, row_end): callables bound to this rank's backend;
    P, Q: this rank's full replicas (torch tensors, updated in place by `update`)."""

    def __init__(self, precompute, update, P, Q, rank=0, world=1, dist=None):
        self.precompute, self.update, self.P, self.Q = precompute, update, P, Q
        self.rank, self.world, self.dist = rank, world, dist
        self.ranges = [row_shard(P.shape[0], rank, world), row_shard(Q.shape[0], rank, world)]
        if world > 1:
            for F in (P, Q):
                assert F.shape[0] % world == 0, "row counts must be divisible by the world size (pad the matrix)"

    def exchange(self, axis):
        if self.world == 1:
            return
        F = self.P if axis == 0 else self.Q
        lo, hi, _ = self.ranges[axis]
        self.dist.all_gather_into_tensor(F, F[lo:hi])

    def half_epoch(self, axis, on_update=None):
        lo, hi, _ = self.ranges[axis]
        self.precompute(axis)
        if on_update:
            on_update(axis, "begin")
        self.update(axis, lo, hi)
        if on_update:
            on_update(axis, "end")
        self.exchange(axis)

    def iteration(self, on_update=None):
        self.half_epoch(0, on_update)
        self.half_epoch(1, on_update)
