"""Device evaluation top-k (csrc/topk.cu) against the NumPy path it replaces
(buffalo/evaluate/base.py:31-42 -> buffalo/parallel/_core.hpp:69-142 quickselect in the reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_topk(P, Q, Qb, k):
    s = P.astype(np.float64) @ Q.astype(np.float64).T
    if Qb is not None:
        s = s + Qb.reshape(1, -1)
    idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(s, idx, axis=1)


@pytest.mark.parametrize("nq,I,d,k,bias", [(7, 300, 20, 10, False), (130, 9000, 128, 50, True), (3, 5000, 33, 700, False),
                                           (5, 40, 8, 64, True), (64, 100_000, 64, 110, False)])
def test_topk_host_matches_numpy(cuda_lib, nq, I, d, k, bias):
    from buffalo_b200 import backend
    rng = np.random.default_rng(nq + I)
    P = rng.normal(size=(nq, d)).astype(np.float32)
    Q = rng.normal(size=(I, d)).astype(np.float32)
    Qb = rng.normal(size=(I, 1)).astype(np.float32) if bias else None
    got = backend.topk_host(P, Q, Qb, k)
    kk = min(k, I)
    assert got.shape == (nq, kk) and got.dtype == np.int32
    ridx, rval = ref_topk(P, Q, Qb, kk)
    s = P.astype(np.float64) @ Q.astype(np.float64).T + (0 if Qb is None else Qb.reshape(1, -1))
    gval = np.take_along_axis(s, got.astype(np.int64), axis=1)
    # same scores in the same (descending) order; indices equal wherever the fp32 scores are not within rounding of a tie
    assert np.allclose(gval, rval, rtol=0, atol=1e-4 * max(1.0, np.abs(rval).max()))
    assert (np.diff(gval, axis=1) <= 1e-4).all()
    assert (got == ridx).mean() > 0.99
    for r in range(nq):
        assert len(set(got[r].tolist())) == kk       # no duplicates


def test_topk_ties_and_device_entry(cuda_lib):
    import torch
    from buffalo_b200 import backend
    # integer factors: many exactly equal scores; among equal scores the smaller index comes first
    rng = np.random.default_rng(0)
    P = rng.integers(0, 2, size=(16, 8)).astype(np.float32)
    Q = rng.integers(0, 2, size=(6000, 8)).astype(np.float32)
    idx, val = backend.topk_device(torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda(), None, 25)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    s = P @ Q.T
    for r in range(16):
        assert np.array_equal(val[r], np.sort(s[r])[::-1][:25])
        assert np.array_equal(s[r][idx[r]], val[r])
        for a, b in zip(range(24), range(1, 25)):
            if val[r][a] == val[r][b]:
                assert idx[r][a] < idx[r][b]


def test_algo_topk_recommendation_uses_device(cuda_lib):
    """Algo._get_topk_recommendation (the call behind topk_recommendation and the validation metrics) goes through
    bfl_topk_host when a GPU is present and returns what the NumPy path returns."""
    from buffalo_b200.algo.base import Algo
    from buffalo_b200.evaluate.base import topk_indices
    rng = np.random.default_rng(5)
    P = rng.normal(size=(40, 24)).astype(np.float32)
    Q = rng.normal(size=(3000, 24)).astype(np.float32)
    class _A(Algo):
        def _get_feature(self, *a, **k):
            return None

        def normalize(self, *a, **k):
            return None

        def get_topk(self, scores, k, sorted=True, num_threads=4):
            return topk_indices(scores, k)
    a = _A.__new__(_A)
    got = Algo._get_topk_recommendation(a, P, Q, None, None, None, 12, 1)
    want = topk_indices(P @ Q.T, 12)
    assert (np.asarray(got) == want).mean() > 0.99
    pool = np.arange(100, 900)
    got = Algo._get_topk_recommendation(a, P, Q, None, None, pool, 12, 1)
    want = pool[topk_indices(P @ Q[pool].T, 12)]
    assert (np.asarray(got) == want).mean() > 0.99
