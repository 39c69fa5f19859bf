"""BPRMF / WARP parity on the GPU through the C ABI against the CPU oracle.

* sampling (Philox draws, seen-item rejection, popularity table) is bit-exact;
* WARP epochs and BPR adagrad/adam epochs only accumulate gradients (warp.cc:156-158, bpr.cc:138-156), so
  with identical draws they are deterministic up to fp32 summation order: gradients, the optimizer step and
  the projected factors are compared element-wise (1e-4 relative);
* plain-SGD BPR is Hogwild in the reference (bpr.cc:157-171, racy by design): exact on a collision-free
  triple list, statistical (same converged loss level, monotone learning) otherwise.
"""
import numpy as np
import pytest

from tests.helpers import init_factors, make_csr, rel_err

pytestmark = pytest.mark.gpu


def sgd_opt(**kw):
    opt = dict(d=32, num_workers=1, optimizer="sgd", use_bias=True, update_i=True, update_j=True, reg_u=0.025,
               reg_i=0.025, reg_j=0.025, reg_b=0.025, lr=0.05, min_lr=0.0001, beta1=0.9, beta2=0.999,
               per_coordinate_normalize=False, num_negative_samples=1, sampling_power=0.0, verify_neg=True,
               random_seed=7, num_iters=4, compute_loss_on_training=True, max_trials=50, threshold=1.0,
               score_func="dot")
    opt.update(kw)
    return opt


def make_pair(kind, opt, P, Q, Qb, indptr, keys, cum=None):
    import oracle
    from buffalo_b200 import backend
    g = backend.CuSGD(kind)
    assert g.init(opt)
    o = oracle.OracleSGD(warp=(kind == "warp"), use_lut=False)
    o.init(opt)
    Pg, Qg, Qbg = P.copy(), Q.copy(), Qb.copy()
    Po, Qo, Qbo = P.copy(), Q.copy(), Qb.copy()
    g.initialize_model(Pg, Qg, Qbg, len(keys))
    o.initialize_model(Po, Qo, Qbo, len(keys))
    if cum is not None:
        g.set_cumulative_table(cum, len(cum))
        o.set_cumulative_table(cum, len(cum))
    g.launch_workers()
    return g, o, (Pg, Qg, Qbg), (Po, Qo, Qbo)


@pytest.mark.parametrize("num_neg,power,verify", [(1, 0.0, True), (3, 0.0, True), (2, 1.0, True), (1, 1.0, False)])
def test_bpr_sampling_bit_exact(cuda_lib, num_neg, power, verify):
    import torch
    U, I, d = 700, 400, 16
    indptr, keys, vals, _ = make_csr(U, I, 9000, seed=21)
    opt = sgd_opt(d=d, num_negative_samples=num_neg, sampling_power=power, verify_neg=verify, optimizer="adagrad")
    P, Q = init_factors(U, d, d, 1, scale=0.1), init_factors(I, d, d, 2, scale=0.1)
    cum = None
    if power > 0:
        cum = np.cumsum(np.bincount(keys, minlength=I).astype(np.int64) ** int(power)).astype(np.int64)  # bpr.py:99-111
    g, o, _, _ = make_pair("bpr", opt, P, Q, np.zeros((I, 1), np.float32), indptr, keys, cum)
    dev = torch.device("cuda:0")
    g.bind_csr(torch.from_numpy(indptr).to(dev), torch.from_numpy(keys).to(dev))
    n = len(keys) * num_neg
    tu, tp, tn = (torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(3))
    g.sample_device(0, U, tu, tp, tn)
    torch.cuda.synchronize()
    ou, op, on = o.sample(0, U, indptr, keys)
    assert np.array_equal(tu.cpu().numpy(), ou) and np.array_equal(tp.cpu().numpy(), op)
    assert np.array_equal(tn.cpu().numpy(), on)
    if verify:
        beg = np.concatenate([[0], indptr[:-1]])
        for s in range(0, n, 97):
            u = ou[s]
            assert on[s] not in keys[beg[u]:indptr[u]]


@pytest.mark.parametrize("optimizer,pcn", [("adagrad", False), ("adam", True)])
def test_bpr_gradient_epochs_match_oracle(cuda_lib, optimizer, pcn):
    U, I, d = 1500, 900, 40
    indptr, keys, vals, _ = make_csr(U, I, 30000, seed=31)
    opt = sgd_opt(d=d, optimizer=optimizer, per_coordinate_normalize=pcn, num_negative_samples=2, lr=0.05)
    P, Q = init_factors(U, d, d, 1, scale=0.3, signed=True), init_factors(I, d, d, 2, scale=0.3, signed=True)
    Qb = init_factors(I, 1, 1, 3, scale=0.1, signed=True)
    g, o, (Pg, Qg, Qbg), (Po, Qo, Qbo) = make_pair("bpr", opt, P, Q, Qb, indptr, keys)
    probe = (np.arange(30, dtype=np.int32), keys[:30].copy(), ((keys[:30] + 1) % I).astype(np.int32))
    for epoch in range(3):
        half = U // 2
        for a, b in ((0, half), (half, U)):
            beg = 0 if a == 0 else int(indptr[a - 1])
            k = np.ascontiguousarray(keys[beg:int(indptr[b - 1])])
            g.add_jobs(a, b, indptr, k)
            o.add_jobs(a, b, indptr, k)
        g.update_parameters()
        o.update_parameters()
        g.wait_until_done()
        assert rel_err(Pg, Po) < 1e-4 and rel_err(Qg, Qo) < 1e-4 and rel_err(Qbg, Qbo) < 1e-4, epoch
        lg, lo = g.compute_loss(*probe), o.compute_loss(*probe)
        assert abs(lg - lo) < 1e-5 * max(1.0, abs(lo))
    assert g.epoch() == 3


def test_bpr_sgd_collision_free_exact(cuda_lib):
    import oracle
    import torch
    from buffalo_b200 import backend
    U, I, d = 600, 1300, 128
    opt = sgd_opt(d=d, optimizer="sgd")
    P, Q = init_factors(U, d, d, 1, scale=0.2, signed=True), init_factors(I, d, d, 2, scale=0.2, signed=True)
    Qb = init_factors(I, 1, 1, 3, scale=0.1, signed=True)
    rng = np.random.default_rng(0)
    us = rng.permutation(U)[:500].astype(np.int32)
    items = rng.permutation(I)[:1000].astype(np.int32)
    ps, ns = items[:500].copy(), items[500:].copy()
    g = backend.CuSGD("bpr")
    g.init(opt)
    dev = torch.device("cuda:0")
    tP, tQ, tQb = (torch.from_numpy(x.copy()).to(dev) for x in (P, Q, Qb))
    g.bind_factors(tP, tQ, tQb, 500)
    g.apply_triples_device(torch.from_numpy(us).to(dev), torch.from_numpy(ps).to(dev), torch.from_numpy(ns).to(dev), 0.05)
    torch.cuda.synchronize()
    o = oracle.sgd_opt_struct(opt)
    Po, Qo, Qbo = P.copy(), Q.copy(), Qb.copy()
    import ctypes as C
    oracle.lib().orc_bpr_update_preupdate(C.byref(o), oracle._f32(Po), oracle._f32(Qo), oracle._f32(Qbo),
                                          oracle._i32(us), oracle._i32(ps), oracle._i32(ns), C.c_int64(500),
                                          C.c_float(0.05))
    assert rel_err(tP.cpu().numpy(), Po) < 1e-5 and rel_err(tQ.cpu().numpy(), Qo) < 1e-5
    assert rel_err(tQb.cpu().numpy(), Qbo) < 1e-5
    assert not np.array_equal(Po, P)


def test_bpr_sgd_converges_like_oracle(cuda_lib):
    """Plain-SGD BPR is Hogwild in the reference (bpr.cc:157-171, racy by design) and the oracle is its sequential
    limit.  The GPU walks contiguous triple ranges per warp (a worker thread's job) with thousands of warps in
    flight, so during the fast transient out of the near-zero initialisation its loss lags the sequential run (a
    lock-step simulation of the same schedule shows up to ~50 % at epoch 1-2) and the two meet again as they
    converge.  Checked here: identical lr schedule, the GPU learns monotonically, and after 8 epochs both reach the
    same loss level (within 25 %; the lock-step model gives 5-18 %)."""
    U, I, d = 20000, 3000, 32
    rng = np.random.default_rng(5)
    A, B = rng.normal(size=(U, 4)).astype(np.float32), rng.normal(size=(I, 4)).astype(np.float32)
    S = A @ B.T   # planted low-rank preference structure
    rows, cols = np.nonzero(S > np.quantile(S[:2000], 0.99))
    order = np.lexsort((cols, rows))
    rows, keys = rows[order], cols[order].astype(np.int32)
    indptr = np.cumsum(np.bincount(rows, minlength=U)).astype(np.int64)
    pu = rng.integers(0, len(keys), 2000)
    probe_u = rows[pu].astype(np.int32)
    probe_p = keys[pu].copy()
    probe_n = rng.integers(0, I, 2000).astype(np.int32)
    import oracle
    epochs = 8
    for seed in (1, 2):
        opt = sgd_opt(d=d, optimizer="sgd", lr=0.1, random_seed=seed, num_iters=epochs, reg_u=0.01, reg_i=0.01,
                      reg_j=0.01, reg_b=0.01)
        P, Q = init_factors(U, d, d, seed, scale=0.05), init_factors(I, d, d, seed + 10, scale=0.05)
        Qb = np.zeros((I, 1), np.float32)
        g, _, (Pg, Qg, Qbg), _ = make_pair("bpr", opt, P, Q, Qb, indptr, keys)
        o = oracle.OracleSGD(warp=False, use_lut=True)
        o.init(opt)
        Po, Qo, Qbo = P.copy(), Q.copy(), Qb.copy()
        o.initialize_model(Po, Qo, Qbo, len(keys))
        hist = []
        for epoch in range(epochs):
            g.add_jobs(0, U, indptr, keys)
            o.add_jobs(0, U, indptr, keys)
            g.update_parameters()
            o.update_parameters()
            lg, lo = g.compute_loss(probe_u, probe_p, probe_n), o.compute_loss(probe_u, probe_p, probe_n)
            hist.append((lg, lo))
            assert abs(g.current_lr() - o.lr) < 1e-12
        lgs = [h[0] for h in hist]
        assert all(b <= a * 1.02 for a, b in zip(lgs, lgs[1:])), hist           # monotone learning
        assert lgs[-1] < 0.5 * lgs[0] and hist[-1][1] < 0.5 * hist[0][1], hist  # both learn
        assert abs(hist[-1][0] - hist[-1][1]) < 0.25 * hist[-1][1], (seed, hist)


@pytest.mark.parametrize("score,optimizer,d", [("dot", "adagrad", 64), ("l2", "adam", 40)])
def test_warp_epochs_match_oracle(cuda_lib, score, optimizer, d):
    import torch
    U, I = 1200, 900
    indptr, keys, vals, _ = make_csr(U, I, 25000, seed=41)
    opt = sgd_opt(d=d, optimizer=optimizer, score_func=score, max_trials=30, lr=0.05, reg_u=0.01, reg_i=0.02, reg_j=0.03,
                  use_bias=False, per_coordinate_normalize=(optimizer == "adam"), num_iters=3)
    P, Q = init_factors(U, d, d, 1, scale=0.3, signed=True), init_factors(I, d, d, 2, scale=0.3, signed=True)
    Qb = np.zeros((I, 1), np.float32)
    g, o, (Pg, Qg, _), (Po, Qo, _) = make_pair("warp", opt, P, Q, Qb, indptr, keys)
    dev = torch.device("cuda:0")
    tt = torch.zeros(len(keys), dtype=torch.int32, device=dev)
    tn = torch.zeros(len(keys), dtype=torch.int32, device=dev)
    g.set_trace(tt, tn)
    for epoch in range(3):
        ot, on = np.zeros(len(keys), np.int32), np.zeros(len(keys), np.int32)
        g.add_jobs(0, U, indptr, keys)
        o.add_jobs(0, U, indptr, keys, trials_out=ot, negs_out=on)
        g.wait_until_done()
        gt, gn = tt.cpu().numpy(), tn.cpu().numpy()
        # rank sampling is bit-exact unless a score sits within fp32 rounding of the margin
        mism = (gt != ot) | (gn != on)
        assert mism.mean() < 2e-3, (epoch, mism.mean())
        gP = g.grad_tensor(0, (U, d)).cpu().numpy()
        if not mism.any():
            assert rel_err(gP, o.gP) < 1e-4
        g.update_parameters()
        o.update_parameters()
        if not mism.any():
            assert rel_err(Pg, Po) < 1e-4 and rel_err(Qg, Qo) < 1e-4
        assert np.linalg.norm(Pg, axis=1).max() <= 1.0 + 1e-5      # warp.cc:196-200
        loss_sum, updates = g.read_stats()
        assert updates > 0
    probe = (np.arange(50, dtype=np.int32) % U, keys[:50].copy(), ((keys[:50] + 3) % I).astype(np.int32))
    assert abs(g.compute_loss(*probe) - o.compute_loss(*probe)) <= 0.021


def test_warp_rejects_sgd_optimizer(cuda_lib):
    from buffalo_b200 import backend
    g = backend.CuSGD("warp")
    assert g.init(sgd_opt(optimizer="sgd")) is False


def _planted(U=12000, I=2500, seed=5):
    """Planted rank-4 preference matrix; one observed item per user (with >= 3 positives) is held out."""
    rng = np.random.default_rng(seed)
    A, B = rng.normal(size=(U, 4)).astype(np.float32), rng.normal(size=(I, 4)).astype(np.float32)
    S = A @ B.T
    rows, cols = np.nonzero(S > np.quantile(S[:2000], 0.985))
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order].astype(np.int32)
    beg = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=U))[:-1]])
    cnt = np.bincount(rows, minlength=U)
    held_pos = np.full(U, -1, np.int64)
    for u in np.nonzero(cnt >= 3)[0]:
        held_pos[u] = beg[u] + rng.integers(0, cnt[u])
    keep = np.ones(len(rows), bool)
    keep[held_pos[held_pos >= 0]] = False
    held_item = np.full(U, -1, np.int32)
    held_item[held_pos >= 0] = cols[held_pos[held_pos >= 0]]
    rows_t, keys = rows[keep], cols[keep]
    indptr = np.cumsum(np.bincount(rows_t, minlength=U)).astype(np.int64)
    return U, I, indptr, np.ascontiguousarray(keys), held_item


def _hr_at_10(P, Q, Qb, indptr, keys, held_item, users):
    """Fraction of sampled users whose held-out item ranks in the top 10 of the unseen items (accuracy@10 of
    buffalo/evaluate/base.py:93 with one ground-truth item = hit rate@10).  Ranking by the device top-k."""
    from buffalo_b200 import backend
    beg = np.concatenate([[0], indptr[:-1]])
    max_seen = int((indptr[users] - beg[users]).max())
    top = backend.topk_host(P[users], Q, Qb, 10 + max_seen)
    hits = 0
    for r, u in enumerate(users):
        seen = set(keys[beg[u]:indptr[u]].tolist())
        ranked = [c for c in top[r].tolist() if c not in seen][:10]
        hits += int(held_item[u] in ranked)
    return hits / len(users)


@pytest.mark.parametrize("kind,optimizer,lr,epochs", [("bpr", "sgd", 0.1, 10), ("warp", "adagrad", 0.1, 6)])
def test_hr10_matches_oracle_over_seeds(cuda_lib, kind, optimizer, lr, epochs):
    """north_star: "BPR/WARP loss trajectory and HR@10 within tolerance under fixed seed".  BPRMF (plain sgd, the
    reference default) and WARP are trained on the same planted matrix by the GPU backend and by the oracle for 5
    seeds each; the mean HR@10 of the two must agree within a band derived from the oracle's own seed-to-seed spread
    (3 combined standard errors, floor 0.02), and both must clearly beat the popularity-free random baseline."""
    import oracle
    U, I, indptr, keys, held_item = _planted()
    users = np.nonzero(held_item >= 0)[0]
    users = np.random.default_rng(0).choice(users, size=1500, replace=False)
    d = 32
    hr_g, hr_o = [], []
    for seed in (1, 2, 3, 4, 5):
        opt = sgd_opt(d=d, optimizer=optimizer, lr=lr, random_seed=seed, num_iters=epochs, reg_u=0.01, reg_i=0.01,
                      reg_j=0.01, reg_b=0.01, use_bias=(kind == "bpr"), max_trials=100)
        P = init_factors(U, d, d, seed, scale=0.05, signed=(kind == "warp"))
        Q = init_factors(I, d, d, seed + 10, scale=0.05, signed=(kind == "warp"))
        Qb = np.zeros((I, 1), np.float32)
        g, _, (Pg, Qg, Qbg), _ = make_pair(kind, opt, P, Q, Qb, indptr, keys)
        o = oracle.OracleSGD(warp=(kind == "warp"), use_lut=(kind == "bpr"))
        o.init(dict(opt, num_workers=8))
        Po, Qo, Qbo = P.copy(), Q.copy(), Qb.copy()
        o.initialize_model(Po, Qo, Qbo, len(keys))
        for _ in range(epochs):
            g.add_jobs(0, U, indptr, keys)
            o.add_jobs(0, U, indptr, keys)
            g.update_parameters()
            o.update_parameters()
        g.wait_until_done()
        hr_g.append(_hr_at_10(Pg, Qg, Qbg if kind == "bpr" else None, indptr, keys, held_item, users))
        hr_o.append(_hr_at_10(Po, Qo, Qbo if kind == "bpr" else None, indptr, keys, held_item, users))
    mg, mo = float(np.mean(hr_g)), float(np.mean(hr_o))
    se = float(np.sqrt(np.var(hr_o, ddof=1) / 5 + np.var(hr_g, ddof=1) / 5))
    band = max(3 * se, 0.02)
    assert mo > 20 * 10.0 / I and mg > 20 * 10.0 / I, (hr_g, hr_o)       # far above chance (10 / I)
    assert abs(mg - mo) <= band, (kind, hr_g, hr_o, band)
