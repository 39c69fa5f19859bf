"""Independent NumPy fp64 restatement of the ALS row solves (TEST INFRASTRUCTURE ONLY).

Written from the maths in SURVEY.md Appendix B, not from buffalo_oracle.c, so that the two
restatements check each other (tests/test_oracle.py).  Everything is float64; the C oracle and
the CUDA kernels compute in float32, so agreement is expected to ~1e-5 relative.

References: lib/algo_impl/als/als.cc:107-209 (direct system), lib/algo.cc:58-82 (manual CG),
als.cc:211-358 (iALS++), als.cc:175-200 + buffalo/algo/als.py:171 (loss).
"""
import numpy as np


def row_slices(indptr_end):
    """indptr holds exclusive end offsets, no leading zero (als.cc:156-157)."""
    beg = np.concatenate([[0], indptr_end[:-1]])
    return beg, indptr_end


def gram(F):
    F = F.astype(np.float64)
    return F.T @ F


def system(Y, G, cols, vals, alpha, reg, adaptive_reg):
    """M = G + a*sum v y y^T + reg*kappa*I ;  b = sum (1 + a v) y   (als.cc:180-202)"""
    Yc = Y[cols].astype(np.float64)
    v = vals.astype(np.float64)
    M = G + alpha * (Yc * v[:, None]).T @ Yc
    kappa = float(len(cols)) if adaptive_reg else 1.0
    M = M + reg * kappa * np.eye(G.shape[0])
    b = ((1.0 + alpha * v)[:, None] * Yc).sum(axis=0)
    return M, b


def manual_cg(M, b, x, iters, eps, tol):
    """lib/algo.cc:58-82"""
    x = x.astype(np.float64).copy()
    r = b - x @ M
    if b @ b < r @ r:
        x[:] = 0.0
        r = b.copy()
    p = r.copy()
    rs_old = r @ r
    for _ in range(iters):
        Ap = p @ M
        a = rs_old / (Ap @ p + eps)
        x += a * p
        r -= a * Ap
        rs_new = r @ r
        if rs_new < tol:
            break
        p = r + (rs_new / (rs_old + eps)) * p
        rs_old = rs_new
    return x


def als_half_epoch(X, Y, indptr_end, keys, vals, opt, axis):
    """Returns (new X as float64, loss numerator, loss denominator) for one half-epoch."""
    X = X.astype(np.float64).copy()
    Yd = Y.astype(np.float64)
    G = gram(Y)
    d = X.shape[1]
    alpha = float(opt.get("alpha", 8.0))
    reg = float(opt.get("reg_u", 0.1) if axis == 0 else opt.get("reg_i", 0.1))
    adaptive = bool(opt.get("adaptive_reg", False))
    optimizer = opt.get("optimizer", "manual_cg")
    if d >= 128:
        optimizer = "ialspp"  # als.cc:46
    eps = float(opt.get("eps", 1e-10))
    tol = float(opt.get("cg_tolerance", 1e-10))
    iters = int(opt.get("num_cg_max_iters", 3))
    bs_opt = min(d, int(opt.get("block_size", 32)))
    compute_loss = bool(opt.get("compute_loss_on_training", True))
    beg, end = row_slices(indptr_end)
    nume = deno = 0.0
    for u in range(X.shape[0]):
        c = keys[beg[u]:end[u]]
        v = vals[beg[u]:end[u]].astype(np.float64)
        n = len(c)
        if n == 0:
            continue
        x = X[u].copy()
        kappa = float(n) if adaptive else 1.0
        if compute_loss:
            if axis == 1:
                dots = Yd[c] @ x
                nume += x @ G @ x - (dots ** 2).sum() + ((dots - 1.0) ** 2 * (1.0 + alpha * v)).sum()
                deno += Y.shape[0] + (alpha * v).sum()
            nume += kappa * reg * (x @ x)
        if optimizer in ("llt", "ldlt"):
            M, b = system(Y, G, c, vals[beg[u]:end[u]], alpha, reg, adaptive)
            X[u] = np.linalg.solve(M, b)
        elif optimizer == "manual_cg":
            M, b = system(Y, G, c, vals[beg[u]:end[u]], alpha, reg, adaptive)
            X[u] = manual_cg(M, b, x, iters, eps, tol)
        elif optimizer == "ialspp":
            Yc = Yd[c]
            yhat = Yc @ x
            for bb in range(0, d, bs_opt):
                bs = bs_opt if bb + bs_opt < d else d - bb
                sl = slice(bb, bb + bs)
                A = G[sl, sl] + reg * np.eye(bs)          # no adaptive reg here (als.cc:278)
                g = x @ G[:, sl] + reg * x[sl] + ((yhat - 1.0) * v * alpha) @ Yc[:, sl]
                H = A + (Yc[:, sl] * (v * alpha)[:, None]).T @ Yc[:, sl]
                dl = np.zeros(bs)
                r = g.copy()
                p = r.copy()
                rsold = r @ r
                if rsold > tol:
                    for _ in range(3):
                        Ap = H @ p
                        step = rsold / (p @ Ap)
                        dl += step * p
                        r -= step * Ap
                        rsnew = r @ r
                        if rsnew < tol:
                            break
                        p = r + (rsnew / rsold) * p
                        rsold = rsnew
                x[sl] -= dl
                yhat -= Yc[:, sl] @ dl
            X[u] = x
        else:
            raise ValueError(optimizer)
    return X, nume, deno


# ---------------------------------------------------------------------------------------------------
# BPRMF / WARP (fp64, written from lib/algo_impl/bpr/bpr.cc:119-188, lib/algo_impl/warp/warp.cc:30-52,137-163,
# 192-226 -- NOT from buffalo_oracle.c) so that the C restatement of the update maths has an independent pin
# ---------------------------------------------------------------------------------------------------
def bpr_logit(x, max_exp=6.0):
    """1 - sigmoid(x) with the reference's clamp (bpr.cc:123-131; the exact function, not the 1000-entry table)."""
    if x > max_exp:
        return 0.0
    if x < -max_exp:
        return 1.0
    return 1.0 / (1.0 + np.exp(x))


def bpr_accumulate(P, Q, Qb, users, positives, negatives, use_bias=True, update_i=True, update_j=True,
                   per_coordinate_normalize=True, num_negative_samples=1):
    """Gradient-accumulating optimizers (bpr.cc:138-156,172-179): P, Q are read-only, so the order of the triples
    does not matter.  `users/positives/negatives` list one entry per (positive, negative draw)."""
    P, Q, Qb = P.astype(np.float64), Q.astype(np.float64), Qb.astype(np.float64)
    gP, gQ, gQb = np.zeros_like(P), np.zeros_like(Q), np.zeros_like(Qb)
    cP, cQ = np.zeros(P.shape[0], np.int64), np.zeros(Q.shape[0], np.int64)
    for t, (u, i, j) in enumerate(zip(users, positives, negatives)):
        x = P[u] @ (Q[i] - Q[j]) + ((Qb[i, 0] - Qb[j, 0]) if use_bias else 0.0)
        lg = bpr_logit(x)
        if per_coordinate_normalize:
            cQ[j] += 1
        gP[u] += lg * (Q[i] - Q[j])
        if update_i:
            gQ[i] += lg * P[u]
            if use_bias:
                gQb[i, 0] += lg
        if update_j:
            gQ[j] -= lg * P[u]
            if use_bias:
                gQb[j, 0] -= lg
        if per_coordinate_normalize and t % num_negative_samples == num_negative_samples - 1:   # once per positive
            cP[u] += 1
            cQ[i] += 1
    return gP, gQ, gQb, cP, cQ


def bpr_sgd(P, Q, Qb, users, positives, negatives, lr, reg_u, reg_i, reg_j, reg_b, use_bias=True, update_i=True,
            update_j=True):
    """Plain SGD, sequential (bpr.cc:157-171).  `g` is a lazy Eigen expression: the user row is updated with the ALREADY
    updated item rows."""
    P, Q, Qb = P.astype(np.float64).copy(), Q.astype(np.float64).copy(), Qb.astype(np.float64).copy()
    for u, i, j in zip(users, positives, negatives):
        x = P[u] @ (Q[i] - Q[j]) + ((Qb[i, 0] - Qb[j, 0]) if use_bias else 0.0)
        lg = bpr_logit(x)
        deriv = lg * P[u]
        if update_i:
            Q[i] += lr * (deriv - reg_i * Q[i])
            if use_bias:
                Qb[i, 0] += lr * (lg - reg_b * Qb[i, 0])
        if update_j:
            Q[j] += lr * (-deriv - reg_j * Q[j])
            if use_bias:
                Qb[j, 0] += lr * (-lg - reg_b * Qb[j, 0])
        P[u] += lr * (lg * (Q[i] - Q[j]) - reg_u * P[u])
    return P, Q, Qb


def bpr_loss(P, Q, Qb, users, positives, negatives, use_bias=True):
    """bpr.cc:227-244"""
    P, Q, Qb = P.astype(np.float64), Q.astype(np.float64), Qb.astype(np.float64)
    tot = 0.0
    for u, i, j in zip(users, positives, negatives):
        x = P[u] @ Q[i] - P[u] @ Q[j] + ((Qb[i, 0] - Qb[j, 0]) if use_bias else 0.0)
        tot += np.log(1.0 + np.exp(-x))
    return tot / len(users)


def warp_accumulate(P, Q, indptr_end, keys, trials, negs, reg_u, reg_i, reg_j, threshold=1.0, score="dot"):
    """Given the trace of the rank-sampling loop (trial count and violating negative per positive; 0 trials = discarded),
    the accumulated gradients, sample counters and partial loss (warp.cc:30-52,150-163)."""
    P, Q = P.astype(np.float64), Q.astype(np.float64)
    I = Q.shape[0]
    gP, gQ = np.zeros_like(P), np.zeros_like(Q)
    cP, cQ = np.zeros(P.shape[0], np.int64), np.zeros(I, np.int64)
    beg = np.concatenate([[0], indptr_end[:-1]])
    rows = np.repeat(np.arange(len(indptr_end)), indptr_end - beg)
    loss, updates = 0.0, 0

    def sc(a, b):
        return a @ b if score == "dot" else -((a - b) @ (a - b))
    for k in range(len(keys)):
        if trials[k] == 0:
            continue
        u, i, j = rows[k], keys[k], negs[k]
        n_seen = int(indptr_end[u] - beg[u])
        phi = np.log(max(1, int((I - n_seen - 1) // int(trials[k]))))
        if score == "dot":
            du, di, dj = phi * (Q[i] - Q[j]), phi * P[u], -phi * P[u]
        else:
            du, di, dj = phi * 2 * (Q[i] - Q[j]), phi * (P[u] - Q[i]), -phi * (P[u] - Q[j])
        gP[u] += du - reg_u * P[u]
        gQ[i] += di - reg_i * Q[i]
        gQ[j] += dj - reg_j * Q[j]
        cP[u] += 1
        cQ[i] += 1
        cQ[j] += 1
        loss += sc(P[u], Q[j]) - sc(P[u], Q[i]) + threshold
        updates += 1
    return gP, gQ, cP, cQ, loss, updates


def warp_project(F):
    """rows are pulled back onto the unit ball (warp.cc:194-200)"""
    F = F.astype(np.float64)
    nrm = np.sqrt((F * F).sum(axis=1))
    return F / np.maximum(1.0, nrm)[:, None]


def warp_loss(P, Q, users, positives, negatives, threshold=1.0, score="dot"):
    """fraction of probe triples that violate the margin (warp.cc:205-223)"""
    P, Q = P.astype(np.float64), Q.astype(np.float64)

    def sc(a, b):
        return a @ b if score == "dot" else -((a - b) @ (a - b))
    return float(np.mean([(sc(P[u], Q[i]) - sc(P[u], Q[j])) < threshold for u, i, j in zip(users, positives, negatives)]))
