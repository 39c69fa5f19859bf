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
