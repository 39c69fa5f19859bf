"""Synthetic inputs shared by the tests (seeded, NumPy only)."""
import numpy as np


def make_csr(num_rows, num_cols, nnz, seed, empty_rows=0, vals="ints", sort_keys=True):
    """Random CSR in the reference layout (buffalo/data/base.py:187-192): indptr = exclusive END
    offsets (no leading zero), keys int32 sorted within a row, vals float32."""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, num_rows, nnz)
    if empty_rows:
        dead = rng.choice(num_rows, size=empty_rows, replace=False)
        rows = rows[~np.isin(rows, dead)]
    cols = rng.integers(0, num_cols, len(rows)).astype(np.int32)
    pairs = np.unique(np.stack([rows, cols], axis=1), axis=0) if sort_keys else np.stack([rows, cols], axis=1)
    rows, cols = pairs[:, 0], pairs[:, 1].astype(np.int32)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    indptr = np.cumsum(np.bincount(rows, minlength=num_rows)).astype(np.int64)
    if vals == "ints":
        v = rng.integers(1, 6, len(rows)).astype(np.float32)
    else:
        v = np.ones(len(rows), dtype=np.float32)
    return indptr, np.ascontiguousarray(cols), v, rows.astype(np.int64)


def transpose_csr(indptr, keys, vals, num_rows, num_cols):
    """colwise copy, sorted by (col, row) like fileio.hpp:330-341."""
    beg = np.concatenate([[0], indptr[:-1]])
    rows = np.repeat(np.arange(num_rows), indptr - beg)
    order = np.lexsort((rows, keys))
    cind = np.cumsum(np.bincount(keys, minlength=num_cols)).astype(np.int64)
    return cind, rows[order].astype(np.int32), np.ascontiguousarray(vals[order])


def init_factors(rows, d, vdim, seed, scale=None, signed=False):
    """abs(N(0, 1/d^2)) like buffalo/algo/als.py:85-86 (scaled up so the problem is well conditioned)."""
    rng = np.random.default_rng(seed)
    s = (1.0 / d ** 2) if scale is None else scale
    F = rng.normal(scale=s, size=(rows, d)).astype(np.float32)
    if not signed:
        F = np.abs(F)
    out = np.zeros((rows, vdim), dtype=np.float32)
    out[:, :d] = F
    return out


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
