"""Synthetic sparse count matrices for tests and benchmarks.

`readme_counts` is the reference's README recipe (/root/reference/README.md:91-100);
`powerlaw_counts` draws Zipf-like user activity / item popularity so that the
degree distribution has the heavy rows and columns real implicit-feedback data has.
Both return (csr, csc) triples of numpy arrays: values, indptr (uint64), indices (uint64),
with duplicates summed and indices sorted inside each row/column, which is what
SciPy's tocsr()/tocsc() hand to the reference (poismf/__init__.py:403-414).
"""
from __future__ import annotations

import numpy as np


def _coo_to_csr_csc(rows, cols, vals, dimA, dimB, dtype):
    key = rows.astype(np.int64) * np.int64(dimB) + cols.astype(np.int64)
    order = np.argsort(key, kind="stable")
    key = key[order]
    vals = vals[order]
    uniq, start = np.unique(key, return_index=True)
    summed = np.add.reduceat(vals.astype(np.float64), start)
    r = (uniq // dimB).astype(np.int64)
    c = (uniq % dimB).astype(np.int64)
    # CSR (already sorted by (row, col))
    indptr_r = np.zeros(dimA + 1, dtype=np.uint64)
    np.cumsum(np.bincount(r, minlength=dimA), out=indptr_r[1:])
    csr = (summed.astype(dtype), indptr_r, c.astype(np.uint64))
    # CSC
    order_c = np.lexsort((r, c))
    indptr_c = np.zeros(dimB + 1, dtype=np.uint64)
    np.cumsum(np.bincount(c, minlength=dimB), out=indptr_c[1:])
    csc = (summed[order_c].astype(dtype), indptr_c, r[order_c].astype(np.uint64))
    return csr, csc


def readme_counts(nusers=100, nitems=1000, nnz=10_000, seed=1, dtype=np.float64):
    """The README smoke shape (BASELINE config #1)."""
    rs = np.random.RandomState(seed)
    u = rs.randint(nusers, size=nnz)
    i = rs.randint(nitems, size=nnz)
    x = 1 + rs.gamma(1, 1, size=nnz).astype(int)
    return _coo_to_csr_csc(u, i, x, nusers, nitems, dtype)


def powerlaw_counts(dimA, dimB, nnz, alpha_a=0.6, alpha_b=0.9, seed=1, dtype=np.float32,
                    chunk=1 << 24):
    """Power-law degree synthetic: p_u ∝ rank^-alpha_a, p_i ∝ rank^-alpha_b.

    Draws `nnz` (u,i) pairs (duplicates are summed, so the number of stored
    non-zeros is slightly below `nnz`), counts x = 1 + Geometric(0.5)-1 ≥ 1.
    Row/column identities are shuffled so that heavy rows are not contiguous.
    """
    rng = np.random.default_rng(seed)
    pa = np.arange(1, dimA + 1, dtype=np.float64) ** (-alpha_a)
    pb = np.arange(1, dimB + 1, dtype=np.float64) ** (-alpha_b)
    cda = np.cumsum(pa); cda /= cda[-1]
    cdb = np.cumsum(pb); cdb /= cdb[-1]
    perm_a = rng.permutation(dimA)
    perm_b = rng.permutation(dimB)
    us, is_, xs = [], [], []
    left = nnz
    while left > 0:
        m = min(left, chunk)
        us.append(perm_a[np.searchsorted(cda, rng.random(m), side="right").clip(max=dimA - 1)])
        is_.append(perm_b[np.searchsorted(cdb, rng.random(m), side="right").clip(max=dimB - 1)])
        xs.append(rng.geometric(0.5, size=m).astype(np.float32))
        left -= m
    u = np.concatenate(us); i = np.concatenate(is_); x = np.concatenate(xs)
    return _coo_to_csr_csc(u, i, x, dimA, dimB, dtype)


def init_factors(dimA, dimB, k, seed=1, dtype=np.float32):
    """A,B = 0.3 + U(0, 0.01), drawn in float64 then cast (poismf/__init__.py:419-425)."""
    rng = np.random.default_rng(seed)
    A = 0.3 + rng.uniform(low=0, high=0.01, size=(dimA, k))
    B = 0.3 + rng.uniform(low=0, high=0.01, size=(dimB, k))
    return np.ascontiguousarray(A.astype(dtype)), np.ascontiguousarray(B.astype(dtype))
