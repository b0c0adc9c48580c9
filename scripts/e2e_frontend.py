"""End-to-end fit of the headline workload from COO triplets: the reference's front end (SciPy
tocsr/tocsc on the host, poismf/__init__.py:402-404) + drop-in run_poismf, against
pmf_b200_fit_coo (conversion on the device), and the drop-in call with the matrix cache on.
Usage: python scripts/e2e_frontend.py [config]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from poismf_b200 import c_funs  # noqa: E402
from scipy.sparse import coo_matrix  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
csr, csc, A0, B0 = bench.make_problem(cfg)
dimA, dimB = A0.shape[0], B0.shape[0]
rng = np.random.default_rng(0)
rows = np.repeat(np.arange(dimA, dtype=np.uint64), np.diff(csr[1].astype(np.int64)))
perm = rng.permutation(rows.shape[0])
rows, cols, vals = rows[perm], csr[2][perm].astype(np.uint64), csr[0][perm]
hp = cfg["hp"]
kw = dict(method=cfg["method"], limit_step=hp.get("limit_step", False), l2_reg=hp["l2_reg"],
          step_size=hp.get("step_size", 1e-7), niter=1, maxupd=hp["maxupd"], early_stop=False, reuse_prev=False)


def t(f, n=3):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t0))
    return min(ts), float(np.mean(ts))


state = {}


def host_front():
    coo = coo_matrix((vals, (rows.astype(np.int64), cols.astype(np.int64))), shape=(dimA, dimB))
    r, c = coo.tocsr(), coo.tocsc()
    state["csr"] = (np.ascontiguousarray(r.data), r.indptr.astype(np.uint64), r.indices.astype(np.uint64))
    state["csc"] = (np.ascontiguousarray(c.data), c.indptr.astype(np.uint64), c.indices.astype(np.uint64))


def dropin():
    A, B = A0.copy(), B0.copy()
    r, c = state["csr"], state["csc"]
    c_funs._run_poismf(r[0], r[2], r[1], c[0], c[2], c[1], A, B, **kw)
    state["AB"] = (A, B)


def coo_fit():
    A, B = A0.copy(), B0.copy()
    c_funs._fit_coo(rows, cols, vals, A, B, **kw)
    state["AB2"] = (A, B)


print("host tocsr+tocsc (min, mean ms):", t(host_front, 2))
print("drop-in run_poismf, pageable host CSR/CSC:", t(dropin))
print("fit_coo (device conversion + sweep), pageable COO:", t(coo_fit))
print("identical factors:", all(np.array_equal(a, b) for a, b in zip(state["AB"], state["AB2"])))
os.environ["POISMF_B200_CACHE_X"] = "1"
print("drop-in run_poismf with POISMF_B200_CACHE_X=1 (matrix resident):", t(dropin))
