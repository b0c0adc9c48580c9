import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import test_boundary as tb
from scipy.sparse import coo_matrix
from oracle.oracle import Ref, Restatement
from conftest import run_device
df = tb._coo_frame()
import pandas as pd
u, um = pd.factorize(df.UserId); i, im = pd.factorize(df.ItemId)
X = coo_matrix((df.Count.to_numpy().astype(np.float32), (u, i)), shape=(len(um), len(im)))
csr_, csc_ = X.tocsr(), X.tocsc()
trip = lambda m: (np.ascontiguousarray(m.data), np.ascontiguousarray(m.indptr, dtype=np.uint64), np.ascontiguousarray(m.indices, dtype=np.uint64))
csr, csc = trip(csr_), trip(csc_)
k = 16
rng = np.random.default_rng(7)
A0 = (0.3 + rng.uniform(0, .01, (len(um), k))).astype(np.float32); B0 = (0.3 + rng.uniform(0, .01, (len(im), k))).astype(np.float32)
orc = Restatement(np.float32)
def llk2(A, B):
    r = np.repeat(np.arange(A.shape[0]), np.diff(csr[1].astype(np.int64)))
    pred = np.einsum("ij,ij->i", A[r].astype(np.float64), B[csr[2].astype(np.int64)].astype(np.float64))
    return float((csr[0] * np.log(pred)).sum()), float(A.sum(0, dtype=np.float64) @ B.sum(0, dtype=np.float64))
for es in (True, False):
  for niter in (1, 3):
    kw = dict(l2_reg=1e3, maxupd=15 * k, numiter=niter, early_stop=es)
    out = {}
    for name, fn in (("strict", lambda A, B: orc.run_poismf(A, B, csr, csc, "tncg", **kw)),
                     ("reffast", lambda A, B: Ref(np.float32, fast=True).run_poismf(A, B, csr, csc, "tncg", nthreads=4, **kw)),
                     ("dev_fast", lambda A, B: run_device(csr, csc, A, B, "tncg", kw)),
                     ("dev_strict", lambda A, B: run_device(csr, csc, A, B, "tncg", kw, flags=1))):
        A, B = A0.copy(), B0.copy(); fn(A, B); out[name] = llk2(A, B) + (float((A == 0).mean()), float((B == 0).mean()))
    print("early_stop", es, "niter", niter)
    for kname, v in out.items(): print(f"   {kname:10s} t1 {v[0]:12.1f} t2 {v[1]:12.1f} llk {v[0]-v[1]:12.1f} zerosA {v[2]:.3f} zerosB {v[3]:.3f}")
