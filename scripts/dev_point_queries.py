"""End-to-end time of the point-query entry points (host buffers in and out) next to the reference's own C
path on the host cores: predict_multiple for n = 1e3 .. 1e7 pairs, single-user topN with and without the
resident item factors.  Factors of the config #2 shape (359k x 160k, k = 50, float32)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poismf_b200 import c_funs
from oracle.oracle import Ref
rng = np.random.default_rng(0)
dimA, dimB, k = 359_000, 160_000, 50
A = np.ascontiguousarray(rng.gamma(0.5, 0.5, size=(dimA, k)).astype(np.float32))
B = np.ascontiguousarray(rng.gamma(0.5, 0.5, size=(dimB, k)).astype(np.float32))
ref = Ref(np.float32, fast=True)
cores = os.cpu_count()
def best(f, reps=5):
    f(); return min(timeit(f) for _ in range(reps))
def timeit(f):
    t0 = time.perf_counter(); f(); return time.perf_counter() - t0
print(f"predict_multiple, {dimA} x {dimB}, k={k}: pairs | device e2e ms | reference ({cores} threads) ms")
for n in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
    ixA = rng.integers(0, dimA, n).astype(np.uint64); ixB = rng.integers(0, dimB, n).astype(np.uint64)
    out = np.empty(n, np.float32)
    t_dev = best(lambda: c_funs._predict_multiple(out, A, B, ixA, ixB))
    t_ref = best(lambda: ref.predict_multiple(A, B, ixA, ixB, nthreads=cores))
    print(f"  {n:10d} | {1e3 * t_dev:9.3f} | {1e3 * t_ref:9.3f}")
none = np.empty(0, np.uint64)
a = np.ascontiguousarray(A[7])
excl = np.sort(rng.choice(dimB, 200, replace=False)).astype(np.uint64)
for label, env in (("upload B per call", None), ("resident B (POISMF_B200_CACHE_FACTORS=1)", "1")):
    if env: os.environ["POISMF_B200_CACHE_FACTORS"] = env
    t_dev = best(lambda: c_funs._call_topN(a, B, none, excl.copy(), top_n=10, output_score=True), reps=10)
    print(f"topN one user, {dimB} items, top-10, 200 exclusions, {label}: device e2e {1e3 * t_dev:.3f} ms")
t_ref = best(lambda: ref.topN(a, B, 10, exclude=excl, nthreads=cores), reps=10)
print(f"topN one user, reference ({cores} threads): {1e3 * t_ref:.3f} ms")
