"""One sweep from init of pg / cg / tncg on a config, device (events) vs the reference build (OpenMP,
all host cores).  Usage: python scripts/method_compare.py [config] [dtype]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from oracle.oracle import Ref  # noqa: E402
from poismf_b200 import make_params  # noqa: E402
from poismf_b200.device import DeviceFit  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "small"]
dtype = np.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else np.float32
csr, csc, A0, B0 = bench.make_problem(cfg, dtype)
dimA, dimB, k = cfg["dimA"], cfg["dimB"], cfg["k"]
nnz = csr[0].shape[0]
cases = {"pg": dict(l2_reg=1e9, step_size=1e-7, maxupd=1), "cg": dict(l2_reg=1e4, maxupd=5, limit_step=True),
         "tncg": dict(l2_reg=1e3, maxupd=15 * k)}
fit = DeviceFit(dimA, dimB, k, dtype, device=0)
stream = torch.cuda.Stream()
fit.set_stream(stream.cuda_stream)
fit.set_csr_csc(csr, csc)
ref = Ref(dtype, fast=True) if Ref.available(dtype, fast=True) else None
for method, hp in cases.items():
    if os.environ.get("ONLY") and method not in os.environ["ONLY"].split(","):
        continue
    params = make_params(method, numiter=1, **hp)
    ts = []
    for it in range(3):
        fit.set_factors(A0, B0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream); fit.sweeps(params); e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    line = f"{method:5s} {np.dtype(dtype).name}: device {min(ts):9.2f} ms/sweep ({nnz / min(ts) / 1e3:8.1f} M nnz/s)"
    if ref is not None:
        A, B = A0.copy(), B0.copy()
        t0 = time.perf_counter()
        ref.run_poismf(A, B, csr, csc, method, numiter=1, nthreads=os.cpu_count(), **hp)
        dt = 1e3 * (time.perf_counter() - t0)
        line += f"   reference ({os.cpu_count()} threads) {dt:9.1f} ms/sweep   ratio {dt / min(ts):6.1f}x"
    print(line, flush=True)
