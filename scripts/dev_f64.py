import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from poismf_b200 import make_params, _lib
from poismf_b200.device import DeviceFit
cfg = bench.CONFIGS["small"]
dtype = np.float64
csr, csc, A0, B0 = bench.make_problem(cfg, dtype)
print("max row", np.diff(csr[1].astype(np.int64)).max(), "max col", np.diff(csc[1].astype(np.int64)).max(), flush=True)
method = sys.argv[1]
hp = {"pg": dict(l2_reg=1e9, step_size=1e-7, maxupd=1), "cg": dict(l2_reg=1e4, maxupd=5, limit_step=True), "tncg": dict(l2_reg=1e3, maxupd=int(os.environ.get("MAXUPD", "750")))}[method]
fit = DeviceFit(cfg["dimA"], cfg["dimB"], cfg["k"], dtype, device=0)
if os.environ.get("TORCHSTREAM"):
    import torch
    stream = torch.cuda.Stream()
    fit.set_stream(stream.cuda_stream)
fit.set_csr_csc(csr, csc)
if not os.environ.get("NOPROF"): fit.set_profiling(True)
fit.set_factors(A0, B0)
for _ in range(int(os.environ.get("REPS", "1"))):
    fit.set_factors(A0, B0)
    fit.sweeps(make_params(method, numiter=1, **hp))
A, B = fit.get_factors()
print(method, "ok", np.isfinite(A).all(), np.isfinite(B).all(), flush=True)
for p in fit.get_profile(): print(p)
