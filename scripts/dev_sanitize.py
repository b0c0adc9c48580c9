"""compute-sanitizer target: one cg sweep (register-tile teams of every size + lock-step path) and one pg sweep
on a small power-law problem.   compute-sanitizer --tool racecheck|memcheck python scripts/dev_sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["POISMF_B200_DENSE_MIN_TOTAL"] = "0"
os.environ["POISMF_B200_DENSE_MIN"] = "600"
from poismf_b200 import make_params
from poismf_b200.device import DeviceFit
from poismf_b200.synth import init_factors, powerlaw_counts
dimA, dimB, k = 3000, 600, 50
csr, csc = powerlaw_counts(dimA, dimB, 120_000, dtype=np.float32, seed=2)
print("max row", np.diff(csr[1].astype(np.int64)).max(), "max col", np.diff(csc[1].astype(np.int64)).max())
A0, B0 = init_factors(dimA, dimB, k, dtype=np.float32)
for method, hp in (("cg", dict(l2_reg=1e3, maxupd=3, limit_step=True)), ("pg", dict(l2_reg=1e6, step_size=1e-6, maxupd=2))):
    fit = DeviceFit(dimA, dimB, k, np.float32)
    fit.set_csr_csc(csr, csc); fit.set_factors(A0, B0)
    fit.sweeps(make_params(method, numiter=1, **hp)); fit.sync()
    A, B = fit.get_factors()
    print(method, "finite", bool(np.isfinite(A).all() and np.isfinite(B).all()))
    fit.close()
