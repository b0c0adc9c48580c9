"""tncg / pg sweeps on the 1/8-scale workload with per-bin device times."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_problem, team_name
from poismf_b200 import make_params
from poismf_b200.device import DeviceFit
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "small"]
csr, csc, A0, B0 = make_problem(cfg)
print("nnz", csr[0].shape[0], flush=True)
fit = DeviceFit(cfg["dimA"], cfg["dimB"], cfg["k"], np.float32)
fit.set_csr_csc(csr, csc)
for method, hp in (("pg", dict(l2_reg=1e9, maxupd=1, step_size=1e-7)), ("tncg", dict(l2_reg=1e3, maxupd=int(sys.argv[2]) if len(sys.argv) > 2 else 50))):
    p = make_params(method, numiter=1, **hp)
    fit.set_factors(A0, B0); fit.set_profiling(True)
    t0 = time.time(); fit.sweeps(p); fit.sync(); dt = time.time() - t0
    print(method, hp, "wall ms", round(1e3 * dt, 2), flush=True)
    for b in fit.get_profile():
        print("   side", b["side"], team_name(b["block_team"]), "cap", b["cap"], "rows", b["nrows"], "nnz", b["nnz"], "ms", round(b["ms"], 3), flush=True)
    fit.set_profiling(False)
