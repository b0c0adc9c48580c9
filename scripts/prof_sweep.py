"""One or more sweeps of a bench config on the device, nothing else (for ncu).
    python scripts/prof_sweep.py [config] [n_sweeps] [method]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_problem
from poismf_b200 import make_params
from poismf_b200.device import DeviceFit
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = dict(CONFIGS[name])
csr, csc, A0, B0 = make_problem(cfg)
fit = DeviceFit(cfg["dimA"], cfg["dimB"], cfg["k"], np.float32)
fit.set_csr_csc(csr, csc); fit.set_factors(A0, B0)
p = make_params(cfg["method"], numiter=n, **cfg["hp"])
fit.sweeps(p); fit.sync()
print("done")
