"""Device time of one sweep of config #2 as a function of maxupd (fixed vs per-iteration cost)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, make_problem
from poismf_b200 import make_params
from poismf_b200.device import DeviceFit
cfg = CONFIGS["c2"]
csr, csc, A0, B0 = make_problem(cfg)
fit = DeviceFit(cfg["dimA"], cfg["dimB"], cfg["k"], np.float32)
st = torch.cuda.Stream(); fit.set_stream(st.cuda_stream); fit.set_csr_csc(csr, csc)
for method, hp_list in (("cg", [dict(l2_reg=1e4, maxupd=m, limit_step=True) for m in (1, 2, 3, 5, 8)]),
                        ("pg", [dict(l2_reg=1e9, maxupd=m, step_size=1e-7) for m in (1, 2, 4)]),
                        ("tncg", [dict(l2_reg=1e3, maxupd=m) for m in (50, 750)])):
    for hp in hp_list:
        p = make_params(method, numiter=1, **hp)
        ts = []
        for rep in range(4):
            fit.set_factors(A0, B0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record(st); fit.sweeps(p); e1.record(st)
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(method, hp, "ms/sweep", round(min(ts[1:]), 3), flush=True)
