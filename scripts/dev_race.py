"""One sweep on a tiny problem (for compute-sanitizer racecheck / synccheck / memcheck).
Usage: python scripts/dev_race.py [f64|f32] [maxupd] [method] [flags]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import problem, run_device
dtype = np.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else np.float32
maxupd = int(sys.argv[2]) if len(sys.argv) > 2 else 30
method = sys.argv[3] if len(sys.argv) > 3 else "tncg"
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
csr, csc, A0, B0, k = problem("pl2k", dtype)
A, B = A0.copy(), B0.copy()
rc = run_device(csr, csc, A, B, method, dict(l2_reg=1e3, maxupd=maxupd, numiter=1, limit_step=True, step_size=1e-4), flags=flags)
print("rc", rc, np.isfinite(A).all(), np.isfinite(B).all(), flush=True)
