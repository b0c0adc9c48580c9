import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import CASES, hyper, problem, run_device
from oracle.oracle import Ref, Restatement
for dt in (np.float64, np.float32):
    for prob in ("pl2k",):
        csr, csc, A0, B0, k = problem(prob, dt)
        orc = Restatement(dt)
        for case in ["cg", "cg_nolimit_w", "tncg", "tncg_w", "tncg_reuse_stop"]:
            method, kw = hyper(case, k)
            if method == "cg": kw["numiter"] = 10
            Ar, Br = A0.copy(), B0.copy(); orc.run_poismf(Ar, Br, csr, csc, method, **kw)
            A2, B2 = A0.copy(), B0.copy(); Ref(dt, fast=True).run_poismf(A2, B2, csr, csc, method, **kw)
            lr = orc.llk(Ar, Br, csr); l2 = orc.llk(A2, B2, csr)
            out = f"{np.dtype(dt).name} {case}: llk_ref {lr:.6e} noise(ref fast vs strict) {abs(l2-lr)/abs(lr):.2e} zeros ref A {np.mean(Ar==0):.3f} B {np.mean(Br==0):.3f} | reffast A {np.mean(A2==0):.3f} B {np.mean(B2==0):.3f} small(<1e-6) reffast A {np.mean(np.abs(A2)<1e-6):.3f} B {np.mean(np.abs(B2)<1e-6):.3f} |"
            for nm, fl in (("strict", 1), ("fast", 0)):
                A, B = A0.copy(), B0.copy(); run_device(csr, csc, A, B, method, kw, flags=fl)
                ld = orc.llk(A, B, csr)
                out += f" {nm}: dllk {abs(ld-lr)/abs(lr):.2e} zeros A {np.mean(A==0):.3f} B {np.mean(B==0):.3f} small A {np.mean(np.abs(A)<1e-6):.3f} B {np.mean(np.abs(B)<1e-6):.3f};"
            print(out, flush=True)
