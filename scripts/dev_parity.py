"""Developer check (GPU): device path vs the CPU checkers on small seeded problems."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import Ref, Restatement
from poismf_b200 import c_funs, FLAG_STRICT, FLAG_NO_CACHED
from poismf_b200.synth import readme_counts, powerlaw_counts, init_factors

only = sys.argv[1] if len(sys.argv) > 1 else None

def stats(X, Y):
    d = np.abs(X - Y)
    nrm = np.linalg.norm(Y, axis=1) + 1e-300
    rowerr = np.linalg.norm(X - Y, axis=1) / nrm
    return d.max() / max(np.abs(Y).max(), 1e-300), np.median(rowerr), rowerr.max()

cases = [
    ("pg", dict(l2_reg=1e9, step_size=1e-7, maxupd=1, numiter=2)),
    ("pg", dict(l2_reg=1e3, step_size=1e-4, maxupd=3, numiter=2, w_mult=2.5, l1_reg=0.1)),
    ("cg", dict(l2_reg=1e4, maxupd=5, numiter=1, limit_step=True)),
    ("cg", dict(l2_reg=1e3, maxupd=5, numiter=2, limit_step=False, w_mult=1.5)),
    ("tncg", dict(l2_reg=1e3, maxupd=None, numiter=1)),
    ("tncg", dict(l2_reg=1e2, maxupd=None, numiter=3, reuse_prev=True, early_stop=True, l1_reg=0.5)),
    ("tncg", dict(l2_reg=1e3, maxupd=None, numiter=2, w_mult=3.0)),
]
for dt in (np.float64, np.float32):
    orc = Ref(dt) if Ref.available(dt) else Restatement(dt)
    for name, gen, k in (("readme", lambda: readme_counts(dtype=dt), 5),
                         ("pl2k", lambda: powerlaw_counts(2000, 800, 60000, dtype=dt), 16),
                         ("pl6k", lambda: powerlaw_counts(6000, 2500, 300000, dtype=dt), 50)):
        csr, csc = gen()
        dimA, dimB = csr[1].shape[0] - 1, csc[1].shape[0] - 1
        maxrow = int(np.diff(csr[1].astype(np.int64)).max()); maxcol = int(np.diff(csc[1].astype(np.int64)).max())
        for method, kw in cases:
            if only and method != only: continue
            kw = dict(kw)
            if kw["maxupd"] is None: kw["maxupd"] = 15 * k
            A0, B0 = init_factors(dimA, dimB, k, dtype=dt)
            Ar, Br = A0.copy(), B0.copy()
            t = time.time(); orc.run_poismf(Ar, Br, csr, csc, method, **kw); tc = time.time() - t
            line = f"{np.dtype(dt).name} {name}(maxrow {maxrow} maxcol {maxcol}) {method} w={kw.get('w_mult',1)} ls={kw.get('limit_step',0)} cpu {tc:.2f}s |"
            for mode, flags in (("strict", FLAG_STRICT), ("fast-direct", FLAG_NO_CACHED), ("fast", 0)):
                A1, B1 = A0.copy(), B0.copy()
                t = time.time()
                rc = c_funs._run_poismf(csr[0], csr[2], csr[1], csc[0], csc[2], csc[1], A1, B1, method=method,
                                        limit_step=kw.get("limit_step", False), l2_reg=kw["l2_reg"], l1_reg=kw.get("l1_reg", 0.),
                                        w_mult=kw.get("w_mult", 1.), step_size=kw.get("step_size", 1e-7), niter=kw["numiter"],
                                        maxupd=kw["maxupd"], early_stop=kw.get("early_stop", False), reuse_prev=kw.get("reuse_prev", False),
                                        flags=flags)
                tg = time.time() - t
                ea = stats(A1, Ar); eb = stats(B1, Br)
                same = np.array_equal(A1, Ar) and np.array_equal(B1, Br)
                line += f" {mode}: rc{rc} {'BITEXACT' if same else f'A max {ea[0]:.1e} med {ea[1]:.1e} | B max {eb[0]:.1e} med {eb[1]:.1e}'} nan={int(np.isnan(A1).any() or np.isnan(B1).any())} {tg:.2f}s;"
            print(line, flush=True)
