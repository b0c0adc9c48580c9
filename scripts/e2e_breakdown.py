"""Phase times of the drop-in run_poismf on the headline workload (POISMF_B200_TIMING),
with page-locked and with pageable host buffers.  Usage: python scripts/e2e_breakdown.py [config]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from poismf_b200 import c_funs  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
csr, csc, A0, B0 = bench.make_problem(cfg)
hA, hB = A0.copy(), B0.copy()


def one(timing):
    hA[...] = A0; hB[...] = B0
    if timing:
        os.environ["POISMF_B200_TIMING"] = "1"
    else:
        os.environ.pop("POISMF_B200_TIMING", None)
    t0 = time.perf_counter()
    c_funs._run_poismf(csr[0], csr[2], csr[1], csc[0], csc[2], csc[1], hA, hB, method=cfg["method"],
                       limit_step=cfg["hp"].get("limit_step", False), l2_reg=cfg["hp"]["l2_reg"],
                       step_size=cfg["hp"].get("step_size", 1e-7), niter=1, maxupd=cfg["hp"]["maxupd"],
                       early_stop=False, reuse_prev=False)
    return 1e3 * (time.perf_counter() - t0)


for mode in ("pageable", "pinned"):
    pin = []
    if mode == "pinned":
        for arr in (csr[0], csr[1], csr[2], csc[0], csc[1], csc[2], hA, hB):
            t = torch.from_numpy(arr)
            torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
            pin.append(t)
    one(False)
    print(f"== {mode}: untimed-phase calls (ms):", " ".join(f"{one(False):.1f}" for _ in range(10)), flush=True)
    sys.stderr.flush()
    print(f"== {mode}: with phase timing: total {one(True):.1f} ms", flush=True)
    ref = (hA.copy(), hB.copy())
    for t in pin:
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
print("checksum", float(ref[0].sum()), float(ref[1].sum()))
