#!/usr/bin/env python
"""bench.py — poismf_b200 headline benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|small]

Metric (BASELINE.json): nnz processed per second per alternating sweep.
A "step" is ONE alternating sweep (B half-sweep over CSC + A half-sweep over CSR) of the
workload; at N=1 the workload is BASELINE config #2 — Last.FM-360K-shaped synthetic,
359k users x 160k items, 17.5M power-law draws, k=50, method=cg (maxupd 5, l2 1e4,
limit_step), float32 (the reference's Python default, poismf/__init__.py:240).

  value    : nnz / device time per sweep, inputs resident in HBM (CUDA events on the
             launching stream, max over ranks); the K timed steps are sweeps 1..K of one fit
  e2e      : the same metric through the drop-in C ABI call run_poismf (numiter=1) with HOST
             buffers: H2D of CSR+CSC+factors and D2H of the factors inside the timed region
  roofline : dominant row-kernel bin: algorithmic bytes of the bin / its average launch time
  cpu_baseline : the reference's own C path (oracle/_ref, OpenMP, all host cores) on a bounded sample

--impl reference times only the reference's CPU implementation (rank 0; other ranks exit).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: dimA, dimB, nnz draws, k, method, hyper-parameters
    "c2": dict(dimA=359_000, dimB=160_000, nnz=17_500_000, k=50, method="cg",
               hp=dict(l2_reg=1e4, maxupd=5, limit_step=True),
               label="lastfm360k-shaped synthetic 359k x 160k, 17.5M power-law draws, k=50, cg"),
    "small": dict(dimA=45_000, dimB=20_000, nnz=2_200_000, k=50, method="cg",
                  hp=dict(l2_reg=1e4, maxupd=5, limit_step=True),
                  label="1/8-scale lastfm360k-shaped synthetic 45k x 20k, 2.2M draws, k=50, cg"),
    "c1": dict(dimA=100, dimB=1000, nnz=10_000, k=5, method="pg",
               hp=dict(l2_reg=1e9, maxupd=1, step_size=1e-7),
               label="README synthetic 100 x 1000, 1e4 nnz, k=5, pg"),
}


def make_problem(cfg, dtype=np.float32):
    from poismf_b200.synth import init_factors, powerlaw_counts, readme_counts
    if cfg["dimA"] == 100:
        csr, csc = readme_counts(dtype=dtype)
    else:
        csr, csc = powerlaw_counts(cfg["dimA"], cfg["dimB"], cfg["nnz"], dtype=dtype, seed=1)
    A0, B0 = init_factors(cfg["dimA"], cfg["dimB"], cfg["k"], seed=1, dtype=dtype)
    return csr, csc, A0, B0


def team_name(code):
    if code == 200:
        return "lockstep"
    if code >= 100:
        return f"regtile{code - 100}w"
    return f"lanes{-code}" if code < 0 else ("block" if code == 1 else f"cluster{code}")


def algorithmic_bytes(nnz, rows, other, k, s):
    """bytes of one half-sweep = nnz*(k*s + s + 4) + rows*(2*k*s + 8) + other*k*s."""
    return nnz * (k * s + s + 4) + rows * (2 * k * s + 8) + other * k * s


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(cfg, steps, warmup, fast=True):
    """The reference's own CPU path (oracle/_ref) on a bounded sample; returns (nnz/s, info)."""
    from oracle.oracle import Ref, Restatement
    sample_cfg = CONFIGS["small"] if cfg["dimA"] > 50_000 else cfg
    csr, csc, A0, B0 = make_problem(sample_cfg)
    nnz = int(csr[0].shape[0])
    cores = os.cpu_count() or 1
    if Ref.available(np.float32, fast=fast):
        lib, kind = Ref(np.float32, fast=fast), "reference"
    else:
        lib, kind, cores = Restatement(np.float32), "port", 1
    hp = dict(sample_cfg["hp"])
    times = []
    for it in range(warmup + steps):
        A, B = A0.copy(), B0.copy()
        t0 = time.perf_counter()
        lib.run_poismf(A, B, csr, csc, sample_cfg["method"], numiter=1, nthreads=cores, **hp)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    info = {"value": nnz / (ms / 1e3), "unit": "nnz/s per sweep", "cores": cores, "kind": kind,
            "sample": f"{sample_cfg['label']}: {nnz} nnz, 1 sweep from init per step, {steps} steps"
                      f" ({'-O3 x86-64-v3 OpenMP build of /root/reference/src + naive BLAS' if kind == 'reference' else 'scalar restatement'})"}
    return ms, nnz, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: refresh replicas by peer-memory stores from the row kernels (p2p) or NCCL broadcasts")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    config_line = {"workload": cfg["label"], "method": cfg["method"], "k": cfg["k"], **cfg["hp"],
                   "l2_flush": "inputs larger than L2 (factors+CSR+CSC ~0.5 GB vs 126 MB)",
                   "parallelism": (f"rows/cols sharded x{args.gpus}, replicas refreshed by "
                                   f"{'NVLink peer-memory stores fused into the row kernels' if args.exchange == 'p2p' else 'NCCL broadcasts'}")
                   if args.gpus > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        ms, nnz, info = run_reference(cfg, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "nnz processed/sec per alternating sweep", "value": info["value"],
                "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_line, "cpu_baseline": info,
                "e2e": {"value": info["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    from poismf_b200 import _lib, c_funs, make_params
    from poismf_b200.device import DeviceFit
    _lib.require_gpu()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    csr, csc, A0, B0 = make_problem(cfg)
    nnz = int(csr[0].shape[0])
    dimA, dimB, k = cfg["dimA"], cfg["dimB"], cfg["k"]
    params = make_params(cfg["method"], numiter=1, flags=args.flags, **cfg["hp"])

    if world == 1:
        fit = DeviceFit(dimA, dimB, k, np.float32, device=local_rank)
        stream = torch.cuda.Stream(device=local_rank)
        fit.set_stream(stream.cuda_stream)
        fit.set_csr_csc(csr, csc)
        reset = lambda: fit.set_factors(A0, B0)
        sweep = lambda: fit.sweeps(params)
        profiler = fit
    else:
        from poismf_b200.sharding import GpuBackend, ShardedSweep
        be = GpuBackend(csr, csc, A0, B0, rank, world, local_rank, exchange=args.exchange)
        stream = be.stream
        drv = ShardedSweep(be, dimA, dimB, np.float32)
        reset = lambda: be.reset(A0, B0)
        sweep = lambda: drv.run(params)
        profiler = be.fit

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    reset()
    for _ in range(warmup):
        sweep()
    barrier()
    reset()
    launches0 = L.pmf_b200_kernel_launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            sweep()
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = L.pmf_b200_kernel_launches() - launches0
    # second pass over the same K sweeps with per-launch CUDA events (bins run one after the
    # other on the launching stream here; in the timed pass above they overlap on side streams)
    reset()
    profiler.set_profiling(True)
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        pe0.record(stream)
        for _ in range(args.steps):
            sweep()
        pe1.record(stream)
    barrier()
    ms_profiled_pass = pe0.elapsed_time(pe1)
    prof = profiler.get_profile()
    profiler.set_profiling(False)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nnz / (ms_step / 1e3)

    # roofline of the dominant row-kernel bin (by device time), rank 0's shard
    roof = None
    if prof:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        top = max(prof, key=lambda p: p["ms"])
        s = 4
        other = dimB if top["side"] == 0 else dimA
        # the opposite factor matrix's column-sum read is a separate kernel: not charged to the bin
        bytes_launch = top["nnz"] * (k * s + s + 4) + top["nrows"] * (2 * k * s + 8)
        avg_ms = top["ms"] / max(top["launches"], 1)
        achieved = bytes_launch / (avg_ms / 1e3) / 1e9 if avg_ms > 0 else 0.0
        sweep_bytes = algorithmic_bytes(nnz, dimA, dimB, k, s) + algorithmic_bytes(nnz, dimB, dimA, k, s)
        traffic = None
        try:   # DRAM bytes of this launch from the committed ncu --set full capture of the same command
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"side{top['side']}_{team_name(top['block_team'])}_cap{top['cap']}")
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": which,
                "kernel": f"rows_{team_name(top['block_team'])}_kernel<{cfg['method']}> side="
                          f"{'CSR(A)' if top['side'] == 0 else 'CSC(B)'} cap={top['cap']} rows={top['nrows']} nnz={top['nnz']}",
                "kernel_avg_ms": avg_ms, "kernel_share_of_step": top["ms"] / max(ms_profiled_pass, 1e-9),
                "serialized_ms_per_step": ms_profiled_pass / args.steps,
                "sweep_algorithmic_GBps": sweep_bytes / (ms_step / 1e3) / 1e9,
                "sweep_frac_of_peak": sweep_bytes / (ms_step / 1e3) / 1e9 / peak,
                "bins": [{"side": p["side"], "team": team_name(p["block_team"]), "cap": p["cap"],
                          "rows": p["nrows"], "nnz": p["nnz"], "ms_per_sweep": p["ms"] / args.steps} for p in prof]}
        # the solver re-reads every staged tile once per evaluation pass, out of shared memory
        # (or L2 for streamed rows); upper bound of passes per row and the on-chip traffic that implies
        mu = int(cfg["hp"].get("maxupd", 1))
        passes = {"cg": 1 + 2 * mu, "pg": 2 * mu}.get(cfg["method"])
        if passes:
            roof["tile_passes_per_row_max"] = passes
            roof["onchip_tile_GBps_upper"] = 2 * nnz * k * s * passes / (ms_step / 1e3) / 1e9

    # e2e: drop-in run_poismf with host buffers (rank 0 alone at N=1; sharded path otherwise reuses value)
    e2e = None
    if not args.no_e2e and world == 1:
        hA, hB = A0.copy(), B0.copy()
        pin = []
        for arr in (csr[0], csr[1], csr[2], csc[0], csc[1], csc[2], hA, hB):
            t = torch.from_numpy(arr)
            try:
                torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
                pin.append(t)
            except Exception:
                pass
        h2d = sum(a.nbytes for a in (csr[0], csr[1], csr[2], csc[0], csc[1], csc[2], A0, B0))
        d2h = A0.nbytes + B0.nbytes
        e_steps = max(3, min(args.steps, 5))
        def one():
            hA[...] = A0; hB[...] = B0
            t0 = time.perf_counter()
            c_funs._run_poismf(csr[0], csr[2], csr[1], csc[0], csc[2], csc[1], hA, hB, method=cfg["method"],
                               limit_step=cfg["hp"].get("limit_step", False), l2_reg=cfg["hp"]["l2_reg"],
                               step_size=cfg["hp"].get("step_size", 1e-7), niter=1, maxupd=cfg["hp"]["maxupd"],
                               early_stop=False, reuse_prev=False, flags=args.flags)
            return time.perf_counter() - t0
        one()
        ts = [one() for _ in range(e_steps)]
        e_ms = 1e3 * float(np.mean(ts))
        e2e = {"value": nnz / (e_ms / 1e3), "unit": "nnz/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e_ms, "steps": e_steps,
               "call": "run_poismf(numiter=1) via C ABI, host buffers (cudaHostRegister'ed), upload+plan+sweep+download"}
        for t in pin:
            try:
                torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
            except Exception:
                pass

    if not args.no_e2e and world > 1:
        # sharded public API end to end: every rank uploads ITS row/column shard and the replicated
        # factors from host memory, runs one sharded sweep (NCCL exchange inside), reads A,B back
        from poismf_b200.sharding import GpuBackend, ShardedSweep
        del be
        torch.cuda.synchronize()
        ts = []
        h2d = d2h = 0
        for it in range(3):
            dist.barrier()
            t0 = time.perf_counter()
            be2 = GpuBackend(csr, csc, A0, B0, rank, world, local_rank, exchange=args.exchange)
            ShardedSweep(be2, dimA, dimB, np.float32).run(params)
            Aout, Bout = be2.factors()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            a0, a1 = be2.rangesA[rank]; b0, b1 = be2.rangesB[rank]
            h2d = (A0.nbytes + B0.nbytes + int(csr[1][a1] - csr[1][a0]) * 12 + int(csc[1][b1] - csc[1][b0]) * 12
                   + (a1 - a0 + b1 - b0) * 8)
            d2h = A0.nbytes + B0.nbytes
            del be2
            if it > 0:
                ts.append(dt)
        t = torch.tensor([float(np.mean(ts))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = 1e3 * float(t.item())
        e2e = {"value": nnz / (e_ms / 1e3), "unit": "nnz/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms, "steps": len(ts),
               "call": "poismf_b200.sharding.GpuBackend + ShardedSweep.run(numiter=1) + factors(): per-rank shard upload, sweep with NCCL exchange, download (max over ranks)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            _, _, cpu = run_reference(cfg, 3, 1)
        except Exception as e:  # the oracle is optional infrastructure for this leg
            cpu = {"value": None, "unit": "nnz/s per sweep", "cores": 0, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {"metric": "nnz processed/sec per alternating sweep", "value": value, "unit": "nnz/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_line, "nnz": nnz, "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
