"""CPU tests of the host-side logic: the C ABI library loads and exports every symbol the
header declares (no compute without a GPU), loud failure without a device, the synthetic
generators, and the sharding layer (world_size 2 over gloo with a CPU stand-in solver)."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest

from conftest import ROOT, has_gpu, hyper, problem


def test_library_exports_every_declared_symbol():
    from poismf_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "poismf_b200.h")).read()
    declared = set(re.findall(r"\b(pmf_b200_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(L, sym), f"{sym} declared in include/poismf_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


@pytest.mark.parametrize("variant", ["double", "float", "double_int"])
def test_host_dropin_exports_reference_prototypes(variant):
    path = os.path.join(ROOT, "poismf_b200", f"libpoismf_host_{variant}.so")
    L = ctypes.CDLL(path)
    for sym in ("run_poismf", "factors_multiple", "factors_single", "predict_multiple", "topN", "get_has_openmp"):
        assert hasattr(L, sym)


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from poismf_b200 import c_funs
    csr, csc, A0, B0, k = problem("readme", np.float64)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        c_funs._run_poismf(csr[0], csr[2], csr[1], csc[0], csc[2], csc[1], A0, B0, method="pg")
    with pytest.raises(RuntimeError):
        c_funs._predict_multiple(np.zeros(1), A0, B0, np.zeros(1, np.uint64), np.zeros(1, np.uint64))
    # the raw C ABI refuses too (return code 1 = the reference's failure code)
    L = ctypes.CDLL(os.path.join(ROOT, "poismf_b200", "libpoismf_host_double.so"))
    L.run_poismf.restype = ctypes.c_int
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    A, B = A0.copy(), B0.copy()
    sz, d = ctypes.c_size_t, ctypes.c_double
    L.run_poismf.argtypes = [ctypes.c_void_p] * 8 + [sz, sz, sz, d, d, d, d, ctypes.c_int, ctypes.c_bool, sz, sz,
                                                      ctypes.c_bool, ctypes.c_bool, ctypes.c_bool, ctypes.c_int]
    rc = L.run_poismf(p(A), p(csr[0]), p(csr[1]), p(csr[2]), p(B), p(csc[0]), p(csc[1]), p(csc[2]),
                      100, 1000, 5, 1e9, 0., 1., 1e-7, 3, False, 1, 1, False, False, True, 1)
    assert rc == 1 and np.array_equal(A, A0)
    # the entry points added around the path refuse as well
    rows = np.repeat(np.arange(100, dtype=np.uint64), np.diff(csr[1].astype(np.int64)))
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        c_funs._fit_coo(rows, csr[2], csr[0], A0.copy(), B0.copy(), method="pg")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        c_funs._coo_to_csr_csc(rows, csr[2], csr[0], 100, 1000)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        c_funs._predict_factors(csr[0][:3], csr[2][:3], B0, B0.sum(0), A0.mean(0))
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        c_funs._topN_batch(A0, B0, top_n=5)
    from poismf_b200 import _lib
    Lb = _lib.lib()
    out = np.zeros(5)
    assert Lb.pmf_b200_factors_single(1, 8, p(out), 5, p(A0[0].copy()), 1, p(csr[0]), p(csr[2]), 3, p(B0),
                                      p(B0.sum(0)), 10, 1e3, 0., 0., 1., 0) == 1
    assert Lb.pmf_b200_fit_coo(1, 8, p(A), p(B), p(rows), p(csr[2]), p(csr[0]), rows.shape[0], 100, 1000, 5,
                               1e9, 0., 1., 1e-7, 3, 0, 1, 1, 0, 0, 0) == 1
    assert np.array_equal(A, A0) and not out.any()


def test_argument_validation_mirrors_wrapper():
    from poismf_b200 import c_funs
    e = np.empty(0)
    with pytest.raises(ValueError, match="no non-zero"):
        c_funs._run_poismf(e, e, e, e, e, e, np.ones((2, 2)), np.ones((2, 2)))


def test_synth_csr_csc_consistent():
    from poismf_b200.synth import powerlaw_counts
    csr, csc = powerlaw_counts(500, 300, 8000, dtype=np.float32)
    for vals, ptr, ind in (csr, csc):
        ptr = ptr.astype(np.int64)
        assert ptr[0] == 0 and ptr[-1] == vals.shape[0] and (np.diff(ptr) >= 0).all()
        for r in range(ptr.shape[0] - 1):
            seg = ind[ptr[r]:ptr[r + 1]].astype(np.int64)
            assert (np.diff(seg) > 0).all()          # sorted, no duplicates
        assert (vals >= 1).all()
    dense = np.zeros((500, 300))
    rows = np.repeat(np.arange(500), np.diff(csr[1].astype(np.int64)))
    dense[rows, csr[2].astype(np.int64)] = csr[0]
    cols = np.repeat(np.arange(300), np.diff(csc[1].astype(np.int64)))
    dense2 = np.zeros((500, 300))
    dense2[csc[2].astype(np.int64), cols] = csc[0]
    assert np.array_equal(dense, dense2)


def test_nnz_balanced_ranges():
    from poismf_b200.sharding import nnz_balanced_ranges, slice_compressed
    csr, csc, A0, B0, k = problem("pl2k", np.float64)
    for parts in (1, 2, 3, 8):
        rg = nnz_balanced_ranges(csr[1], parts)
        assert rg[0][0] == 0 and rg[-1][1] == A0.shape[0]
        assert all(rg[i][1] == rg[i + 1][0] for i in range(parts - 1))
        ptr = csr[1].astype(np.int64)
        rg_nnz = nnz_balanced_ranges(csr[1], parts, cost_aware=False)
        loads = [ptr[e] - ptr[b] for b, e in rg_nnz]
        assert max(loads) <= ptr[-1] / parts + np.diff(ptr).max()
        from poismf_b200.sharding import row_cost
        cost = row_cost(np.diff(ptr))
        cl = [cost[b:e].sum() for b, e in rg]
        assert max(cl) <= cost.sum() / parts + cost.max() + 1e-9
    v, p, i = slice_compressed(csr, 10, 20)
    assert p[0] == 0 and p[-1] == v.shape[0] == i.shape[0]
    assert nnz_balanced_ranges(np.array([0, 5]), 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]   # more parts than rows: empty tails


def test_devpool_size_classes(tmp_path):
    """The device block cache hands out size classes: never smaller than the request, at most 12.5 %
    (and 256 MB) larger, and requests of one class map to one block size (host-only check of
    poismf_b200/csrc/devpool.h; no CUDA call is made)."""
    import shutil
    import subprocess
    cxx = shutil.which("g++", path="/usr/bin:" + os.environ.get("PATH", "")) or shutil.which("c++")
    cuda_inc = "/usr/local/cuda/include"
    if cxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("needs a host C++ compiler and the CUDA headers")
    src = tmp_path / "t.cpp"
    src.write_text('#include "devpool.h"\n#include <cstdio>\n#include <cstdlib>\n'
                   'int main(int c, char** v) { for (int i = 1; i < c; i++) '
                   'printf("%zu\\n", pmf::DevPool::size_class((size_t)strtoull(v[i], 0, 10))); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.run([cxx, "-std=c++17", "-I", cuda_inc, "-I", os.path.join(ROOT, "poismf_b200", "csrc"),
                    str(src), "-o", str(exe)], check=True)
    req = [1, 511, 512, 513, 5000, 2**20, 2**20 + 1, 72_000_000, 3 * 2**30, 48 * 2**30 + 12345]
    got = [int(x) for x in subprocess.run([str(exe)] + [str(r) for r in req], check=True,
                                          capture_output=True, text=True).stdout.split()]
    for r, g in zip(req, got):
        assert g >= max(r, 512) and g - r <= max(512, r // 8, 0) and g - r <= 256 * 2**20, (r, g)
    assert got[0] == got[1] == got[2] == 512 and got[3] == 1024 and got[5] == 2**20


def test_user_ranges_and_row_cost():
    from poismf_b200.sharding import row_cost, user_ranges
    for n, parts in ((10, 3), (7, 8), (0, 2), (1000, 8)):
        rg = user_ranges(n, parts)
        assert rg[0][0] == 0 and rg[-1][1] == n and all(a[1] == b[0] for a, b in zip(rg, rg[1:]))
        sizes = [hi - lo for lo, hi in rg]
        assert max(sizes) - min(sizes) <= 1
    c = row_cost(np.array([0, 1, 1000, 1001, 16000, 16001]))
    assert c[0] == 0 and np.all(np.diff(c) > 0)          # empty rows cost nothing; cost grows with length


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, case, q):
    import torch.distributed as dist
    import torch
    from oracle.oracle import Restatement
    from poismf_b200 import _lib, make_params
    from poismf_b200.sharding import ShardedSweep, nnz_balanced_ranges
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dt = np.float64
    csr, csc, A0, B0, k = problem("pl2k", dt)
    method, kw = hyper(case, k)
    orc = Restatement(dt)
    A, B = A0.copy(), B0.copy()
    rA, rB = nnz_balanced_ranges(csr[1], world), nnz_balanced_ranges(csc[1], world)

    class CpuStandIn:
        """half-sweep = the oracle restricted to this rank's rows (test infrastructure)."""
        def half_sweep(self, side, params, step, cdiv):
            lo, hi = (rA if side == _lib.SIDE_CSR else rB)[rank]
            # one-sided update of the local rows: emulate with a masked copy of the matrix
            mat = csr if side == _lib.SIDE_CSR else csc
            vals, ptr, ind = mat
            ptr2 = ptr.copy().astype(np.int64)
            keep = np.zeros(vals.shape[0], bool); keep[ptr2[lo]:ptr2[hi]] = True
            M, F = (A, B) if side == _lib.SIDE_CSR else (B, A)
            before = M.copy()
            self._one_side(side, M, F, params, step)
            M[:lo] = before[:lo]; M[hi:] = before[hi:]
            # tncg early stop (src/poismf.c:393-396): LOCAL rows with non-zeros that moved less than 1e-4
            nz = np.diff(ptr2)[lo:hi] > 0
            d = before[lo:hi] - M[lo:hi]
            moved = np.zeros(hi - lo)
            for c in range(d.shape[1]):                       # the reference's left-to-right dot
                moved = moved + d[:, c] * d[:, c]
            return int(((moved <= 1e-4) & nz).sum())
        def allreduce_int(self, v):
            t = torch.tensor([int(v)], dtype=torch.int64)
            dist.all_reduce(t)
            return int(t.item())
        def _one_side(self, side, M, F, params, step):
            import ctypes as C
            r = C.c_double
            lib = orc.lib
            fn = lib.oracle_half_sweep
            fn.restype = C.c_int
            mat = csr if side == _lib.SIDE_CSR else csc
            fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                           C.c_size_t, C.c_size_t, r, r, r, r, C.c_int, C.c_size_t, C.c_int, C.c_int]
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            fn(params.method, p(M), p(F), p(mat[0]), p(mat[1]), p(mat[2]), M.shape[0], F.shape[0], M.shape[1],
               params.l2_reg, params.l1_reg, params.w_mult, step, int(side == _lib.SIDE_CSR), params.maxupd,
               params.limit_step, params.reuse_prev)
        def exchange(self, side):
            M, ranges = (A, rA) if side == _lib.SIDE_CSR else (B, rB)
            t = torch.from_numpy(M)
            for r_, (lo, hi) in enumerate(ranges):
                if hi > lo:
                    dist.broadcast(t[lo:hi], src=r_)

    params = make_params(method, **kw)
    ShardedSweep(CpuStandIn(), A.shape[0], B.shape[0], dt).run(params)
    Af, Bf = A0.copy(), B0.copy()
    orc.run_poismf(Af, Bf, csr, csc, method, **kw)
    q.put((rank, bool(np.array_equal(A, Af) and np.array_equal(B, Bf))))
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["pg", "cg", "tncg", "tncg_reuse_stop"])
def test_sharded_sweep_gloo_world2(case):
    """Two CPU ranks, each updating only its own nnz-balanced row/column range and exchanging
    slices, reproduce the unsharded oracle bit for bit — including tncg's early stop, whose count of
    unchanged rows is summed over the ranks before it is compared with the dimension
    (src/poismf.c:393-403, :606): with local counts the ranks would disagree on when to stop."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _gloo_topn_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle.oracle import Restatement
    from poismf_b200.sharding import topn_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    n_users, n_items, k, n_top = 37, 500, 6, 9
    A = rng.gamma(1, 1, size=(n_users, k)); B = rng.gamma(1, 1, size=(n_items, k))
    users = rng.permutation(n_users)[:30].astype(np.uint64)
    lens = rng.integers(0, 20, users.shape[0])
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    eix = np.concatenate([np.sort(rng.choice(n_items, int(m), replace=False)) for m in lens]).astype(np.uint64)
    orc = Restatement(np.float64)

    def cpu_scorer(A_, B_, u_, p_, i_, n_, s_):            # test infrastructure: the oracle, user by user
        ids = np.empty((u_.shape[0], n_), np.uint64); sc = np.empty((u_.shape[0], n_))
        for j, usr in enumerate(u_):
            ex = i_[int(p_[j]):int(p_[j + 1])]
            _, ids[j], sc[j] = orc.topN(np.ascontiguousarray(A_[int(usr)]), B_, n_, exclude=ex if ex.size else None)
        return ids, sc

    ids, sc = topn_sharded(A, B, n_top, users=users, excl_ptr=ptr, excl_ix=eix, output_score=True, rank=rank,
                           world=world, scorer=cpu_scorer)
    ids1, sc1 = topn_sharded(A, B, n_top, users=users, excl_ptr=ptr, excl_ix=eix, output_score=True, scorer=cpu_scorer)
    # gather=False: a rank keeps the lists of its own share of the users
    from poismf_b200.sharding import user_ranges
    ids_l, sc_l = topn_sharded(A, B, n_top, users=users, excl_ptr=ptr, excl_ix=eix, output_score=True, rank=rank,
                               world=world, scorer=cpu_scorer, gather=False)
    lo, hi = user_ranges(users.shape[0], world)[rank]
    local_ok = np.array_equal(ids_l, ids1[lo:hi]) and np.array_equal(sc_l, sc1[lo:hi])
    q.put((rank, bool(np.array_equal(ids, ids1) and np.array_equal(sc, sc1) and ids.shape == (30, n_top) and local_ok)))
    dist.destroy_process_group()


def test_topn_user_sharding_gloo_world2():
    """topN shards by users with B replicated: two ranks' gathered lists == one rank's."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_topn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res



def test_shard_ranges_are_remembered_per_offsets_array():
    """GpuBackend.load computes the cost-balanced shard boundaries once per offsets array (address, length, sampled
    content); another array, or the same buffer with other content, gets its own."""
    from poismf_b200.sharding import GpuBackend, nnz_balanced_ranges, row_cost
    be = GpuBackend.__new__(GpuBackend)          # host logic only: no device
    be.world = 3
    rng = np.random.default_rng(0)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(0, 50, 5000))]).astype(np.uint64)
    r1 = be._ranges(ptr)
    assert r1 == nnz_balanced_ranges(ptr, 3) and be._ranges(ptr) is r1
    ptr2 = ptr.copy()
    assert be._ranges(ptr2) == r1 and be._ranges(ptr2) is not r1
    ptr[1:] += np.arange(1, ptr.shape[0], dtype=np.uint64) * 40          # same buffer, other content
    r3 = be._ranges(ptr)
    assert r3 == nnz_balanced_ranges(ptr, 3) and r3 != r1
    # the tunable cost model reproduces the default one
    n = np.array([0, 5, 600, 2000, 20000])
    base = row_cost(n)
    os.environ["POISMF_B200_ROW_COST"] = "10,1000:1,16000:1"
    try:
        assert np.array_equal(row_cost(n), base)
    finally:
        del os.environ["POISMF_B200_ROW_COST"]
