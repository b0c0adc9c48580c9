// poismf_b200 — shared device-side definitions for the row-update kernels (sm_100a).
//
// Vocabulary (follows the reference's domain, /root/reference/src/poismf.c):
//   half-sweep : update of every row of one factor matrix with the other held fixed
//   row        : one user (CSR side) or one item (CSC side); its "tile" is the set of
//                gathered rows of the FIXED factor matrix, one per non-zero
//   team       : the group of threads that cooperates on one row (8/16/32 lanes of a warp, a
//                CTA, or a thread-block cluster — cluster_team.cuh)
//
// Numerics modes (template parameter STRICT):
//   STRICT = true  : mimics the reference built with sequential BLAS and no FMA
//                    contraction (oracle/_ref strict build): every reduction is
//                    summed left-to-right by one thread, a*b+c is two roundings.
//                    Used by the parity tests.
//   STRICT = false : FMA + tree reductions; same algorithm, different rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>

#define PMF_DEVINL __device__ __forceinline__

namespace pmf {

enum Method : int { M_TNCG = 1, M_CG = 2, M_PG = 3 };  // == reference enum, src/poismf.h:225

// ---------------------------------------------------------------------------
// Arithmetic policy
// ---------------------------------------------------------------------------
template <class real> struct RealTraits;
template <> struct RealTraits<float> {
    static constexpr int V = 4;  // elements per 16-byte vector
    PMF_DEVINL static float eps() { return FLT_EPSILON; }
    PMF_DEVINL static float huge() { return __int_as_float(0x7f800000); }
};
template <> struct RealTraits<double> {
    static constexpr int V = 2;
    PMF_DEVINL static double eps() { return DBL_EPSILON; }
    PMF_DEVINL static double huge() { return __longlong_as_double(0x7ff0000000000000LL); }
};

PMF_DEVINL float mul_rn(float a, float b) { return __fmul_rn(a, b); }
PMF_DEVINL double mul_rn(double a, double b) { return __dmul_rn(a, b); }
PMF_DEVINL float add_rn(float a, float b) { return __fadd_rn(a, b); }
PMF_DEVINL double add_rn(double a, double b) { return __dadd_rn(a, b); }
PMF_DEVINL float sub_rn(float a, float b) { return __fsub_rn(a, b); }
PMF_DEVINL double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// a*b + c : one rounding (fast) or two (strict, like gcc -ffp-contract=off)
template <bool STRICT, class real> PMF_DEVINL real mad(real a, real b, real c)
{
    if (STRICT) return add_rn(mul_rn(a, b), c);
    return fma(a, b, c);
}
// products/sums that must never be contracted by nvcc when STRICT
template <bool STRICT, class real> PMF_DEVINL real mul(real a, real b)
{
    if (STRICT) return mul_rn(a, b);
    return a * b;
}
template <bool STRICT, class real> PMF_DEVINL real add(real a, real b)
{
    if (STRICT) return add_rn(a, b);
    return a + b;
}
template <bool STRICT, class real> PMF_DEVINL real sub(real a, real b)
{
    if (STRICT) return sub_rn(a, b);
    return a - b;
}

// The reference evaluates log() in double even in the float build (no tgmath;
// src/poismf.c:204,262) and accumulates `lsum += x * log(p)` with the product
// kept in double:  lsum = (real)((double)lsum + (double)x * log((double)p)).
// STRICT reproduces this; fast mode uses the native-width log.
template <bool STRICT> PMF_DEVINL float xlogp_acc(float acc, float x, float p)
{
    if (STRICT) return (float)__dadd_rn((double)acc, __dmul_rn((double)x, log((double)p)));
    return fmaf(x, logf(p), acc);
}
template <bool STRICT> PMF_DEVINL double xlogp_acc(double acc, double x, double p)
{
    if (STRICT) return __dadd_rn(acc, __dmul_rn(x, log(p)));
    return fma(x, log(p), acc);
}
// The per-nnz term alone (for two-level summation in fast mode).  Float uses the hardware
// log2 (MUFU.LG2, |abs err| < 4e-7 on [0.5,2], <= 2 ulp elsewhere): the term feeds only the
// line-search comparisons of the fast path, whose float noise floor is far above that.
PMF_DEVINL float xlogp(float x, float p) { return x * __logf(p); }
PMF_DEVINL double xlogp(double x, double p) { return x * log(p); }

PMF_DEVINL bool is_bad(float v) { return isnan(v) || isinf(v); }
PMF_DEVINL bool is_bad(double v) { return isnan(v) || isinf(v); }

// ---------------------------------------------------------------------------
// Teams
// ---------------------------------------------------------------------------
// A team is the group of threads that cooperates on one row.  It exposes
//   rank()/size()/sync()      : split of the per-non-zero loops and of the "owner" k-loops
//   sum/min/max/bcast0        : team-wide reductions (block-level ones cost two barriers)
//   kbegin()/kstride()/ksum() : k-vector reductions WITHIN one warp (or sub-warp): whichever warp
//                               walks the shared k-vector reduces it with shuffles, no barrier
//                               (k <= 256: at most 8 elements per lane).  CTA teams run their
//                               k-scalar phases on warp 0 (k_leader) and broadcast (kslots).
//   nnz_sum*/nnz_vec_sum      : sums over the row's non-zeros (span the cluster for gangs)
// All reductions return the SAME bits to every member: the solver's control flow is
// executed redundantly by all members and must not diverge.
template <class T> PMF_DEVINL T warp_sum(T v, unsigned mask = 0xffffffffu, int width = 32)
{
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <class T> PMF_DEVINL T warp_min(T v, unsigned mask = 0xffffffffu, int width = 32)
{
    for (int o = width >> 1; o > 0; o >>= 1) { T w = __shfl_xor_sync(mask, v, o); v = w < v ? w : v; }
    return v;
}
template <class T> PMF_DEVINL T warp_max(T v, unsigned mask = 0xffffffffu, int width = 32)
{
    for (int o = width >> 1; o > 0; o >>= 1) { T w = __shfl_xor_sync(mask, v, o); v = w > v ? w : v; }
    return v;
}

// W lanes of a warp (W = 8, 16 or 32) per row: 32/W short rows advance side by side in one
// warp.  All synchronisation uses the sub-warp's own lane mask, so sub-warps of one warp may
// follow different control flow.
template <int W> struct SubWarpTeam {
    int lane;        // rank within the team
    unsigned mask;   // lanes of this team within the warp
    PMF_DEVINL explicit SubWarpTeam(void* /*scratch*/)
    {
        const int wl = threadIdx.x & 31;
        lane = wl & (W - 1);
        mask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (wl & ~(W - 1)));
    }
    PMF_DEVINL int rank() const { return lane; }
    PMF_DEVINL int size() const { return W; }
    PMF_DEVINL void sync() const { __syncwarp(mask); }
    template <class T> PMF_DEVINL T sum(T v) const { return warp_sum(v, mask, W); }
    template <class T> PMF_DEVINL T min(T v) const { return warp_min(v, mask, W); }
    template <class T> PMF_DEVINL T max(T v) const { return warp_max(v, mask, W); }
    template <class T> PMF_DEVINL T bcast0(T v) const { return __shfl_sync(mask, v, 0, W); }
    PMF_DEVINL int kbegin() const { return lane; }
    PMF_DEVINL int kstride() const { return W; }
    template <class T> PMF_DEVINL T ksum(T v) const { return warp_sum(v, mask, W); }
    template <class T> PMF_DEVINL T kmin(T v) const { return warp_min(v, mask, W); }
    template <class T> PMF_DEVINL T kmax(T v) const { return warp_max(v, mask, W); }
    // the k-scalar phase of the solvers is executed by the "k leader" (here: the whole team)
    static constexpr bool k_bcast = false;
    PMF_DEVINL bool k_leader() const { return true; }
    template <class T> PMF_DEVINL T* kslots() const { return nullptr; }
    static constexpr bool is_gang = false;
    PMF_DEVINL bool owns_row() const { return true; }
    PMF_DEVINL unsigned crank_() const { return 0; }
    PMF_DEVINL unsigned csize_() const { return 1; }
    template <class T> PMF_DEVINL T nnz_sum(T v) const { return sum(v); }
    template <class T, int N> PMF_DEVINL void nnz_sum_n(T (&v)[N]) const
    {
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = sum(v[j]);
    }
    template <class real> PMF_DEVINL void nnz_vec_sum(real*, int) const {}
};
using WarpTeam = SubWarpTeam<32>;

// CTA-level reductions shared by BlockTeam and ClusterTeam.  `red` = 34 doubles of shared
// memory (64 slots used by sum_n as floats/doubles... sized below).  Second stage by shuffles.
struct BlockOps {
    double* red;
    PMF_DEVINL int rank() const { return threadIdx.x; }
    PMF_DEVINL int size() const { return blockDim.x; }
    PMF_DEVINL void sync() const { __syncthreads(); }
    PMF_DEVINL int kbegin() const { return threadIdx.x & 31; }
    PMF_DEVINL int kstride() const { return 32; }
    template <class T> PMF_DEVINL T ksum(T v) const { return warp_sum(v); }
    template <class T> PMF_DEVINL T kmin(T v) const { return warp_min(v); }
    template <class T> PMF_DEVINL T kmax(T v) const { return warp_max(v); }
    // the k-scalar phase of the solvers runs on warp 0 only; its results are broadcast through
    // 12 shared slots (bytes 544..640 of the team scratch) and one barrier
    static constexpr bool k_bcast = true;
    PMF_DEVINL bool k_leader() const { return threadIdx.x < 32; }
    template <class T> PMF_DEVINL T* kslots() const { return reinterpret_cast<T*>(red + 68); }
    // up to 4 sums at once with one pair of barriers (blockDim <= 512: at most 16 warps)
    template <class T, int N> PMF_DEVINL void sum_n(T (&v)[N]) const
    {
        static_assert(N <= 4, "at most 4 values per block reduction");
        T* r = (T*)red;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = warp_sum(v[j]);
        __syncthreads();   // protect `red` against the previous reduction's readers
        if (lane == 0)
#pragma unroll
            for (int j = 0; j < N; j++) r[j * 16 + w] = v[j];
        __syncthreads();
        // slot j*16 + w; lane l folds slots l (values 0,1) and l+32 (values 2,3) in 16-lane groups
        T p0 = ((lane & 15) < nw && (lane >> 4) < N) ? r[lane] : (T)0;
        T p1 = (N > 2 && (lane & 15) < nw && 2 + (lane >> 4) < N) ? r[lane + 32] : (T)0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            if (N > 2) p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        }
        v[0] = __shfl_sync(0xffffffffu, p0, 0);
        if (N > 1) v[1] = __shfl_sync(0xffffffffu, p0, 16);
        if (N > 2) v[2] = __shfl_sync(0xffffffffu, p1, 0);
        if (N > 3) v[3] = __shfl_sync(0xffffffffu, p1, 16);
    }
    template <class T> PMF_DEVINL T sum(T v) const
    {
        T a[1] = {v};
        sum_n(a);
        return a[0];
    }
    template <class T, class OP> PMF_DEVINL T reduce_idem(T v, OP op) const   // idempotent ops only
    {
        T* r = (T*)red;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) r[w] = v;
        __syncthreads();
        T part = r[(lane & 15) < nw ? (lane & 15) : 0];
        part = op(part, __shfl_xor_sync(0xffffffffu, part, 8));
        part = op(part, __shfl_xor_sync(0xffffffffu, part, 4));
        part = op(part, __shfl_xor_sync(0xffffffffu, part, 2));
        part = op(part, __shfl_xor_sync(0xffffffffu, part, 1));
        return part;
    }
    template <class T> PMF_DEVINL T min(T v) const { return reduce_idem(v, [](T a, T b) { return b < a ? b : a; }); }
    template <class T> PMF_DEVINL T max(T v) const { return reduce_idem(v, [](T a, T b) { return b > a ? b : a; }); }
    template <class T> PMF_DEVINL T bcast0(T v) const
    {
        T* r = (T*)red;
        __syncthreads();
        if (threadIdx.x == 0) r[66] = v;
        __syncthreads();
        return r[66];
    }
};

// One CTA per row.  `scratch` points at >= 34 doubles of shared memory.
struct BlockTeam : BlockOps {
    PMF_DEVINL explicit BlockTeam(void* scratch) { red = (double*)scratch; }
    static constexpr bool is_gang = false;
    PMF_DEVINL bool owns_row() const { return true; }
    PMF_DEVINL unsigned crank_() const { return 0; }
    PMF_DEVINL unsigned csize_() const { return 1; }
    template <class T> PMF_DEVINL T nnz_sum(T v) const { return sum(v); }
    template <class T, int N> PMF_DEVINL void nnz_sum_n(T (&v)[N]) const { sum_n(v); }
    template <class real> PMF_DEVINL void nnz_vec_sum(real*, int) const {}
};

// ---------------------------------------------------------------------------
// Small k-vector helpers on shared-memory vectors (length k, team-strided)
// ---------------------------------------------------------------------------
// Sequential (reference-order) dot when STRICT: member 0 sums, everyone gets it.
template <bool STRICT, class real, class Team>
PMF_DEVINL real vdot(const Team& tm, const real* x, const real* y, int k)
{
    if (STRICT) {
        real s = 0;
        if (tm.rank() == 0)
            for (int i = 0; i < k; i++) s = add_rn(s, mul_rn(x[i], y[i]));
        return tm.bcast0(s);
    }
    real s = 0;
    for (int i = tm.kbegin(); i < k; i += tm.kstride()) s = fma(x[i], y[i], s);
    s = tm.ksum(s);
    // CTA / cluster teams: every warp computes the sum redundantly and at its own pace; nobody may
    // overwrite x or y before the slowest warp has read them (a warp that saw a half-updated vector
    // would leave the replicated control flow: found by racecheck in tn_linesearch, r1)
    if (Team::k_bcast) tm.sync();
    return s;
}
// sqrt(sum x^2) — the shim's nrm2 (oracle/blas_shim.c)
template <bool STRICT, class real, class Team>
PMF_DEVINL real vnrm2(const Team& tm, const real* x, int k)
{
    return sqrt(vdot<STRICT>(tm, x, x, k));
}
// y += a*x  (each element touched by exactly one member: order-free)
template <bool STRICT, class real, class Team>
PMF_DEVINL void vaxpy(const Team& tm, real a, const real* x, real* y, int k)
{
    for (int i = tm.rank(); i < k; i += tm.size()) y[i] = mad<STRICT>(a, x[i], y[i]);
}
template <class real, class Team>
PMF_DEVINL void vcopy(const Team& tm, const real* x, real* y, int k)
{
    for (int i = tm.rank(); i < k; i += tm.size()) y[i] = x[i];
}
template <class real, class Team>
PMF_DEVINL void vfill(const Team& tm, real* y, real v, int k)
{
    for (int i = tm.rank(); i < k; i += tm.size()) y[i] = v;
}

}  // namespace pmf
