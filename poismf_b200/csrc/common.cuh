// poismf_b200 — shared device-side definitions for the row-update kernels (sm_100a).
//
// Vocabulary (follows the reference's domain, /root/reference/src/poismf.c):
//   half-sweep : update of every row of one factor matrix with the other held fixed
//   row        : one user (CSR side) or one item (CSC side); its "tile" is the set of
//                gathered rows of the FIXED factor matrix, one per non-zero
//   team       : the group of threads that cooperates on one row (a warp, or a CTA)
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
// The per-nnz term alone (for two-level summation in fast mode)
PMF_DEVINL float xlogp(float x, float p) { return x * logf(p); }
PMF_DEVINL double xlogp(double x, double p) { return x * log(p); }

PMF_DEVINL bool is_bad(float v) { return isnan(v) || isinf(v); }
PMF_DEVINL bool is_bad(double v) { return isnan(v) || isinf(v); }

// ---------------------------------------------------------------------------
// Teams
// ---------------------------------------------------------------------------
// A team exposes: rank(), size(), sync(), bcast-capable reductions.  All
// reductions return the SAME bits to every member (solver control flow is
// executed redundantly by all members and must not diverge).
struct WarpTeam {
    int lane;
    PMF_DEVINL explicit WarpTeam(void* /*scratch*/) : lane(threadIdx.x & 31) {}
    PMF_DEVINL int rank() const { return lane; }
    PMF_DEVINL int size() const { return 32; }
    PMF_DEVINL void sync() const { __syncwarp(); }
    template <class T> PMF_DEVINL T sum(T v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    template <class T> PMF_DEVINL T min(T v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
        return v;
    }
    template <class T> PMF_DEVINL T max(T v) const
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
        return v;
    }
    template <class T> PMF_DEVINL T bcast0(T v) const { return __shfl_sync(0xffffffffu, v, 0); }
};

// One CTA per row.  `scratch` points at >= 34 doubles of shared memory.
struct BlockTeam {
    double* red;
    PMF_DEVINL explicit BlockTeam(void* scratch) : red((double*)scratch) {}
    PMF_DEVINL int rank() const { return threadIdx.x; }
    PMF_DEVINL int size() const { return blockDim.x; }
    PMF_DEVINL void sync() const { __syncthreads(); }
    template <class T, class OP> PMF_DEVINL T reduce(T v, OP op) const
    {
        T* r = (T*)red;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();  // protect `red` against the previous reduction's readers
        if (lane == 0) r[w] = v;
        __syncthreads();
        T acc = r[0];
        for (int i = 1; i < nw; i++) acc = op(acc, r[i]);  // same order on every thread
        return acc;
    }
    template <class T> PMF_DEVINL T sum(T v) const { return reduce(v, [](T a, T b) { return a + b; }); }
    template <class T> PMF_DEVINL T min(T v) const { return reduce(v, [](T a, T b) { return b < a ? b : a; }); }
    template <class T> PMF_DEVINL T max(T v) const { return reduce(v, [](T a, T b) { return b > a ? b : a; }); }
    template <class T> PMF_DEVINL T bcast0(T v) const
    {
        T* r = (T*)red;
        __syncthreads();
        if (threadIdx.x == 0) r[33] = v;
        __syncthreads();
        return r[33];
    }
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
    for (int i = tm.rank(); i < k; i += tm.size()) s = fma(x[i], y[i], s);
    return tm.sum(s);
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
