// poismf_b200 — register-tile row solvers (fast numerics, float32): pg and the cached-line-search cg.
//
// Same algorithm as solve_pg / solve_cg_cached (solver_pg_cg.cuh; the reference's
// src/poismf.c:126-188 and src/nonnegcg.c:177-346 with the cached line search its own TODO
// describes, src/poismf.c:191-193), laid out for rows of up to 512 non-zeros so that NOTHING the
// inner loop touches lives in shared memory:
//
//   * the row's tile (the gathered rows {F_j}) is loaded ONCE per half-sweep, with 16-byte loads,
//     straight into REGISTERS.  A warp is viewed as 8 x 4 lanes (tg = lane / 4, ig = lane % 4):
//     lane (tg, ig) holds, for each of its TPL non-zeros t = (warp, s, tg), the 16-byte chunks
//     ig, ig + 4, ig + 8, ... of F_t  (NC chunks: k <= 16 NC).
//   * <v, F_t> : every lane multiplies its chunks, the 4 ig-lanes of a non-zero fold their partial
//     sums with 3 shuffles in a reduce-scatter that leaves lane ig with the dot product of
//     non-zero slot ig — the lane that then owns that non-zero's scalars (x_t, p_t, q_t, c_t) in
//     the O(n) line search.
//   * sum_t c_t F_t : every lane accumulates its TPL non-zeros into its 4 NC components, the 8
//     tg-lanes fold with a reduce-scatter (2NC + NC + NC/2 shuffles) that leaves each lane OWNING
//     NC/2 components of every k-vector of the solver (x, g, g_prev, d, d_prev, csum): all k-vector
//     arithmetic is then NC/2 elements per lane in registers, its reductions plain warp butterflies.
//   * teams of NW > 1 warps (one CTA per row) fold the per-warp partial k-vectors and the sums over
//     non-zeros through a few hundred bytes of shared memory and ONE barrier per reduction; every
//     warp then holds identical k-vectors and executes the same control flow, so no broadcast is
//     needed.
//
// Per CG iteration a warp issues ~8 NC TPL FMAs for the two tile passes, ~40 shuffles and two
// packed butterflies — an order of magnitude fewer instructions per non-zero than the
// shared-memory teams of kernels.cuh, which remain the path for strict numerics, double
// precision, w_mult != 1, cg without limit_step, tncg and rows longer than 512.
#pragma once
#include "kernels.cuh"

namespace pmf {

constexpr unsigned RT_FULL = 0xffffffffu;

template <int NW> struct RtShared {
    float fold[NW][4][8][4][4];   // per warp: [chunk slot j][tg][ig][4]: the 8 tg-lanes' partial k-vectors of the gaxpy pass
    float part[2][NW][64];        // per-warp folded partial k-vectors, double-buffered (NW > 1)
    float red[2][8][NW];          // per-warp partial sums over non-zeros (<= 8 at a time), double-buffered
    float dsh[64];                // the search direction, written by the k-phase warp, read in chunk layout by all
    float ksc[8];                 // k-scalars of the iteration: <g,d>, |d|^2, |g|^2, <csum + 2 l2 x, d>, max step
    int next_row;
};

PMF_DEVINL float rt_redux_min(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

// sums of N = 2, 4 or 8 values over the 32 lanes of a warp with a transposing butterfly: after the call
// v[0] of the lanes whose TOP log2(N) lane bits spell j holds the sum of value j  (N-1 + 5-log2(N) shuffles
// instead of 5 N)
template <int N> PMF_DEVINL void rt_packed_sum(float (&v)[N], int lane)
{
    static_assert(N == 2 || N == 4 || N == 8, "2, 4 or 8 values");
    if (N >= 8) {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(RT_FULL, send, 16);
        }
    }
    if (N >= 4) {
        constexpr int M = N >= 8 ? 8 : 16;
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(RT_FULL, send, M);
        }
    }
    {
        constexpr int M = N >= 8 ? 4 : (N >= 4 ? 8 : 16);
        const bool up = (lane & M) != 0;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(RT_FULL, send, M);
#pragma unroll
        for (int o = M >> 1; o > 0; o >>= 1) v[0] += __shfl_xor_sync(RT_FULL, v[0], o);
    }
}

// <v, F_t> for the lane's non-zero slot (see the header comment); v in the lane's chunk layout
template <int NC, int TPL>
PMF_DEVINL float rt_dots(const float4 (&T)[TPL][NC], const float4 (&v)[NC], int ig)
{
    constexpr int PS = TPL <= 1 ? 1 : (TPL <= 2 ? 2 : 4);
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < TPL; s++) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NC; j++) acc = vdot4(T[s][j], v[j], acc);
        a[s] = acc;
    }
    if (PS == 4) {
        const bool b1 = (ig & 2) != 0, b0 = (ig & 1) != 0;
        float k0 = b1 ? a[2] : a[0], k1 = b1 ? a[3] : a[1];
        const float s0 = b1 ? a[0] : a[2], s1 = b1 ? a[1] : a[3];
        k0 += __shfl_xor_sync(RT_FULL, s0, 2);
        k1 += __shfl_xor_sync(RT_FULL, s1, 2);
        float kk = b0 ? k1 : k0;
        const float ss = b0 ? k0 : k1;
        kk += __shfl_xor_sync(RT_FULL, ss, 1);
        return kk;
    } else if (PS == 2) {
        const bool b0 = (ig & 1) != 0;
        float kk = b0 ? a[1] : a[0];
        const float ss = b0 ? a[0] : a[1];
        kk += __shfl_xor_sync(RT_FULL, ss, 1);
        kk += __shfl_xor_sync(RT_FULL, kk, 2);
        return kk;
    } else {
        float kk = a[0];
        kk += __shfl_xor_sync(RT_FULL, kk, 1);
        kk += __shfl_xor_sync(RT_FULL, kk, 2);
        return kk;
    }
}

// out[i] = (sum over THIS WARP's non-zeros of c_t F_t)[component owned by this lane, i].  The 8 tg-lanes'
// partial vectors are folded through the warp's 2 KB of shared memory (4 STS.128 + 8 LDS.64 per lane, fixed
// order) rather than by shuffles: a third of the instructions of a transposing butterfly.
template <int NC, int TPL, int NW>
PMF_DEVINL void rt_gaxpy(const float4 (&T)[TPL][NC], float c, float (&out)[NC / 2], RtShared<NW>& sh, int warp, int lane)
{
    constexpr int OWN = NC / 2;
    const int ig = lane & 3, tg = lane >> 2;
    float cs[TPL];
#pragma unroll
    for (int s = 0; s < TPL; s++) cs[s] = __shfl_sync(RT_FULL, c, (lane & ~3) | s);
    __syncwarp();                                   // the previous fold's readers are done
#pragma unroll
    for (int j = 0; j < NC; j++) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < TPL; s++) vfma(a, cs[s], T[s][j]);
        *reinterpret_cast<float4*>(&sh.fold[warp][j][tg][ig][0]) = a;      // 32 lanes: 512 contiguous bytes
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < OWN; i++) out[i] = 0.f;
    // owned components: flat elements tg*OWN + i of this lane's 4 NC (chunk slot (tg*OWN + i) / 4)
#pragma unroll
    for (int t = 0; t < 8; t++) {
        if (OWN == 2) {
            const float2 v = *reinterpret_cast<const float2*>(&sh.fold[warp][tg >> 1][t][ig][(tg & 1) * 2]);
            out[0] += v.x; out[1] += v.y;
        } else {
#pragma unroll
            for (int i = 0; i < OWN; i++) out[i] += sh.fold[warp][(tg * OWN + i) >> 2][t][ig][(tg * OWN + i) & 3];
        }
    }
}

// the lane's chunk layout of the k-vector in sh.dsh (component order)
template <int NC, int NW> PMF_DEVINL void rt_read_dir(const RtShared<NW>& sh, float4 (&out)[NC], int ig)
{
#pragma unroll
    for (int j = 0; j < NC; j++) out[j] = *reinterpret_cast<const float4*>(&sh.dsh[4 * (ig + 4 * j)]);
}

// resident CTAs per SM the kernels are compiled for: 16 warps per SM when a lane holds 4 tile rows
// (128 registers), 20 with fewer
#ifndef PMF_RT_WARPS64
#define PMF_RT_WARPS64 20
#endif
template <int NC, int TPL, int NW> struct RtCfg {
    static constexpr int tile_regs = 4 * NC * TPL;
    static constexpr int warps_per_sm = tile_regs >= 64 ? PMF_RT_WARPS64 : (tile_regs >= 48 ? 20 : 24);
    static constexpr int min_ctas = warps_per_sm / NW < 1 ? 1 : warps_per_sm / NW;
};

template <int NC, int TPL, int NW, int METHOD>
__global__ void __launch_bounds__(32 * NW, RtCfg<NC, TPL, NW>::min_ctas) rows_regtile_kernel(const SideParams<float> P)
{
    static_assert(NC % 2 == 0 && NC <= 4 && TPL >= 1 && TPL <= 4, "unsupported register tile");
    constexpr int OWN = NC / 2;
    constexpr int PS = TPL <= 1 ? 1 : (TPL <= 2 ? 2 : 4);
    __shared__ RtShared<NW> sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ig = lane & 3, tg = lane >> 2;
    const int L = P.ldf >> 2, k = P.k, ldf = P.ldf;
    const HalfSweepConsts<float>& hc = P.hc;
    int buf = 0;

    int idx;
    if (NW == 1) {
        idx = lane == 0 ? atomicAdd(P.counter, 1) : 0;
        idx = __shfl_sync(RT_FULL, idx, 0);
    } else {
        if (threadIdx.x == 0) sh.next_row = atomicAdd(P.counter, 1);
        __syncthreads();
        idx = sh.next_row;
    }
    while (idx < P.nrows) {
        int nxt = 0;
        if (threadIdx.x == 0) nxt = atomicAdd(P.counter, 1);      // consumed after this row: latency hidden
        const int row = P.rows[idx];
        const long long beg = P.ptr[row];
        const int n = (int)(P.ptr[row + 1] - beg);
        float* Mrow = P.M + (size_t)row * ldf;

        // ---- the tile, once, into registers ---------------------------------------------------
        float4 T[TPL][NC];
        {
            int id[TPL];
#pragma unroll
            for (int s = 0; s < TPL; s++) {
                const int t = warp * (8 * TPL) + s * 8 + tg;
                id[s] = t < n ? __ldg(P.ind + beg + t) : -1;
            }
#pragma unroll
            for (int s = 0; s < TPL; s++) {
                const float4* fr = reinterpret_cast<const float4*>(P.F + (size_t)(id[s] < 0 ? 0 : id[s]) * ldf);
#pragma unroll
                for (int j = 0; j < NC; j++) {
                    const int c = ig + 4 * j;
                    T[s][j] = (id[s] >= 0 && c < L) ? __ldg(fr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        // the non-zero whose scalars this lane owns
        const int sl = ig & (PS - 1);
        const int t_ls = warp * (8 * TPL) + sl * 8 + tg;
        const bool act = ig < PS && sl < TPL && t_ls < n;
        const float xval = act ? __ldg(P.xv + beg + t_ls) : 0.f;
        // the row being solved: chunk layout (for the first tile pass) and owned components
        float4 vd[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int c = ig + 4 * j;
            vd[j] = c < L ? *reinterpret_cast<const float4*>(Mrow + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float xo[OWN], cs[OWN];
        int gi[OWN];
#pragma unroll
        for (int i = 0; i < OWN; i++) {
            const int e = tg * OWN + i;
            gi[i] = 4 * (ig + 4 * (e >> 2)) + (e & 3);
            xo[i] = gi[i] < k ? Mrow[gi[i]] : 0.f;
            cs[i] = gi[i] < k ? P.csum[gi[i]] : 0.f;
        }

        // k-vector phases are executed by warp 0 only (NW > 1: the other warps wait at the barrier instead
        // of repeating them); its lanes own the row's components
        const bool kwarp = warp == 0;
        auto team_sync = [&]() { if (NW > 1) __syncthreads(); else __syncwarp(); };
        // fold of the per-warp partial k-vectors (fixed order), by the k-phase warp
        auto fold_parts = [&](float (&v)[OWN]) {
            if (NW > 1) {
#pragma unroll
                for (int i = 0; i < OWN; i++) sh.part[buf][warp][lane * OWN + i] = v[i];
                __syncthreads();
                if (kwarp) {
#pragma unroll
                    for (int i = 0; i < OWN; i++) v[i] = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; w++)
#pragma unroll
                        for (int i = 0; i < OWN; i++) v[i] += sh.part[buf][w][lane * OWN + i];
                }
                buf ^= 1;
            }
        };
        // the k-phase warp publishes a k-vector (owned layout -> component order) for everybody's chunk layout
        auto publish = [&](const float (&v)[OWN]) {
            if (kwarp) {
#pragma unroll
                for (int i = 0; i < OWN; i++) sh.dsh[gi[i]] = v[i];       // gi[0..OWN) are consecutive components
            }
        };

        if (METHOD == M_PG) {
            // ---- pg (src/poismf.c:172-185); cs = pre-scaled column sums -----------------------
            for (int u = 0; u < hc.maxupd; u++) {
                const float p = rt_dots<NC, TPL>(T, vd, ig);
                const float c = act ? xval / p : 0.f;
                float g[OWN];
                rt_gaxpy<NC, TPL, NW>(T, c, g, sh, warp, lane);
                fold_parts(g);
                if (kwarp) {
#pragma unroll
                    for (int i = 0; i < OWN; i++) {
                        float v = fmaf(hc.step_w, g[i], xo[i]);
                        v += cs[i];
                        v *= hc.cdiv;
                        xo[i] = (v > 0.f) ? v : 0.f;
                    }
                }
                if (u + 1 < hc.maxupd) {
                    team_sync();            // the previous readers of dsh are done
                    publish(xo);
                    team_sync();
                    rt_read_dir<NC, NW>(sh, vd, ig);
                }
            }
        } else {
            // ---- cg, cached line search (solve_cg_cached) ------------------------------------
            const float tol = 1e-2f, c_ls = 0.01f;
            const int max_ls = 20, maxnfeval = 150;
            const int maxiter = hc.maxupd <= 0 ? INT32_MAX : hc.maxupd;
            const float xl = xval * 0.693147180559945f;            // x_t ln 2: x log p = xl * log2 p
            float p = rt_dots<NC, TPL>(T, vd, ig);                  // p_t = <x, F_t>
            float cft = act ? __fdividef(-xval, p) : 0.f;           // gradient coefficients
            float fcur, regx;
            {
                float r[4] = {act ? xl * __log2f(p) : 0.f, 0.f, 0.f, 0.f};
                if (kwarp) {
#pragma unroll
                    for (int i = 0; i < OWN; i++) { r[1] = fmaf(cs[i], xo[i], r[1]); r[2] = fmaf(xo[i], xo[i], r[2]); }
                }
                rt_packed_sum<4>(r, lane);                          // lanes 0 / 8 / 16: the three sums of this warp
                if ((lane & 7) == 0 && lane < 24) sh.red[buf][lane >> 3][warp] = r[0];
                team_sync();
                float ls = 0.f;
#pragma unroll
                for (int w = 0; w < NW; w++) ls += sh.red[buf][0][w];
                regx = fmaf(hc.l2, sh.red[buf][2][0], sh.red[buf][1][0]);
                fcur = regx - ls * hc.w;                            // nonnegcg.c:191
                buf ^= 1;
            }
            float go[OWN], gpo[OWN], dpo[OWN], dn[OWN];
#pragma unroll
            for (int i = 0; i < OWN; i++) { gpo[i] = 0.f; dpo[i] = 0.f; go[i] = 0.f; dn[i] = 0.f; }
            float gprev_sq = 0.f, fnew = 0.f;
            int nfe = 1;
            bool stop = is_bad(fcur);
            for (int it = 0; it < maxiter && !stop; it++) {
                // gradient at x (:231): one tile pass, folded onto csum + 2 l2 x
                rt_gaxpy<NC, TPL, NW>(T, cft, go, sh, warp, lane);
                fold_parts(go);
                if (kwarp) {
#pragma unroll
                    for (int i = 0; i < OWN; i++) go[i] = fmaf(hc.two_l2, xo[i], cs[i]) + go[i];
                    // direction (:236-261) and every k-scalar of this iteration
                    float theta = 0.f, beta = 0.f;
                    if (it > 0) {
                        float tb[2] = {0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < OWN; i++)
                            if (!(xo[i] <= 0.f)) {
                                tb[0] = fmaf(go[i], dpo[i], tb[0]);
                                tb[1] = fmaf(go[i], go[i] - gpo[i], tb[1]);
                            }
                        rt_packed_sum<2>(tb, lane);                 // lanes < 16: theta's sum, lanes >= 16: beta's
                        const float t0 = __shfl_sync(RT_FULL, tb[0], 0), t1 = __shfl_sync(RT_FULL, tb[0], 16);
                        const float inv = 1.f / gprev_sq;
                        theta = t0 * inv;
                        beta = t1 * inv;
                    }
                    float sc[4] = {0.f, 0.f, 0.f, 0.f};    // <g,d>, |d|^2, |g|^2, <csum + 2 l2 x, d>
                    float m = 1.f;
#pragma unroll
                    for (int i = 0; i < OWN; i++) {
                        const float xi = xo[i], g_i = go[i];
                        float di = (xi <= 0.f && g_i >= 0.f) ? 0.f : -g_i;
                        if (it > 0 && !(xi <= 0.f)) di += beta * dpo[i] - theta * (g_i - gpo[i]);
                        dn[i] = di;
                        sc[0] = fmaf(g_i, di, sc[0]); sc[1] = fmaf(di, di, sc[1]); sc[2] = fmaf(g_i, g_i, sc[2]);
                        sc[3] = fmaf(fmaf(hc.two_l2, xi, cs[i]), di, sc[3]);
                        if (di < 0.f) { const float r = -xi / di; m = (r < m) ? r : m; }        // limit_step (:272-279)
                    }
                    rt_packed_sum<4>(sc, lane);                     // lanes 0 / 8 / 16 / 24 hold the four sums
                    m = rt_redux_min(m);
                    if ((lane & 7) == 0) sh.ksc[lane >> 3] = sc[0];
                    if (lane == 0) sh.ksc[4] = m;
                }
                team_sync();                 // (also: the previous readers of dsh are past their tile pass)
                const float gd = sh.ksc[0], dsq = sh.ksc[1], gg = sh.ksc[2], lin = sh.ksc[3], smax = sh.ksc[4];
                if (fabsf(gd) <= tol) break;                                                 // :264-269 (floats compare alike in double)
                publish(dn);
                team_sync();
                rt_read_dir<NC, NW>(sh, vd, ig);
                const float q = rt_dots<NC, TPL>(T, vd, ig);                                 // q_t = <d, F_t>

                // line search (:297-327): trial steps smax 4^-j, FOUR per reduction round; lane group j = lane / 8
                // ends with trial j's sum; the first acceptable trial in sequence order wins
                bool accepted = false;
                float step = 0.f;
                const float l2dd = hc.l2 * dsq;
                auto freg = [&](float sj) { return fmaf(sj, fmaf(sj, l2dd, lin), regx); };
                auto pow4 = [](int e) { return __int_as_float((127 - 2 * e) << 23); };       // 4^-e, exact
                for (int base = 0; base < max_ls && !accepted && !stop; base += 4) {
                    float lsv[4];
                    {
                        float sj = smax * pow4(base);
#pragma unroll
                        for (int j = 0; j < 4; j++) { lsv[j] = xl * __log2f(fmaf(sj, q, p)); sj *= 0.25f; }
                    }
                    if (!act) { lsv[0] = 0.f; lsv[1] = 0.f; lsv[2] = 0.f; lsv[3] = 0.f; }
                    rt_packed_sum<4>(lsv, lane);
                    const int jmine = lane >> 3;
                    float tot = lsv[0];
                    if (NW > 1) {
                        if ((lane & 7) == 0) sh.red[buf][jmine][warp] = tot;
                        __syncthreads();
                        tot = 0.f;
#pragma unroll
                        for (int w = 0; w < NW; w++) tot += sh.red[buf][jmine][w];
                        buf ^= 1;
                    }
                    const float sj = smax * pow4(base + jmine);
                    const float fj = freg(sj) - tot * hc.w;
                    const bool okj = !is_bad(fj) && fj <= fcur - c_ls * sj * dsq;
                    const unsigned okmask = __ballot_sync(RT_FULL, okj);
                    const int firstok = okmask ? (__ffs(okmask) - 1) >> 3 : 4;
                    const int fails_allowed = maxnfeval - nfe;      // the fails_allowed-th failed trial stops the solver
                    int last;
                    if (firstok < 4 && firstok < fails_allowed) { accepted = true; nfe += firstok; last = firstok; }
                    else if (fails_allowed <= 4 && fails_allowed <= firstok) { stop = true; nfe += fails_allowed; last = fails_allowed - 1; }
                    else { nfe += 4; last = 3; }
                    fnew = __shfl_sync(RT_FULL, fj, 8 * last);
                    step = __shfl_sync(RT_FULL, sj, 8 * last);
                }
                if (stop && !accepted) break;                                                // :317-320
                if (accepted) {
                    regx = freg(step);
                    if (kwarp) {
#pragma unroll
                        for (int i = 0; i < OWN; i++) {
                            const float v = fmaf(step, dn[i], xo[i]);
                            xo[i] = (v >= hc.clip_thr) ? v : 0.f;
                        }
                    }
                    p = fmaf(step, q, p);
                    cft = act ? __fdividef(-xval, p) : 0.f;
                }
                fcur = fnew;                                                                 // :328 (Q4)
                gprev_sq = gg;                                                               // :332
#pragma unroll
                for (int i = 0; i < OWN; i++) { gpo[i] = go[i]; dpo[i] = dn[i]; }            // :335-339
            }
        }

        // ---- the solved row: own replica and (fused exchange) every peer's ----------------------
        if (warp == 0) {
#pragma unroll
            for (int i = 0; i < OWN; i++)
                if (gi[i] < k) {
                    Mrow[gi[i]] = xo[i];
                    for (int q = 0; q < P.npeers; q++) P.peerM[q][(size_t)row * ldf + gi[i]] = xo[i];
                }
        }
        if (NW == 1) idx = __shfl_sync(RT_FULL, nxt, 0);
        else {
            __syncthreads();
            if (threadIdx.x == 0) sh.next_row = nxt;
            __syncthreads();
            idx = sh.next_row;
        }
    }
}

}  // namespace pmf
