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
    float part[2][NW][32][4];     // per-warp partial k-vectors (OWN <= 4 floats per lane), double-buffered
    float red[2][4][NW];          // per-warp partial sums over non-zeros (<= 4 at a time), double-buffered
    int next_row;
};

// one step of a reduce-scatter over the lane pairs (lane, lane ^ mask): the lane whose bit is set
// keeps the upper half of v[0..N), its partner the lower half
template <int N> PMF_DEVINL void rt_rs_step(float* v, int mask, bool upper)
{
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float send = upper ? v[i] : v[i + N / 2];
        const float keep = upper ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(RT_FULL, send, mask);
    }
}

template <int N> PMF_DEVINL void rt_warp_sum(float (&v)[N])
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < N; j++) v[j] += __shfl_xor_sync(RT_FULL, v[j], o);
}

// sums over the row's non-zeros: warp butterfly, then (NW > 1) the per-warp partials through shared
// memory, folded by every warp in the same (butterfly) order
template <int NW, int N>
PMF_DEVINL void rt_team_sum(float (&v)[N], RtShared<NW>& sh, int& buf, int warp, int lane)
{
    static_assert(N <= 4, "at most 4 sums per reduction");
    rt_warp_sum(v);
    if (NW > 1) {
        if (lane == 0)
#pragma unroll
            for (int j = 0; j < N; j++) sh.red[buf][j][warp] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = lane < NW ? sh.red[buf][j][lane] : 0.f;
#pragma unroll
        for (int o = NW >> 1; o > 0; o >>= 1)
#pragma unroll
            for (int j = 0; j < N; j++) v[j] += __shfl_xor_sync(RT_FULL, v[j], o);
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = __shfl_sync(RT_FULL, v[j], 0);
        buf ^= 1;
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

// out[i] = (sum over the team's non-zeros of c_t F_t)[component owned by this lane, i]
template <int NC, int TPL, int NW>
PMF_DEVINL void rt_gaxpy(const float4 (&T)[TPL][NC], float c, float (&out)[NC / 2], RtShared<NW>& sh, int& buf,
                         int warp, int lane)
{
    constexpr int OWN = NC / 2;
    float cs[TPL];
#pragma unroll
    for (int s = 0; s < TPL; s++) cs[s] = __shfl_sync(RT_FULL, c, (lane & ~3) | s);
    float v[4 * NC];
#pragma unroll
    for (int j = 0; j < NC; j++) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < TPL; s++) vfma(a, cs[s], T[s][j]);
        v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
    rt_rs_step<4 * NC>(v, 16, (lane & 16) != 0);
    rt_rs_step<2 * NC>(v, 8, (lane & 8) != 0);
    rt_rs_step<NC>(v, 4, (lane & 4) != 0);
    if (NW > 1) {
#pragma unroll
        for (int i = 0; i < OWN; i++) sh.part[buf][warp][lane][i] = v[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < OWN; i++) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; w++) s += sh.part[buf][w][lane][i];
            out[i] = s;
        }
        buf ^= 1;
    } else {
#pragma unroll
        for (int i = 0; i < OWN; i++) out[i] = v[i];
    }
}

// the lane's chunk layout of a k-vector from its owned layout
template <int NC> PMF_DEVINL void rt_allgather(const float (&own)[NC / 2], float4 (&out)[NC], int ig)
{
    constexpr int OWN = NC / 2;
    float v[4 * NC];
#pragma unroll
    for (int e = 0; e < 4 * NC; e++) v[e] = __shfl_sync(RT_FULL, own[e % OWN], ((e / OWN) << 2) | ig);
#pragma unroll
    for (int j = 0; j < NC; j++) out[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// resident CTAs per SM the kernels are compiled for: 16 warps per SM when a lane holds 4 tile rows
// (128 registers), 20 with fewer
template <int NC, int TPL, int NW> struct RtCfg {
    static constexpr int tile_regs = 4 * NC * TPL;
    static constexpr int warps_per_sm = tile_regs >= 64 ? 16 : (tile_regs >= 48 ? 20 : 24);
    static constexpr int min_ctas = warps_per_sm / NW < 1 ? 1 : warps_per_sm / NW;
};

template <int NC, int TPL, int NW, int METHOD>
__global__ void __launch_bounds__(32 * NW, RtCfg<NC, TPL, NW>::min_ctas) rows_regtile_kernel(const SideParams<float> P)
{
    static_assert(NC % 2 == 0 && NC <= 8 && TPL >= 1 && TPL <= 4, "unsupported register tile");
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

        if (METHOD == M_PG) {
            // ---- pg (src/poismf.c:172-185); cs = pre-scaled column sums -----------------------
            for (int u = 0; u < hc.maxupd; u++) {
                const float p = rt_dots<NC, TPL>(T, vd, ig);
                const float c = act ? xval / p : 0.f;
                float g[OWN];
                rt_gaxpy<NC, TPL, NW>(T, c, g, sh, buf, warp, lane);
#pragma unroll
                for (int i = 0; i < OWN; i++) {
                    float v = fmaf(hc.step_w, g[i], xo[i]);
                    v += cs[i];
                    v *= hc.cdiv;
                    xo[i] = (v > 0.f) ? v : 0.f;
                }
                if (u + 1 < hc.maxupd) rt_allgather<NC>(xo, vd, ig);
            }
        } else {
            // ---- cg, cached line search (solve_cg_cached) ------------------------------------
            const float tol = 1e-2f, decr = 0.25f, c_ls = 0.01f;
            const int max_ls = 20, maxnfeval = 150;
            const int maxiter = hc.maxupd <= 0 ? INT32_MAX : hc.maxupd;
            float p = rt_dots<NC, TPL>(T, vd, ig);                  // p_t = <x, F_t>
            float cft = act ? -xval / p : 0.f;                      // gradient coefficients
            float fcur, regx;
            {
                float r[3] = {act ? xlogp(xval, p) : 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < OWN; i++) { r[1] = fmaf(cs[i], xo[i], r[1]); r[2] = fmaf(xo[i], xo[i], r[2]); }
                float ls[1] = {r[0]};
                rt_team_sum<NW, 1>(ls, sh, buf, warp, lane);
                float kk[2] = {r[1], r[2]};
                rt_warp_sum(kk);
                regx = fmaf(hc.l2, kk[1], kk[0]);
                fcur = regx - ls[0] * hc.w;                         // nonnegcg.c:191
            }
            float go[OWN], gpo[OWN], dpo[OWN], dn[OWN];
#pragma unroll
            for (int i = 0; i < OWN; i++) { gpo[i] = 0.f; dpo[i] = 0.f; }
            float gprev_sq = 0.f, fnew = 0.f;
            int nfe = 1;
            bool stop = is_bad(fcur);
            for (int it = 0; it < maxiter && !stop; it++) {
                // gradient at x (:231): one tile pass, folded onto csum + 2 l2 x
                rt_gaxpy<NC, TPL, NW>(T, cft, go, sh, buf, warp, lane);
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
                    rt_warp_sum(tb);
                    theta = tb[0] / gprev_sq;
                    beta = tb[1] / gprev_sq;
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
                rt_warp_sum(sc);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const float w2 = __shfl_xor_sync(RT_FULL, m, o); m = w2 < m ? w2 : m; }
                const float gd = sc[0], dsq = sc[1], gg = sc[2], lin = sc[3], smax = m;
                if (fabs((double)gd) <= (double)tol) break;                                  // :264-269

                rt_allgather<NC>(dn, vd, ig);
                const float q = rt_dots<NC, TPL>(T, vd, ig);                                 // q_t = <d, F_t>

                // line search (:297-327): first trial alone, then four at a time
                float step = smax;
                bool accepted = false;
                const float l2dd = hc.l2 * dsq;
                auto freg = [&](float sj) { return fmaf(sj, fmaf(sj, l2dd, lin), regx); };
                {
                    float ls[1] = {act ? xlogp(xval, fmaf(step, q, p)) : 0.f};
                    rt_team_sum<NW, 1>(ls, sh, buf, warp, lane);
                    fnew = freg(step) - ls[0] * hc.w;
                    if (!is_bad(fnew) && fnew <= fcur - c_ls * step * dsq) accepted = true;
                    else { nfe++; if (nfe >= maxnfeval) stop = true; }
                }
                int lsn = 1;
                while (!accepted && !stop && lsn < max_ls) {
                    constexpr int NB = 4;
                    const int nb = (max_ls - lsn) < NB ? (max_ls - lsn) : NB;
                    float steps[NB], lsv[NB];
                    {
                        float sj = step * decr;
#pragma unroll
                        for (int j = 0; j < NB; j++) { steps[j] = sj; sj *= decr; }
                    }
#pragma unroll
                    for (int j = 0; j < NB; j++) lsv[j] = act ? xlogp(xval, fmaf(steps[j], q, p)) : 0.f;
                    rt_team_sum<NW, NB>(lsv, sh, buf, warp, lane);
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        if (j < nb && !accepted && !stop) {
                            fnew = freg(steps[j]) - lsv[j] * hc.w;
                            if (!is_bad(fnew) && fnew <= fcur - c_ls * steps[j] * dsq) { accepted = true; step = steps[j]; }
                            else { nfe++; if (nfe >= maxnfeval) stop = true; }
                        }
                    }
                    if (!accepted) { step = steps[nb - 1]; lsn += nb; }
                }
                if (stop && !accepted) break;                                                // :317-320
                if (accepted) {
                    regx = freg(step);
#pragma unroll
                    for (int i = 0; i < OWN; i++) {
                        const float v = fmaf(step, dn[i], xo[i]);
                        xo[i] = (v >= hc.clip_thr) ? v : 0.f;
                    }
                    p = fmaf(step, q, p);
                    cft = act ? -xval / p : 0.f;
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
