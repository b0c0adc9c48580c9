// poismf_b200 — proximal-gradient and non-negative CG row solvers (device).
//
//   pg : /root/reference/src/poismf.c:126-133,172-185
//   cg : /root/reference/src/nonnegcg.c:177-346 called as src/poismf.c:315-320
//        (tol 1e-2, maxnfeval 150, maxiter = maxupd, decr .25, c .01, max_ls 20)
//
// All members of the team execute the scalar control flow redundantly on
// identical values; only the k-vector and per-non-zero loops are split.
#pragma once
#include "rowops.cuh"

namespace pmf {

// k-vectors of one row's solver, all in shared memory, each kp reals, pads zero.
template <class real> struct CgVecs {
    real *x, *g0, *g1, *d0, *d1, *xnew, *csum;
};

// ---- pg -------------------------------------------------------------------
// x: the row (shared), shift: pre-scaled column sums (shared, see sweep driver),
// buf: shared scratch vector.
template <bool STRICT, class real, class Team>
PMF_DEVINL void solve_pg(const Team& tm, const RowView<real>& rv, const HalfSweepConsts<real>& hc,
                         real* x, const real* shift, real* buf)
{
    const int k = rv.k;
    for (int u = 0; u < hc.maxupd; u++) {
        dots<STRICT>(tm, rv, x, rv.pa);
        for (int t = tm.rank(); t < rv.n; t += tm.size()) rv.pb[t] = rv.xv[t] / rv.pa[t];
        vfill(tm, buf, (real)0, rv.kp);
        tm.sync();
        gaxpy<STRICT>(tm, rv, rv.pb, buf);
        for (int i = tm.rank(); i < k; i += tm.size()) {
            real v = mad<STRICT>(hc.step_w, buf[i], x[i]);   // :177
            v = add<STRICT>(v, shift[i]);                     // :181
            v = mul<STRICT>(v, hc.cdiv);                      // :182
            x[i] = (v > (real)0) ? v : (real)0;               // :183-184 (NaN -> 0)
        }
        tm.sync();
    }
}

// ---- cg (direct: every objective evaluation re-reads the tile) ---------------
template <bool STRICT, class real, class Team>
PMF_DEVINL void solve_cg(const Team& tm, const RowView<real>& rv_in, const HalfSweepConsts<real>& hc,
                         const CgVecs<real>& vv)
{
    RowView<real> rv = rv_in;
    const int k = rv.k;
    const real tol = (real)1e-2, decr = (real)0.25, c_ls = (real)0.01;
    const int max_ls = 20, maxnfeval = 150;
    real* x = vv.x;
    real *g = vv.g0, *d = vv.d0, *gprev = nullptr, *dprev = nullptr;
    real* xnew = vv.xnew;
    const real* csum = vv.csum;
    real gprev_sq = 0, fnew = 0;
    int nfe = 1;
    const int maxiter = hc.maxupd <= 0 ? INT32_MAX : hc.maxupd;

    real fcur = eval_f_cg<STRICT>(tm, rv, csum, hc, x, rv.pa);    // nonnegcg.c:191
    if (is_bad(fcur)) return;                                      // :223-226
    bool have_p = true;   // rv.pa == <x, F_t> for the current x

    for (int it = 0; it < maxiter; it++) {
        if (!have_p) { dots<STRICT>(tm, rv, x, rv.pa); have_p = true; }
        eval_g_cg<STRICT>(tm, rv, csum, hc, x, rv.pa, g);          // :231

        for (int i = tm.rank(); i < k; i += tm.size())             // :236-239
            d[i] = (x[i] <= (real)0 && g[i] >= (real)0) ? (real)0 : -g[i];
        tm.sync();
        if (it > 0) {                                              // :242-261
            real theta = 0, beta = 0;
            if (STRICT) {
                if (tm.rank() == 0)
                    for (int i = 0; i < k; i++) {
                        if (!(x[i] <= (real)0)) {
                            theta = add_rn(theta, mul_rn(g[i], dprev[i]));
                            beta = add_rn(beta, mul_rn(g[i], sub_rn(g[i], gprev[i])));
                        } else {  // `+= 0.`
                            theta = add_rn(theta, (real)0);
                            beta = add_rn(beta, (real)0);
                        }
                    }
                theta = tm.bcast0(theta);
                beta = tm.bcast0(beta);
            } else {
                for (int i = tm.rank(); i < k; i += tm.size())
                    if (!(x[i] <= (real)0)) {
                        theta = fma(g[i], dprev[i], theta);
                        beta = fma(g[i], g[i] - gprev[i], beta);
                    }
                theta = tm.sum(theta);
                beta = tm.sum(beta);
            }
            theta /= gprev_sq;
            beta /= gprev_sq;
            for (int i = tm.rank(); i < k; i += tm.size()) {
                if (!(x[i] <= (real)0)) {
                    const real corr = sub<STRICT>(mul<STRICT>(beta, dprev[i]),
                                                  mul<STRICT>(theta, sub<STRICT>(g[i], gprev[i])));
                    d[i] = add<STRICT>(d[i], corr);
                }
            }
            tm.sync();
        }

        const real gd = vdot<STRICT>(tm, g, d, k);                 // :264-269
        if (fabs((double)gd) <= (double)tol) return;

        real smax;                                                 // :272-288
        if (hc.limit_step) {
            real m = (real)1;
            for (int i = tm.rank(); i < k; i += tm.size())
                if (d[i] < (real)0) { const real r = -x[i] / d[i]; m = (r < m) ? r : m; }
            smax = tm.min(m);
        } else {
            real m = (real)0;
            for (int i = tm.rank(); i < k; i += tm.size())
                if (d[i] < (real)0) { const real r = -x[i] / d[i]; m = (r > m) ? r : m; }
            m = tm.max(m);
            const double cand = 0.99 * (double)m;
            smax = (real)(cand < 1.0 ? cand : 1.0);
        }

        const real dsq = vdot<STRICT>(tm, d, d, k);                // :295
        real step = smax;
        bool accepted = false;

        {
            for (int ls = 0; ls < max_ls; ls++) {                   // :297-327
                for (int i = tm.rank(); i < k; i += tm.size()) {
                    real v = mad<STRICT>(step, d[i], x[i]);
                    if (hc.limit_step) v = ((double)v >= 1e-15) ? v : (real)0;
                    else v = (v > (real)0) ? v : (real)0;
                    xnew[i] = v;
                }
                tm.sync();
                fnew = eval_f_cg<STRICT>(tm, rv, csum, hc, xnew, rv.pc);
                if (!is_bad(fnew) &&
                    fnew <= sub<STRICT>(fcur, mul<STRICT>(mul<STRICT>(c_ls, step), dsq))) {
                    vcopy(tm, xnew, x, k);
                    real* tmp = rv.pa; rv.pa = rv.pc; rv.pc = tmp;   // pa now matches the new x
                    tm.sync();
                    accepted = true;
                    break;
                }
                nfe++;
                if (nfe >= maxnfeval) return;
                step = mul<STRICT>(step, decr);
            }
        }
        (void)accepted;
        fcur = fnew;                                               // :328 (Q4: even if no trial passed)
        gprev_sq = vdot<STRICT>(tm, g, g, k);                      // :332
        dprev = d; gprev = g;                                      // :335-339
        d = (d == vv.d0) ? vv.d1 : vv.d0;
        g = (g == vv.g0) ? vv.g1 : vv.g0;
    }
}

// ---- cg, cached line search (fast numerics, limit_step only) -------------------
// The line search re-uses p_t = <x,F_t> and q_t = <d,F_t>, so a trial costs O(n) instead of
// O(n*k) — the optimisation the reference's own TODO describes (src/poismf.c:191-193,
// nonnegcg.c:291-294); legal with limit_step because the clip then moves a coordinate by
// less than 1e-15.  Trial steps form the fixed geometric sequence step, step/4, ...: they
// are evaluated in batches (1, then 4 at a time) with ONE reduction over the non-zeros per
// batch, and the first acceptable one in sequence order wins, as in the sequential search.
// k-vector reductions are done redundantly per warp (no barriers); one tile pass for the
// gradient and one for q per CG iteration.
template <class real, class Team>
PMF_DEVINL void solve_cg_cached(const Team& tm, const RowView<real>& rv, const HalfSweepConsts<real>& hc,
                                const CgVecs<real>& vv)
{
    const int k = rv.k, n = rv.n;
    const real tol = (real)1e-2, decr = (real)0.25, c_ls = (real)0.01;
    const int max_ls = 20, maxnfeval = 150;
    real* x = vv.x;
    real *g = vv.g0, *d = vv.d0, *gprev = vv.g1, *dprev = vv.d1;
    const real* csum = vv.csum;
    const int kb = tm.kbegin(), ks = tm.kstride();
    const int maxiter = hc.maxupd <= 0 ? INT32_MAX : hc.maxupd;

    dots<false>(tm, rv, x, rv.pa);                                // p_t = <x, F_t>
    real fcur, regx;          // regx = <csum,x> + l2 |x|^2 at the current x, carried across iterations
    {
        real ls = 0;
        for (int t = tm.rank(); t < n; t += tm.size()) {
            const real xt = rv.xv[t], pt = rv.pa[t];
            ls += xlogp(xt, pt);
            rv.pb[t] = -xt / pt;            // gradient coefficients for the first iteration
        }
        ls = tm.nnz_sum(ls);
        real reg = 0, sq = 0;
        for (int i = kb; i < k; i += ks) { const real xi = x[i]; reg = fma(csum[i], xi, reg); sq = fma(xi, xi, sq); }
        reg = tm.ksum(reg); sq = tm.ksum(sq);
        regx = fma(hc.l2, sq, reg);
        fcur = regx - ls * hc.w;                                  // nonnegcg.c:191
    }
    if (is_bad(fcur)) return;
    int nfe = 1;
    real gprev_sq = 0, fnew = 0;

    for (int it = 0; it < maxiter; it++) {
        // ---- gradient at x (:231): one pass over the tile.  The coefficients -x_t/p_t were
        // refreshed when p was (f0 above / the accepted step below); w == 1 starts the
        // accumulation from csum + 2 l2 x inside the fold (x pads are zero, so are csum's).
        tm.sync();
        if (hc.w == (real)1) {
            gaxpy<false>(tm, rv, rv.pb, g, [&](int i) { return fma(hc.two_l2, x[i], csum[i]); });
        } else {
            for (int i = tm.rank(); i < k; i += tm.size()) g[i] = 0;
            tm.sync();
            gaxpy<false>(tm, rv, rv.pb, g);
            for (int i = tm.rank(); i < k; i += tm.size()) g[i] = fma(hc.two_l2, x[i], fma(g[i], hc.w, csum[i]));
            tm.sync();
        }
        // ---- direction (:236-261) and every k-scalar of this iteration.  Executed by the k
        // leader (the whole (sub-)warp team, or warp 0 of a CTA team, which then broadcasts).
        // Besides <g,d>, |d|^2, |g|^2 and the step bound it collects lin = <csum + 2 l2 x, d>, from
        // which the regulariser of every trial point follows in O(1):
        //   <csum,x+s d> + l2 |x+s d|^2 = regx + s lin + s^2 l2 |d|^2
        // (the clip at 1e-15 moves this by < 1e-15 |csum|: inside the cached search's tolerance)
        real gd = 0, dsq = 0, gg = 0, smax = (real)1, lin = 0;
        if (tm.k_leader()) {
            real theta = 0, beta = 0;
            if (it > 0) {
                for (int i = kb; i < k; i += ks)
                    if (!(x[i] <= (real)0)) {
                        const real gi = g[i];
                        theta = fma(gi, dprev[i], theta);
                        beta = fma(gi, gi - gprev[i], beta);
                    }
                theta = tm.ksum(theta) / gprev_sq;
                beta = tm.ksum(beta) / gprev_sq;
            }
            real m = (real)1;
            for (int i = kb; i < k; i += ks) {
                const real xi = x[i], gi = g[i], ci = csum[i];
                real di = (xi <= (real)0 && gi >= (real)0) ? (real)0 : -gi;
                if (it > 0 && !(xi <= (real)0)) di += beta * dprev[i] - theta * (gi - gprev[i]);
                d[i] = di;
                gd = fma(gi, di, gd); dsq = fma(di, di, dsq); gg = fma(gi, gi, gg);
                lin = fma(fma(hc.two_l2, xi, ci), di, lin);
                if (di < (real)0) { const real r = -xi / di; m = (r < m) ? r : m; }  // limit_step (:272-279)
            }
            gd = tm.ksum(gd); dsq = tm.ksum(dsq); gg = tm.ksum(gg);
            lin = tm.ksum(lin);
            smax = tm.kmin(m);
            if (Team::k_bcast && kb == 0) {
                real* ksl = tm.template kslots<real>();
                ksl[0] = gd; ksl[1] = dsq; ksl[2] = gg; ksl[3] = lin; ksl[4] = smax;
            }
        }
        tm.sync();                                                 // d (and the scalars) visible to everyone
        if (Team::k_bcast) {
            const real* ksl = tm.template kslots<real>();
            gd = ksl[0]; dsq = ksl[1]; gg = ksl[2]; lin = ksl[3]; smax = ksl[4];
        }
        if (fabs((double)gd) <= (double)tol) return;               // :264-269

        dots<false>(tm, rv, d, rv.pc);                             // q_t = <d, F_t>

        // ---- line search (:297-327): first trial alone, then four at a time
        real step = smax;
        bool accepted = false;
        const real l2dd = hc.l2 * dsq;
        auto freg = [&](real sj) {   // <csum,x'> + l2 |x'|^2 at x' = x + sj d
            return fma(sj, fma(sj, l2dd, lin), regx);
        };
        {
            real lsum = 0;
            for (int t = tm.rank(); t < n; t += tm.size())
                lsum += xlogp(rv.xv[t], fma(step, rv.pc[t], rv.pa[t]));
            lsum = tm.nnz_sum(lsum);
            fnew = freg(step) - lsum * hc.w;
            if (!is_bad(fnew) && fnew <= fcur - c_ls * step * dsq) accepted = true;
            else { nfe++; if (nfe >= maxnfeval) return; }
        }
        int ls = 1;
        while (!accepted && ls < max_ls) {
            constexpr int NB = 4;
            const int nb = (max_ls - ls) < NB ? (max_ls - ls) : NB;
            real steps[NB], lsv[NB];
            {
                real sj = step * decr;
#pragma unroll
                for (int j = 0; j < NB; j++) { steps[j] = sj; sj *= decr; lsv[j] = 0; }
            }
            for (int t = tm.rank(); t < n; t += tm.size()) {
                const real pt = rv.pa[t], qt = rv.pc[t], xt = rv.xv[t];
#pragma unroll
                for (int j = 0; j < NB; j++) lsv[j] += xlogp(xt, fma(steps[j], qt, pt));
            }
            tm.nnz_sum_n(lsv);
            bool stop_all = false;
#pragma unroll
            for (int j = 0; j < NB; j++) {
                if (j < nb && !accepted && !stop_all) {
                    fnew = freg(steps[j]) - lsv[j] * hc.w;
                    if (!is_bad(fnew) && fnew <= fcur - c_ls * steps[j] * dsq) { accepted = true; step = steps[j]; }
                    else { nfe++; if (nfe >= maxnfeval) stop_all = true; }
                }
            }
            if (stop_all && !accepted) return;                      // :317-320
            if (!accepted) { step = steps[nb - 1]; ls += nb; }
        }
        if (accepted) {
            regx = freg(step);
            for (int i = tm.rank(); i < k; i += tm.size()) {
                const real v = fma(step, d[i], x[i]);
                x[i] = (v >= hc.clip_thr) ? v : (real)0;
            }
            for (int t = tm.rank(); t < n; t += tm.size()) {
                const real pt = fma(step, rv.pc[t], rv.pa[t]);
                rv.pa[t] = pt;
                rv.pb[t] = -rv.xv[t] / pt;
            }
        }
        fcur = fnew;                                               // :328 (Q4)
        gprev_sq = gg;                                             // :332
        real* tv = d; d = dprev; dprev = tv;                       // :335-339
        tv = g; g = gprev; gprev = tv;
    }
}

}  // namespace pmf
