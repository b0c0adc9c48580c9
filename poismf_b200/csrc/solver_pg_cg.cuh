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

// ---- cg -------------------------------------------------------------------
// CACHED (fast mode, limit_step only): the line search re-uses p_t = <x,F_t> and
// q_t = <d,F_t> so that a trial costs O(n) instead of O(n*k) — the optimisation
// the reference's own TODO describes (src/poismf.c:191-193, nonnegcg.c:291-294).
template <bool STRICT, bool CACHED, class real, class Team>
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

        if (CACHED) {
            // q_t = <d, F_t> once per CG iteration; p_trial = p + step*q
            dots<false>(tm, rv, d, rv.pc);
            for (int ls = 0; ls < max_ls; ls++) {
                real reg = 0, sq = 0;
                for (int i = tm.rank(); i < k; i += tm.size()) {
                    real v = fma(step, d[i], x[i]);
                    v = ((double)v >= 1e-15) ? v : (real)0;
                    xnew[i] = v;
                    reg = fma(csum[i], v, reg);
                    sq = fma(v, v, sq);
                }
                real lsum = 0;
                for (int t = tm.rank(); t < rv.n; t += tm.size())
                    lsum += xlogp(rv.xv[t], fma(step, rv.pc[t], rv.pa[t]));
                reg = tm.sum(reg); sq = tm.sum(sq); lsum = tm.sum(lsum);
                fnew = fma(hc.l2, sq, reg) - lsum * hc.w;
                if (!is_bad(fnew) && fnew <= fcur - c_ls * step * dsq) {
                    tm.sync();
                    vcopy(tm, xnew, x, k);
                    for (int t = tm.rank(); t < rv.n; t += tm.size())
                        rv.pa[t] = fma(step, rv.pc[t], rv.pa[t]);
                    tm.sync();
                    accepted = true;
                    break;
                }
                nfe++;
                if (nfe >= maxnfeval) return;
                step *= decr;
                tm.sync();
            }
        } else {
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

}  // namespace pmf
