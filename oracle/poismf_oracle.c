/* ============================================================================
 * TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the poismf hot path.
 *
 * This file restates, in plain sequential C, the algorithm of the reference
 * (david-cortes/poismf) for the alternating sweep, predict_multiple and topN.
 * It is imported only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, as the CHECKER.  The product library
 * (libpoismf_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity pinning: the reference ships no golden vectors, so this
 * restatement is pinned against the reference ITSELF, compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/ (strict build:
 * -O2 -ffp-contract=off + oracle/blas_shim.c).  tests/test_oracle_vs_ref.py
 * asserts bit-identical factors for pg/cg/tncg in double and float, and the
 * fixtures under tests/golden/ (made by tests/golden/make_golden.py from
 * oracle/_ref) pin it where /root/reference is absent (the GPU box).
 *
 * Every function cites the reference file:line it follows.  All level-1
 * "BLAS" operations are the naive left-to-right loops of oracle/blas_shim.c.
 * Floating-point expression shapes (which sub-expressions are evaluated in
 * double even when real==float: the reference never uses tgmath)
 * are kept, because in float they change the rounding.
 *
 * Build: see oracle/Makefile (-DORACLE_FLOAT selects real = float).
 * ==========================================================================*/
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_FLOAT
typedef float real;
#define REAL_EPS FLT_EPSILON
#define REAL_HUGE HUGE_VALF
#else
typedef double real;
#define REAL_EPS DBL_EPSILON
#define REAL_HUGE HUGE_VAL
#endif
typedef uint64_t ix_t; /* == size_t of the Python build, src/poismf.h:76 */
static const real LOW = 0; /* the only lower bound used (zeros_tncg, src/poismf.c:485); typed `real` so that
                              expressions involving it stay in `real` arithmetic as in the reference */

/* -------------------------------------------------------------------------
 * Sequential vector kernels (oracle/blas_shim.c order)
 * -----------------------------------------------------------------------*/
static real vdot(int n, const real *x, const real *y)
{
    real s = 0;
    for (int i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}
static void vaxpy(int n, real a, const real *x, real *y)
{
    for (int i = 0; i < n; i++) y[i] += a * x[i];
}
static void vscal(int n, real a, real *x)
{
    for (int i = 0; i < n; i++) x[i] *= a;
}
static real vnrm2(int n, const real *x)
{
    real s = 0;
    for (int i = 0; i < n; i++) s += x[i] * x[i];
#ifdef ORACLE_FLOAT
    return sqrtf(s);
#else
    return sqrt(s);
#endif
}

/* One row's slice of the sparse matrix plus the constants of its sub-problem
 * (the reference's `fdata` closure, src/poismf.h:121-130). */
typedef struct {
    const real *F;     /* the FIXED factor matrix, row-major [other_dim x k] */
    const real *csum;  /* column sums (+l1), or this row's weighted version  */
    const real *xval;  /* the row's non-zero values                          */
    const ix_t *xind;  /* the row's non-zero indices                          */
    ix_t nnz;
    real l2, w;
    int k;
} rowprob;

/* src/poismf.c:194-208  calc_fun_single (objective used by cg) */
static real obj_cg(const real *a, const rowprob *p)
{
    int k = p->k;
    real reg = vdot(k, p->csum, a);
    reg += p->l2 * vdot(k, a, a);
    real ls = 0.;
    for (ix_t t = 0; t < p->nnz; t++)
        ls += p->xval[t] * log(vdot(k, a, p->F + p->xind[t] * (size_t)k));
    return reg - ls * p->w;
}

/* src/poismf.c:210-223 (w==1) and :225-240 (w!=1)  gradient used by cg */
static void grad_cg(const real *a, real *g, const rowprob *p)
{
    int k = p->k;
    if (p->w == 1.) {
        memcpy(g, p->csum, sizeof(real) * (size_t)k);
        vaxpy(k, 2. * p->l2, a, g);
        for (ix_t t = 0; t < p->nnz; t++) {
            const real *f = p->F + p->xind[t] * (size_t)k;
            vaxpy(k, -p->xval[t] / vdot(k, a, f), f, g);
        }
    } else {
        memset(g, 0, sizeof(real) * (size_t)k);
        for (ix_t t = 0; t < p->nnz; t++) {
            const real *f = p->F + p->xind[t] * (size_t)k;
            vaxpy(k, -p->xval[t] / vdot(k, a, f), f, g);
        }
        vscal(k, p->w, g);
        vaxpy(k, 1., p->csum, g);
        vaxpy(k, 2. * p->l2, a, g);
    }
}

/* src/poismf.c:242-273  calc_fun_and_grad (tncg; objective has NO l2 term, Q3) */
static real obj_grad_tn(const real *a, real *g, const rowprob *p)
{
    int k = p->k;
    real ls = 0;
    memset(g, 0, sizeof(real) * (size_t)k);
    for (ix_t t = 0; t < p->nnz; t++) {
        const real *f = p->F + p->xind[t] * (size_t)k;
        real pred = vdot(k, a, f);
        vaxpy(k, -p->xval[t] / pred, f, g);
        ls += p->xval[t] * log(pred);
    }
    if (p->w != 1.) vscal(k, p->w, g);
    vaxpy(k, 1., p->csum, g);
    real reg = vdot(k, p->csum, a);
    vaxpy(k, 2. * p->l2, a, g);
    return reg - ls * p->w;
}

/* Exposed single evaluations, for known-answer tests of the device kernels. */
void oracle_eval(const real *a, const real *F, const real *csum,
                 const real *xval, const ix_t *xind, ix_t nnz, int k,
                 real l2, real w,
                 real *f_cg, real *g_cg, real *f_tn, real *g_tn)
{
    rowprob p = { F, csum, xval, xind, nnz, l2, w, k };
    *f_cg = obj_cg(a, &p);
    grad_cg(a, g_cg, &p);
    *f_tn = obj_grad_tn(a, g_tn, &p);
}

/* -------------------------------------------------------------------------
 * pg row update — src/poismf.c:126-133 (calc_grad_pgd) + :172-185
 * `shift` is the pre-scaled column-sum vector the caller prepared.
 * -----------------------------------------------------------------------*/
static void pg_row(real *a, const rowprob *p, const real *shift, real step_w,
                   real cdiv, size_t maxupd, real *buf)
{
    int k = p->k;
    for (size_t u = 0; u < maxupd; u++) {
        memset(buf, 0, sizeof(real) * (size_t)k);
        for (ix_t t = 0; t < p->nnz; t++) {
            const real *f = p->F + p->xind[t] * (size_t)k;
            vaxpy(k, p->xval[t] / vdot(k, f, a), f, buf);
        }
        vaxpy(k, step_w, buf, a);
        vaxpy(k, 1., shift, a);
        vscal(k, cdiv, a);
        for (int i = 0; i < k; i++) a[i] = (a[i] > 0.) ? a[i] : 0.;
    }
}

/* -------------------------------------------------------------------------
 * cg row solver — src/nonnegcg.c:177-346 (Li 2013 modified PRP, non-negative)
 * Called as src/poismf.c:315-320: tol 1e-2, maxnfeval 150, decr .25, c .01,
 * max_ls 20.  Returns iterations done through *niter_out (may be NULL).
 * -----------------------------------------------------------------------*/
void oracle_cg_solve(real *x, const rowprob *p, size_t maxiter, int limit_step,
                     real *buf5k, size_t *niter_out, size_t *nfeval_out, real *f_out)
{
    const real tol = 1e-2, decr = 0.25, c_ls = 0.01;
    const size_t max_ls = 20, maxnfeval = 150;
    int n = p->k;
    real *gbuf = buf5k, *dbuf = buf5k + 2 * (size_t)n, *xnew = buf5k + 4 * (size_t)n;
    real *g = gbuf, *d = dbuf, *gprev = NULL, *dprev = NULL;
    int flip = 0;
    real gprev_sq = 0, theta, beta, fnew = 0;
    real fcur = obj_cg(x, p);
    size_t nfe = 1, it = 0;

    if (isnan(fcur) || isinf(fcur)) goto done;           /* :223-226 */
    if (maxiter <= 0) maxiter = INT32_MAX;

    for (it = 0; it < maxiter; it++) {
        grad_cg(x, g, p);                                 /* :231 */
        for (int i = 0; i < n; i++)                       /* :236-239 */
            d[i] = (x[i] <= 0. && g[i] >= 0.) ? 0. : -g[i];
        if (it > 0) {                                     /* :242-261 */
            theta = 0; beta = 0;
            for (int i = 0; i < n; i++) {
                theta += (x[i] <= 0.) ? 0. : g[i] * dprev[i];
                beta  += (x[i] <= 0.) ? 0. : g[i] * (g[i] - gprev[i]);
            }
            theta /= gprev_sq;
            beta /= gprev_sq;
            for (int i = 0; i < n; i++)
                d[i] += (x[i] <= 0.) ? 0. : beta * dprev[i] - theta * (g[i] - gprev[i]);
        }
        real gd = vdot(n, g, d);                          /* :264-269 */
        if (fabs(gd) <= tol) goto done;

        real smax;                                        /* :272-288 */
        if (limit_step) {
            smax = 1.;
            for (int i = 0; i < n; i++)
                if (d[i] < 0.) smax = fmin(smax, -x[i] / d[i]);
        } else {
            smax = 0.;
            for (int i = 0; i < n; i++)
                if (d[i] < 0.) smax = fmax(smax, -x[i] / d[i]);
            smax = fmin(1., 0.99 * smax);
        }

        real dsq = vdot(n, d, d);                         /* :295 */
        real step = smax;
        for (size_t ls = 0; ls < max_ls; ls++) {          /* :297-327 */
            memcpy(xnew, x, sizeof(real) * (size_t)n);
            vaxpy(n, step, d, xnew);
            if (limit_step) {
                for (int i = 0; i < n; i++) xnew[i] = (xnew[i] >= 1e-15) ? xnew[i] : 0.;
            } else {
                for (int i = 0; i < n; i++) xnew[i] = (xnew[i] > 0.) ? xnew[i] : 0.;
            }
            fnew = obj_cg(xnew, p);
            if (!isinf(fnew) && !isnan(fnew)) {
                if (fnew <= fcur - c_ls * step * dsq) {
                    memcpy(x, xnew, sizeof(real) * (size_t)n);
                    break;
                }
            }
            nfe++;
            if (nfe >= maxnfeval) goto done;
            step *= decr;
        }
        fcur = fnew;                                      /* :328 (Q4) */
        gprev_sq = vdot(n, g, g);                         /* :332 */
        dprev = d; gprev = g;                             /* :335-339 */
        flip ^= 1;
        d = dbuf + (flip ? n : 0);
        g = gbuf + (flip ? n : 0);
    }
done:
    if (niter_out) *niter_out = it;
    if (nfeval_out) *nfeval_out = nfe;
    if (f_out) *f_out = fcur;
}

/* -------------------------------------------------------------------------
 * tncg row solver — src/tnc.c (TNC 1.3 trimmed to lower bound 0), called as
 * src/poismf.c:383-391.  Restated around one state struct; the numbered
 * comments give the reference lines each block follows.
 * -----------------------------------------------------------------------*/
typedef struct {
    int n;
    const rowprob *prob;
    real *xscale, *xoffset;              /* tnc.c:361-376               */
    real *oldg, *g, *tmp, *diagb, *pk;   /* tnc.c:622-626               */
    real *sk, *yk, *sr, *yr;             /* tnc.c:627-630               */
    real *r, *v, *zk, *emat, *gv;        /* tnc.c:1220-1224             */
    real *w0, *w1, *w2;                  /* shared 3n scratch: msolve hg/hyr/hyk (:1480-1482),
                                            HVP xv (:1409), precond bsk (:1605),
                                            line search temp/tempgfull/newgfull (:1700-1702) */
    int *pivot;
    int nfeval, maxnfeval;
} tn_t;

static void tn_project(int n, real *x, const int *pivot)           /* :1015 */
{
    for (int i = 0; i < n; i++) if (pivot[i] != 0) x[i] = 0.0;
}
static void tn_coerce(int n, real *x)                               /* :466 */
{
    for (int i = 0; i < n; i++) x[i] = (x[i] < 0.) ? 0. : x[i];
}
static void tn_unscale(int n, real *x, const real *xs, const real *xo) /* :482 */
{
    for (int i = 0; i < n; i++) x[i] = x[i] * xs[i] + xo[i];
}
static void tn_scaleg(int n, real *g, const real *xs, real fscale)  /* :504 */
{
    for (int i = 0; i < n; i++) g[i] *= xs[i] * fscale;
}

/* tnc.c:1533-1575  ssbfgs / ssbfgs2 (out may alias hjv) */
static void tn_ssbfgs(int n, real gamma, const real *sj, const real *hjv,
                      const real *hjyj, real yjsj, real yjhyj, real vsj,
                      real vhyj, real *out)
{
    real beta, delta;
    if (yjsj == 0.0) { delta = 0.0; beta = 0.0; }
    else {
        delta = (gamma * yjhyj / yjsj + 1.0) * vsj / yjsj - gamma * vhyj / yjsj;
        beta = -gamma * vsj / yjsj;
    }
    for (int i = 0; i < n; i++) out[i] = gamma * hjv[i] + delta * sj[i] + beta * hjyj[i];
}

/* tnc.c:1444-1528  msolve: y = (2-step self-scaled BFGS)^-1 g */
static void tn_msolve(tn_t *s, const real *g, real *y, int upd1, real yksk,
                      real yrsr, int lreset)
{
    int n = s->n;
    if (upd1) { for (int i = 0; i < n; i++) y[i] = g[i] / s->diagb[i]; return; }
    real gsk = vdot(n, g, s->sk);
    real *hg = s->w0, *hyr = s->w1, *hyk = s->w2;
    if (lreset) {
        for (int i = 0; i < n; i++) {
            real rd = 1.0 / s->diagb[i];
            hg[i] = g[i] * rd; hyk[i] = s->yk[i] * rd;
        }
        real ykhyk = vdot(n, s->yk, hyk);
        real ghyk = vdot(n, g, hyk);
        tn_ssbfgs(n, 1.0, s->sk, hg, hyk, yksk, ykhyk, gsk, ghyk, y);
    } else {
        for (int i = 0; i < n; i++) {
            real rd = 1.0 / s->diagb[i];
            hg[i] = g[i] * rd; hyk[i] = s->yk[i] * rd; hyr[i] = s->yr[i] * rd;
        }
        real gsr = vdot(n, g, s->sr);
        real ghyr = vdot(n, g, hyr);
        real yrhyr = vdot(n, s->yr, hyr);
        tn_ssbfgs(n, 1.0, s->sr, hg, hyr, yrsr, yrhyr, gsr, ghyr, hg);
        real yksr = vdot(n, s->yk, s->sr);
        real ykhyr = vdot(n, s->yk, hyr);
        tn_ssbfgs(n, 1.0, s->sr, hyk, hyr, yrsr, yrhyr, yksr, ykhyr, hyk);
        real ykhyk = vdot(n, hyk, s->yk);
        real ghyk = vdot(n, hyk, g);
        tn_ssbfgs(n, 1.0, s->sk, hg, hyk, yksk, ykhyk, gsk, ghyk, y);
    }
}

/* tnc.c:1580-1658  initPreconditioner */
static void tn_init_precond(tn_t *s, int lreset, real yksk, real yrsr, int upd1)
{
    int n = s->n;
    real *diagb = s->diagb, *emat = s->emat, *bsk = s->w0;
    if (upd1) { memcpy(emat, diagb, sizeof(real) * (size_t)n); return; }
    real sds, srds, yrsk, td;
    if (lreset) {
        for (int i = 0; i < n; i++) bsk[i] = diagb[i] * s->sk[i];
        sds = vdot(n, s->sk, bsk);
        if (yksk == 0.0) yksk = 1.0;
        if (sds == 0.0) sds = 1.0;
        for (int i = 0; i < n; i++) {
            td = diagb[i];
            emat[i] = td - td * td * s->sk[i] * s->sk[i] / sds + s->yk[i] * s->yk[i] / yksk;
        }
    } else {
        for (int i = 0; i < n; i++) bsk[i] = diagb[i] * s->sr[i];
        sds = vdot(n, s->sr, bsk);
        srds = vdot(n, s->sk, bsk);
        yrsk = vdot(n, s->yr, s->sk);
        if (yrsr == 0.0) yrsr = 1.0;
        if (sds == 0.0) sds = 1.0;
        for (int i = 0; i < n; i++) {
            td = diagb[i];
            bsk[i] = td * s->sk[i] - bsk[i] * srds / sds + s->yr[i] * yrsk / yrsr;
            emat[i] = td - td * td * s->sr[i] * s->sr[i] / sds + s->yr[i] * s->yr[i] / yrsr;
        }
        sds = vdot(n, s->sk, bsk);
        if (yksk == 0.0) yksk = 1.0;
        if (sds == 0.0) sds = 1.0;
        for (int i = 0; i < n; i++)
            emat[i] -= bsk[i] * bsk[i] / sds + s->yk[i] * s->yk[i] / yksk;
    }
}

/* tnc.c:1388-1435  finite-difference Hessian-vector product */
static void tn_hvp(tn_t *s, const real *x, real fscale, real accuracy, real xnorm)
{
    int n = s->n;
    real *xv = s->w0, *gv = s->gv, f;
    real delta = accuracy * (xnorm + 1.0);
    for (int i = 0; i < n; i++) xv[i] = x[i] + delta * s->v[i];
    tn_unscale(n, xv, s->xscale, s->xoffset);
    tn_coerce(n, xv);
    f = obj_grad_tn(xv, gv, s->prob); (void)f;
    tn_scaleg(n, gv, s->xscale, fscale);
    real dinv = 1.0 / delta;
    for (int i = 0; i < n; i++) gv[i] = (gv[i] - s->g[i]) * dinv;
    for (int i = 0; i < n; i++) if (s->xscale[i] == 0.0) gv[i] = 0.0;   /* :1432 */
}

/* tnc.c:1162-1341  tnc_direction: preconditioned linear CG on the Newton system */
static void tn_direction(tn_t *s, real *zsol, const real *x, int maxCGit,
                         int upd1, real yksk, real yrsr, int lreset, real fscale,
                         real accuracy, real gnorm, real xnorm)
{
    int n = s->n;
    real *g = s->g, *r = s->r, *v = s->v, *zk = s->zk, *gv = s->gv;
    const int *pivot = s->pivot;
    real alpha, beta, qold, qnew, rhsnrm, tol, vgv, rz, rzold, qtest, pr, gtp;

    if (maxCGit == 0) {
        for (int i = 0; i < n; i++) zsol[i] = -g[i];
        tn_project(n, zsol, pivot);
        return;
    }
    rhsnrm = gnorm; tol = 1e-12; qold = 0.0; rzold = 0.0;
    tn_init_precond(s, lreset, yksk, yrsr, upd1);
    for (int i = 0; i < n; i++) { r[i] = -g[i]; v[i] = 0.0; zsol[i] = 0.0; }

    for (int k = 0; k < maxCGit; k++) {
        tn_project(n, r, pivot);
        tn_msolve(s, r, zk, upd1, yksk, yrsr, lreset);
        tn_project(n, zk, pivot);
        rz = vdot(n, r, zk);
        if ((rz / rhsnrm < tol) || (s->nfeval >= (s->maxnfeval - 1))) {
            if (k == 0) {
                for (int i = 0; i < n; i++) zsol[i] = -g[i];
                tn_project(n, zsol, pivot);
            }
            break;
        }
        beta = (k == 0) ? 0.0 : rz / rzold;
        for (int i = 0; i < n; i++) v[i] = zk[i] + beta * v[i];
        tn_project(n, v, pivot);
        tn_hvp(s, x, fscale, accuracy, xnorm);
        ++s->nfeval;
        tn_project(n, gv, pivot);
        vgv = vdot(n, v, gv);
        if (vgv / rhsnrm < tol) {
            if (k == 0) {
                tn_msolve(s, g, zsol, upd1, yksk, yrsr, lreset);
                for (int i = 0; i < n; i++) zsol[i] = -zsol[i];
                tn_project(n, zsol, pivot);
            }
            break;
        }
        {   /* :1347-1362 diagonalScaling(n, emat, v, gv, r) */
            real vr = 1.0 / vdot(n, v, r);
            real vgv2 = 1.0 / vdot(n, v, gv);
            for (int i = 0; i < n; i++) {
                s->emat[i] += -r[i] * r[i] * vr + gv[i] * gv[i] * vgv2;
                s->emat[i] = (s->emat[i] <= 1e-6) ? 1. : s->emat[i];
            }
        }
        alpha = rz / vgv;
        vaxpy(n, alpha, v, zsol);
        vaxpy(n, -alpha, gv, r);
        gtp = vdot(n, zsol, g);
        pr = vdot(n, r, zsol);
        qnew = (gtp + pr) * 0.5;
        qtest = (k + 1) * (1.0 - qold / qnew);
        if (qtest <= 0.5) break;
        if (gtp > 0.0) { vaxpy(n, -alpha, v, zsol); break; }
        qold = qnew;
        rzold = rz;
    }
    memcpy(s->diagb, s->emat, sizeof(real) * (size_t)n);          /* :1329 */
}

/* The Gill-Murray step-length machine, tnc.c:1822-2154 (getptcInit/getptcIter).
 * State is held in one struct instead of 17 pointer arguments. */
typedef struct {
    real reltol, abstol, tnytol, xbnd, u, fu, gu, xmin, fmin, gmin, xw, fw, gw,
         a, b, oldf, b1, scxbnd, e, step, factor, gtest1, gtest2, tol;
    int braktd;
} gm_t;
enum { GM_OK = 0, GM_EVAL = 1, GM_EINVAL = 2, GM_FAIL = 3 };

static int gm_init(gm_t *q, real eta, real rmu)                    /* :1822-1888 */
{
    if (q->u <= 0.0 || q->xbnd <= q->tnytol || q->gu > 0.0) return GM_EINVAL;
    if (q->xbnd < q->abstol) q->abstol = q->xbnd;
    q->tol = q->abstol;
    q->a = 0.0; q->xw = 0.0; q->xmin = 0.0;
    q->oldf = q->fu; q->fmin = q->fu; q->fw = q->fu;
    q->gw = q->gu; q->gmin = q->gu;
    q->step = q->u; q->factor = 5.0; q->braktd = 0;
    q->scxbnd = q->xbnd;
    q->b = q->scxbnd + q->reltol * fabs(q->scxbnd) + q->abstol;
    q->e = q->b + q->b;
    q->b1 = q->b;
    q->gtest1 = -rmu * q->gu;
    q->gtest2 = -eta * q->gu;
    if (q->step >= q->scxbnd) {
        q->step = q->scxbnd;
        q->scxbnd -= (q->reltol * fabs(q->xbnd) + q->abstol) / (1.0 + q->reltol);
    }
    q->u = q->step;
    if (fabs(q->step) < q->tol && q->step < 0.0) q->u = -(q->tol);
    if (fabs(q->step) < q->tol && q->step >= 0.0) q->u = q->tol;
    return GM_EVAL;
}

static int gm_iter(gm_t *q, real big, real rtsmll, real fpresn)   /* :1890-2154 */
{
    real abgw, absr, p, qq, r, s, scale, denom, a1, d1, d2, sumsq, abgmin,
         chordm, chordu, xmidpt, twotol;
    int convrg, skip_to_check = 0;

    if (q->fu <= q->fmin) {
        chordu = q->oldf - (q->xmin + q->u) * q->gtest1;
        if (q->fu > chordu) {
            chordm = q->oldf - q->xmin * q->gtest1;
            q->gu = -(q->gmin);
            denom = chordm - q->fmin;
            if (fabs(denom) < 1e-15) {
                denom = 1e-15;
                if (chordm - q->fmin < 0.0) denom = -denom;
            }
            if (q->xmin != 0.0) q->gu = q->gmin * (chordu - q->fu) / denom;
            q->fu = 0.5 * q->u * (q->gmin + q->gu) + q->fmin;
            if (q->fu < q->fmin) q->fu = q->fmin;
        } else {
            q->fw = q->fmin; q->fmin = q->fu;
            q->gw = q->gmin; q->gmin = q->gu;
            q->xmin += q->u; q->a -= q->u; q->b -= q->u;
            q->xw = -(q->u); q->scxbnd -= q->u;
            if (q->gu <= 0.0) q->a = 0.0;
            else { q->b = 0.0; q->braktd = 1; }
            q->tol = fabs(q->xmin) * q->reltol + q->abstol;
            skip_to_check = 1;
        }
    }
    if (!skip_to_check) {
        if (q->u < 0.0) q->a = q->u;
        else { q->b = q->u; q->braktd = 1; }
        q->xw = q->u; q->fw = q->fu; q->gw = q->gu;
    }

    twotol = q->tol + q->tol;
    xmidpt = 0.5 * (q->a + q->b);
    convrg = (fabs(xmidpt) <= twotol - 0.5 * (q->b - q->a)) ||
             (fabs(q->gmin) <= q->gtest2 && q->fmin < q->oldf &&
              ((fabs(q->xmin - q->xbnd) > q->tol) || (!q->braktd)));
    if (convrg) {
        if (q->xmin != 0.0) return GM_OK;
        if (fabs(q->oldf - q->fw) <= fpresn) return GM_FAIL;
        q->tol = 0.1 * q->tol;
        if (q->tol < q->tnytol) return GM_FAIL;
        q->reltol = 0.1 * q->reltol;
        q->abstol = 0.1 * q->abstol;
        twotol = 0.1 * twotol;
    }

    r = 0.0; qq = 0.0; s = 0.0;
    if (fabs(q->e) > q->tol) {
        int minimum_found = 0;
        r = 3.0 * (q->fmin - q->fw) / q->xw + q->gmin + q->gw;
        absr = fabs(r);
        qq = absr;
        if (q->gw != 0.0 && q->gmin != 0.0) {
            abgw = fabs(q->gw);
            abgmin = fabs(q->gmin);
            s = sqrt(abgmin) * sqrt(abgw);
            if (q->gw / abgw * q->gmin > 0.0) {
                if (r >= s || r <= -s) {
                    qq = sqrt(fabs(r + s)) * sqrt(fabs(r - s));
                } else {
                    r = 0.0; qq = 0.0;
                    minimum_found = 1;
                }
            } else {
                sumsq = 1.0; p = 0.0;
                if (absr >= s) {
                    if (absr > rtsmll) p = absr * rtsmll;
                    if (s >= p) { real value = s / absr; sumsq = 1.0 + value * value; }
                    scale = absr;
                } else {
                    if (s > rtsmll) p = s * rtsmll;
                    if (absr >= p) { real value = absr / s; sumsq = 1.0 + value * value; }
                    scale = s;
                }
                sumsq = sqrt(sumsq);
                qq = big;
                if (scale < big / sumsq) qq = scale * sumsq;
            }
        }
        if (!minimum_found) {
            if (q->xw < 0.0) qq = -qq;
            s = q->xw * (q->gmin - r - qq);
            qq = q->gw - q->gmin + qq + qq;
            if (qq > 0.0) s = -s;
            if (qq <= 0.0) qq = -qq;
            r = q->e;
            if (q->b1 != q->step || q->braktd) q->e = q->step;
        }
    }

    /* MinimumFound: */
    a1 = q->a;
    q->b1 = q->b;
    q->step = xmidpt;
    if ((!q->braktd) || ((q->a == 0.0 && q->xw < 0.0) || (q->b == 0.0 && q->xw > 0.0))) {
        if (q->braktd) {
            d1 = q->xw;
            d2 = q->a;
            if (q->a == 0.0) d2 = q->b;
            q->u = -d1 / d2;
            q->step = 5.0 * d2 * (0.1 + 1.0 / q->u) / 11.0;
            if (q->u < 1.0) q->step = 0.5 * d2 * sqrt(q->u);
        } else {
            q->step = -(q->factor) * q->xw;
            if (q->step > q->scxbnd) q->step = q->scxbnd;
            if (q->step != q->scxbnd) q->factor = 5.0 * q->factor;
        }
        if (q->step <= 0.0) a1 = q->step;
        if (q->step > 0.0) q->b1 = q->step;
    }

    if (fabs(s) <= fabs(0.5 * qq * r) || s <= qq * a1 || s >= qq * q->b1) {
        q->e = q->b - q->a;
    } else {
        q->step = s / qq;
        if (q->step - q->a < twotol || q->b - q->step < twotol) {
            if (xmidpt <= 0.0) q->step = -(q->tol);
            else q->step = q->tol;
        }
    }
    if (q->step >= q->scxbnd) {
        q->step = q->scxbnd;
        q->scxbnd -= (q->reltol * fabs(q->xbnd) + q->abstol) / (1.0 + q->reltol);
    }
    q->u = q->step;
    if (fabs(q->step) < q->tol && q->step < 0.0) q->u = -(q->tol);
    if (fabs(q->step) < q->tol && q->step >= 0.0) q->u = q->tol;
    return GM_EVAL;
}

enum { LSR_OK = 0, LSR_MAXFUN = 1, LSR_FAIL = 2 };

/* tnc.c:1664-1813  linearSearch */
static int tn_linesearch(tn_t *s, real fscale, real eta, real ftol, real xbnd,
                         const real *p, real *x, real *f, real *alpha, real *gfull)
{
    int n = s->n, itcnt = 0, itest;
    const int maxlsit = 64;
    real *temp = s->w0, *tempg = s->w1, *newg = s->w2;
    real big, fpresn, rtsmll, ualpha, pe, xnorm, rteps;
    gm_t q;

    memcpy(temp, gfull, sizeof(real) * (size_t)n);
    tn_scaleg(n, temp, s->xscale, fscale);
    q.gu = vdot(n, temp, p);
    memcpy(temp, x, sizeof(real) * (size_t)n);
    tn_project(n, temp, s->pivot);
    xnorm = vnrm2(n, temp);

    rteps = sqrt(REAL_EPS);
    pe = vnrm2(n, p) + REAL_EPS;
    q.reltol = rteps * (xnorm + 1.0) / pe;
    q.abstol = -REAL_EPS * (1.0 + fabs(*f)) / (q.gu - REAL_EPS);
    q.tnytol = REAL_EPS * (xnorm + 1.0) / pe;
    rtsmll = REAL_EPS;
    big = 1.0 / (REAL_EPS * REAL_EPS);
    fpresn = ftol;
    q.u = *alpha; q.fu = *f; q.fmin = *f; q.xbnd = xbnd;
    q.xmin = *alpha;  /* the reference aliases *alpha with getptc's xmin; init overwrites it */

    itest = gm_init(&q, eta, 1e-4);
    if (itest == GM_EVAL) *alpha = q.xmin; /* xmin was zeroed through the alias (:1849) */

    while (itest == GM_EVAL) {
        if ((++itcnt > maxlsit) || (s->nfeval >= s->maxnfeval)) break;
        ualpha = *alpha + q.u;
        for (int i = 0; i < n; i++) temp[i] = x[i] + ualpha * p[i];
        tn_unscale(n, temp, s->xscale, s->xoffset);
        tn_coerce(n, temp);
        q.fu = obj_grad_tn(temp, tempg, s->prob);
        ++s->nfeval;
        q.fu *= fscale;
        memcpy(temp, tempg, sizeof(real) * (size_t)n);
        tn_scaleg(n, temp, s->xscale, fscale);
        q.gu = vdot(n, temp, p);
        itest = gm_iter(&q, big, rtsmll, fpresn);
        *alpha = q.xmin;
        if (*alpha == ualpha) memcpy(newg, tempg, sizeof(real) * (size_t)n);
    }

    if (itest == GM_OK) {
        *f = q.fmin;
        vaxpy(n, *alpha, p, x);
        memcpy(gfull, newg, sizeof(real) * (size_t)n);
        return LSR_OK;
    } else if (itcnt > maxlsit) return LSR_FAIL;
    else if (itest != GM_EVAL) return LSR_FAIL;
    return LSR_MAXFUN;
}

/* tnc.c:251-463 (tnc) + :554-993 (tnc_minimize), with the fixed arguments of
 * src/poismf.c:383-391: messages 0, eta .25, stepmx 10, accuracy 0, fmin 0,
 * ftol 1e-4, xtol -1, pgtol -1, rescale 1.3, low = 0, up = +inf.
 * work: 20*n reals, iwork: n ints.  Returns the tnc_rc code. */
int oracle_tnc_solve(real *x, const rowprob *prob, int maxCGit, int maxnfeval,
                     real *work, int *iwork, real *gfull, real *f_out,
                     int *nfeval_out, int *niter_out)
{
    int n = prob->k, rc, niter = 0;
    real eta = 0.25, stepmx = 10., accuracy = 0., fmin = 0., ftol = 1e-4,
         xtol = -1., pgtol = -1., rescale = 1.3;
    real fscale, rteps, f;
    tn_t S; tn_t *s = &S;
    s->n = n; s->prob = prob; s->pivot = iwork; s->nfeval = 0; s->maxnfeval = maxnfeval;
    {
        real *w = work;
        s->xscale = w; w += n; s->xoffset = w; w += n;
        s->oldg = w; w += n; s->g = w; w += n; s->tmp = w; w += n; s->diagb = w; w += n;
        s->pk = w; w += n; s->sk = w; w += n; s->yk = w; w += n; s->sr = w; w += n;
        s->yr = w; w += n; s->r = w; w += n; s->v = w; w += n; s->zk = w; w += n;
        s->emat = w; w += n; s->gv = w; w += n; s->w0 = w; w += n; s->w1 = w; w += n;
        s->w2 = w; w += n;
    }
    tn_coerce(n, x);                                               /* :323 */
    if (maxnfeval < 1) { rc = 3; goto finish_noscale; }            /* :325 */
    f = obj_grad_tn(x, gfull, prob);                               /* :341 */
    s->nfeval++;
    fscale = 1.0;
    for (int i = 0; i < n; i++) {                                  /* :393-394 */
        s->xscale[i] = 1.0 + fabs(x[i]);
        s->xoffset[i] = x[i];
    }
    rteps = sqrt(REAL_EPS);                                        /* :402-436 */
    if (stepmx < rteps * 10.0) stepmx = 1.0e1;
    if (eta < 0.0 || eta >= 1.0) eta = 0.25;
    if (rescale < 0) rescale = 1.3;
    if (maxCGit < 0) { maxCGit = n / 2; if (maxCGit < 1) maxCGit = 1; else if (maxCGit > 50) maxCGit = 50; }
    if (maxCGit > n) maxCGit = n;
    if (accuracy <= REAL_EPS) accuracy = rteps;
    if (ftol < 0.0) ftol = accuracy;
    if (pgtol < 0.0) pgtol = 1e-2 * sqrt(accuracy);
    if (xtol < 0.0) xtol = rteps;

    /* ---- tnc_minimize (:554-993) ---- */
    {
        real fLastReset, difnew, epsred, oldgtp, difold, oldf, xnorm, newscale,
             gnorm, ustpmax, fLastConstraint, spe, yrsr, yksk, alpha = 0.0;
        real *g = s->g, *temp = s->tmp, *pk = s->pk;
        int *pivot = s->pivot;
        int icycle, oldnfeval, lreset, newcon, upd1, remcon;

        difnew = 0.0; epsred = 0.05; upd1 = 1; icycle = n - 1; newcon = 1;
        lreset = 0; yrsr = 0.0; yksk = 0.0;

        for (int i = 0; i < n; i++)                                /* scalex :492 */
            if (s->xscale[i] > 0.0) x[i] = (x[i] - s->xoffset[i]) / s->xscale[i];
        f *= fscale;

        for (int i = 0; i < n; i++) {                              /* setConstraints :513 (low = 0) */
            if (s->xscale[i] == 0.0) pivot[i] = 2;
            else if (x[i] * s->xscale[i] + s->xoffset[i] - LOW <= REAL_EPS * 10.0 * (fabs(LOW) + 1.0))
                pivot[i] = -1;
            else pivot[i] = 0;
        }
        memcpy(g, gfull, sizeof(real) * (size_t)n);
        tn_scaleg(n, g, s->xscale, fscale);
        for (int i = 0; i < n; i++)                                /* :670-674 */
            if (-pivot[i] * g[i] < 0.0) pivot[i] = 0;
        tn_project(n, g, pivot);
        gnorm = vnrm2(n, g);
        fLastConstraint = f; fLastReset = f;
        for (int i = 0; i < n; i++) s->diagb[i] = 1.0;

        for (;;) {
            if (vnrm2(n, g) <= pgtol * fscale) {                   /* :700 */
                memcpy(g, gfull, sizeof(real) * (size_t)n);
                tn_project(n, g, pivot);
                rc = 0; break;
            }
            if (s->nfeval >= maxnfeval) { rc = 3; break; }         /* :715 */

            newscale = vnrm2(n, g);                                /* :721-746 */
            if ((newscale > REAL_EPS) && (fabs(log10(newscale)) > rescale)) {
                newscale = 1.0 / newscale;
                f *= newscale; fscale *= newscale; gnorm *= newscale;
                fLastConstraint *= newscale; fLastReset *= newscale; difnew *= newscale;
                for (int i = 0; i < n; i++) g[i] *= newscale;
                for (int i = 0; i < n; i++) s->diagb[i] = 1.0;
                upd1 = 1; icycle = n - 1; newcon = 1;
            }

            memcpy(temp, x, sizeof(real) * (size_t)n);             /* :748-751 */
            tn_project(n, temp, pivot);
            xnorm = vnrm2(n, temp);
            oldnfeval = s->nfeval;

            tn_direction(s, pk, x, maxCGit, upd1, yksk, yrsr, lreset, fscale,
                         accuracy, gnorm, xnorm);                  /* :754 */

            if (!newcon) {                                         /* :770-785 */
                if (!lreset) {
                    vaxpy(n, 1., s->sk, s->sr);
                    vaxpy(n, 1., s->yk, s->yr);
                    icycle++;
                } else {
                    memcpy(s->sr, s->sk, sizeof(real) * (size_t)n);
                    memcpy(s->yr, s->yk, sizeof(real) * (size_t)n);
                    fLastReset = f;
                    icycle = 1;
                }
            }
            memcpy(s->oldg, g, sizeof(real) * (size_t)n);          /* :787-789 */
            oldf = f;
            oldgtp = vdot(n, pk, g);
            ustpmax = stepmx / (vnrm2(n, pk) + REAL_EPS);          /* :792 */

            spe = ustpmax;                                         /* stepMax :1041 (low=0, up=inf) */
            for (int i = 0; i < n; i++) {
                if ((pivot[i] == 0) && (pk[i] != 0.0)) {
                    real t;
                    if (pk[i] < 0.0) {
                        t = (LOW - s->xoffset[i]) / s->xscale[i] - x[i];
                        if (t > spe * pk[i]) spe = t / pk[i];
                    } else {
                        t = (REAL_HUGE - s->xoffset[i]) / s->xscale[i] - x[i];
                        if (t < spe * pk[i]) spe = t / pk[i];
                    }
                }
            }

            if (spe > 0.0) {
                int lsrc;
                {   /* initialStep :1368 */
                    real fm = fmin / fscale;
                    real d = fabs(f - fm);
                    alpha = 1.0;
                    if (d * 2.0 <= -(oldgtp) && d >= REAL_EPS) alpha = d * -2.0 / oldgtp;
                    if (alpha >= spe) alpha = spe;
                }
                lsrc = tn_linesearch(s, fscale, eta, ftol, spe, pk, x, &f, &alpha, gfull);
                if (lsrc == LSR_FAIL) { rc = 4; break; }           /* :818 */
                if (alpha >= 0.9 * ustpmax) stepmx *= 1e2;         /* :824 */
                if (alpha - spe >= -REAL_EPS * 10.0) newcon = 1;   /* :833 */
                else {
                    if (lsrc != LSR_OK) { rc = (lsrc == LSR_MAXFUN) ? 3 : 4; break; }
                    newcon = 0;
                }
            } else newcon = 1;

            if (newcon) {                                          /* :855-863 addConstraint :1072 */
                int added = 0;
                for (int i = 0; i < n; i++) {
                    if ((pivot[i] == 0) && (pk[i] != 0.0)) {
                        if (pk[i] < 0.0) {
                            real tolc = REAL_EPS * 10.0 * (fabs(LOW) + 1.0);
                            if (x[i] * s->xscale[i] + s->xoffset[i] - LOW <= tolc) {
                                pivot[i] = -1;
                                x[i] = (LOW - s->xoffset[i]) / s->xscale[i];
                                added = 1;
                            }
                        }
                    }
                }
                if (!added) {
                    if (s->nfeval == oldnfeval) { rc = 6; break; }
                }
                fLastConstraint = f;
            }
            niter++;

            difold = difnew;                                       /* :875-887 */
            difnew = oldf - f;
            if (icycle == 1) {
                if (difnew > difold * 2.0) epsred += epsred;
                if (difnew < difold * 0.5) epsred *= 0.5;
            }
            memcpy(g, gfull, sizeof(real) * (size_t)n);            /* :889-894 */
            tn_scaleg(n, g, s->xscale, fscale);
            memcpy(temp, g, sizeof(real) * (size_t)n);
            tn_project(n, temp, pivot);
            gnorm = vnrm2(n, temp);

            {   /* removeConstraint :1113 */
                real pgtolfs = pgtol * fscale;
                remcon = 0;
                if (!(((fLastConstraint - f) <= (oldgtp * -0.5)) && (gnorm > pgtolfs))) {
                    int imax = -1; real cmax = 0.0;
                    for (int i = 0; i < n; i++) {
                        if (pivot[i] == 2) continue;
                        real t = -pivot[i] * g[i];
                        if (t < cmax) { cmax = t; imax = i; }
                    }
                    if (imax != -1) { pivot[imax] = 0; remcon = 1; }
                }
            }
            if (remcon) {                                          /* :901-907 */
                memcpy(temp, g, sizeof(real) * (size_t)n);
                tn_project(n, temp, pivot);
                gnorm = vnrm2(n, temp);
                fLastConstraint = f;
            }
            if (!remcon && !newcon) {                              /* :909-929 */
                if (fabs(difnew) <= ftol * fscale) { rc = 1; break; }
                if (alpha * vnrm2(n, pk) <= xtol) { rc = 2; break; }
            }
            tn_project(n, g, pivot);                               /* :931 */

            if (!newcon) {                                         /* :940-962 */
                for (int i = 0; i < n; i++) {
                    s->yk[i] = g[i] - s->oldg[i];
                    s->sk[i] = alpha * pk[i];
                }
                yksk = vdot(n, s->yk, s->sk);
                if (icycle == (n - 1) || difnew < epsred * (fLastReset - f)) lreset = 1;
                else {
                    yrsr = vdot(n, s->yr, s->sr);
                    lreset = (yrsr <= 0.0) ? 1 : 0;
                }
                upd1 = 0;
            }
        }
        tn_unscale(n, x, s->xscale, s->xoffset);                   /* :971-973 */
        tn_coerce(n, x);
        f /= fscale;
    }
    if (f_out) *f_out = f;
    if (nfeval_out) *nfeval_out = s->nfeval;
    if (niter_out) *niter_out = niter;
    return rc;
finish_noscale:
    if (nfeval_out) *nfeval_out = s->nfeval;
    if (niter_out) *niter_out = 0;
    return rc;
}

/* -------------------------------------------------------------------------
 * Sweep driver — src/poismf.c:435-632 (signal handling omitted: test infra)
 * -----------------------------------------------------------------------*/
static void colsum(real *out, const real *M, size_t nrow, size_t ncol)  /* :77-83 */
{
    memset(out, 0, sizeof(real) * ncol);
    for (size_t r = 0; r < nrow; r++)
        for (size_t c = 0; c < ncol; c++) out[c] += M[r * ncol + c];
}

/* :85-123  per-row weighted sums: (w-1)*sum_{j in row} F_j + csum */
static void weighted_sums(const real *F, const real *csum, real *out,
                          const ix_t *ind, const ix_t *ptr, size_t dim, size_t k, real w)
{
    memset(out, 0, dim * k * sizeof(real));
    for (size_t r = 0; r < dim; r++)
        for (ix_t t = ptr[r]; t < ptr[r + 1]; t++)
            vaxpy((int)k, 1., F + ind[t] * k, out + r * k);
    real nw = w - 1.;
    for (size_t i = 0; i < dim * k; i++) out[i] *= nw;
    for (size_t r = 0; r < dim; r++) vaxpy((int)k, 1., csum, out + r * k);
}

/* One half-sweep: update every row of `M` (dim rows) holding `F` fixed. */
static int half_sweep(int method, real *M, const real *F, const real *xv,
                      const ix_t *ptr, const ix_t *ind, size_t dim, size_t k,
                      const real *csum, const real *csum_w, real l2, real w,
                      real step, real cdiv, size_t maxupd, int limit_step,
                      int reuse_prev, int early_stop, real *buf, int *ibuf)
{
    size_t n_unchanged = 0;
    for (size_t r = 0; r < dim; r++) {
        real *row = M + r * k;
        ix_t nz = ptr[r + 1] - ptr[r];
        if (nz == 0) { memset(row, 0, k * sizeof(real)); continue; }   /* Q6 */
        rowprob p = { F, (w != 1.) ? csum_w + r * k : csum, xv + ptr[r], ind + ptr[r],
                      nz, l2, w, (int)k };
        if (method == 3) {                                  /* pg  :139-188 */
            pg_row(row, &p, p.csum, step * w, cdiv, maxupd, buf);
        } else if (method == 2) {                           /* cg  :275-322 */
            oracle_cg_solve(row, &p, maxupd, limit_step, buf, NULL, NULL, NULL);
        } else {                                            /* tncg :324-404 */
            real *prev = buf + 21 * k, *gfull = buf + 20 * k, fval;
            int maxCGit = (int)fmax(1., fmin(50., (real)k / 2.));
            if (early_stop) memcpy(prev, row, k * sizeof(real));
            if (!reuse_prev) for (size_t i = 0; i < k; i++) row[i] = 1e-3;
            oracle_tnc_solve(row, &p, maxCGit, (int)maxupd, buf, ibuf, gfull, &fval, NULL, NULL);
            if (early_stop) {
                vaxpy((int)k, -1., row, prev);
                n_unchanged += vdot((int)k, prev, prev) <= 1e-4;
            }
        }
    }
    if (method == 1 && early_stop)
        return ((double)n_unchanged / (double)dim) >= .95;
    return 0;
}

/* Same argument meaning and return codes as run_poismf (src/poismf.h:226-233);
 * method: 1 tncg, 2 cg, 3 pg. */
int oracle_run_poismf(real *A, const real *Xr, const ix_t *Xr_indptr, const ix_t *Xr_indices,
                      real *B, const real *Xc, const ix_t *Xc_indptr, const ix_t *Xc_indices,
                      size_t dimA, size_t dimB, size_t k,
                      real l2_reg, real l1_reg, real w_mult, real step_size,
                      int method, int limit_step, size_t numiter, size_t maxupd,
                      int early_stop, int reuse_prev)
{
    size_t dmax = dimA > dimB ? dimA : dimB;
    real *csum = (real *)malloc(sizeof(real) * k);
    real *buf = (real *)malloc(sizeof(real) * 22 * k);
    int *ibuf = (int *)malloc(sizeof(int) * k);
    real *csum_w = (w_mult != 1.) ? (real *)malloc(sizeof(real) * k * dmax) : NULL;
    real cdiv, neg_step = -step_size;
    int stopA = 0, stopB = 0;
    if (!csum || !buf || !ibuf || (w_mult != 1. && !csum_w)) {
        free(csum); free(buf); free(ibuf); free(csum_w); return 1;
    }
    for (size_t it = 0; it < numiter; it++) {
        cdiv = 1. / (1. + 2. * l2_reg * step_size);                  /* :511 */
        /* ---- B half-sweep over CSC (:512-557) ---- */
        colsum(csum, A, dimA, k);
        if (l1_reg > 0.) for (size_t c = 0; c < k; c++) csum[c] += l1_reg;
        if (w_mult != 1.) weighted_sums(A, csum, csum_w, Xc_indices, Xc_indptr, dimB, k, w_mult);
        if (method == 3) {
            if (w_mult == 1.) vscal((int)k, neg_step, csum);
            else for (size_t i = 0; i < dimB * k; i++) csum_w[i] *= neg_step;
        }
        if (!(method == 1 && stopB))
            stopB = half_sweep(method, B, A, Xc, Xc_indptr, Xc_indices, dimB, k, csum, csum_w,
                               l2_reg, w_mult, step_size, cdiv, maxupd, limit_step,
                               reuse_prev, early_stop && method == 1, buf, ibuf);
        if (method == 3) { step_size *= 0.5; neg_step = -step_size; }   /* :532 (Q2) */
        /* ---- A half-sweep over CSR (:562-604) ---- */
        colsum(csum, B, dimB, k);
        if (l1_reg > 0.) for (size_t c = 0; c < k; c++) csum[c] += l1_reg;
        if (w_mult != 1.) weighted_sums(B, csum, csum_w, Xr_indices, Xr_indptr, dimA, k, w_mult);
        if (method == 3) {
            if (w_mult == 1.) vscal((int)k, neg_step, csum);
            else for (size_t i = 0; i < dimA * k; i++) csum_w[i] *= neg_step;
            vscal((int)k, neg_step, csum);                             /* :577 (Q1) */
        }
        if (!(method == 1 && stopA))
            stopA = half_sweep(method, A, B, Xr, Xr_indptr, Xr_indices, dimA, k, csum, csum_w,
                               l2_reg, w_mult, step_size, cdiv, maxupd, limit_step,
                               reuse_prev, early_stop && method == 1, buf, ibuf);
        if (stopA && stopB) break;                                    /* :606 */
    }
    free(csum); free(buf); free(ibuf); free(csum_w);
    return 0;
}


/* One half-sweep in isolation (what the sharding layer drives): update every row of M
 * with F fixed, including the column-sum preparation of src/poismf.c:511-526 (B side,
 * is_A_side=0) or :562-577 (A side, is_A_side=1: pg scales the sums twice, Q1).
 * `step_size` is the CURRENT pg step (already halved for the A side by the caller). */
int oracle_half_sweep(int method, real *M, const real *F, const real *xv, const ix_t *ptr,
                      const ix_t *ind, size_t dim, size_t other, size_t k,
                      real l2_reg, real l1_reg, real w_mult, real step_size, int is_A_side,
                      size_t maxupd, int limit_step, int reuse_prev)
{
    real *csum = (real *)malloc(sizeof(real) * k);
    real *buf = (real *)malloc(sizeof(real) * 22 * k);
    int *ibuf = (int *)malloc(sizeof(int) * k);
    real *csum_w = (w_mult != 1.) ? (real *)malloc(sizeof(real) * k * dim) : NULL;
    real neg_step = -step_size;
    /* cnst_div uses the step of the START of the sweep (:511): un-halve it on the A side */
    real step0 = (method == 3 && is_A_side) ? (real)(step_size * 2.) : step_size;
    real cdiv = 1. / (1. + 2. * l2_reg * step0);
    colsum(csum, F, other, k);
    if (l1_reg > 0.) for (size_t c = 0; c < k; c++) csum[c] += l1_reg;
    if (w_mult != 1.) weighted_sums(F, csum, csum_w, ind, ptr, dim, k, w_mult);
    if (method == 3) {
        if (w_mult == 1.) vscal((int)k, neg_step, csum);
        else for (size_t i = 0; i < dim * k; i++) csum_w[i] *= neg_step;
        if (is_A_side) vscal((int)k, neg_step, csum);
    }
    half_sweep(method, M, F, xv, ptr, ind, dim, k, csum, csum_w, l2_reg, w_mult, step_size, cdiv,
               maxupd, limit_step, reuse_prev, 0, buf, ibuf);
    free(csum); free(buf); free(ibuf); free(csum_w);
    return 0;
}


/* src/pred.c:66-199  factors_multiple: factors for new rows, B and Bsum (already +l1) fixed.
 * pg: niter proximal steps with halving step size, Bsum scaled by -step ONCE per iteration
 * (w==1), while for w!=1 the weighted sums are scaled by -step0 at :126 and AGAIN by the current
 * -step at :160 (kept as is); cg: one solve with maxupd*niter iterations; tncg: one solve,
 * starting from Amean iff reuse_mean. */
int oracle_factors_multiple(real *A, const real *B, const real *Bsum, const real *Amean,
                            const real *Xr, const ix_t *Xr_indptr, const ix_t *Xr_indices,
                            int k_int, size_t dimA, real l2_reg, real w_mult, real step_size,
                            size_t niter, size_t maxupd, int method, int limit_step, int reuse_mean)
{
    size_t k = (size_t)k_int;
    real *buf = (real *)malloc(sizeof(real) * 22 * k);
    int *ibuf = (int *)malloc(sizeof(int) * k);
    real *scaled = (real *)malloc(sizeof(real) * k);
    real *csum_w = NULL, *csum_w_scaled = NULL;
    if (w_mult != 1.) {
        csum_w = (real *)malloc(sizeof(real) * k * dimA);
        weighted_sums(B, Bsum, csum_w, Xr_indices, Xr_indptr, dimA, k, w_mult);
        if (method == 3) {
            for (size_t i = 0; i < dimA * k; i++) csum_w[i] *= -step_size;           /* :126 */
            csum_w_scaled = (real *)malloc(sizeof(real) * k * dimA);
        }
    }
    if (reuse_mean || method != 1)                                                  /* :144-147 */
        for (size_t r = 0; r < dimA; r++) memcpy(A + r * k, Amean, k * sizeof(real));
    if (method == 3) {
        for (size_t it = 0; it < niter; it++) {                                     /* :152-167 */
            if (w_mult == 1.) { memcpy(scaled, Bsum, sizeof(real) * k); vscal(k_int, -step_size, scaled); }
            else { memcpy(csum_w_scaled, csum_w, sizeof(real) * k * dimA);
                   for (size_t i = 0; i < dimA * k; i++) csum_w_scaled[i] *= -step_size; }
            real cdiv = 1. / (1. + 2. * l2_reg * step_size);
            half_sweep(3, A, B, Xr, Xr_indptr, Xr_indices, dimA, k, scaled, csum_w_scaled, l2_reg, w_mult,
                       step_size, cdiv, maxupd, 0, 0, 0, buf, ibuf);
            step_size *= 0.5;
        }
    } else if (method == 2) {                                                       /* :171-178 */
        half_sweep(2, A, B, Xr, Xr_indptr, Xr_indices, dimA, k, Bsum, csum_w, l2_reg, w_mult,
                   step_size, 0, maxupd * niter, limit_step, 0, 0, buf, ibuf);
    } else {                                                                        /* :180-188 */
        half_sweep(1, A, B, Xr, Xr_indptr, Xr_indices, dimA, k, Bsum, csum_w, l2_reg, w_mult,
                   step_size, 0, maxupd, 0, reuse_mean, 0, buf, ibuf);
    }
    free(buf); free(ibuf); free(scaled); free(csum_w); free(csum_w_scaled);
    return 0;
}

/* src/pred.c:201-304  factors_single: one new row by tncg.  Bsum carries the OLD l1 already; the
 * difference l1_new - l1_old is added only when positive (:218, :254-257), AFTER the w_mult
 * adjustment (:241-247); an empty row gives zeros (:212-215); start = Amean iff reuse_mean, else 1e-3. */
int oracle_factors_single(real *out, size_t k, const real *Amean, int reuse_mean,
                          const real *X, const ix_t *X_ind, size_t nnz, const real *B, const real *Bsum,
                          int maxupd, real l2_reg, real l1_new, real l1_old, real w_mult)
{
    if (nnz == 0) { memset(out, 0, k * sizeof(real)); return 0; }
    real l1_reg = l1_new - l1_old;
    real *pass = (real *)malloc(sizeof(real) * k);
    if (w_mult != 1.) {
        memset(pass, 0, sizeof(real) * k);
        for (size_t t = 0; t < nnz; t++) vaxpy((int)k, 1., B + X_ind[t] * k, pass);
        vscal((int)k, w_mult - 1., pass);
        vaxpy((int)k, 1., Bsum, pass);
    } else {
        memcpy(pass, Bsum, sizeof(real) * k);
    }
    if (l1_reg > 0.) for (size_t i = 0; i < k; i++) pass[i] += l1_reg;
    if (reuse_mean) memcpy(out, Amean, k * sizeof(real));
    else for (size_t i = 0; i < k; i++) out[i] = 1e-3;
    rowprob p = { B, pass, X, X_ind, (ix_t)nnz, l2_reg, w_mult, (int)k };
    real *buf = (real *)malloc(sizeof(real) * 22 * k);
    int *ibuf = (int *)malloc(sizeof(int) * k);
    int maxCGit = (int)k / 2;                                   /* tnc.c:413-421 (maxCGit = -1) */
    if (maxCGit < 1) maxCGit = 1; else if (maxCGit > 50) maxCGit = 50;
    real fval;
    oracle_tnc_solve(out, &p, maxCGit, maxupd, buf, ibuf, buf + 20 * k, &fval, NULL, NULL);
    free(buf); free(ibuf); free(pass);
    return 0;
}

/* Single-row entry points for known-answer tests (cg / tncg solvers alone). */
void oracle_cg_row(real *a, const real *F, const real *csum, const real *xval,
                   const ix_t *xind, ix_t nnz, int k, real l2, real w,
                   size_t maxupd, int limit_step, size_t *niter, size_t *nfeval, real *fval)
{
    rowprob p = { F, csum, xval, xind, nnz, l2, w, k };
    real *buf = (real *)malloc(sizeof(real) * 5 * (size_t)k);
    oracle_cg_solve(a, &p, maxupd, limit_step, buf, niter, nfeval, fval);
    free(buf);
}
int oracle_tnc_row(real *a, const real *F, const real *csum, const real *xval,
                   const ix_t *xind, ix_t nnz, int k, real l2, real w,
                   int maxupd, int *niter, int *nfeval, real *fval)
{
    rowprob p = { F, csum, xval, xind, nnz, l2, w, k };
    real *buf = (real *)malloc(sizeof(real) * 22 * (size_t)k);
    int *ibuf = (int *)malloc(sizeof(int) * (size_t)k);
    int maxCGit = (int)fmax(1., fmin(50., (real)k / 2.));
    int rc = oracle_tnc_solve(a, &p, maxCGit, maxupd, buf, ibuf, buf + 20 * (size_t)k, fval, nfeval, niter);
    free(buf); free(ibuf);
    return rc;
}

/* src/pred.c:42-64  predict_multiple */
void oracle_predict_multiple(real *out, const real *A, const real *B,
                             const ix_t *ixA, const ix_t *ixB, size_t n, int k)
{
    for (size_t i = 0; i < n; i++)
        out[i] = vdot(k, A + ixA[i] * (size_t)k, B + ixB[i] * (size_t)k);
}

/* src/topN.c:112-284  topN for one user.  Scores are the sequential dots of the
 * reference's three scoring branches (all equal to <a, B_j>); the ranking is
 * "descending score", ties in unspecified order in the reference (qsort), here
 * broken by ascending item id.  Return codes as the reference: 0 ok, 2 invalid. */
typedef struct { real s; ix_t j; } scored;
static int cmp_scored(const void *pa, const void *pb)
{
    const scored *a = (const scored *)pa, *b = (const scored *)pb;
    if (a->s != b->s) return (a->s < b->s) ? 1 : -1;
    return (a->j > b->j) - (a->j < b->j);
}
int oracle_topN(const real *a_vec, const real *B, int k,
                const ix_t *include_ix, size_t n_include,
                const ix_t *exclude_ix, size_t n_exclude,
                ix_t *outp_ix, real *outp_score, size_t n_top, size_t n)
{
    if (n_include == 0) include_ix = NULL;
    if (n_exclude == 0) exclude_ix = NULL;
    if (include_ix && exclude_ix) return 2;                          /* :124-128 */
    if (n_top == 0) return 2;
    if (n_exclude > n - n_top) return 2;
    if (n_include > n) return 2;
    size_t n_take = include_ix ? n_include : n - n_exclude, m = 0;
    scored *c = (scored *)malloc(sizeof(scored) * (n_take ? n_take : 1));
    char *mask = NULL;
    if (!c) return 1;
    if (include_ix) {
        for (size_t t = 0; t < n_include; t++) { c[m].j = include_ix[t]; m++; }
    } else {
        mask = (char *)calloc(n, 1);
        if (!mask) { free(c); return 1; }
        for (size_t t = 0; t < n_exclude; t++) mask[exclude_ix[t]] = 1;
        for (size_t j = 0; j < n; j++) if (!mask[j]) { if (m < n_take) c[m].j = j; m++; }
        free(mask);
        if (m != n_take) { free(c); return 2; }   /* duplicate ids in exclude list */
    }
    for (size_t t = 0; t < n_take; t++) c[t].s = vdot(k, a_vec, B + c[t].j * (size_t)k);
    qsort(c, n_take, sizeof(scored), cmp_scored);
    for (size_t t = 0; t < n_top && t < n_take; t++) {
        outp_ix[t] = c[t].j;
        if (outp_score) outp_score[t] = c[t].s;
    }
    free(c);
    return 0;
}

/* Poisson log-likelihood used by the harness (the reference declares eval_llk,
 * src/poismf.h:258-269, but never defines it).  Always in double:
 *   llk = sum_nz x*log<a_i,b_j> - <colsum(A), colsum(B)>             */
double oracle_llk(const real *A, const real *B, const real *Xr, const ix_t *Xr_indptr,
                  const ix_t *Xr_indices, size_t dimA, size_t dimB, size_t k)
{
    double acc = 0;
    for (size_t r = 0; r < dimA; r++)
        for (ix_t t = Xr_indptr[r]; t < Xr_indptr[r + 1]; t++) {
            double d = 0;
            for (size_t c = 0; c < k; c++) d += (double)A[r * k + c] * (double)B[Xr_indices[t] * k + c];
            acc += (double)Xr[t] * log(d);
        }
    double cross = 0;
    for (size_t c = 0; c < k; c++) {
        double sa = 0, sb = 0;
        for (size_t r = 0; r < dimA; r++) sa += A[r * k + c];
        for (size_t r = 0; r < dimB; r++) sb += B[r * k + c];
        cross += sa * sb;
    }
    return acc - cross;
}
