/* poismf_b200 — host drop-in layer.
 *
 * Defines the reference's own C entry points with the reference's own prototypes
 * (/root/reference/src/poismf.h:170, :226-233, :240-247, :250-257, :270-289) and forwards
 * them to the CUDA library through the C ABI of include/poismf_b200.h.  It takes
 * the place of src/poismf.c, src/pred.c and src/topN.c in a wrapper build
 * (INTEGRATION.md); like them it is compiled once per value type:
 *
 *     (default)        real_t = double, sparse_ix = size_t   (Python cfuns_double)
 *     -DUSE_FLOAT      real_t = float,  sparse_ix = size_t   (Python cfuns_float)
 *     -DPMF_INDEX_INT  real_t = double, sparse_ix = int      (R, -D_FOR_R)
 *
 * `nthreads` is accepted and ignored (the device decides its own parallelism).
 * There is no CPU fallback: without a CUDA device run_poismf returns 1.
 */
#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>
#include "poismf_b200.h"

#ifdef USE_FLOAT
typedef float real_t;
#define PMF_DTYPE PMF_F32
#else
typedef double real_t;
#define PMF_DTYPE PMF_F64
#endif
#if defined(PMF_INDEX_INT) || defined(_FOR_R)
typedef int sparse_ix;
#else
typedef size_t sparse_ix;
#endif
#define PMF_IXB ((int)sizeof(sparse_ix))

typedef enum Method { tncg = 1, cg = 2, pg = 3 } Method;   /* src/poismf.h:225 */

/* src/poismf.c:55-62.  The GPU path has no OpenMP; report "true" so that callers
 * which only use it to warn about single-threaded builds stay quiet. */
bool get_has_openmp(void) { return true; }

/* src/poismf.c:435-632 */
int run_poismf(
    real_t *restrict A, real_t *restrict Xr, sparse_ix *restrict Xr_indptr, sparse_ix *restrict Xr_indices,
    real_t *restrict B, real_t *restrict Xc, sparse_ix *restrict Xc_indptr, sparse_ix *restrict Xc_indices,
    const size_t dimA, const size_t dimB, const size_t k,
    const real_t l2_reg, const real_t l1_reg, const real_t w_mult, real_t step_size,
    const Method method, const bool limit_step, const size_t numiter, const size_t maxupd,
    const bool early_stop, const bool reuse_prev,
    const bool handle_interrupt, const int nthreads)
{
    (void)nthreads;
    return pmf_b200_run_poismf(PMF_DTYPE, PMF_IXB, A, Xr, Xr_indptr, Xr_indices, B, Xc, Xc_indptr, Xc_indices,
                               dimA, dimB, k, (double)l2_reg, (double)l1_reg, (double)w_mult, (double)step_size,
                               (int)method, (int)limit_step, numiter, maxupd, (int)early_stop, (int)reuse_prev,
                               (int)handle_interrupt, 0);
}

#ifndef PMF_NO_PREDICT   /* define when src/pred.c is kept whole in the wrapper build */
/* src/pred.c:42-64.  The reference takes no dimensions (callers validate the ids,
 * poismf/__init__.py:815); the device copy needs them, so they are recovered from
 * the ids themselves: only rows up to the largest id are uploaded. */
void predict_multiple(
    real_t *restrict out,
    real_t *restrict A, real_t *restrict B,
    sparse_ix *ixA, sparse_ix *ixB,
    size_t n, int k,
    int nthreads)
{
    (void)nthreads;
    size_t dimA = 0, dimB = 0;
    for (size_t i = 0; i < n; i++) {
        if ((size_t)ixA[i] + 1 > dimA) dimA = (size_t)ixA[i] + 1;
        if ((size_t)ixB[i] + 1 > dimB) dimB = (size_t)ixB[i] + 1;
    }
    if (n == 0) return;
    (void)pmf_b200_predict_multiple(PMF_DTYPE, PMF_IXB, out, A, B, ixA, ixB, n, k, dimA, dimB);
}
#endif

#ifndef PMF_NO_FACTORS   /* define when src/pred.c is kept whole in the wrapper build */
/* src/pred.c:66-199.  B's row count is not an argument of the reference: it is recovered from the
 * largest item id among the new rows (rows of B beyond it are never read). */
int factors_multiple(
    real_t *A, real_t *B,
    real_t *Bsum, real_t *Amean,
    real_t *Xr, sparse_ix *Xr_indptr, sparse_ix *Xr_indices,
    int k, size_t dimA,
    real_t l2_reg, real_t w_mult,
    real_t step_size, size_t niter, size_t maxupd,
    Method method, bool limit_step, bool reuse_mean,
    int nthreads)
{
    (void)nthreads;
    size_t dimB = 0;
    const size_t nnz = (size_t)Xr_indptr[dimA] - (size_t)Xr_indptr[0];
    for (size_t i = 0; i < nnz; i++)
        if ((size_t)Xr_indices[i] + 1 > dimB) dimB = (size_t)Xr_indices[i] + 1;
    if (dimB == 0) dimB = 1;
    return pmf_b200_factors_multiple(PMF_DTYPE, PMF_IXB, A, B, Bsum, Amean, Xr, Xr_indptr, Xr_indices, k, dimA, dimB,
                                     (double)l2_reg, (double)w_mult, (double)step_size, niter, maxupd,
                                     (int)method, (int)limit_step, (int)reuse_mean, 0);
}

/* src/pred.c:201-304.  Single-row inference is a latency path: a wrapper that serves one user at a
 * time may prefer to keep the reference's CPU factors_single (compile with -DPMF_NO_FACTORS and keep
 * src/pred.c, src/tnc.c); this one runs the same tncg solve on the device. */
int factors_single(
    real_t *restrict out, size_t k,
    real_t *restrict Amean, bool reuse_mean,
    real_t *restrict X, sparse_ix X_ind[], size_t nnz,
    real_t *restrict B, real_t *restrict Bsum,
    int maxupd, real_t l2_reg, real_t l1_new, real_t l1_old,
    real_t w_mult)
{
    return pmf_b200_factors_single(PMF_DTYPE, PMF_IXB, out, k, Amean, (int)reuse_mean, X, X_ind, nnz, B, Bsum,
                                   maxupd, (double)l2_reg, (double)l1_new, (double)l1_old, (double)w_mult, 0);
}
#endif

/* src/topN.c:112-284 */
int topN(
    real_t *restrict a_vec, real_t *restrict B, int k,
    sparse_ix *restrict include_ix, size_t n_include,
    sparse_ix *restrict exclude_ix, size_t n_exclude,
    sparse_ix *restrict outp_ix, real_t *restrict outp_score,
    size_t n_top, size_t n, int nthreads)
{
    (void)nthreads;
    return pmf_b200_topN(PMF_DTYPE, PMF_IXB, a_vec, B, k, include_ix, n_include, exclude_ix, n_exclude,
                         outp_ix, outp_score, n_top, n);
}
