/* TEST INFRASTRUCTURE ONLY — never linked into the product library.
 *
 * Naive, strictly sequential level-1/2 BLAS used to build the reference's own
 * C sources (which ship no BLAS: prototypes only, /root/reference/src/poismf.h:136-152)
 * into oracle/_ref/.  The summation order is the textbook left-to-right one so
 * that it can be mimicked exactly by the device "strict" mode.
 * Select the value type with -DUSE_FLOAT exactly as the reference does
 * (/root/reference/src/poismf.h:91-109).
 */
#include <math.h>
#include <stddef.h>

#ifdef USE_FLOAT
  typedef float real_t;
  #define FN(x) cblas_s##x
#else
  typedef double real_t;
  #define FN(x) cblas_d##x
#endif

real_t FN(dot)(const int n, const real_t *x, const int incx, const real_t *y, const int incy)
{
    real_t acc = 0;
    for (int i = 0; i < n; i++) acc += x[(size_t)i * incx] * y[(size_t)i * incy];
    return acc;
}

void FN(axpy)(const int n, const real_t alpha, const real_t *x, const int incx, real_t *y, const int incy)
{
    for (int i = 0; i < n; i++) y[(size_t)i * incy] += alpha * x[(size_t)i * incx];
}

void FN(scal)(const int n, const real_t alpha, real_t *x, const int incx)
{
    for (int i = 0; i < n; i++) x[(size_t)i * incx] *= alpha;
}

real_t FN(nrm2)(const int n, const real_t *x, const int incx)
{
    real_t acc = 0;
    for (int i = 0; i < n; i++) acc += x[(size_t)i * incx] * x[(size_t)i * incx];
#ifdef USE_FLOAT
    return sqrtf(acc);
#else
    return sqrt(acc);
#endif
}

/* Row-major, no-transpose only: the single call site is
 * /root/reference/src/topN.c:219-223 (order=101, trans=111). */
void FN(gemv)(const int order, const int trans, const int m, const int n,
              const real_t alpha, const real_t *a, const int lda,
              const real_t *x, const int incx, const real_t beta,
              real_t *y, const int incy)
{
    (void)order; (void)trans;
    for (int r = 0; r < m; r++) {
        real_t acc = 0;
        for (int c = 0; c < n; c++) acc += a[(size_t)r * lda + c] * x[(size_t)c * incx];
        y[(size_t)r * incy] = alpha * acc + ((beta == 0) ? 0 : beta * y[(size_t)r * incy]);
    }
}
