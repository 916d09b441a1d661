/* TEST INFRASTRUCTURE ONLY -- table-driven O(n^2) DCT/DST standing in for
 * libfftw3 in the oracle build (see fftw3.h in this directory).
 *
 * Definitions (FFTW manual, "1d Real-even DFTs (DCTs)" / "Real-odd DFTs"):
 *   REDFT10: Y[k] = 2 * sum_{j=0}^{n-1} X[j] cos(pi (j+1/2) k / n)
 *   REDFT01: Y[k] = X[0] + 2 * sum_{j=1}^{n-1} X[j] cos(pi j (k+1/2) / n)
 *   RODFT01: Y[k] = (-1)^k X[n-1] + 2 * sum_{j=0}^{n-2} X[j] sin(pi (j+1)(k+1/2) / n)
 * Tables are cached per (n, kind); rows are evaluated as dense dot products so
 * gcc vectorises them (the CPU-baseline timing must not be inflated by a
 * cos()-per-term shim).
 */
#include "fftw3.h"
#ifndef S2PB_FFTW_ALT
#define S2PB_FFTW_ALT 0   /* 1, 2: alternative summation orders, only for oracle/_ref/mgm_alt{1,2} (scripts/nodata_spread.py) */
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct s2pb_fftw_plan_s {
    int n;
    fftw_r2r_kind kind;
    double *in, *out;
    const double *table; /* n*n, row k = coefficients applied to X[0..n-1] */
};

#define MAXCACHE 64
static struct { int n; fftw_r2r_kind kind; double *t; } cache[MAXCACHE];
static int ncache = 0;

static const double *get_table(int n, fftw_r2r_kind kind)
{
    const double *found = NULL;
#pragma omp critical(s2pb_fftw_cache)
    {
        for (int i = 0; i < ncache; i++)
            if (cache[i].n == n && cache[i].kind == kind) found = cache[i].t;
        if (!found) {
            double *t = (double *)malloc(sizeof(double) * (size_t)n * n);
            const double pi = 3.14159265358979323846264338327950288;
            for (int k = 0; k < n; k++)
                for (int j = 0; j < n; j++) {
                    double v;
                    if (kind == FFTW_REDFT10)
                        v = 2.0 * cos(pi * (j + 0.5) * k / n);
                    else if (kind == FFTW_REDFT01)
                        v = (j == 0) ? 1.0 : 2.0 * cos(pi * j * (k + 0.5) / n);
                    else /* RODFT01 */
                        v = (j == n - 1) ? ((k & 1) ? -1.0 : 1.0)
                                         : 2.0 * sin(pi * (j + 1) * (k + 0.5) / n);
                    t[(size_t)k * n + j] = v;
                }
            if (ncache < MAXCACHE) {
                cache[ncache].n = n; cache[ncache].kind = kind; cache[ncache].t = t;
                ncache++;
            }
            found = t;
        }
    }
    return found;
}

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

fftw_plan fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind, unsigned flags)
{
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->n = n; p->kind = kind; p->in = in; p->out = out;
    p->table = get_table(n, kind);
    return p;
}

void fftw_execute(const fftw_plan p)
{
    const int n = p->n;
    const double *restrict x = p->in;
    double *restrict y = p->out;
    for (int k = 0; k < n; k++) {
        const double *restrict row = p->table + (size_t)k * n;
#if S2PB_FFTW_ALT == 1      /* same products, summed from the far end: "another fftw build" for the spread study */
        double acc = 0.0;
        for (int j = n - 1; j >= 0; j--) acc += row[j] * x[j];
        y[k] = acc;
#elif S2PB_FFTW_ALT == 2    /* 80-bit accumulation, rounded once: a near-exact transform */
        long double acc = 0.0L;
        for (int j = 0; j < n; j++) acc += (long double)row[j] * (long double)x[j];
        y[k] = (double)acc;
#else
        double acc = 0.0;
        for (int j = 0; j < n; j++) acc += row[j] * x[j];
        y[k] = acc;
#endif
    }
}

void fftw_destroy_plan(fftw_plan p) { free(p); }
