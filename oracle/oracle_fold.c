/*
 * TEST INFRASTRUCTURE (see tls_oracle.c).  Phase fold of core.py:15-18.
 *
 * numba compiles foldfast with fastmath=True and LLVM rewrites time/period as
 * time*(1/period): one reciprocal, one rounded product, no FMA (SURVEY.md §0.2,
 * verified bit-for-bit on sampled periods).  This file is compiled with
 * -ffp-contract=off and without -ffast-math so the product is rounded before
 * floor() and the subtraction, which fixes the sort order of the oracle.
 */
#include <math.h>
#include <stdint.h>

void tls_oracle_fold(const double *t, int64_t n, double period, double *phase)
{
    const double r = 1.0 / period;
    for (int64_t k = 0; k < n; k++) {
        double x = t[k] * r;
        phase[k] = x - floor(x);
    }
}

/* core.py:9-12 (fold with T0), same rewrite: (t-T0)*(1/P) - floor(.) */
void tls_oracle_fold_t0(const double *t, int64_t n, double period, double t0, double *phase)
{
    const double r = 1.0 / period;
    for (int64_t k = 0; k < n; k++) {
        double x = (t[k] - t0) * r;
        phase[k] = x - floor(x);
    }
}
