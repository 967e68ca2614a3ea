/*
 * TEST INFRASTRUCTURE — CPU restatement ("oracle") of the TLS period-search hot
 * path.  NOT part of the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * Restates, function by function, /root/reference/transitleastsquares:
 *   fold_phases          core.py:15-18   (foldfast; numba fastmath turns t/P into t*(1/P))
 *   stable_argsort       core.py:120     (numpy argsort kind="mergesort": stable)
 *   t14_fraction         grid.py:9-32    (T14)
 *   search_one_period    core.py:96-188  (search_period)
 *     window means       helpers.py:70-73 (running_mean via cumulative sums), core.py:167
 *     out-of-transit     core.py:79-93   (sequential add/remove recurrence)
 *     edge correction    core.py:21-25
 *     template scan      core.py:28-76   (lowest_residuals_in_this_duration)
 *   tls_oracle_search    main.py:140-185 (the period loop; OpenMP instead of a process pool)
 *
 * Pinned against the reference's own numba path run in the build container:
 * tests/golden/*.npz (made by oracle/make_golden.py), see tests/test_oracle.py.
 *
 * One deliberate deviation that only helps the CPU baseline: the cumulative sum
 * behind running_mean is computed once per period instead of once per duration
 * (same additions in the same order, so identical values).
 *
 * Build: see oracle/Makefile (fold kept free of FMA contraction so the phases
 * are bit-identical to numba's).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HOT __attribute__((target_clones("avx2,fma", "default")))

typedef struct {
    const double *signal;   /* flat template samples, all rows back to back   */
    const int64_t *offset;  /* [rows] start of row r in signal                */
    const int64_t *length;  /* [rows] trimmed template length L_r             */
    const int64_t *width;   /* [rows] width_in_samples W_r                    */
    const double *overshoot;/* [rows]                                         */
    int64_t rows;
} oracle_templates;

typedef struct {
    double transit_depth_min, R_star_min, R_star_max, M_star_min, M_star_max, T0_fit_margin;
} oracle_params;

/* tls_constants.py:20-25, 71, 78 */
static const double kG = 6.673e-11, kRsun = 695508000.0, kRjup = 69911000.0;
static const double kMsun = 1.989e30, kDay = 86400.0, kSignalDepth = 0.5, kFracMax = 0.12;

/* grid.py:9-32 */
static double t14_fraction(double R_s, double M_s, double P, int small)
{
    double Ps = P * kDay, R = kRsun * R_s, M = kMsun * M_s;
    double cube = pow((4 * Ps) / (M_PI * kG * M), 1.0 / 3);
    double t14 = small ? R * cube : (R + 2 * kRjup) * cube;
    double frac = t14 / Ps;
    return frac > kFracMax ? kFracMax : frac;
}

/* core.py:15-18 — compiled in oracle_fold.c with -ffp-contract=off */
void tls_oracle_fold(const double *t, int64_t n, double period, double *phase);

/* core.py:120 — bottom-up merge sort on (phase, index): stable */
static void stable_argsort(const double *key, int64_t n, int64_t *idx, int64_t *tmp)
{
    for (int64_t i = 0; i < n; i++) idx[i] = i;
    /* insertion-sort runs of 16, then merge */
    const int64_t run = 16;
    for (int64_t lo = 0; lo < n; lo += run) {
        int64_t hi = lo + run < n ? lo + run : n;
        for (int64_t i = lo + 1; i < hi; i++) {
            int64_t v = idx[i];
            double kv = key[v];
            int64_t j = i - 1;
            while (j >= lo && key[idx[j]] > kv) { idx[j + 1] = idx[j]; j--; }
            idx[j + 1] = v;
        }
    }
    int64_t *src = idx, *dst = tmp;
    for (int64_t w = run; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) dst[o++] = (key[src[b]] < key[src[a]]) ? src[b++] : src[a++];
            while (a < mid) dst[o++] = src[a++];
            while (b < hi) dst[o++] = src[b++];
        }
        int64_t *s = src; src = dst; dst = s;
    }
    if (src != idx) memcpy(idx, src, (size_t)n * sizeof(int64_t));
}

typedef struct {
    double *phase, *flux, *dys, *pdata, *pinv, *csum, *mean, *ootr;
    int64_t *idx, *tmp;
} scratch_t;

static int scratch_alloc(scratch_t *s, int64_t n, int64_t m)
{
    size_t e = (size_t)(n + m + 2);
    s->phase = malloc(e * 8); s->flux = malloc(e * 8); s->dys = malloc(e * 8);
    s->pdata = malloc(e * 8); s->pinv = malloc(e * 8); s->csum = malloc(e * 8);
    s->mean = malloc(e * 8); s->ootr = malloc(e * 8);
    s->idx = malloc(e * 8); s->tmp = malloc(e * 8);
    return s->phase && s->flux && s->dys && s->pdata && s->pinv && s->csum && s->mean &&
           s->ootr && s->idx && s->tmp;
}
static void scratch_free(scratch_t *s)
{
    free(s->phase); free(s->flux); free(s->dys); free(s->pdata); free(s->pinv);
    free(s->csum); free(s->mean); free(s->ootr); free(s->idx); free(s->tmp);
}

/* core.py:21-25 pieces: sum((1-x)^2 * w) */
HOT static double weighted_sq_sum(const double *x, const double *w, int64_t n)
{
    double acc = 0;
#pragma omp simd reduction(+ : acc)
    for (int64_t i = 0; i < n; i++) acc += ((1 - x[i]) * (1 - x[i])) * w[i];
    return acc;
}

/* core.py:79-93 */
HOT static void out_of_transit(const double *data, const double *inv, int64_t len, int64_t W,
                               double fullsum, double *chi2)
{
    double window = weighted_sq_sum(data, inv, W);
    chi2[0] = fullsum - window;
    for (int64_t i = 1; i < len - W + 1; i++) {
        int64_t vis = i - 1, invis = i - 1 + W;
        double add = (1 - data[vis]) * (1 - data[vis]) * inv[vis];
        double rem = (1 - data[invis]) * (1 - data[invis]) * inv[invis];
        chi2[i] = chi2[i - 1] + add - rem;
    }
}

/* core.py:28-76 */
HOT static void lowest_residuals(const double *mean, int64_t n_mean, double depth_min,
                                 const double *data, int64_t W, const double *signal, int64_t L,
                                 const double *inv, double overshoot, const double *ootr,
                                 double edge, int64_t datapoints, double margin,
                                 double *best_out, double *depth_out)
{
    double best = (double)datapoints, best_depth = 0;
    int64_t xth = 1;
    if (margin > 0 && (double)W > margin) {
        double inv_margin = 1 / margin;
        xth = (int64_t)((double)W / inv_margin);
        if (xth < 1) xth = 1;
    }
    for (int64_t i = 0; i < n_mean; i++) {
        if (mean[i] > depth_min && (i % xth == 0)) {
            const double *d = data + i, *w = inv + i;
            double target = mean[i] * overshoot;
            double scale = kSignalDepth / target;
            double rs = 1 / scale;
            double acc = 0;
#pragma omp simd reduction(+ : acc)
            for (int64_t j = 0; j < L; j++) {
                double sigi = (1 - signal[j]) * rs;
                double r = d[j] - (1 - sigi);
                acc += (r * r) * w[j];
            }
            double stat = acc + ootr[i] - edge;
            if (stat < best) { best = stat; best_depth = 1 - target; }
        }
    }
    *best_out = best; *depth_out = best_depth;
}

/* core.py:96-188.  uniq[] = ascending unique widths, first_row[] = first row with that width */
static void search_one_period(double period, const double *t, const double *y, const double *dy,
                              int64_t n, double span, const oracle_templates *tp,
                              const int64_t *uniq, const int64_t *first_row, int64_t n_uniq,
                              int64_t M, const oracle_params *prm, scratch_t *s,
                              double *chi2_out, int64_t *row_out, double *depth_out)
{
    tls_oracle_fold(t, n, period, s->phase);
    stable_argsort(s->phase, n, s->idx, s->tmp);
    for (int64_t k = 0; k < n; k++) { s->flux[k] = y[s->idx[k]]; s->dys[k] = dy[s->idx[k]]; }

    /* core.py:126-132 patching */
    int64_t len = n + M;
    for (int64_t k = 0; k < len; k++) {
        int64_t src = k < n ? k : k - n;
        double e = s->dys[src];
        s->pinv[k] = 1 / (e * e);
        s->pdata[k] = s->flux[src];
    }
    /* core.py:21-25 */
    double regular = 0;
    for (int64_t k = 0; k < n; k++)
        regular += ((1 - s->flux[k]) * (1 - s->flux[k])) * 1 / (s->dys[k] * s->dys[k]);
    double fullsum = weighted_sq_sum(s->pdata, s->pinv, len);
    double edge = fullsum - regular;

    /* core.py:143-156 admissible widths */
    double dmax = t14_fraction(prm->R_star_max, prm->M_star_max, period, 0);
    double dmin = t14_fraction(prm->R_star_min, prm->M_star_min, period, 1);
    double naive = span / period;
    double corr = (naive + 1) / naive;
    int64_t wmin = (int64_t)floor(dmin * (double)n);
    int64_t wmax = (int64_t)ceil(dmax * (double)n * corr);

    /* helpers.py:70-73 cumulative sum with a leading zero */
    s->csum[0] = 0;
    for (int64_t k = 0; k < len; k++) s->csum[k + 1] = s->csum[k] + s->pdata[k];

    double best = INFINITY, best_depth = 0;
    int64_t best_row = 0;
    for (int64_t u = 0; u < n_uniq; u++) {
        int64_t W = uniq[u];
        if (W < wmin || W > wmax) continue;
        int64_t row = first_row[u];
        int64_t n_mean = len - W + 1;
        double fw = (double)W;
        for (int64_t i = 0; i < n_mean; i++) s->mean[i] = 1 - (s->csum[i + W] - s->csum[i]) / fw;
        out_of_transit(s->pdata, s->pinv, len, W, fullsum, s->ootr);
        double this_best, this_depth;
        lowest_residuals(s->mean, n_mean, prm->transit_depth_min, s->pdata, W,
                         tp->signal + tp->offset[row], tp->length[row], s->pinv,
                         tp->overshoot[row], s->ootr, edge, n, prm->T0_fit_margin,
                         &this_best, &this_depth);
        if (this_best < best) { best = this_best; best_row = row; best_depth = this_depth; }
    }
    *chi2_out = best; *row_out = best_row; *depth_out = best_depth;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* main.py:140-185: all periods; results in the order of periods[] */
int tls_oracle_search(const double *t, const double *y, const double *dy, int64_t n,
                      const double *periods, int64_t n_periods, const oracle_templates *tp,
                      const oracle_params *prm, int threads, double *chi2, int64_t *row,
                      double *depth)
{
    if (n < 1 || tp->rows < 1) return -1;
    /* core.py:113-116 */
    int64_t *uniq = malloc((size_t)tp->rows * 8), *first_row = malloc((size_t)tp->rows * 8);
    memcpy(uniq, tp->width, (size_t)tp->rows * 8);
    qsort(uniq, (size_t)tp->rows, 8, cmp_i64);
    int64_t n_uniq = 0;
    for (int64_t r = 0; r < tp->rows; r++)
        if (n_uniq == 0 || uniq[r] != uniq[n_uniq - 1]) uniq[n_uniq++] = uniq[r];
    for (int64_t u = 0; u < n_uniq; u++) {
        int64_t r = 0;
        while (tp->width[r] != uniq[u]) r++;   /* core.py:163-165 */
        first_row[u] = r;
    }
    int64_t M = uniq[n_uniq - 1];
    if (M % 2 != 0) M += 1;
    double tmin = t[0], tmax = t[0];
    for (int64_t k = 1; k < n; k++) { if (t[k] < tmin) tmin = t[k]; if (t[k] > tmax) tmax = t[k]; }
    double span = tmax - tmin;

    int failed = 0;
#ifdef _OPENMP
    if (threads < 1) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        scratch_t s;
        int ok = scratch_alloc(&s, n, M);
        if (!ok) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp barrier
        if (!failed) {
#pragma omp for schedule(dynamic, 4)
            for (int64_t p = 0; p < n_periods; p++)
                search_one_period(periods[p], t, y, dy, n, span, tp, uniq, first_row, n_uniq, M,
                                  prm, &s, &chi2[p], &row[p], &depth[p]);
        }
        scratch_free(&s);
    }
    free(uniq); free(first_row);
    return failed ? -2 : 0;
}

int tls_oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
