/*
 * tlsb200.h — C ABI of the B200 (sm_100a) TLS period-search library (libtlsb200.so).
 *
 * The reference (hippke/tls 1.0.31) has no FFI/plugin interface; its seam for this
 * path is the Python call
 *     core.search_period(period, t, y, dy, transit_depth_min, R_star_min, R_star_max,
 *                        M_star_min, M_star_max, lc_arr, lc_cache_overview, T0_fit_margin)
 * (/root/reference/transitleastsquares/core.py:96-188) made once per trial period from
 * main.py:142-160 (pool) and main.py:165-183 (serial).  Every argument except `period`
 * is loop-invariant, so the replacement is ONE batched call per search.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer; the library
 *     copies host->device and never keeps a host pointer after return;
 *   - every function returns 0 on success or a negative tlsb_status; the message of
 *     the last failure on the calling thread is tlsb_last_error();
 *   - there is NO CPU fallback: without a CUDA device the calls fail with
 *     TLSB_ERR_CUDA.
 */
#ifndef TLSB200_H
#define TLSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TLSB_OK = 0,
    TLSB_ERR_ARG = -1,    /* bad argument (null pointer, n < 3, no templates, ...) */
    TLSB_ERR_CUDA = -2,   /* CUDA runtime error or no device                       */
    TLSB_ERR_ALLOC = -3,  /* host or device allocation failed                      */
    TLSB_ERR_STATE = -4   /* handle used before its inputs were set                */
} tlsb_status;

/* The light curve after validate_inputs (validate.py:9-46): t, y, dy are f64[n], dy > 0.
 * Replaces the t=, y=, dy= arguments of search_period (core.py:98-100). */
typedef struct {
    const double *t;
    const double *y;
    const double *dy;
    int64_t n;
} tlsb_lightcurve;

/* The template bank, flattened.  Replaces lc_arr (ragged object array, transit.py:159)
 * and lc_cache_overview (struct array {duration, width_in_samples, overshoot},
 * transit.py:108-111) of search_period (core.py:106-107).  Row r owns
 * signal[offset[r] .. offset[r]+length[r]) with length[r] <= width[r]. */
typedef struct {
    const double *signal;
    const int64_t *offset;
    const int64_t *length;
    const int64_t *width;
    const double *overshoot;
    int64_t rows;
} tlsb_templates;

/* The scalar arguments of search_period (core.py:101-105,108). */
typedef struct {
    double transit_depth_min;
    double R_star_min;
    double R_star_max;
    double M_star_min;
    double M_star_max;
    double T0_fit_margin;
} tlsb_params;

/* Where to run.  devices == NULL or n_devices <= 0 means "current device".  With
 * several ordinals the periods are dealt round-robin to the GPUs by host threads of
 * this one process (the one-process-per-GPU + NCCL path lives above the ABI, in
 * tls_b200/distributed.py, and uses the handle API below). */
typedef struct {
    const int32_t *devices;
    int32_t n_devices;
} tlsb_exec;

/*
 * One search: replaces the whole period loop of main.py:140-185.
 * Outputs are in the order of periods[] (main.py:190-196 sorts afterwards):
 *   chi2_out[p]   = search_period(...)[1]   minimum chi^2, N if nothing was fitted,
 *                                           +inf if no duration was admissible
 *   row_out[p]    = search_period(...)[2]   row of the template bank
 *   depth_out[p]  = search_period(...)[3]   1 - depth of the best model (0 if none)
 *   t0_index_out  = optional (may be NULL): index into the phase-sorted samples of the
 *                   window start of the best model, -1 if none.  The reference discards
 *                   this (stats.py:136-138).
 */
int tlsb_search_periods(const tlsb_lightcurve *lc, const double *periods, int64_t n_periods,
                        const tlsb_templates *tp, const tlsb_params *prm, const tlsb_exec *ex,
                        double *chi2_out, int64_t *row_out, double *depth_out,
                        int64_t *t0_index_out);

/* ---- handle API: inputs stay resident in HBM across searches (multi-planet reruns,
 * batches, benchmarking the kernels without the copies). ---- */
typedef struct tlsb_handle tlsb_handle;

int tlsb_create(tlsb_handle **out, int32_t device);
int tlsb_destroy(tlsb_handle *h);
/* host -> device; each may be called again to replace that input */
int tlsb_set_lightcurve(tlsb_handle *h, const tlsb_lightcurve *lc);
int tlsb_set_templates(tlsb_handle *h, const tlsb_templates *tp, const tlsb_params *prm);
int tlsb_set_periods(tlsb_handle *h, const double *periods, int64_t n_periods);
/* The three setters in one call, every upload asynchronous on `cuda_stream` (the stream the following
 * tlsb_search_async will use) and no synchronisation: the caller's buffers must stay valid and unchanged until that
 * stream has passed the uploads (e.g. until the results of the search have been read).  Any of lc, tp (with prm),
 * periods may be NULL to keep what the handle has.  This is the per-step upload of the multi-GPU end-to-end path
 * (one process per GPU re-sends light curve, bank and its period shard: main.py:140-163 pickles the same per task). */
int tlsb_set_inputs_async(tlsb_handle *h, void *cuda_stream, const tlsb_lightcurve *lc, const tlsb_templates *tp,
                          const tlsb_params *prm, const double *periods, int64_t n_periods);
/* Multi-GPU: undo the interleaved period partition ON THE DEVICE.  gathered_dev is the all-gathered record buffer,
 * rank-major: world shards of 3 * ceil(n/world) + 1 words, shard r holding periods r, r + world, ... in the layout of
 * tlsb_search_async.  out_dev receives 3 * n_periods + 1 words: chi2 | depth | packed in the job's period order
 * (main.py:190-196) and the SUM of the shards' status words.  Asynchronous on `cuda_stream`; needs no handle. */
int tlsb_unshard_records(const void *gathered_dev, int64_t n_periods, int32_t world, void *out_dev, void *cuda_stream);
/* Launch the plan + search kernels on `cuda_stream` (a cudaStream_t, NULL = default
 * stream); asynchronous.  Results go to the handle's device buffer, or, when `records_dev`
 * is not NULL, to that device buffer of 3*n_periods + 1 8-byte words: three planes
 * chi2 (f64), depth (f64), packed (int64: row in the low 32 bits, t0 index in the high),
 * then ONE status word (int64).  status != 0 means the device-side T14 limits
 * (grid.py:9-32, core.py:143-156) of that many periods fell within 1e-9 (relative) of an
 * integer, where the device pow() cannot be trusted to round like the host libm: the
 * consumer must then call tlsb_resolve_plan (tlsb_get_results does this by itself for the
 * handle's own buffer). */
int tlsb_search_async(tlsb_handle *h, void *cuda_stream, void *records_dev);
/* 0 = plan on the device (default); 1 = exact plan on the host (libm pow, bit-identical to
 * the reference's T14); 2 = device plan that flags every period (exercises the repair path);
 * 3 = as 2 and the device plan is deliberately wrong for every 7th period (tests). */
int tlsb_set_plan_mode(tlsb_handle *h, int32_t mode);
/* After a search on `cuda_stream` whose status word (see tlsb_search_async) is non-zero: recompute the
 * flagged periods' admissible widths on the host (core.py:143-156 with libm, bit-identical to the
 * reference), search again ONLY the periods whose range differs from the device's, and clear the
 * status word.  records_dev as given to tlsb_search_async (NULL = the handle's buffer).
 * Synchronises.  tlsb_get_results and tlsb_search_batch do this themselves. */
int tlsb_resolve_plan(tlsb_handle *h, void *cuda_stream, void *records_dev);
/* How many searches had to be redone completely with the exact host plan (more flagged periods
 * than the plan kernel lists), and how many single periods were re-searched by the repair path. */
int64_t tlsb_plan_fallback_count(const tlsb_handle *h);
int64_t tlsb_plan_repair_count(const tlsb_handle *h);
/* Wait for the stream and copy the handle's own result buffer to the host.  TLSB_ERR_STATE if the most recent
 * tlsb_search_async was given its own records_dev (the handle's buffer would hold an older search). */
int tlsb_get_results(tlsb_handle *h, void *cuda_stream, double *chi2_out, int64_t *row_out,
                     double *depth_out, int64_t *t0_index_out);
/* Number of kernels this library launched for the most recent tlsb_search_async. */
int64_t tlsb_last_launch_count(const tlsb_handle *h);
/* Average device time [ms] per launch of the main search kernel between the CUDA events the
 * library records around it on its launching stream (0 if nothing ran). Synchronises. */
double tlsb_last_search_kernel_ms(tlsb_handle *h);
/* 1 if the most recent search ran with the folded light curve resident in shared memory,
 * 0 if it streamed it through global scratch. */
int32_t tlsb_last_path_resident(const tlsb_handle *h);
/* Which kernel layout the most recent search used: 1 = resident (folded curve in shared memory),
 * 2 = tiled (phase A in global scratch, phase B from shared-memory chunks staged with bulk async
 * copies), 3 = streaming (everything through L1/L2; last resort for windows wider than a chunk). */
int32_t tlsb_last_path(const tlsb_handle *h);
/* Chunk capacity [doubles per staged array] of the most recent tiled search (0 otherwise). */
int32_t tlsb_last_chunk(const tlsb_handle *h);
/* T0 candidates one lane carried through the tap loop in the most recent search (7 with equal
 * weights, 5 with per-point weights or when 7 would cost too many offsets per chunk). */
int32_t tlsb_last_block(const tlsb_handle *h);
/* Tiled path: how many of the unique template widths (ascending) the most recent search took from
 * staged chunks; the remaining, widest ones were searched from the L2 scratch.  Equals the number of
 * unique widths on the other paths. */
int32_t tlsb_last_tiled_widths(const tlsb_handle *h);
/* Tiled path: the fold is sorted on chip, one phase segment at a time (segment_capacity keys per
 * segment, n_segments segments; both 0 when the sort runs in global scratch).  A period whose
 * phases cluster so strongly that a segment overflows is sorted in global scratch instead;
 * global_sort_periods counts those of the most recent search (synchronises).  Any pointer may be NULL. */
int tlsb_last_sort_info(tlsb_handle *h, int32_t *segment_capacity, int32_t *n_segments,
                        int64_t *global_sort_periods);
/* Force a layout (tests, experiments): path 0 = automatic (default), 1..3 as above; a search
 * fails with TLSB_ERR_ARG if the forced layout cannot hold the inputs.  chunk_doubles > 0 caps
 * the tiled path's chunk capacity (never below the widest window) so that small inputs exercise
 * several chunks; chunk_doubles < 0 caps it at exactly -chunk_doubles, so that the widest widths
 * no longer fit a chunk and take the pass that reads the folded curve from L2 instead. */
int tlsb_set_path(tlsb_handle *h, int32_t path, int32_t chunk_doubles);
/* Equal weights (dy=None): every gate survivor first gets an fp32 correlation with a rigorous error bound, and only
 * the candidates whose chi2 lower bound does not exceed the smallest upper bound seen so far are evaluated in fp64
 * (core.py:57-74 restated).  mode 1 (default) = filter on, mode 0 = every survivor through the exact
 * evaluation; results are bit-identical either way (tests/test_gpu_filter.py).  count_stats != 0 makes the next
 * searches count survivors and finalists on the device. */
int tlsb_set_filter(tlsb_handle *h, int32_t mode, int32_t count_stats);
/* Counters of the most recent search when counting was on (synchronises): gate survivors that took the fp32 pass,
 * finalists evaluated in fp64, and how many of those did not fit the finalist queue.  Any pointer may be NULL. */
int tlsb_last_filter_stats(tlsb_handle *h, int64_t *candidates, int64_t *finalists, int64_t *overflows);
/* Launch shape of the most recent search kernel (any pointer may be NULL). */
int tlsb_last_layout(const tlsb_handle *h, int32_t *threads, int32_t *ctas_per_sm, int32_t *queue_capacity,
                     int64_t *smem_bytes);

/* ---- final_T0_fit (stats.py:135-204, called from main.py:273-283) ----
 * After the period search the reference scans `n_trials` mid-transit epochs at the best
 * period: fold(t, period, Tx) (core.py:9-12), stable argsort (stats.py:173), roll the sorted
 * flux by int(dur/2)+1 (stats.py:186-190), sum (flux - model)^2 / weight^2 over the first
 * `dur` slots and (flux - 1)^2 / weight^2 over the rest (stats.py:193-195), where the
 * reference's weight is the rolled flux rolled once more (stats.py:191 — dy has no effect;
 * kept).  One CUDA launch does all trials against the handle's resident light curve.
 *   model_in[dur]   1 - (1 - signal) / (SIGNAL_DEPTH / (1 - depth))          (stats.py:141-143)
 *   trials[n]       numpy.linspace(min(t), min(t) + period, points)            (stats.py:154-156)
 *   residuals_out   residuals_total of every trial, in trial order            (stats.py:195)
 *   best_index_out  optional: first index of the minimum (strict '<', stats.py:200-202), -1 if
 *                   no residual is below +inf.  T0 = trials[best_index].
 * Synchronous: returns after the results are on the host. */
int tlsb_final_t0_fit(tlsb_handle *h, void *cuda_stream, const double *model_in, int64_t dur,
                      double period, const double *trials, int64_t n_trials,
                      double *residuals_out, int64_t *best_index_out);
/* Same, one-shot with HOST light-curve buffers (device < 0: current device). */
int tlsb_final_t0_fit_lc(const tlsb_lightcurve *lc, int32_t device, const double *model_in,
                         int64_t dur, double period, const double *trials, int64_t n_trials,
                         double *residuals_out, int64_t *best_index_out);
/* Device time [ms] of the most recent T0-fit kernel of this handle (CUDA events). */
double tlsb_last_t0_fit_ms(const tlsb_handle *h);

/* ---- spectra (stats.py:105-132 + helpers.running_median, helpers.py:93-108) ----
 * chi2 (ascending-period order, as main.py:190-196 leaves it) -> SR = min(chi2)/chi2,
 * SDE_raw, power_raw, and the median-detrended, re-normalised `power` with its SDE; all on
 * the device, batched over `n_curves` rows of `n_periods` values each (HOST buffers in and out).
 *   median_window   the reference's `kernel` (oversampling_factor * SDE_MEDIAN_KERNEL_SIZE, made
 *                   odd, stats.py:115-117); detrending applies when n_periods > 2 * window
 *   SR_out, power_raw_out, power_out   [n_curves][n_periods], any may be NULL
 *   SDE_raw_out, SDE_out               [n_curves]
 *   argmax_out      optional [n_curves]: first index of the maximum of `power` (main.py:271) */
int tlsb_spectra(int32_t device, const double *chi2, int64_t n_periods, int64_t n_curves,
                 int64_t median_window, double *SR_out, double *power_raw_out, double *power_out,
                 double *SDE_raw_out, double *SDE_out, int64_t *argmax_out);

/* ---- batches (BASELINE.json config "1,000 independent K2-like light curves") ----
 * Several light curves of the SAME length stay resident on one handle; they share the handle's
 * period grid and template bank (the reference builds both from the time span and the sample
 * count only, main.py:53-88, so curves of one campaign share them).
 *   t   f64[n] when shared_t != 0 (one time axis for all curves), else f64[n_curves][n]
 *   y, dy   f64[n_curves][n]
 * tlsb_select_lightcurve picks the curve that tlsb_search_async and tlsb_final_t0_fit work on
 * (0 after tlsb_set_lightcurves / tlsb_set_lightcurve). */
int tlsb_set_lightcurves(tlsb_handle *h, const double *t, const double *y, const double *dy, int64_t n,
                         int64_t n_curves, int32_t shared_t);
int tlsb_select_lightcurve(tlsb_handle *h, int64_t index);
int64_t tlsb_lightcurve_count(const tlsb_handle *h);
/* The whole batch in one call: for every resident curve the plan + search kernels (the period
 * loop of main.py:140-185), then stats.spectra for all curves at once (stats.py:105-132), all
 * queued on `cuda_stream` with ONE synchronisation at the end.  Outputs (HOST buffers):
 *   chi2_out, row_out, depth_out, t0_index_out   [n_curves][n_periods], order of periods[], any may be NULL
 *   power_out               [n_curves][n_periods] in ASCENDING-period order (results.power), may be NULL
 *   SDE_raw_out, SDE_out    [n_curves]
 *   best_period_index_out   optional [n_curves]: index into periods[] of the highest `power` peak
 *                           (main.py:271-272: period = periods[argmax(power)]) */
int tlsb_search_batch(tlsb_handle *h, void *cuda_stream, int64_t median_window, double *chi2_out,
                      int64_t *row_out, double *depth_out, int64_t *t0_index_out, double *power_out,
                      double *SDE_raw_out, double *SDE_out, int64_t *best_period_index_out);

const char *tlsb_last_error(void);
const char *tlsb_version(void);
int32_t tlsb_device_count(void);

/* The calling thread's current CUDA device (what a negative `device` argument resolves to), or -1. */
int32_t tlsb_current_device(void);

#ifdef __cplusplus
}
#endif
#endif /* TLSB200_H */
