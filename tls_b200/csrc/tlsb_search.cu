// tlsb_search.cu — the TLS period/duration/T0 grid search as sm_100a CUDA + its C ABI.
//
// What one trial period costs in the reference (core.py:96-188, per period):
//   fold (core.py:15-18) -> stable argsort (core.py:120) -> gathers (:121-123) -> patch (:126-132)
//   -> T14 limits and the admissible widths (:143-156, grid.py:9-32)
//   -> per admissible duration W: running_mean (helpers.py:70-73), out_of_transit_residuals
//   (core.py:79-93), lowest_residuals_in_this_duration (core.py:28-76) -> min over durations.
//
// Two kernels per search:
//
//   tlsb_plan_kernel    per period: the T14 limits -> admissible range of unique widths, and a
//                       counting sort of the periods by cost (most expensive first).
//   tlsb_search_kernel  persistent CTAs, each takes one period at a time and keeps everything
//                       on chip when the folded curve fits shared memory ("resident" path), or
//                       in a per-CTA global scratch that stays in L2 ("streaming" path, any N):
//     A. fold in fp64 with the reciprocal-multiply form numba emits, bucket the phases
//        (histogram -> scan -> scatter), rank inside the bucket by (phase, index) = a stable
//        sort, gather d = 1-y and w = 1/dy^2 to their sorted slots, wrap the first M samples
//        to the end, block-scan d into cumulative sums, block-reduce T = sum w d^2.
//     B. With d = 1-y, D = mean*overshoot and q_j = (1-signal_j)/SIGNAL_DEPTH the reference's
//        statistic is algebraically
//            chi2_i(W) = T + D^2 * sum_j q_j^2 w_{i+j} - 2 D * sum_j q_j (w d)_{i+j}
//                          - sum_{k=L..W-1} (w d^2)_{i+k}
//        (the sum over the window of w d^2 cancels against out_of_transit_residuals and the
//        edge correction).  B1 gates every candidate offset from two
//        cumulative-sum reads (mean_i > transit_depth_min) and appends the surviving blocks of
//        kBlock neighbouring candidates to a CTA-wide queue; B2 runs the register-blocked,
//        software-pipelined tap loop on full warps of survivors.
//     C. lexicographic (chi2, width order, offset) block arg-min = the reference's strict-<
//        tie rules (core.py:71, :183), sentinel N / +inf handling (core.py:46, :139-140).
//
// No tensor cores: there is no dense contraction here (per-offset depth, gate and stride).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/tlsb200.h"
#include "tlsb_internal.h"

namespace {

// R = kB (template parameter of the kernels): consecutive T0 candidates one lane carries through the
// tap loop.  Odd (neighbouring lanes sit R*stride doubles apart in shared memory): 7 when all weights
// are equal (one correlation: the register window fits), 5 with unequal weights (two correlations).
constexpr int kBlockMax = 7;          // host-side slack (template padding, array slack) is sized for the largest R
constexpr int kSub = 4;               // sub-tiles of 32 blocks a warp gates per queue reservation
__host__ __device__ constexpr int tile_size(int kb) { return 32 * kb * kSub; }  // candidates one warp gates at a time
constexpr int kPadGroups = 3;         // slack (in groups of kBlock steps) behind templates and patched arrays
constexpr int kScanItems = 5;         // items per thread per scan tile (odd: conflict-free in smem)
constexpr int kResScanItems = 19;     // ... of the resident kernel: one tile covers 19 * 256 = 4864 samples (cfg-1: one tile, 3 barriers)
#ifndef TLSB_RES_HSCAN
#define TLSB_RES_HSCAN 19  // one tile for cfg-1's 4,322 histogram words, like kResScanItems (5 -> 17: 2.611 -> 2.579 ms per cfg-1 grid)
#endif
#ifndef TLSB_RES_SORT_U
#define TLSB_RES_SORT_U 4
#endif
constexpr int kResHScanItems = TLSB_RES_HSCAN;  // ... of the bucket-histogram scan of the resident kernel
constexpr int kResSortU = TLSB_RES_SORT_U;      // independent key chains per thread in the resident kernel's fold / scatter / rank loops
constexpr int kSegPerThread = 16;   // keys of one segment a thread keeps in registers (S <= 16 * threads)
constexpr int kSegScanItems = 17;   // scan tile of the on-chip sort: 17 * threads > S, so one tile and three barriers per scan
constexpr int kMaxSegments = 64;     // phase segments of the on-chip sort of the tiled path
constexpr int kPlanThreads = 1024;
constexpr int kPlanBins = 1024;
constexpr int kUnsureCap = 4096;     // uncertain periods the plan kernel lists (more: the whole plan is redone on the host)
constexpr unsigned kFull = 0xffffffffu;
constexpr double kSignalDepth = 0.5;  // tls_constants.py:71
constexpr double kPlanEps = 1e-9;     // relative distance to an integer below which the device plan is "uncertain"

}  // namespace

namespace tlsb {  // shared with the other translation units (tlsb_internal.h)
thread_local std::string g_error;

int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}
}  // namespace tlsb

namespace {
using tlsb::fail;
using tlsb::g_error;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(TLSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// ------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------
// Per unique width (ascending), core.py:113 / :163-165.  One record so that a lane can fetch
// everything about "its" width with a few shared-memory loads.
struct WidthRec {
    int W;        // width in samples
    int L;        // template length L <= W
    int X;        // T0 stride (core.py:50-55)
    int row;      // first row of the bank with that width
    int q;        // offset of the (zero padded) template in tq
    int ncand;    // candidates: offsets i = c*X, c in [0, ncand)
    int tiles;    // ceil(ncand / kTile)
    int cum;      // tiles of all wider widths (the sweep runs wide -> narrow)
    double os;    // overshoot
    double invW;  // 1 / W
    double sq2;   // sum_j q_j^2 (the quadratic term when all weights are equal)
};

struct PlanArgs {
    const double *periods;
    int P;
    const WidthRec *rec;
    int nU;
    int N;
    double span;                                        // max(t) - min(t), core.py:148
    double R_star_min, R_star_max, M_star_min, M_star_max;
    double eps;                                         // kPlanEps (or huge: test mode)
    int *ulo, *uhi, *order, *bin_of;                    // [P]
    int *gbins;                                         // [kPlanBins + 2] cost histogram, uncertain periods, finished CTAs (zero between launches)
    int *unsure_list;                                   // [kUnsureCap] the first uncertain periods
    int sabotage;                                       // tests: drop the widest admissible width of every 7th period
    long long *status;                                  // records word 3P: number of uncertain periods
};

struct SearchArgs {
    // light curve, prepared once per curve by prepare_kernel
    const double *t;      // [N]
    const double *dval;   // [N] 1 - y
    const double *wval;   // [N] 1 / dy^2
    int N;
    const double *tq;     // flat q_j = (1 - signal_j) / SIGNAL_DEPTH, each template zero padded
    const WidthRec *rec;  // [nU]
    int nU;
    int M;                // patch length (max width, made even) core.py:114-116
    int pad;              // readable slack behind the patched arrays
    // periods
    const double *periods;
    const int *ulo;       // [P] admissible unique-width index range [ulo, uhi)
    const int *uhi;
    const int *order;     // [P] processing order (most expensive first)
    int P;
    double depth_min;
    double w0;            // the common weight 1/dy^2 when every dy is the same (dy=None), else unused
    // outputs: three planes of P 8-byte words
    double *out_chi2;
    double *out_depth;
    long long *out_packed;
    // scheduling
    int *counter;         // [2] next period, finished CTAs
    int qcap;             // capacity of the CTA-wide survivor queue
    // streaming path scratch
    unsigned char *scratch;
    size_t scratch_per_cta;
    int NB;               // number of phase buckets
    int chunk;            // tiled path: doubles per staged array (cs / w / wd) in shared memory
    int seg_cap;          // tiled path, on-chip sort: elements per phase segment (0: sort in global scratch)
    int n_seg;            // number of phase segments (<= kMaxSegments)
    int n_tiled;          // tiled path: unique widths [0, n_tiled) are searched from staged chunks, the rest from L2
};

// tls_constants.py:20-25,78 and grid.py:9-32 (T14); same operation order on host and device
__host__ __device__ inline double t14_fraction(double R_s, double M_s, double P, bool small)
{
    const double G = 6.673e-11, R_sun = 695508000.0, R_jup = 69911000.0, M_sun = 1.989e30;
    const double pi = 3.141592653589793;
    const double Ps = P * 86400.0, R = R_sun * R_s, Ms = M_sun * M_s;
    const double cube = pow((4 * Ps) / (pi * G * Ms), 1.0 / 3);
    const double t14 = small ? R * cube : (R + 2 * R_jup) * cube;
    const double frac = t14 / Ps;
    return frac > 0.12 ? 0.12 : frac;
}

// Admissible width range per period (core.py:143-156) and the processing order.  The T14 limits
// (two fp64 pow() per period) are spread over many CTAs; every CTA adds its periods to a global
// histogram of cost bins, and the LAST CTA to finish scans the bins and scatters the periods
// into the processing order (most expensive first), then clears the bins for the next launch.
// The device pow() may differ from the host libm in the last bits; a period whose limits sit
// within eps of an integer is counted in *status and the host then redoes the plan exactly.
__global__ void __launch_bounds__(kPlanThreads) tlsb_plan_kernel(const PlanArgs a)
{
    __shared__ int bins[kPlanBins];
    __shared__ int warp_tot[32];
    __shared__ int last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nU = a.nU;
    const int total_tiles = nU > 0 ? a.rec[0].cum + a.rec[0].tiles : 0;
    const double Nd = (double)a.N;
    for (int p = blockIdx.x * kPlanThreads + tid; p < a.P; p += gridDim.x * kPlanThreads) {
        const double period = a.periods[p];
        const double dmax = t14_fraction(a.R_star_max, a.M_star_max, period, false);
        const double dmin = t14_fraction(a.R_star_min, a.M_star_min, period, true);
        const double naive = a.span / period;
        const double corr = (naive + 1) / naive;
        const double xlo = dmin * Nd, xhi = dmax * Nd * corr;
        const double wmin_f = floor(xlo), wmax_f = ceil(xhi);
        const bool unsure = fabs(xlo - rint(xlo)) <= a.eps * fmax(1.0, fabs(xlo)) ||
                            fabs(xhi - rint(xhi)) <= a.eps * fmax(1.0, fabs(xhi)) || !(xlo == xlo) || !(xhi == xhi);
        // first u with W >= wmin_f, one past the last u with W <= wmax_f
        int lo = 0, n = nU;
        while (n > 0) {
            const int half = n >> 1;
            if ((double)a.rec[lo + half].W < wmin_f) { lo += half + 1; n -= half + 1; } else n = half;
        }
        int hi = 0;
        n = nU;
        while (n > 0) {
            const int half = n >> 1;
            if ((double)a.rec[hi + half].W <= wmax_f) { hi += half + 1; n -= half + 1; } else n = half;
        }
        if (!(wmax_f >= wmin_f) || hi < lo) hi = lo;  // NaN / empty
        if (a.sabotage && p % 7 == 3 && hi > lo) hi -= 1;
        a.ulo[p] = lo;
        a.uhi[p] = hi;
        const int cost = hi > lo ? a.rec[lo].cum + a.rec[lo].tiles - a.rec[hi - 1].cum : 0;
        int bin = (int)(((long long)cost * kPlanBins) / (total_tiles + 1));
        bin = kPlanBins - 1 - (bin < kPlanBins ? bin : kPlanBins - 1);  // expensive periods first
        a.bin_of[p] = bin;
        atomicAdd(&a.gbins[bin], 1);
        if (unsure) {
            const int at = atomicAdd(&a.gbins[kPlanBins], 1);
            if (at < kUnsureCap) a.unsure_list[at] = p;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(&a.gbins[kPlanBins + 1], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    // exclusive scan of the 1024 bins (one per thread)
    const int mine = *(volatile int *)&a.gbins[tid];
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int wv = warp_tot[lane];
        int wi = wv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(kFull, wi, off);
            if (lane >= off) wi += o;
        }
        warp_tot[lane] = wi - wv;
    }
    __syncthreads();
    bins[tid] = warp_tot[wid] + incl - mine;
    __syncthreads();
    for (int p = tid; p < a.P; p += kPlanThreads) a.order[atomicAdd(&bins[__ldcg(a.bin_of + p)], 1)] = p;
    if (tid == 0) {
        *a.status = (long long)*(volatile int *)&a.gbins[kPlanBins];
        a.gbins[kPlanBins] = 0;
        a.gbins[kPlanBins + 1] = 0;
    }
    a.gbins[tid] = 0;  // self-cleaning: the next launch needs no memset
}

__device__ __forceinline__ double fold_phase(double t, double r)
{
    // core.py:15-18 as compiled by numba fastmath: t*(1/P) - floor(t*(1/P)); the product is
    // rounded on its own (never fused into the subtraction).
    double x = __dmul_rn(t, r);
    return x - floor(x);
}

__device__ __forceinline__ int bucket_of(double phase, int NB)
{
    int b = __double2int_rz(phase * (double)NB);
    return b < NB - 1 ? b : NB - 1;
}

// In-place block-wide inclusive scan of data[0..n) (all threads must call).
template <int kT, typename T, int kScanItems = ::kScanItems>
__device__ void block_inclusive_scan(T *data, int n, T *warp_tot /* [kT/32+1] shared */)
{
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    T carry = T(0);
    for (int base = 0; base < n; base += kT * kScanItems) {
        const int first = base + tid * kScanItems;
        T v[kScanItems];
        T run = T(0);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            T x = (first + k < n) ? data[first + k] : T(0);
            run += x;
            v[k] = run;
        }
        // warp scan of the per-thread totals
        T incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            T o = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            T wv = (lane < kW) ? warp_tot[lane] : T(0);
            T wi = wv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                T o = __shfl_up_sync(kFull, wi, off);
                if (lane >= off) wi += o;
            }
            if (lane < kW) warp_tot[lane] = wi - wv;  // exclusive warp offsets
            if (lane == kW - 1) warp_tot[kW] = wi;    // tile total
        }
        __syncthreads();
        T excl = __shfl_up_sync(kFull, incl, 1);  // exclusive prefix of this thread inside its warp
        if (lane == 0) excl = T(0);
        const T offset = carry + warp_tot[wid] + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (first + k < n) data[first + k] = v[k] + offset;
        carry += warp_tot[kW];
        __syncthreads();
    }
}

// Second half of the sort: skey/sid hold the keys grouped by bucket (any order inside a bucket); rank every key
// inside its bucket by (phase, index) = numpy's stable mergesort order, and gather src1 (src2) to the sorted
// slots of dst1 (dst2).  Bucket b spans [H[b-1], H[b]) with kShift = 0 (H[-1] = 0), [H[b], H[b+1]) with
// kShift = 1.  Ends WITHOUT a barrier.
template <int kT, typename idx_t, bool kTwo, int kU, int kShift>
__device__ __forceinline__ void rank_gather(int N, int NB, const int *H, const double *skey, const idx_t *sid,
                                            const double *__restrict__ src1, const double *__restrict__ src2,
                                            double *dst1, double *dst2)
{
    const int tid = threadIdx.x;
    for (int q0 = tid; q0 < N; q0 += kT * kU) {
        double key[kU], v1[kU], v2[kU];
        int id[kU], lo[kU], hi[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int q = q0 + u * kT;
            key[u] = q < N ? skey[q] : 0.0;
            id[u] = q < N ? (int)sid[q] : 0;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {  // gathers issued early: they overlap the ranking loops
            v1[u] = __ldcs(src1 + id[u]);
            v2[u] = kTwo ? __ldcs(src2 + id[u]) : 0.0;
            const int bk = bucket_of(key[u], NB);
            lo[u] = (bk + kShift) ? H[bk - 1 + kShift] : 0;
            hi[u] = H[bk + kShift];
        }
        // the kU ranking loops run in lockstep so that their loads are in flight together
        int rank[kU], longest = 0;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            rank[u] = lo[u];
            if (q0 + u * kT >= N) hi[u] = lo[u];
            longest = max(longest, hi[u] - lo[u]);
        }
        for (int s = 0; s < longest; ++s) {
            double ks[kU];
            int is[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int at = lo[u] + s < hi[u] ? lo[u] + s : lo[u];  // a harmless re-read once this chain is done
                ks[u] = skey[at];
                is[u] = (int)sid[at];
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (lo[u] + s < hi[u])
                    rank[u] += (ks[u] < key[u]) || (ks[u] == key[u] && is[u] < id[u]);  // (phase, index): the stable order
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (q0 + u * kT < N) {
                dst1[rank[u]] = v1[u];
                if (kTwo) dst2[rank[u]] = v2[u];
            }
        }
    }
}

// Phase fold + stable sort of one trial (core.py:15-18 + :120-123, or stats.py:172-176 with an
// epoch): histogram of NB phase buckets -> block scan -> scatter (any order inside a bucket) ->
// rank inside the bucket by (phase, index) = numpy's stable mergesort order; src1 (and src2) are
// gathered to their sorted slots in dst1 (dst2).  dst1 doubles as the store of the unsorted
// phases until the ranking step; skey/sid/H are scratch.  Ends WITHOUT a barrier.
template <int kT, typename idx_t, bool kTwo, bool kEpoch, int kU = 4, int kHScanItems = ::kScanItems>
__device__ __forceinline__ void fold_sort_gather(const double *__restrict__ t, double T0, double r, int N, int NB,
                                                 int *H, double *skey, idx_t *sid,
                                                 const double *__restrict__ src1, const double *__restrict__ src2,
                                                 double *dst1, double *dst2, int *scan_scratch)
{
    // kU independent load chains per thread (the streaming layouts sort in L2/HBM)
    const int tid = threadIdx.x;
    for (int b = tid; b <= NB; b += kT) H[b] = 0;
    __syncthreads();
    double *ph_unsorted = dst1;  // [N], free until the ranking step writes the sorted values
    for (int k0 = tid; k0 < N; k0 += kT * kU) {
        double tv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) tv[u] = (k0 + u * kT < N) ? __ldcs(t + k0 + u * kT) : 0.0;  // streamed
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int k = k0 + u * kT;
            if (k < N) {
                const double ph = fold_phase(kEpoch ? tv[u] - T0 : tv[u], r);
                ph_unsorted[k] = ph;
                atomicAdd(&H[bucket_of(ph, NB) + 1], 1);
            }
        }
    }
    __syncthreads();
    // inclusive scan of H[0..NB] (H[0] = 0): H[b] = number of keys in buckets < b
    block_inclusive_scan<kT, int, kHScanItems>(H, NB + 1, scan_scratch);
    for (int k0 = tid; k0 < N; k0 += kT * kU) {
        double ph[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) ph[u] = (k0 + u * kT < N) ? ph_unsorted[k0 + u * kT] : 0.0;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int k = k0 + u * kT;
            if (k < N) {
                const int pos = atomicAdd(&H[bucket_of(ph[u], NB)], 1);  // any order inside the bucket
                skey[pos] = ph[u];
                sid[pos] = (idx_t)k;
            }
        }
    }
    __syncthreads();  // now H[b] = end of bucket b; the unsorted phases are dead
    rank_gather<kT, idx_t, kTwo, kU, 0>(N, NB, H, skey, sid, src1, src2, dst1, dst2);
}

// After the sort: cs1[0..N) holds the sorted d = 1 - y (cs1 = cs + 1).  Wrap the first M samples to
// the end (core.py:126-132), then ONE pass turns d into its inclusive cumulative sum in place
// (helpers.py:70-73), writes wd = w * d and returns this thread's share of T = sum_{k<N} w d^2.
// With begin > 0 the pass resumes at position `begin` with the running sum `carry` (the samples
// there already hold their d; nothing is wrapped).
template <int kT, bool kUniformW, int kScanItems = ::kScanItems>
__device__ __forceinline__ double wrap_weight_scan(double *cs1, double *w, double *wd, double w0, int N, int NM,
                                                   int NMP, double *warp_tot /* [kT/32 + 1] shared */,
                                                   int begin = 0, double carry = 0.0)
{
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = N + tid; k < NMP; k += kT) {
        if (k < NM) {
            if (begin == 0) {
                cs1[k] = cs1[k - N];
                if (!kUniformW) w[k] = w[k - N];
            }
        } else {  // slack read (never used) by the unguarded tap groups
            wd[k] = 0.0;
            if (!kUniformW) w[k] = 0.0;
        }
    }
    __syncthreads();
    double tpart = 0.0;
    for (int base = begin; base < NM; base += kT * kScanItems) {
        const int first = base + tid * kScanItems;
        double v[kScanItems];
        double run = 0.0;
        double dv[kScanItems], wv[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {  // every load in flight before the arithmetic
            const int e = first + k < NM ? first + k : NM - 1;
            dv[k] = cs1[e];
            wv[k] = kUniformW ? w0 : w[e];
        }
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = first + k;
            double d = 0.0;
            if (e < NM) {
                d = dv[k];
                const double x = wv[k] * d;
                wd[e] = x;
                if (e < N) tpart = fma(x, d, tpart);
            }
            run += d;
            v[k] = run;
        }
        double incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const double wv = (lane < kW) ? warp_tot[lane] : 0.0;
            double wi = wv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double o = __shfl_up_sync(kFull, wi, off);
                if (lane >= off) wi += o;
            }
            if (lane < kW) warp_tot[lane] = wi - wv;
            if (lane == kW - 1) warp_tot[kW] = wi;
        }
        __syncthreads();
        double excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 0.0;
        const double offset = carry + warp_tot[wid] + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (first + k < NM) cs1[first + k] = v[k] + offset;
        carry += warp_tot[kW];
        __syncthreads();
    }
    return tpart;
}

struct Best {
    double chi2;
    double D;
    int u;
    int i;
};

__device__ __forceinline__ bool better(double c, int u, int i, const Best &b)
{
    return (c < b.chi2) || (c == b.chi2 && (u < b.u || (u == b.u && i < b.i)));
}

// The tap loop for one block of kBlock candidates of width record `wr`, window starts
// i0 + r*X (r < kBlock).  With the stride X the taps split into X residue classes
// j = X*a + b; inside one class candidate r at step m = a + r reads sample i0 + b + X*m, so
// every staged sample (w, w*d) feeds all kBlock candidates and the template value loaded at
// step m is reused from registers for the next kBlock-1 steps.  Steps go in unguarded groups
// of kBlock, and the loads of group g+1 are issued before the arithmetic of group g (software
// pipeline): templates are zero padded in tq and the patched arrays have slack behind them,
// so ramp-in/ramp-out and the one-group overshoot need no predicates.
// kUniformW: all weights equal (dy=None) -> only B = sum q_j (w d)_{i+j} is accumulated.
template <int kBlock, bool kUnit, bool kUniformW>
__device__ __forceinline__ void tap_block(const WidthRec &wr, const double *__restrict__ tq,
                                          const double *__restrict__ w, const double *__restrict__ wd,
                                          int c0, double (&A)[kBlock], double (&B)[kBlock])
{
    const int L = wr.L, X = kUnit ? 1 : wr.X;
    const int i0 = c0 * X;
#pragma unroll
    for (int r = 0; r < kBlock; ++r) { A[r] = 0.0; B[r] = 0.0; }
    const int nb = X < L ? X : L;
    // Strided widths: neighbouring lanes sit kBlock*X doubles apart, which maps onto only
    // 16/gcd(X,16) of the 16 eight-byte banks.  Lanes therefore walk the X residue classes in
    // ROTATED order, starting at b0 = (block / (16/g)) mod g (a function of the candidate, not of
    // the lane, so results do not depend on queue order): the lanes that share a bank through
    // the stride get distinct residues and the half-warp is conflict free again.
    int b0 = 0;
    if (!kUnit && nb == X) {
        const int g = min(X & -X, 16);
        b0 = ((c0 / kBlock) / (16 / g)) & (g - 1);
    }
    const int groups = ((L + X - 1) / X + 2 * kBlock - 2) / kBlock;  // ceil((taps + kBlock-1) / kBlock), widest class
    for (int t = 0; t < nb; ++t) {
        int b = b0 + t;
        if (b >= nb) b -= nb;
        const double *__restrict__ qp = tq + wr.q + b;
        const double *__restrict__ wp = w + i0 + b;
        const double *__restrict__ wdp = wd + i0 + b;
        double qw[kBlock], pw[kBlock];  // circular: the value loaded at step m lives in slot m % kBlock
#pragma unroll
        for (int r = 0; r < kBlock; ++r) { qw[r] = 0.0; pw[r] = 0.0; }
        double qk[2][kBlock], wv[2][kBlock], wdv[2][kBlock];
        auto load = [&](int buf) {
#pragma unroll
            for (int mm = 0; mm < kBlock; ++mm) {
                qk[buf][mm] = __ldg(qp + mm * X);
                wdv[buf][mm] = wdp[mm * X];
                if (!kUniformW) wv[buf][mm] = wp[mm * X];
            }
            qp += kBlock * X;
            wp += kBlock * X;
            wdp += kBlock * X;
        };
        auto compute = [&](int buf) {
#pragma unroll
            for (int mm = 0; mm < kBlock; ++mm) {
                qw[mm] = qk[buf][mm];
                if (!kUniformW) pw[mm] = qk[buf][mm] * qk[buf][mm];
#pragma unroll
                for (int r = 0; r < kBlock; ++r) {
                    const int slot = (mm - r + kBlock) % kBlock;  // loaded r steps ago
                    B[r] = fma(qw[slot], wdv[buf][mm], B[r]);
                    if (!kUniformW) A[r] = fma(pw[slot], wv[buf][mm], A[r]);
                }
            }
        };
        load(0);
#pragma unroll 1
        for (int g = 0; g < groups; g += 2) {
            load(1);
            compute(0);
            if (g + 1 >= groups) break;
            load(0);
            compute(1);
        }
    }
}

// Samples L..W-1 of a window are in neither the in-transit nor the out-of-transit sum when a
// template was trimmed to L < W (rare: L == W for the limb-darkened templates).
// w d^2 = (w d)^2 / w.
template <bool kUniformW>
__device__ __noinline__ double untouched_tail(const double *w, const double *wd, double w0, int from, int to)
{
    double rest = 0.0;
#pragma unroll 1
    for (int k = from; k < to; ++k) rest += wd[k] * wd[k] / (kUniformW ? w0 : w[k]);
    return rest;
}


// B1 for one block: bit rr of the result is set when candidate c0 + rr of a width passes the gate
// mean_i > transit_depth_min (core.py:58), mean from two cumulative-sum reads (helpers.py:70-73).
// kUnit: stride 1 (most widths) - the 2 * kBlock loads get immediate offsets.
template <int kBlock, bool kUnit>
__device__ __forceinline__ int gate_block(const double *cs, int c0, int c_end, int W, int X, double invW,
                                          double depth_min)
{
    const int Xs = kUnit ? 1 : X;
    int mask = 0;
    if (c0 + kBlock <= c_end) {  // straight line: all loads in flight, then the compares
        const double *__restrict__ lo = cs + (size_t)c0 * Xs;
        const double *__restrict__ hi = lo + W;
        double mean[kBlock];
#pragma unroll
        for (int rr = 0; rr < kBlock; ++rr) mean[rr] = (hi[rr * Xs] - lo[rr * Xs]) * invW;
#pragma unroll
        for (int rr = 0; rr < kBlock; ++rr) mask |= (mean[rr] > depth_min ? 1 : 0) << rr;
    } else {  // the last, partial block of this width (or nothing)
        for (int rr = 0; rr < kBlock; ++rr) {
            const int c = c0 + rr;
            if (c < c_end) {
                const int i = c * Xs;
                if ((cs[i + W] - cs[i]) * invW > depth_min) mask |= 1 << rr;
            }
        }
    }
    return mask;
}

// After the tap loop: chi2 of the block's surviving candidates (bit `rr` of mask), the block's own minimum first
// (same width, ascending offsets: strict '<' keeps the earliest), then ONE lexicographic comparison against the
// lane's running best.  The cumulative sums of all kBlock candidates are loaded up front, unconditionally (the
// arrays have slack behind them), so that the loads are in flight together instead of one per taken branch.
template <int kBlock, bool kUnit, bool kUniformW>
__device__ __forceinline__ void block_min(const WidthRec &wr, const double *cs, const double *w, const double *wd,
                                          double w0, double T, int i0, int mask, int u, const double (&A)[kBlock],
                                          const double (&B)[kBlock], Best &best)
{
    double lo[kBlock], hi[kBlock];
    const int X = kUnit ? 1 : wr.X;
    const double *__restrict__ p = cs + i0;
    const double *__restrict__ ph = p + wr.W;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        lo[rr] = p[rr * X];
        hi[rr] = ph[rr * X];
    }
    double blk_chi = INFINITY, blk_D = 0.0;
    int blk_i = -1;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        const int i = i0 + rr * X;
        const double mean = (hi[rr] - lo[rr]) * wr.invW;
        const double D = mean * wr.os;
        const double Aq = kUniformW ? w0 * wr.sq2 : A[rr];
        double chi = T + D * (D * Aq - 2.0 * B[rr]);
        const bool on = (mask >> rr) & 1;
        if (wr.L < wr.W && on) chi -= untouched_tail<kUniformW>(w, wd, w0, i + wr.L, i + wr.W);
        if (on && chi < blk_chi) { blk_chi = chi; blk_D = D; blk_i = i; }
    }
    if (blk_i >= 0 && better(blk_chi, u, blk_i, best)) { best.chi2 = blk_chi; best.D = blk_D; best.u = u; best.i = blk_i; }
}

template <int kT, bool kResident, bool kUniformW, int kBlock>
__global__ void __launch_bounds__(kT, (kT <= 384 ? 2 : 1)) tlsb_search_kernel(const __grid_constant__ SearchArgs a)
{
    constexpr int kW = kT / 32;
    constexpr int kTile = tile_size(kBlock);
    using idx_t = typename std::conditional<kResident, unsigned short, unsigned int>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, M = a.M, NM = N + M, NB = a.NB, nU = a.nU;
    const int NMP = NM + a.pad;

    // ---- carve memory ------------------------------------------------------------------
    // cs   : (NM+1) doubles   cumulative sums of d, cs[0] = 0      [the sort keys while sorting]
    // w    : NMP doubles      weights (not kept when all weights are equal)
    // wd   : NMP doubles      w*d                                   [the sorted d before that]
    // queue: qcap int2        survivor blocks                       [resident: H + sid while sorting]
    const size_t cs_elems = (size_t)(NM + 2) & ~(size_t)1;
    double *cs, *w, *wd;
    idx_t *sid;
    int *H;
    int2 *queue;
    unsigned char *tail;
    if (kResident) {
        cs = reinterpret_cast<double *>(smem_raw);
        w = cs + cs_elems;
        wd = kUniformW ? w : w + NMP;
        queue = reinterpret_cast<int2 *>(wd + NMP);
        H = reinterpret_cast<int *>(queue);
        sid = reinterpret_cast<idx_t *>(H + NB + 1);
        tail = reinterpret_cast<unsigned char *>(queue + a.qcap);
    } else {
        unsigned char *g = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
        cs = reinterpret_cast<double *>(g);
        w = cs + cs_elems;
        wd = kUniformW ? w : w + NMP;
        sid = reinterpret_cast<idx_t *>(wd + NMP);
        queue = reinterpret_cast<int2 *>(smem_raw);
        H = reinterpret_cast<int *>(queue + a.qcap);
        tail = smem_raw + (size_t)a.qcap * 8 + (((size_t)(NB + 1) * 4 + 15) & ~(size_t)15);
    }
    double *skey = wd;  // the sort keys borrow the wd area; the sorted d go straight to cs[1..N]
    WidthRec *rec = reinterpret_cast<WidthRec *>(tail);                       // [nU]
    double *red_d = reinterpret_cast<double *>(rec + nU);                     // [2*kW + 2]
    int *red_i = reinterpret_cast<int *>(red_d + 2 * kW + 2);                 // [2*kW]
    int *s_next = red_i + 2 * kW;  // [4] period slot, "tiles left" flag, queue fill, queue head

    for (int k = tid; k < nU * (int)(sizeof(WidthRec) / 4); k += kT)
        reinterpret_cast<int *>(rec)[k] = reinterpret_cast<const int *>(a.rec)[k];

    const unsigned lt_mask = (1u << lane) - 1u;
    const double depth_min = a.depth_min;
    const int qstop = a.qcap - kW * 32 * kSub;  // gating pauses here: every warp can still add one tile

    for (;;) {
        if (tid == 0) {
            s_next[0] = atomicAdd(a.counter, 1);
            s_next[1] = 0;
            s_next[2] = 0;
            s_next[3] = 0;
        }
        __syncthreads();
        const int slot_p = s_next[0];
        if (slot_p >= a.P) break;
        const int p = a.order[slot_p];
        const double period = a.periods[p];
        const double r = 1.0 / period;
        const int ulo = a.ulo[p], uhi = a.uhi[p];

        if (ulo >= uhi) {  // core.py:139-140,158-160: nothing admissible -> inf, row 0, depth 0
            if (tid == 0) {
                a.out_chi2[p] = INFINITY;
                a.out_depth[p] = 0.0;
                a.out_packed[p] = (long long)0 | ((long long)(unsigned)-1 << 32);
            }
            __syncthreads();
            continue;
        }

        // ---- A. fold + stable bucket-rank sort + gather --------------------------------
        fold_sort_gather<kT, idx_t, !kUniformW, false, (kResident ? kResSortU : 4), (kResident ? kResHScanItems : kScanItems)>(
            a.t, 0.0, r, N, NB, H, skey, sid, a.dval, a.wval, cs + 1, w,
                                                       reinterpret_cast<int *>(red_d));
        if (tid == 0) cs[0] = 0.0;
        __syncthreads();  // the sorted d sit in cs[1..N]; the keys (in the wd area) are dead
        double tpart = wrap_weight_scan<kT, kUniformW, (kResident ? kResScanItems : kScanItems)>(cs + 1, w, wd, a.w0, N, NM,
                                                                                                 NMP, red_d);
#pragma unroll
        for (int off = 16; off; off >>= 1) tpart += __shfl_xor_sync(kFull, tpart, off);
        if (lane == 0) red_d[kW + 1 + wid] = tpart;
        __syncthreads();
        double T = 0.0;
        for (int k = 0; k < kW; ++k) T += red_d[kW + 1 + k];  // T = sum w d^2 over the unpatched curve; fixed order

        // ---- B. gate + survivor compaction + tap loop ----------------------------------------
        // B1: warp `wid` gates tiles wid, wid+kW, ... of the sweep (wide widths first) from two
        //     cumulative-sum reads per candidate and appends the surviving blocks to the
        //     CTA-wide queue (one reservation per tile, so a tile's survivors stay together and
        //     the queue is nearly sorted by width).
        // B2: warps grab 32 consecutive queue entries - almost always one width, so template
        //     loads broadcast and the lanes run in lockstep - and run the tap loop.
        // The queue is bounded; B1/B2 alternate until all tiles are gated.
        Best best;
        best.chi2 = (double)N;  // core.py:46: a model must beat N to count
        best.D = 0.0;
        best.u = -1;  // "no model yet": loses every tie, so a candidate must be strictly below N
        best.i = -1;

        const int tile_end = rec[ulo].cum + rec[ulo].tiles;
        int g_next = rec[uhi - 1].cum + wid;
        int cur_u = uhi - 1;
        int u_begin = rec[cur_u].cum, u_end = u_begin + rec[cur_u].tiles;  // tile range of width cur_u
        for (;;) {
            // B1
            while (g_next < tile_end) {
                int fill = 0;
                if (lane == 0) fill = *(volatile int *)&s_next[2];
                if (__shfl_sync(kFull, fill, 0) >= qstop) break;
                const int g = g_next;
                g_next += kW;
                while (g >= u_end) {
                    --cur_u;
                    u_begin = u_end;
                    u_end = u_begin + rec[cur_u].tiles;
                }
                const int u = cur_u;
                const int W = rec[u].W, X = rec[u].X, ncand = rec[u].ncand;
                const double invW = rec[u].invW;
                const int c_tile = (g - u_begin) * kTile + lane * kBlock;
                int masks[kSub];
                unsigned votes[kSub];
                int total = 0;
                if (X == 1) {
#pragma unroll
                    for (int sb = 0; sb < kSub; ++sb)
                        masks[sb] = gate_block<kBlock, true>(cs, c_tile + sb * 32 * kBlock, ncand, W, 1, invW, depth_min);
                } else {
#pragma unroll
                    for (int sb = 0; sb < kSub; ++sb)
                        masks[sb] = gate_block<kBlock, false>(cs, c_tile + sb * 32 * kBlock, ncand, W, X, invW, depth_min);
                }
#pragma unroll
                for (int sb = 0; sb < kSub; ++sb) {
                    votes[sb] = __ballot_sync(kFull, masks[sb] != 0);
                    total += __popc(votes[sb]);
                }
                if (total) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_next[2], total);
                    base = __shfl_sync(kFull, base, 0);
#pragma unroll
                    for (int sb = 0; sb < kSub; ++sb) {
                        if (masks[sb])
                            queue[base + __popc(votes[sb] & lt_mask)] =
                                make_int2(c_tile + sb * 32 * kBlock, u | (masks[sb] << 16));
                        base += __popc(votes[sb]);
                    }
                }
            }
            if (lane == 0 && g_next < tile_end) s_next[1] = 1;  // this warp has tiles left
            __syncthreads();
            const int qfill = s_next[2];
            const bool more = s_next[1] != 0;
            // B2
            for (;;) {
                int h = 0;
                if (lane == 0) h = atomicAdd(&s_next[3], 32);
                h = __shfl_sync(kFull, h, 0);
                if (h >= qfill) break;
                if (h + lane < qfill) {
                    const int2 e = queue[h + lane];
                    const int u = e.y & 0xffff, mask = e.y >> 16;
                    const WidthRec wr = rec[u];
                    const int i0 = e.x * wr.X;
                    double A[kBlock], B[kBlock];
                    if (wr.X == 1) {
                        tap_block<kBlock, true, kUniformW>(wr, a.tq, w, wd, e.x, A, B);
                        block_min<kBlock, true, kUniformW>(wr, cs, w, wd, a.w0, T, i0, mask, u, A, B, best);
                    } else {
                        tap_block<kBlock, false, kUniformW>(wr, a.tq, w, wd, e.x, A, B);
                        block_min<kBlock, false, kUniformW>(wr, cs, w, wd, a.w0, T, i0, mask, u, A, B, best);
                    }
                }
            }
            if (!more) break;
            __syncthreads();  // everyone has left B2 before the queue is reused
            if (tid == 0) { s_next[1] = 0; s_next[2] = 0; s_next[3] = 0; }
            __syncthreads();
        }

        // ---- C. block arg-min with the reference's tie order ---------------------------
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            Best o;
            o.chi2 = __shfl_xor_sync(kFull, best.chi2, off);
            o.D = __shfl_xor_sync(kFull, best.D, off);
            o.u = __shfl_xor_sync(kFull, best.u, off);
            o.i = __shfl_xor_sync(kFull, best.i, off);
            if (better(o.chi2, o.u, o.i, best)) best = o;
        }
        __syncthreads();  // everyone is done reading red_d (T) and the queue before they are reused
        if (lane == 0) {
            red_d[wid] = best.chi2;
            red_d[kW + wid] = best.D;
            red_i[wid] = best.u;
            red_i[kW + wid] = best.i;
        }
        __syncthreads();
        if (wid == 0) {
            Best b2;
            b2.chi2 = (double)N; b2.D = 0.0; b2.u = -1; b2.i = -1;
            if (lane < kW) {
                b2.chi2 = red_d[lane];
                b2.D = red_d[kW + lane];
                b2.u = red_i[lane];
                b2.i = red_i[kW + lane];
            }
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                Best o;
                o.chi2 = __shfl_xor_sync(kFull, b2.chi2, off);
                o.D = __shfl_xor_sync(kFull, b2.D, off);
                o.u = __shfl_xor_sync(kFull, b2.u, off);
                o.i = __shfl_xor_sync(kFull, b2.i, off);
                if (better(o.chi2, o.u, o.i, b2)) b2 = o;
            }
            if (lane == 0) {
                if (b2.u >= 0) {
                    a.out_chi2[p] = b2.chi2;
                    a.out_depth[p] = 1.0 - b2.D;  // core.py:74
                    a.out_packed[p] = (long long)(unsigned)rec[b2.u].row | ((long long)b2.i << 32);
                } else {  // every duration returned the sentinel: first admissible row, depth 0
                    a.out_chi2[p] = (double)N;
                    a.out_depth[p] = 0.0;
                    a.out_packed[p] = (long long)(unsigned)rec[ulo].row | ((long long)(unsigned)-1 << 32);
                }
            }
        }
        __syncthreads();
    }

    // last CTA out resets the scheduler so the next launch needs no memset
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(a.counter + 1, 1);
        if (done == (int)gridDim.x - 1) {
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tiled path (light curves too long for the resident path): phase A runs in a per-CTA global
// scratch; phase B walks the folded curve in POSITION CHUNKS.  One elected thread stages
// cs[a0, a0+C), wd[a0, a0+C) (and w) of the chunk into shared memory with 1-D bulk async
// copies (TMA, cp.async.bulk -> mbarrier complete_tx), then the same gate / survivor queue /
// register-blocked tap loop as the resident kernel runs from shared memory for every
// candidate block that STARTS inside [a0, a0+TP), TP = C - (widest admissible window + the
// tap loop's overshoot).  Each folded sample is read from L2 once per chunk instead of twice
// per admissible width.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TLSB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TLSB_DONE;\n"
        "bra TLSB_WAIT;\n"
        "TLSB_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// what a candidate block of width record wr may read behind its start offset
__host__ __device__ inline int window_need(int W, int X, int kb) { return W + kPadGroups * kb * X + 2; }

// Phase A of the tiled path ON CHIP.  The light curve does not fit shared memory, but a slice of
// it does: the phase axis is cut into n_seg equal segments; ONE pass over t folds every sample
// and appends (phase, index) to its segment's list in global scratch (warp-aggregated smem
// counters); then each segment is sorted entirely in shared memory - histogram over fine phase
// buckets with the arrival index of every key kept (so the scatter needs no second round of
// atomics), block scan, rank inside the bucket by (phase, index) - its d = 1-y (and w) gathered
// to their sorted slots, and the segment emitted in phase order: wd = w*d, T, the cumulative sum
// continued from the previous segment, and the first M samples stashed behind position N for the
// wrap (core.py:126-132).  Returns false (nothing consumed) if a segment overflows its capacity -
// strongly clustered phases - and the caller then sorts in global scratch instead.
template <int kT, bool kUniformW>
__device__ __forceinline__ bool sort_on_chip(const SearchArgs &a, double r, unsigned char *area, int *cnt,
                                             double *gkey, unsigned *gid, double *cs1, double *w, double *wd,
                                             int nmp_even, double *red_d, double &tpart_out)
{
    constexpr int kU = 4;                  // independent chains of the rank / gather loop
    constexpr int kUP = 4;                 // ... of the partition pass
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, M = a.M, NM = N + M, S = a.seg_cap, ns = a.n_seg;
    const unsigned lt_mask = (1u << lane) - 1u;
    double *val_s = reinterpret_cast<double *>(area);            // [S] sorted d of the segment
    double *wv_s = val_s + S;                                    // [S] sorted w (unequal weights only)
    double *skey_s = kUniformW ? wv_s : wv_s + S;                // [S] keys, bucket order
    int *H = reinterpret_cast<int *>(skey_s + S);                // [S + 2] fine-bucket histogram
    unsigned *sid_s = reinterpret_cast<unsigned *>(H + S + 2);   // [S] sample ids, bucket order
    const double dns = (double)ns, dS = (double)S;

    for (int j = tid; j <= ns; j += kT) cnt[j] = 0;
    __syncthreads();
    // ---- partition: one pass over t ------------------------------------------------------------
    for (int kb = wid * 32; kb < N; kb += kT * kUP) {  // warp-uniform bounds: every lane reaches the match
        double tv[kUP];
#pragma unroll
        for (int u = 0; u < kUP; ++u) tv[u] = (kb + u * kT + lane < N) ? __ldcs(a.t + kb + u * kT + lane) : 0.0;
#pragma unroll
        for (int u = 0; u < kUP; ++u) {
            const int k = kb + u * kT + lane;
            int sg = -1;
            double ph = 0.0;
            if (k < N) {
                ph = fold_phase(tv[u], r);
                sg = min(ns - 1, __double2int_rz(__dmul_rn(ph, dns)));
            }
            const unsigned peers = __match_any_sync(kFull, sg);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader && sg >= 0) base = atomicAdd(&cnt[sg], __popc(peers));
            base = __shfl_sync(kFull, base, leader);
            if (sg >= 0) {
                const int slot = base + __popc(peers & lt_mask);
                if (slot < S) {
                    gkey[(size_t)sg * S + slot] = ph;
                    gid[(size_t)sg * S + slot] = (unsigned)k;
                }
            }
        }
    }
    __syncthreads();
    int worst = 0;
    for (int j = 0; j < ns; ++j) worst = max(worst, cnt[j]);
    if (worst > S) return false;

    auto fine = [&](double ph, int j) {  // monotone in ph inside segment j
        // explicit roundings: the three passes must map a key to the same bucket (no FMA contraction)
        const double x = __dsub_rn(__dmul_rn(ph, dns), (double)j);
        const int fb = __double2int_rz(__dmul_rn(x, dS));
        return fb < 0 ? 0 : (fb < S - 1 ? fb : S - 1);
    };
    double tpart = 0.0, carry = 0.0;
    int off = 0;
    for (int j = 0; j < ns; ++j) {
        const int nj = cnt[j];
        if (nj == 0) continue;
        const double *lk = gkey + (size_t)j * S;
        const unsigned *li = gid + (size_t)j * S;
        for (int b = tid; b <= S; b += kT) H[b] = 0;
        __syncthreads();
        // histogram: every key of this thread stays in registers together with its bucket and its
        // arrival index inside the bucket, so the scatter below needs neither a reload nor atomics
        double ph[kSegPerThread];
        unsigned id[kSegPerThread], where[kSegPerThread];
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            const int q = tid + i * kT;
            ph[i] = q < nj ? lk[q] : 0.0;
            id[i] = q < nj ? li[q] : 0u;
        }
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            if (tid + i * kT < nj) {
                const int fb = fine(ph[i], j);
                where[i] = ((unsigned)atomicAdd(&H[fb + 1], 1) << 16) | (unsigned)fb;
            }
        }
        __syncthreads();
        block_inclusive_scan<kT, int, kSegScanItems>(H, S + 1, reinterpret_cast<int *>(red_d));  // H[b] = keys in buckets < b
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            if (tid + i * kT < nj) {
                const int pos = H[where[i] & 0xffffu] + (int)(where[i] >> 16);
                skey_s[pos] = ph[i];
                sid_s[pos] = id[i];
            }
        }
        __syncthreads();
        for (int q0 = tid; q0 < nj; q0 += kT * kU) {  // rank inside the bucket, gather to sorted slots
            double key[kU], v1[kU], v2[kU];
            unsigned sidq[kU];
            int lo[kU], hi[kU], rank[kU], longest = 0;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int q = q0 + u * kT < nj ? q0 + u * kT : 0;
                key[u] = skey_s[q];
                sidq[u] = sid_s[q];
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                v1[u] = __ldcs(a.dval + sidq[u]);
                v2[u] = kUniformW ? 0.0 : __ldcs(a.wval + sidq[u]);
                const int fb = fine(key[u], j);
                lo[u] = H[fb];
                hi[u] = q0 + u * kT < nj ? H[fb + 1] : lo[u];
                rank[u] = lo[u];
                longest = max(longest, hi[u] - lo[u]);
            }
            for (int s2 = 0; s2 < longest; ++s2) {  // the kU ranking loops in lockstep
                double ks[kU];
                unsigned is[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int at = lo[u] + s2 < hi[u] ? lo[u] + s2 : lo[u];
                    ks[u] = skey_s[at];
                    is[u] = sid_s[at];
                }
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    if (lo[u] + s2 < hi[u])
                        rank[u] += (ks[u] < key[u]) || (ks[u] == key[u] && is[u] < sidq[u]);  // (phase, index)
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (q0 + u * kT < nj) {
                    val_s[rank[u]] = v1[u];
                    if (!kUniformW) wv_s[rank[u]] = v2[u];
                }
            }
        }
        __syncthreads();
        for (int q = tid; q < nj; q += kT) {  // emit: wd, T, w, and the samples the wrap repeats
            const int pos = off + q;
            const double d = val_s[q];
            const double wv = kUniformW ? a.w0 : wv_s[q];
            const double x = wv * d;
            wd[pos] = x;
            tpart = fma(x, d, tpart);
            if (!kUniformW) w[pos] = wv;
            if (pos < M) {
                cs1[N + pos] = d;
                if (!kUniformW) w[N + pos] = wv;
            }
        }
        __syncthreads();
        block_inclusive_scan<kT, double, kSegScanItems>(val_s, nj, red_d);
        for (int q = tid; q < nj; q += kT) cs1[off + q] = carry + val_s[q];
        carry += val_s[nj - 1];
        off += nj;
        __syncthreads();
    }
    // positions N .. NM-1 hold the wrapped d: continue the cumulative sum, weight them, zero the slack
    wrap_weight_scan<kT, kUniformW, kSegScanItems>(cs1, w, wd, a.w0, N, NM, nmp_even, red_d, N, carry);
    tpart_out = tpart;
    return true;
}

template <int kT, bool kUniformW, int kBlock>
__global__ void __launch_bounds__(kT, (kT <= 256 ? 2 : 1)) tlsb_search_tiled_kernel(const __grid_constant__ SearchArgs a)
{
    constexpr int kW = kT / 32;
    constexpr int kTile = tile_size(kBlock);
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, M = a.M, NM = N + M, NB = a.NB, nU = a.nU;
    const int NMP = NM + a.pad;
    const int C = a.chunk;

    // ---- global scratch of this CTA: cs | [w] | wd | sid  (every array 16-byte aligned) -------
    const size_t cs_elems = ((size_t)NM + 2) & ~(size_t)1;
    const size_t nmp_even = ((size_t)NMP + 1) & ~(size_t)1;
    unsigned char *g = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
    double *cs = reinterpret_cast<double *>(g);
    double *w = cs + cs_elems;
    double *wd = kUniformW ? w : w + nmp_even;
    unsigned *sid = reinterpret_cast<unsigned *>(wd + nmp_even);
    double *skey = wd;  // the sort keys borrow the wd area; the sorted d go straight to cs[1..N]

    // ---- shared: queue | chunk of cs | [chunk of w] | chunk of wd | records, tables, scratch ----
    int2 *queue = reinterpret_cast<int2 *>(smem_raw);
    double *cs_s = reinterpret_cast<double *>(queue + a.qcap);
    double *w_s = cs_s + C;
    double *wd_s = kUniformW ? w_s : w_s + C;
    int *H = reinterpret_cast<int *>(cs_s);  // phase A only: the histogram borrows the chunk area
    WidthRec *rec = reinterpret_cast<WidthRec *>(wd_s + C);                   // [nU]
    double *red_d = reinterpret_cast<double *>(rec + nU);                     // [2*kW + 2]
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(red_d + 2 * kW + 2);
    int *red_i = reinterpret_cast<int *>(bar + 1);                            // [2*kW]
    int *s_next = red_i + 2 * kW;  // [8] period slot, "tiles left" flag, queue fill, queue head, chunk tiles
    int *ch_lo = s_next + 8;       // [nU] first candidate of the chunk, per width
    int *ch_hi = ch_lo + nU;       // [nU] one past the last
    int *ch_tiles = ch_hi + nU;    // [nU]
    int *seg_cnt = ch_tiles + nU;  // [kMaxSegments + 1] on-chip sort: keys per phase segment
    // segment lists of the on-chip sort, behind the arrays above
    double *gkey = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(sid) + (((size_t)N * 4 + 15) & ~(size_t)15));
    unsigned *gid = reinterpret_cast<unsigned *>(gkey + (size_t)a.n_seg * a.seg_cap);

    for (int k = tid; k < nU * (int)(sizeof(WidthRec) / 4); k += kT)
        reinterpret_cast<int *>(rec)[k] = reinterpret_cast<const int *>(a.rec)[k];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned parity = 0;

    const unsigned lt_mask = (1u << lane) - 1u;
    const double depth_min = a.depth_min;
    const int qstop = a.qcap - kW * 32 * kSub;

    for (;;) {
        if (tid == 0) s_next[0] = atomicAdd(a.counter, 1);
        __syncthreads();
        const int slot_p = s_next[0];
        if (slot_p >= a.P) break;
        const int p = a.order[slot_p];
        const double period = a.periods[p];
        const double r = 1.0 / period;
        const int ulo = a.ulo[p], uhi = a.uhi[p];

        if (ulo >= uhi) {  // core.py:139-140,158-160
            if (tid == 0) {
                a.out_chi2[p] = INFINITY;
                a.out_depth[p] = 0.0;
                a.out_packed[p] = (long long)0 | ((long long)(unsigned)-1 << 32);
            }
            __syncthreads();
            continue;
        }

        // ---- A. fold + stable sort + gather, wrap, w*d, T, cumulative sums (global scratch) ----
        double tpart = 0.0;
        if (tid == 0) cs[0] = 0.0;
        bool on_chip = false;
        if (a.seg_cap > 0)
            on_chip = sort_on_chip<kT, kUniformW>(a, r, reinterpret_cast<unsigned char *>(cs_s), seg_cnt, gkey, gid,
                                                  cs + 1, w, wd, (int)nmp_even, red_d, tpart);
        if (!on_chip) {  // clustered phases (or no room for segments): sort in the global scratch
            if (tid == 0 && a.seg_cap > 0) atomicAdd(a.counter + 4, 1);
            fold_sort_gather<kT, unsigned, !kUniformW, false, 8>(a.t, 0.0, r, N, NB, H, skey, sid, a.dval, a.wval,
                                                                 cs + 1, w, reinterpret_cast<int *>(red_d));
            __syncthreads();
            tpart = wrap_weight_scan<kT, kUniformW>(cs + 1, w, wd, a.w0, N, NM, (int)nmp_even, red_d);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) tpart += __shfl_xor_sync(kFull, tpart, off);
        if (lane == 0) red_d[kW + 1 + wid] = tpart;
        __syncthreads();
        double T = 0.0;
        for (int k = 0; k < kW; ++k) T += red_d[kW + 1 + k];
        fence_proxy_async();  // this thread's global writes -> visible to the bulk copies below

        // ---- B. chunks ---------------------------------------------------------------------
        Best best;
        best.chi2 = (double)N;
        best.D = 0.0;
        best.u = -1;
        best.i = -1;
        // Widths [ulo, uT) are searched from staged chunks; the few widest ones whose window leaves too few start
        // offsets in a chunk (unequal weights on a 4-year curve: three staged arrays) are searched afterwards
        // straight from this CTA's L2 scratch, as one "chunk" that spans the whole folded curve.
        const int uT = min(uhi, a.n_tiled);

        // candidate range and tiles of the widths [ua, ub) for start offsets [a0, a0 + span): tables + s_next[4]
        auto build_tables = [&](int a0, int span, int ua, int ub) {
            if (wid == kW - 1) {
                int total = 0;
                for (int base = 0; base < ub - ua; base += 32) {
                    const int idx = base + lane;
                    int tiles = 0;
                    if (idx < ub - ua) {
                        const int u = ub - 1 - idx;
                        const int X = rec[u].X;
                        int lo = (a0 + X - 1) / X;
                        int hi = (a0 + span + X - 1) / X;
                        if (hi > rec[u].ncand) hi = rec[u].ncand;
                        if (hi < lo) hi = lo;
                        tiles = (hi - lo + kTile - 1) / kTile;
                        ch_lo[u] = lo;
                        ch_hi[u] = hi;
                        ch_tiles[u] = tiles;
                    }
#pragma unroll
                    for (int off = 16; off; off >>= 1) tiles += __shfl_xor_sync(kFull, tiles, off);
                    total += tiles;
                }
                if (lane == 0) s_next[4] = total;
            }
        };

        // gate + taps over the tables' tiles, widths ub-1 downwards; csb / wb / wdb are indexable by global offsets
        auto sweep = [&](const double *csb, const double *wb, const double *wdb, int ub) {
            const int tile_end = s_next[4];
            int g_next = wid;
            int cur_u = ub - 1;
            int u_begin = 0, u_end = ch_tiles[cur_u];
            for (;;) {
                // B1: gate
                while (g_next < tile_end) {
                    int fill = 0;
                    if (lane == 0) fill = *(volatile int *)&s_next[2];
                    if (__shfl_sync(kFull, fill, 0) >= qstop) break;
                    const int gt = g_next;
                    g_next += kW;
                    while (gt >= u_end) {
                        --cur_u;
                        u_begin = u_end;
                        u_end = u_begin + ch_tiles[cur_u];
                    }
                    const int u = cur_u;
                    const int W = rec[u].W, X = rec[u].X, c_end = ch_hi[u];
                    const double invW = rec[u].invW;
                    const int c_tile = ch_lo[u] + (gt - u_begin) * kTile + lane * kBlock;
                    int masks[kSub];
                    unsigned votes[kSub];
                    int total = 0;
                    if (X == 1) {
#pragma unroll
                        for (int sb = 0; sb < kSub; ++sb)
                            masks[sb] = gate_block<kBlock, true>(csb, c_tile + sb * 32 * kBlock, c_end, W, 1, invW, depth_min);
                    } else {
#pragma unroll
                        for (int sb = 0; sb < kSub; ++sb)
                            masks[sb] = gate_block<kBlock, false>(csb, c_tile + sb * 32 * kBlock, c_end, W, X, invW, depth_min);
                    }
#pragma unroll
                    for (int sb = 0; sb < kSub; ++sb) {
                        votes[sb] = __ballot_sync(kFull, masks[sb] != 0);
                        total += __popc(votes[sb]);
                    }
                    if (total) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s_next[2], total);
                        base = __shfl_sync(kFull, base, 0);
#pragma unroll
                        for (int sb = 0; sb < kSub; ++sb) {
                            if (masks[sb])
                                queue[base + __popc(votes[sb] & lt_mask)] =
                                    make_int2(c_tile + sb * 32 * kBlock, u | (masks[sb] << 16));
                            base += __popc(votes[sb]);
                        }
                    }
                }
                if (lane == 0 && g_next < tile_end) s_next[1] = 1;
                __syncthreads();
                const int qfill = s_next[2];
                const bool more = s_next[1] != 0;
                // B2: taps
                for (;;) {
                    int h = 0;
                    if (lane == 0) h = atomicAdd(&s_next[3], 32);
                    h = __shfl_sync(kFull, h, 0);
                    if (h >= qfill) break;
                    if (h + lane < qfill) {
                        const int2 e = queue[h + lane];
                        const int u = e.y & 0xffff, mask = e.y >> 16;
                        const WidthRec wr = rec[u];
                        const int i0 = e.x * wr.X;
                        double A[kBlock], B[kBlock];
                        if (wr.X == 1) {
                            tap_block<kBlock, true, kUniformW>(wr, a.tq, wb, wdb, e.x, A, B);
                            block_min<kBlock, true, kUniformW>(wr, csb, wb, wdb, a.w0, T, i0, mask, u, A, B, best);
                        } else {
                            tap_block<kBlock, false, kUniformW>(wr, a.tq, wb, wdb, e.x, A, B);
                            block_min<kBlock, false, kUniformW>(wr, csb, wb, wdb, a.w0, T, i0, mask, u, A, B, best);
                        }
                    }
                }
                if (!more) break;
                __syncthreads();
                if (tid == 0) { s_next[1] = 0; s_next[2] = 0; s_next[3] = 0; }
                __syncthreads();
            }
        };

        if (ulo < uT) {
            const int TP = (C - window_need(rec[uT - 1].W, rec[uT - 1].X, kBlock)) & ~1;
            const int i_last = NM - rec[ulo].W;  // the narrowest admissible width has the most offsets
            for (int a0 = 0; a0 <= i_last; a0 += TP) {
                fence_proxy_async();
                __syncthreads();  // phase A / the previous chunk are done with the staging area and the tables
                if (tid == 0) {
                    const int len_cs = min(C, (int)cs_elems - a0);
                    const int len_wd = min(C, (int)nmp_even - a0);
                    mbar_expect_tx(bar, 8u * (unsigned)(len_cs + (kUniformW ? 1 : 2) * len_wd));
                    bulk_copy_g2s(cs_s, cs + a0, 8u * (unsigned)len_cs, bar);
                    if (!kUniformW) bulk_copy_g2s(w_s, w + a0, 8u * (unsigned)len_wd, bar);
                    bulk_copy_g2s(wd_s, wd + a0, 8u * (unsigned)len_wd, bar);
                    s_next[1] = 0;
                    s_next[2] = 0;
                    s_next[3] = 0;
                }
                build_tables(a0, TP, ulo, uT);
                __syncthreads();
                mbar_wait(bar, parity);
                parity ^= 1u;
                sweep(cs_s - a0, w_s - a0, wd_s - a0, uT);
            }
        }
        if (uT < uhi) {  // the widest widths: gate and taps read the scratch through L1/L2
            __syncthreads();  // the last chunk's sweep is done with the queue and the tables
            if (tid == 0) { s_next[1] = 0; s_next[2] = 0; s_next[3] = 0; }
            build_tables(0, 1 << 30, max(ulo, uT), uhi);
            __syncthreads();
            sweep(cs, w, wd, uhi);
        }

        // ---- C. block arg-min with the reference's tie order ---------------------------
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            Best o;
            o.chi2 = __shfl_xor_sync(kFull, best.chi2, off);
            o.D = __shfl_xor_sync(kFull, best.D, off);
            o.u = __shfl_xor_sync(kFull, best.u, off);
            o.i = __shfl_xor_sync(kFull, best.i, off);
            if (better(o.chi2, o.u, o.i, best)) best = o;
        }
        __syncthreads();
        if (lane == 0) {
            red_d[wid] = best.chi2;
            red_d[kW + wid] = best.D;
            red_i[wid] = best.u;
            red_i[kW + wid] = best.i;
        }
        __syncthreads();
        if (wid == 0) {
            Best b2;
            b2.chi2 = (double)N; b2.D = 0.0; b2.u = -1; b2.i = -1;
            if (lane < kW) {
                b2.chi2 = red_d[lane];
                b2.D = red_d[kW + lane];
                b2.u = red_i[lane];
                b2.i = red_i[kW + lane];
            }
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                Best o;
                o.chi2 = __shfl_xor_sync(kFull, b2.chi2, off);
                o.D = __shfl_xor_sync(kFull, b2.D, off);
                o.u = __shfl_xor_sync(kFull, b2.u, off);
                o.i = __shfl_xor_sync(kFull, b2.i, off);
                if (better(o.chi2, o.u, o.i, b2)) b2 = o;
            }
            if (lane == 0) {
                if (b2.u >= 0) {
                    a.out_chi2[p] = b2.chi2;
                    a.out_depth[p] = 1.0 - b2.D;  // core.py:74
                    a.out_packed[p] = (long long)(unsigned)rec[b2.u].row | ((long long)b2.i << 32);
                } else {
                    a.out_chi2[p] = (double)N;
                    a.out_depth[p] = 0.0;
                    a.out_packed[p] = (long long)(unsigned)rec[ulo].row | ((long long)(unsigned)-1 << 32);
                }
            }
        }
        fence_proxy_async();  // chunk reads (generic proxy) before the next period's bulk copies
        __syncthreads();
    }

    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(a.counter + 1, 1);
        if (done == (int)gridDim.x - 1) {
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------
// final_T0_fit (stats.py:135-204): at the best period, every trial epoch Tx folds the light
// curve with fold(t, period, Tx) (core.py:9-12), sorts it stably (stats.py:173), rolls the
// sorted flux by dur/2+1 (stats.py:186-190), and sums the weighted residuals against the
// in-transit model (first dur samples) and against 1 (the rest).  The reference overwrites
// its weights with a SECOND roll of the rolled flux (stats.py:191), so the weight of slot k is
// 1 / flux_sorted[k - 2*shift]^2 and dy plays no part; that quirk is kept.
// One CTA per trial epoch, persistent; same fold + bucket-rank sort as the search kernel.
struct T0Args {
    const double *t;       // [N]
    const double *y;       // [N]
    int N;
    const double *trials;  // [n_trials] numpy.linspace(min(t), min(t)+period, points), made on the host
    int n_trials;
    const double *model;   // [dur] 1 - (1 - signal) / (SIGNAL_DEPTH / (1 - depth)), stats.py:141-143
    int dur;
    int shift;             // int(dur / 2) + 1
    double period;
    double *residuals;     // [n_trials]
    int *counter;          // [2] next trial, finished CTAs
    int NB;
    unsigned char *scratch;
    size_t scratch_per_cta;
};

template <int kT, bool kResident>
__global__ void __launch_bounds__(kT, (kT <= 256 ? 2 : 1)) tlsb_t0fit_kernel(const __grid_constant__ T0Args a)
{
    constexpr int kW = kT / 32;
    using idx_t = typename std::conditional<kResident, unsigned short, unsigned int>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, NB = a.NB;
    const size_t n_even = ((size_t)N + 1) & ~(size_t)1;

    double *skey, *ys;
    idx_t *sid;
    int *H;
    unsigned char *tail;
    if (kResident) {
        skey = reinterpret_cast<double *>(smem_raw);
        ys = skey + n_even;
        H = reinterpret_cast<int *>(ys + n_even);
        sid = reinterpret_cast<idx_t *>(H + ((NB + 2) & ~1));
        tail = reinterpret_cast<unsigned char *>(sid) + (((size_t)N * sizeof(idx_t) + 15) & ~(size_t)15);
    } else {
        unsigned char *g = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
        skey = reinterpret_cast<double *>(g);
        ys = skey + n_even;
        sid = reinterpret_cast<idx_t *>(ys + n_even);
        H = reinterpret_cast<int *>(smem_raw);
        tail = smem_raw + (((size_t)(NB + 1) * 4 + 15) & ~(size_t)15);
    }
    double *red_d = reinterpret_cast<double *>(tail);  // [kW + 2]
    int *s_next = reinterpret_cast<int *>(red_d + kW + 2);

    const double r = 1.0 / a.period;
    const int dur = a.dur, sh1 = a.shift % N, sh2 = (2 * (a.shift % N)) % N;
    for (;;) {
        if (tid == 0) s_next[0] = atomicAdd(a.counter, 1);
        __syncthreads();
        const int trial = s_next[0];
        if (trial >= a.n_trials) break;
        const double T0 = a.trials[trial];
        fold_sort_gather<kT, idx_t, false, true>(a.t, T0, r, N, NB, H, skey, sid, a.y, nullptr, ys, nullptr,
                                                 reinterpret_cast<int *>(red_d));
        __syncthreads();
        double part = 0.0;
        for (int k = tid; k < N; k += kT) {
            int k1 = k - sh1, k2 = k - sh2;
            if (k1 < 0) k1 += N;
            if (k2 < 0) k2 += N;
            const double flux = ys[k1], wsrc = ys[k2];
            const double ref = k < dur ? __ldg(a.model + k) : 1.0;
            const double diff = flux - ref;
            part += (diff * diff) / (wsrc * wsrc);  // stats.py:193-194
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(kFull, part, off);
        if (lane == 0) red_d[wid] = part;
        __syncthreads();
        if (tid == 0) {
            double total = 0.0;
            for (int k = 0; k < kW; ++k) total += red_d[k];  // fixed order: deterministic
            a.residuals[trial] = total;
        }
        __syncthreads();
    }
    if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(a.counter + 1, 1);
        if (done == (int)gridDim.x - 1) {
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
    }
}

// d = 1 - y, w = 1/dy^2 (core.py:127 computes 1/dy**2 the same way), once per light curve.
__global__ void tlsb_prepare_kernel(const double *__restrict__ y, const double *__restrict__ dy,
                                    double *__restrict__ dval, double *__restrict__ wval, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        dval[k] = 1.0 - y[k];
        const double e = dy[k];
        wval[k] = 1.0 / (e * e);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// How one search is laid out on the SM (chosen per search from N, M, the bank and the device).
struct Layout {
    bool resident = false;
    bool tiled = false;    // not resident: phase B from shared-memory chunks staged by bulk async copies
    int chunk = 0;         // doubles per staged array
    int kb = 5;            // candidates per lane (block size R)
    int seg_cap = 0;       // on-chip sort of the tiled path: segment capacity (0 = off) and count
    int n_seg = 0;
    int n_tiled = 0;       // tiled path: widths [0, n_tiled) fit a chunk with enough start offsets left
    int threads = 256;     // 256 (two CTAs per SM) or 512 (one)
    int ctas_per_sm = 2;
    int qcap = 4096;
    int NB = 0;
    size_t smem = 0;
    size_t scratch_per_cta = 0;
};

}  // namespace

struct tlsb_handle {
    int device = 0;
    int num_sms = 0;
    size_t max_smem = 0;     // per CTA (opt-in)
    size_t smem_per_sm = 0;
    // light curve
    int N = 0;
    double span = 0.0;
    bool uniform_w = false;  // every dy identical (dy=None -> std(y) everywhere, validate.py:39-40)
    double w0 = 0.0;         // 1/dy^2 in that case
    DevBuf t, y, dy, dval, wval;   // n_curves light curves back to back (t: one copy when shared)
    bool have_lc = false;
    int n_curves = 1;              // tlsb_set_lightcurves
    bool shared_t = true;          // every curve uses the same time stamps
    int cur = 0;                   // the curve tlsb_search_async / tlsb_final_t0_fit work on
    std::vector<double> c_span, c_w0;
    std::vector<char> c_uniform;
    bool dev_plan_valid = false;   // ulo/uhi/order on the device match (periods, templates, span)
    double dev_plan_span = 0.0;
    DevBuf asc_order, brec, bchi, bSR, bpr, bpw, bscal, bamax;  // batch pipeline
    std::vector<int> h_asc_order;
    bool asc_valid = false;        // asc_order matches the current periods (made on demand by the batch call)
    bool defer_sync = false;       // one-shot call: the caller's buffers outlive the whole call, setters need not wait
    std::vector<double> h_tq;      // host copy of tq (keeps the upload source alive without a synchronisation)
    // templates
    tlsb_params prm{};
    int nU = 0, M = 0, pad = 0;
    std::vector<WidthRec> recs;   // unique widths, ascending
    DevBuf tq, d_rec;
    bool have_tp = false;
    bool recs_stale = true;       // ncand/tiles/cum depend on N + M
    int rec_kb = 0;               // ... and on the block size R the tiles were counted for
    // periods
    int P = 0;
    std::vector<double> h_periods;
    DevBuf periods, ulo, uhi, order, bin_of;
    bool have_periods = false;
    int path_mode = 0;            // 0 auto, 1 resident, 2 tiled, 3 streaming (tlsb_set_path)
    int chunk_cap = 0;            // tiled path: cap of the chunk capacity in doubles (tests), 0 = none
    int plan_mode = 0;            // 0 device plan, 1 exact host plan, 2 device plan flagging every period (tests)
    bool host_plan_valid = false;
    // outputs / scheduling / scratch
    DevBuf out, counter, scratch, plan_bins, unsure;
    // final_T0_fit
    DevBuf t0_trials, t0_model, t0_resid;
    bool t0_resident = false;
    double t0_ms = 0.0;
    // bookkeeping
    int64_t launches = 0;
    int64_t fallbacks = 0;        // searches redone completely with the exact host plan
    int64_t repairs = 0;          // periods re-searched because their exact limits differed from the device's
    Layout layout;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

namespace {

int upload(DevBuf &buf, const void *src, size_t bytes, cudaStream_t s = nullptr)
{
    if (buf.ensure(bytes ? bytes : 8)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    if (bytes) CUDA_TRY(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, s));
    return 0;
}

// candidates and scheduler tiles per width (depend on N + M and on the block size R), wide -> narrow prefix
int refresh_records(tlsb_handle *h, int kb, cudaStream_t s)
{
    const int kTile = tile_size(kb);
    int cum = 0;
    for (int u = h->nU - 1; u >= 0; --u) {
        WidthRec &wr = h->recs[u];
        wr.ncand = (h->N + h->M - wr.W) / wr.X + 1;  // offsets i = c*X, i in [0, N+M-W]
        wr.tiles = (wr.ncand + kTile - 1) / kTile;
        wr.cum = cum;
        cum += wr.tiles;
    }
    int rc;
    if ((rc = upload(h->d_rec, h->recs.data(), sizeof(WidthRec) * (size_t)h->nU, s))) return rc;  // ordered behind earlier launches on s
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(s));  // h->recs itself stays alive and is only rewritten here
    h->recs_stale = false;
    h->rec_kb = kb;
    h->host_plan_valid = false;
    h->dev_plan_valid = false;
    return 0;
}

// The exact plan on the host (libm pow, bit-identical to the reference's T14): admissible
// unique-width range per period (core.py:143-156) + processing order.  Used when the device
// plan reports a limit too close to an integer to trust its pow(), and by plan_mode 1.
void exact_range(const tlsb_handle *h, double period, int *lo, int *hi)
{
    const double dmax = t14_fraction(h->prm.R_star_max, h->prm.M_star_max, period, false);
    const double dmin = t14_fraction(h->prm.R_star_min, h->prm.M_star_min, period, true);
    const double naive = h->span / period;
    const double corr = (naive + 1) / naive;
    const double wmin_f = std::floor(dmin * (double)h->N);
    const double wmax_f = std::ceil(dmax * (double)h->N * corr);
    int a = 0;
    while (a < h->nU && (double)h->recs[a].W < wmin_f) ++a;
    int b = h->nU;
    while (b > a && (double)h->recs[b - 1].W > wmax_f) --b;
    if (!(wmax_f >= wmin_f)) b = a;  // NaN / empty
    *lo = a;
    *hi = b;
}

int host_plan(tlsb_handle *h)
{
    const int P = h->P;
    std::vector<int> lo(P), hi(P), order(P);
    for (int p = 0; p < P; ++p) exact_range(h, h->h_periods[p], &lo[p], &hi[p]);
    std::iota(order.begin(), order.end(), 0);
    // most expensive first: cost ~ number of candidate tiles in the admissible range
    auto cost = [&](int p) {
        return hi[p] > lo[p] ? h->recs[lo[p]].cum + h->recs[lo[p]].tiles - h->recs[hi[p] - 1].cum : 0;
    };
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost(x) > cost(y); });
    int rc;
    if ((rc = upload(h->ulo, lo.data(), sizeof(int) * P))) return rc;
    if ((rc = upload(h->uhi, hi.data(), sizeof(int) * P))) return rc;
    if ((rc = upload(h->order, order.data(), sizeof(int) * P))) return rc;
    CUDA_TRY(cudaStreamSynchronize(nullptr));  // the vectors above go out of scope
    h->host_plan_valid = true;
    h->dev_plan_valid = false;
    return 0;
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

size_t tail_bytes(int nU, int threads)
{
    const int kW = threads / 32;
    return (size_t)nU * sizeof(WidthRec) + (size_t)(2 * kW + 2) * 8 + (size_t)2 * kW * 4 + 16;
}

size_t resident_smem_bytes(int N, int M, int pad, int nU, bool uniform, int qcap, int threads)
{
    const size_t NM = (size_t)N + M, NMP = NM + pad;
    const size_t cs = ((NM + 2) & ~(size_t)1) * 8;
    return cs + (uniform ? 1 : 2) * NMP * 8 + (size_t)qcap * 8 + tail_bytes(nU, threads);
}

// Pick the on-chip layout: two 256-thread CTAs per SM when two folded curves fit the SM's
// shared memory, else one 512-thread CTA, else the streaming path (global scratch in L2).
Layout choose_layout(const tlsb_handle *h)
{
    Layout best;
    const int N = h->N;
    // block size R: 7 candidates per lane when all weights are equal, 5 with two correlations (registers)
    const char *kbe = std::getenv("TLSB_BLOCK");  // experiments: force 5
    const int kb_pref = (h->uniform_w && !(kbe && std::atoi(kbe) == 5)) ? 7 : 5;
    best.kb = kb_pref;
    if (N < 65536 && h->path_mode <= 1) {
        int tries[2][2] = {{256, 2}, {512, 1}};
        // experiments: TLSB_THREADS=320 (or 384) -> two CTAs of that size per SM at 96 (80) registers per thread, i.e.
        // 20 (24) warps per SM instead of 16; pair it with TLSB_BLOCK=5.  Equal weights only.
        if (const char *te = std::getenv("TLSB_THREADS")) {
            const int tt = std::atoi(te);
            if ((tt == 320 || tt == 384) && h->uniform_w) tries[0][0] = tt;
        }
        const int qcaps[3] = {4096, 3584, 3072};
        for (const auto &t : tries) {
            for (int qcap : qcaps) {
                const size_t bytes = resident_smem_bytes(N, h->M, h->pad, h->nU, h->uniform_w, qcap, t[0]);
                if (bytes > h->max_smem) continue;
                if ((bytes + 1024) * (size_t)t[1] > h->smem_per_sm) continue;
                // the sort borrows the queue: H (NB+1 ints) + sid (N u16)
                const long long room = (long long)qcap * 8 - 2LL * N - 8;
                if (room < 4LL * 64) continue;
                int NB = (int)std::min<long long>(N, room / 4 - 1);
                if (NB < N / 16) continue;
                best.resident = true;
                best.threads = t[0];
                best.ctas_per_sm = t[1];
                best.qcap = qcap;
                best.NB = NB;
                best.smem = bytes;
                return best;
            }
        }
    }
    // Tiled path: phase A in global scratch, phase B from staged chunks.  The chunk must hold the
    // widest window of the bank plus a useful number of start offsets.
    const size_t NM = (size_t)N + h->M, NMP = NM + h->pad;
    const size_t cs = ((NM + 2) & ~(size_t)1) * 8;
    const size_t nmp_even = (NMP + 1) & ~(size_t)1;
    const int narr = h->uniform_w ? 2 : 3;
    int need_max = 0, need5 = 0;
    for (const WidthRec &wr : h->recs) {
        need_max = std::max(need_max, window_need(wr.W, wr.X, kb_pref));
        need5 = std::max(need5, window_need(wr.W, wr.X, 5));
    }
    const char *force = std::getenv("TLSB_TILED");  // "0": never, "256"/"512": force that CTA size (experiments)
    const int forced = force ? std::atoi(force) : -1;
    if (forced != 0 && h->path_mode != 3) {
        const int tries[2][3] = {{256, 2, 3072}, {512, 1, 4096}};  // threads, CTAs per SM, queue entries
        for (const auto &t : tries) {
            if (forced > 0 && forced != t[0]) continue;
            size_t per_cta = std::min(h->max_smem, h->smem_per_sm / (size_t)t[1] - 1024);
            if (const char *cap = std::getenv("TLSB_SMEM_KB"))  // experiments: leave part of the SM's 256 KB to L1
                per_cta = std::min(per_cta, (size_t)std::atoi(cap) * 1024 / (size_t)t[1]);
            const size_t fixed = (size_t)t[2] * 8 + tail_bytes(h->nU, t[0]) + (size_t)h->nU * 12 + 128 + (size_t)(kMaxSegments + 2) * 4;
            if (per_cta <= fixed) continue;
            long long C = (long long)(((per_cta - fixed) / (8 * (size_t)narr)) & ~(size_t)1);
            const bool exact_cap = h->chunk_cap < 0;  // tests: cap the chunk exactly; widths that do not fit take the L2 pass
            if (h->chunk_cap > 0) C = std::min<long long>(C, std::max<long long>(h->chunk_cap, need_max + 64) & ~1LL);
            if (exact_cap) C = std::min<long long>(C, (long long)(-h->chunk_cap) & ~1LL);
            int kb = kb_pref, n_tiled = h->nU;
            long long TP = 0;
            bool fits = C > need5;
            if (fits) {
                TP = C - need_max;
                if (kb > 5 && 5 * TP < 4 * (C - need5)) {  // the longer overshoot would cost > 20 % of the offsets per chunk
                    kb = 5;
                    TP = C - need5;
                }
                if (TP < (h->chunk_cap != 0 ? 2 : 256)) fits = false;
                // One big CTA per SM: a width that would leave less than half of a chunk as start offsets (every sample
                // staged more than twice) takes the L2 pass even though it fits (cfg-2: 22.0 -> 20.7 ms per 6,000
                // periods with 7 of 66 widths moved; TLSB_TP_FRAC = percent, experiments).
                if (fits && t[1] == 1 && h->chunk_cap == 0) {
                    const char *fr = std::getenv("TLSB_TP_FRAC");
                    const long long tp_min = C * (fr ? std::atoi(fr) : 50) / 100;
                    int n = 0;
                    for (const WidthRec &wr : h->recs) {
                        if (C - window_need(wr.W, wr.X, kb) < tp_min) break;
                        ++n;
                    }
                    if (4 * n >= 3 * h->nU && n < h->nU) {
                        n_tiled = n;
                        TP = C - window_need(h->recs[(size_t)n - 1].W, h->recs[(size_t)n - 1].X, kb);
                    }
                }
            }
            if (!fits) {
                // The widest windows leave (almost) no start offsets in a chunk - e.g. three staged arrays for a
                // 4-year curve with unequal weights.  One big CTA per SM then tiles the widths that do fit and
                // searches the few widest ones straight from its L2 scratch (kernel: uT).
                if (t[1] != 1 && !exact_cap) continue;
                kb = 5;
                const long long tp_min = exact_cap ? 64 : std::max<long long>(1024, C / 2);
                n_tiled = 0;
                for (const WidthRec &wr : h->recs) {
                    if (C - window_need(wr.W, wr.X, 5) < tp_min) break;
                    ++n_tiled;
                }
                if (n_tiled < 1 || (!exact_cap && 4 * n_tiled < 3 * h->nU)) continue;
                TP = C - window_need(h->recs[(size_t)n_tiled - 1].W, h->recs[(size_t)n_tiled - 1].X, 5);
            } else if (h->chunk_cap == 0 && t[1] == 2 && forced < 0 && (TP < 2048 || 5 * TP < 3 * C)) {
                continue;  // two CTAs per SM only when most of a chunk is start offsets (halo below ~40 %); one big CTA otherwise
            }
            best.resident = false;
            best.tiled = true;
            best.kb = kb;
            best.threads = t[0];
            best.ctas_per_sm = t[1];
            best.qcap = t[2];
            best.chunk = (int)C;
            best.n_tiled = n_tiled;
            best.NB = (int)std::min<long long>(N, (long long)narr * C * 2 - 2);
            best.smem = (size_t)t[2] * 8 + (size_t)narr * (size_t)C * 8 + tail_bytes(h->nU, t[0]) + (size_t)h->nU * 12 + 128 +
                        (size_t)(kMaxSegments + 2) * 4;
            best.scratch_per_cta = cs + (size_t)(narr - 1) * nmp_even * 8 + align16((size_t)N * 4);
            // on-chip sort: segments of S keys sorted in the chunk area; 1.5x head room over N / n_seg
            const char *oc = std::getenv("TLSB_ONCHIP_SORT");  // "0" disables (experiments)
            const size_t area = (size_t)narr * (size_t)C * 8;
            long long S = (long long)(((area - 64) / (h->uniform_w ? 24 : 32)) & ~(size_t)1);
            S = std::min<long long>(S, (long long)kSegPerThread * t[0]);
            const long long ns = S > 0 ? (3LL * N + 2 * S - 1) / (2 * S) : 0;
            if (!(oc && std::atoi(oc) == 0) && S >= 64 && S <= 65534 && ns >= 1 && ns <= kMaxSegments) {
                best.seg_cap = (int)S;
                best.n_seg = (int)ns;
                best.scratch_per_cta += (size_t)ns * (size_t)S * 12 + 16;
            }
            best.scratch_per_cta = (best.scratch_per_cta + 255) & ~(size_t)255;
            return best;
        }
    }
    // Last resort (a window wider than shared memory can stage): everything through L1/L2.
    best.resident = false;
    best.tiled = false;
    best.threads = 256;
    best.ctas_per_sm = 2;
    best.qcap = 4096;
    const size_t fixed = (size_t)best.qcap * 8 + tail_bytes(h->nU, best.threads) + 64;
    const size_t per_cta = std::min(h->max_smem, h->smem_per_sm / 2 - 1024);
    const size_t budget = per_cta > fixed ? per_cta - fixed : 0;
    best.NB = (int)std::min<size_t>((size_t)N, budget / 4 > 2 ? budget / 4 - 2 : 0);
    best.smem = (size_t)best.qcap * 8 + align16((size_t)(best.NB + 1) * 4) + tail_bytes(h->nU, best.threads);
    best.scratch_per_cta = (cs + (h->uniform_w ? 1 : 2) * NMP * 8 + (size_t)N * 4 + 255) & ~(size_t)255;
    return best;
}

template <typename K>
cudaError_t launch_search(K kernel, const SearchArgs &a, int grid, int threads, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

// plan (unless the exact host plan is in force) + search, asynchronous on `s`
// `only` / `n_only`: search just these periods (device array of indices) with the plan that is already
// on the device - used to repair the few periods whose device-side limits were uncertain.
int enqueue_search(tlsb_handle *h, cudaStream_t s, void *records_dev, bool exact_plan, const int *only = nullptr,
                   int n_only = 0)
{
    int rc;
    const Layout lay = choose_layout(h);
    if ((h->recs_stale || h->rec_kb != lay.kb) && (rc = refresh_records(h, lay.kb, s))) return rc;
    const int P = h->P;
    double *rec_words = reinterpret_cast<double *>(records_dev);
    long long *status = reinterpret_cast<long long *>(rec_words + 3 * (size_t)P);
    if (h->ulo.ensure(sizeof(int) * (size_t)P) || h->uhi.ensure(sizeof(int) * (size_t)P) ||
        h->order.ensure(sizeof(int) * (size_t)P) || h->bin_of.ensure(sizeof(int) * (size_t)P))
        return fail(TLSB_ERR_ALLOC, "device allocation failed");
    h->launches = 0;
    if (only) {
        // keep ulo/uhi as they are
    } else if (exact_plan) {
        if (!h->host_plan_valid && (rc = host_plan(h))) return rc;
        CUDA_TRY(cudaMemsetAsync(status, 0, 8, s));
    } else if (h->dev_plan_valid && h->dev_plan_span == h->span && h->plan_mode == 0) {
        CUDA_TRY(cudaMemsetAsync(status, 0, 8, s));  // same periods, bank and span as the previous launch
    } else {
        PlanArgs pa{};
        pa.periods = h->periods.as<double>(); pa.P = P; pa.rec = h->d_rec.as<WidthRec>(); pa.nU = h->nU;
        pa.N = h->N; pa.span = h->span;
        pa.R_star_min = h->prm.R_star_min; pa.R_star_max = h->prm.R_star_max;
        pa.M_star_min = h->prm.M_star_min; pa.M_star_max = h->prm.M_star_max;
        pa.eps = h->plan_mode >= 2 ? 1e300 : kPlanEps;
        pa.sabotage = h->plan_mode == 3 ? 1 : 0;
        pa.ulo = h->ulo.as<int>(); pa.uhi = h->uhi.as<int>(); pa.order = h->order.as<int>();
        pa.bin_of = h->bin_of.as<int>(); pa.status = status;
        pa.gbins = h->plan_bins.as<int>();
        pa.unsure_list = h->unsure.as<int>();
        const int plan_grid = std::max(1, std::min(h->num_sms, (P + kPlanThreads - 1) / kPlanThreads));
        tlsb_plan_kernel<<<plan_grid, kPlanThreads, 0, s>>>(pa);
        CUDA_TRY(cudaGetLastError());
        h->host_plan_valid = false;
        h->launches += 1;
        h->dev_plan_valid = false;  // becomes valid only once its status word has been seen clean (batch)
        h->dev_plan_span = h->span;
    }

    h->layout = lay;
    if (h->path_mode == 1 && !lay.resident) return fail(TLSB_ERR_ARG, "tlsb_set_path: the folded curve does not fit shared memory (resident path)");
    if (h->path_mode == 2 && !lay.tiled) return fail(TLSB_ERR_ARG, "tlsb_set_path: the widest window does not fit a shared-memory chunk (tiled path)");
    SearchArgs a{};
    const size_t cur_off = (size_t)h->cur * (size_t)h->N;
    a.t = h->t.as<double>() + (h->shared_t ? 0 : cur_off);
    a.dval = h->dval.as<double>() + cur_off; a.wval = h->wval.as<double>() + cur_off; a.N = h->N;
    a.tq = h->tq.as<double>(); a.rec = h->d_rec.as<WidthRec>(); a.nU = h->nU; a.M = h->M; a.pad = h->pad;
    a.periods = h->periods.as<double>(); a.ulo = h->ulo.as<int>(); a.uhi = h->uhi.as<int>();
    a.order = only ? only : h->order.as<int>(); a.P = only ? n_only : P;
    a.depth_min = h->prm.transit_depth_min; a.w0 = h->w0;
    a.out_chi2 = rec_words;
    a.out_depth = rec_words + P;
    a.out_packed = reinterpret_cast<long long *>(rec_words + 2 * (size_t)P);
    a.counter = h->counter.as<int>();
    a.qcap = lay.qcap;
    a.NB = lay.NB;
    a.chunk = lay.chunk;
    a.seg_cap = lay.seg_cap;
    a.n_seg = lay.n_seg;
    a.n_tiled = lay.tiled ? lay.n_tiled : h->nU;
    const int grid = std::min(only ? n_only : P, h->num_sms * lay.ctas_per_sm);
    if (!lay.resident) {
        if (lay.NB < 1) return fail(TLSB_ERR_ARG, "too many distinct template widths for shared memory");
        a.scratch_per_cta = lay.scratch_per_cta;
        if (h->scratch.ensure(a.scratch_per_cta * (size_t)grid)) return fail(TLSB_ERR_ALLOC, "device allocation failed (scratch)");
        a.scratch = h->scratch.as<unsigned char>();
    }
    CUDA_TRY(cudaMemsetAsync(h->counter.as<int>() + 4, 0, 4, s));  // periods whose on-chip sort overflowed
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    const bool uni = h->uniform_w;
    cudaError_t le = cudaSuccess;
#define TLSB_GO(K) le = launch_search(K, a, grid, lay.threads, lay.smem, s)
    if (lay.resident) {
        if (lay.threads == 256) {
            if (uni && lay.kb == 7) TLSB_GO((tlsb_search_kernel<256, true, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_kernel<256, true, true, 5>));
            else TLSB_GO((tlsb_search_kernel<256, true, false, 5>));
        } else if (lay.threads == 320) {  // experiments (TLSB_THREADS)
            if (lay.kb == 7) TLSB_GO((tlsb_search_kernel<320, true, true, 7>));
            else TLSB_GO((tlsb_search_kernel<320, true, true, 5>));
        } else if (lay.threads == 384) {
            if (lay.kb == 7) TLSB_GO((tlsb_search_kernel<384, true, true, 7>));
            else TLSB_GO((tlsb_search_kernel<384, true, true, 5>));
        } else {
            if (uni && lay.kb == 7) TLSB_GO((tlsb_search_kernel<512, true, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_kernel<512, true, true, 5>));
            else TLSB_GO((tlsb_search_kernel<512, true, false, 5>));
        }
    } else if (lay.tiled) {
        if (lay.threads == 256) {
            if (uni && lay.kb == 7) TLSB_GO((tlsb_search_tiled_kernel<256, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_tiled_kernel<256, true, 5>));
            else TLSB_GO((tlsb_search_tiled_kernel<256, false, 5>));
        } else {
            if (uni && lay.kb == 7) TLSB_GO((tlsb_search_tiled_kernel<512, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_tiled_kernel<512, true, 5>));
            else TLSB_GO((tlsb_search_tiled_kernel<512, false, 5>));
        }
    } else {
        if (uni && lay.kb == 7) TLSB_GO((tlsb_search_kernel<256, false, true, 7>));
        else if (uni) TLSB_GO((tlsb_search_kernel<256, false, true, 5>));
        else TLSB_GO((tlsb_search_kernel<256, false, false, 5>));
    }
#undef TLSB_GO
    CUDA_TRY(le);
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    h->launches += 1;
    h->timed = true;
    return 0;
}


// The plan kernel flagged `count` periods whose T14 limits sit too close to an integer for the
// device pow() to be trusted.  Recompute just those on the host (libm, bit-identical to the
// reference), patch the device plan where it differs, and list the periods that changed.
// Returns 1 if there are more flagged periods than the kernel could list (caller: whole exact plan).
int find_changed_periods(tlsb_handle *h, cudaStream_t s, long long count, std::vector<int> *changed)
{
    changed->clear();
    if (count > kUnsureCap) return 1;
    const int n = (int)count;
    std::vector<int> list((size_t)n), lo((size_t)h->P), hi((size_t)h->P);
    CUDA_TRY(cudaMemcpyAsync(list.data(), h->unsure.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(lo.data(), h->ulo.p, sizeof(int) * (size_t)h->P, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(hi.data(), h->uhi.p, sizeof(int) * (size_t)h->P, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int k = 0; k < n; ++k) {
        const int p = list[(size_t)k];
        int a, b;
        exact_range(h, h->h_periods[(size_t)p], &a, &b);
        if (a != lo[(size_t)p] || b != hi[(size_t)p]) {
            CUDA_TRY(cudaMemcpyAsync(h->ulo.as<int>() + p, &a, sizeof(int), cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(h->uhi.as<int>() + p, &b, sizeof(int), cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaStreamSynchronize(s));  // a, b live on this stack frame
            changed->push_back(p);
        }
    }
    if (!changed->empty())  // the list buffer doubles as the processing order of the repair launch
        CUDA_TRY(cudaMemcpyAsync(h->unsure.p, changed->data(), sizeof(int) * changed->size(), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return 0;
}

// Repair the records of the current curve after a search whose status word was `count` != 0.
int resolve_records(tlsb_handle *h, cudaStream_t s, void *records_dev, long long count)
{
    std::vector<int> changed;
    int rc = find_changed_periods(h, s, count, &changed);
    if (rc < 0) return rc;
    long long *status = reinterpret_cast<long long *>(reinterpret_cast<double *>(records_dev) + 3 * (size_t)h->P);
    if (rc == 1) {  // too many to list: the whole plan on the host, everything again
        h->fallbacks += 1;
        return enqueue_search(h, s, records_dev, true);
    }
    if (!changed.empty()) {
        h->repairs += (int64_t)changed.size();
        const int64_t before = h->launches;
        if ((rc = enqueue_search(h, s, records_dev, false, h->unsure.as<int>(), (int)changed.size()))) return rc;
        h->launches += before;
    }
    CUDA_TRY(cudaMemsetAsync(status, 0, 8, s));
    return 0;
}

}  // namespace

extern "C" {

const char *tlsb_last_error(void) { return g_error.c_str(); }
const char *tlsb_version(void) { return "tlsb200 0.2 (sm_100a)"; }

int32_t tlsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t tlsb_current_device(void)
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return d;
}

int tlsb_create(tlsb_handle **out, int32_t device)
{
    if (!out) return fail(TLSB_ERR_ARG, "tlsb_create: out is NULL");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(TLSB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return fail(TLSB_ERR_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    tlsb_handle *h = new (std::nothrow) tlsb_handle();
    if (!h) return fail(TLSB_ERR_ALLOC, "out of host memory");
    h->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->max_smem = prop.sharedMemPerBlockOptin;
    h->smem_per_sm = prop.sharedMemPerMultiprocessor;
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    if (h->counter.ensure(32)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    CUDA_TRY(cudaMemset(h->counter.p, 0, 32));  // [0,1] search kernel, [2,3] T0-fit kernel
    if (h->plan_bins.ensure((kPlanBins + 2) * 4)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    CUDA_TRY(cudaMemset(h->plan_bins.p, 0, (kPlanBins + 2) * 4));
    if (h->unsure.ensure(kUnsureCap * 4)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    *out = h;
    return 0;
}

int tlsb_destroy(tlsb_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (DevBuf *b : {&h->t, &h->y, &h->dy, &h->dval, &h->wval, &h->tq, &h->d_rec, &h->periods, &h->ulo,
                      &h->uhi, &h->order, &h->bin_of, &h->out, &h->counter, &h->scratch, &h->t0_trials, &h->t0_model,
                      &h->t0_resid, &h->plan_bins, &h->unsure, &h->asc_order, &h->brec, &h->bchi, &h->bSR, &h->bpr, &h->bpw, &h->bscal, &h->bamax})
        b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}

static int set_curves(tlsb_handle *h, const double *t, const double *y, const double *dy, int64_t n64,
                      int64_t n_curves, bool shared_t)
{
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = (int)n64;
    const size_t total = (size_t)n * (size_t)n_curves, bytes = sizeof(double) * total;
    int rc;
    if ((rc = upload(h->t, t, shared_t ? sizeof(double) * (size_t)n : bytes))) return rc;
    if ((rc = upload(h->y, y, bytes))) return rc;
    if ((rc = upload(h->dy, dy, bytes))) return rc;
    if (h->dval.ensure(bytes) || h->wval.ensure(bytes)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    tlsb_prepare_kernel<<<(unsigned)((total + 255) / 256), 256>>>(h->y.as<double>(), h->dy.as<double>(),
                                                                 h->dval.as<double>(), h->wval.as<double>(), total);
    CUDA_TRY(cudaGetLastError());
    h->c_span.assign((size_t)n_curves, 0.0);
    h->c_w0.assign((size_t)n_curves, 0.0);
    h->c_uniform.assign((size_t)n_curves, 0);
    for (int64_t c = 0; c < n_curves; ++c) {
        const double *tc = shared_t ? t : t + (size_t)c * n, *dc = dy + (size_t)c * n;
        if (!shared_t || c == 0) {
            double tmin = tc[0], tmax = tc[0];  // core.py:148: max(t) - min(t)
            for (int k = 1; k < n; ++k) {
                tmin = std::min(tmin, tc[k]);
                tmax = std::max(tmax, tc[k]);
            }
            h->c_span[c] = tmax - tmin;
        } else
            h->c_span[c] = h->c_span[0];
        bool uniform = true;
        for (int k = 1; k < n && uniform; ++k) uniform = dc[k] == dc[0];
        h->c_uniform[c] = uniform ? 1 : 0;
        h->c_w0[c] = 1.0 / (dc[0] * dc[0]);
    }
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->n_curves = (int)n_curves;
    h->shared_t = shared_t;
    h->cur = 0;
    h->uniform_w = h->c_uniform[0] != 0;
    h->w0 = h->c_w0[0];
    h->N = n;
    h->span = h->c_span[0];
    h->have_lc = true;
    h->recs_stale = true;
    h->host_plan_valid = false;
    h->dev_plan_valid = false;
    return 0;
}

int tlsb_set_lightcurve(tlsb_handle *h, const tlsb_lightcurve *lc)
{
    if (!h || !lc || !lc->t || !lc->y || !lc->dy) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurve: NULL argument");
    if (lc->n < 3 || lc->n > (int64_t)1 << 28) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurve: need 3 <= n <= 2^28 samples");
    return set_curves(h, lc->t, lc->y, lc->dy, lc->n, 1, true);
}

int tlsb_set_lightcurves(tlsb_handle *h, const double *t, const double *y, const double *dy, int64_t n,
                         int64_t n_curves, int32_t shared_t)
{
    if (!h || !t || !y || !dy) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurves: NULL argument");
    if (n < 3 || n > (int64_t)1 << 28) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurves: need 3 <= n <= 2^28 samples");
    if (n_curves < 1 || n_curves > 65535 || n * n_curves > (int64_t)1 << 31)
        return fail(TLSB_ERR_ARG, "tlsb_set_lightcurves: need 1 <= n_curves <= 65535 and n * n_curves <= 2^31");
    return set_curves(h, t, y, dy, n, n_curves, shared_t != 0);
}

int tlsb_select_lightcurve(tlsb_handle *h, int64_t index)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_select_lightcurve: NULL handle");
    if (!h->have_lc) return fail(TLSB_ERR_STATE, "tlsb_select_lightcurve: no light curves set");
    if (index < 0 || index >= h->n_curves) return fail(TLSB_ERR_ARG, "tlsb_select_lightcurve: index out of range");
    h->cur = (int)index;
    h->uniform_w = h->c_uniform[(size_t)index] != 0;
    h->w0 = h->c_w0[(size_t)index];
    if (h->span != h->c_span[(size_t)index]) {
        h->span = h->c_span[(size_t)index];
        h->host_plan_valid = false;
    }
    return 0;
}

int64_t tlsb_lightcurve_count(const tlsb_handle *h) { return h && h->have_lc ? h->n_curves : 0; }

int tlsb_set_templates(tlsb_handle *h, const tlsb_templates *tp, const tlsb_params *prm)
{
    if (!h || !tp || !prm || !tp->signal || !tp->offset || !tp->length || !tp->width || !tp->overshoot)
        return fail(TLSB_ERR_ARG, "tlsb_set_templates: NULL argument");
    if (tp->rows < 1) return fail(TLSB_ERR_ARG, "tlsb_set_templates: empty template bank");
    CUDA_TRY(cudaSetDevice(h->device));
    const int R = (int)tp->rows;
    // unique widths ascending, first row with each width (core.py:113, :163-165)
    std::vector<int64_t> uniq(tp->width, tp->width + R);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    const int nU = (int)uniq.size();
    if (nU > 65535) return fail(TLSB_ERR_ARG, "tlsb_set_templates: more than 65535 distinct widths");
    std::vector<WidthRec> recs(nU);
    std::vector<double> tq;
    int xmax = 1;
    for (int u = 0; u < nU; ++u) {
        int r = 0;
        while (tp->width[r] != uniq[u]) ++r;
        const int64_t W = uniq[u], L = tp->length[r];
        if (W < 1 || L < 1 || L > W || W > (int64_t)1 << 28)
            return fail(TLSB_ERR_ARG, "tlsb_set_templates: need 1 <= length <= width");
        WidthRec &wr = recs[u];
        wr.W = (int)W;
        wr.L = (int)L;
        wr.row = r;
        wr.q = (int)tq.size();
        wr.os = tp->overshoot[r];
        wr.invW = 1.0 / (double)W;
        // core.py:50-55 stride of the T0 scan
        int xth = 1;
        const double margin = prm->T0_fit_margin;
        if (margin > 0 && (double)W > margin) {
            const double inv_margin = 1 / margin;
            xth = (int)((double)W / inv_margin);
            if (xth < 1) xth = 1;
        }
        wr.X = xth;
        xmax = std::max(xmax, xth);
        wr.ncand = 0;  // need N: refresh_records
        wr.tiles = 0;
        wr.cum = 0;
        const double *s = tp->signal + tp->offset[r];
        double sq2 = 0.0;
        for (int64_t j = 0; j < L; ++j) {
            const double q = (1 - s[j]) / kSignalDepth;  // core.py:61-68
            tq.push_back(q);
            sq2 = std::fma(q, q, sq2);
        }
        wr.sq2 = sq2;
        for (int j = 0; j < xth * kPadGroups * kBlockMax; ++j) tq.push_back(0.0);  // ramp-out + pipeline overshoot
    }
    int M = recs[nU - 1].W;  // core.py:114-116
    if (M % 2 != 0) M += 1;
    int rc;
    h->h_tq.swap(tq);  // stays alive behind the asynchronous upload
    if ((rc = upload(h->tq, h->h_tq.data(), h->h_tq.size() * 8))) return rc;
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->recs.swap(recs);
    h->pad = kPadGroups * kBlockMax * xmax;
    h->nU = nU;
    h->M = M;
    h->prm = *prm;
    h->have_tp = true;
    h->recs_stale = true;
    h->host_plan_valid = false;
    h->dev_plan_valid = false;
    return 0;
}

int tlsb_set_periods(tlsb_handle *h, const double *periods, int64_t n_periods)
{
    if (!h || (!periods && n_periods > 0)) return fail(TLSB_ERR_ARG, "tlsb_set_periods: NULL argument");
    if (n_periods < 0 || n_periods > (int64_t)1 << 30) return fail(TLSB_ERR_ARG, "tlsb_set_periods: bad count");
    CUDA_TRY(cudaSetDevice(h->device));
    h->P = (int)n_periods;
    h->h_periods.assign(periods, periods + n_periods);
    int rc;
    if ((rc = upload(h->periods, periods, sizeof(double) * (size_t)n_periods))) return rc;
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->asc_valid = false;
    h->have_periods = true;
    h->host_plan_valid = false;
    h->dev_plan_valid = false;
    return 0;
}

int tlsb_set_plan_mode(tlsb_handle *h, int32_t mode)
{
    if (!h || mode < 0 || mode > 3) return fail(TLSB_ERR_ARG, "tlsb_set_plan_mode: mode must be 0..3");
    h->plan_mode = mode;
    return 0;
}

int tlsb_set_path(tlsb_handle *h, int32_t path, int32_t chunk_doubles)
{
    if (!h || path < 0 || path > 3) return fail(TLSB_ERR_ARG, "tlsb_set_path: path must be 0..3");
    h->path_mode = path;
    h->chunk_cap = chunk_doubles;
    return 0;
}

int tlsb_search_async(tlsb_handle *h, void *cuda_stream, void *records_dev)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_search_async: NULL handle");
    if (!h->have_lc || !h->have_tp || !h->have_periods)
        return fail(TLSB_ERR_STATE, "tlsb_search_async: light curve, templates and periods must be set first");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    h->launches = 0;
    h->timed = false;
    if (h->P == 0) return 0;
    if (h->M > h->N) return fail(TLSB_ERR_ARG, "widest template is longer than the light curve");
    if (!records_dev) {
        if (h->out.ensure(((size_t)h->P * 3 + 1) * 8)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
        records_dev = h->out.p;
    }
    return enqueue_search(h, s, records_dev, h->plan_mode == 1);
}

int tlsb_get_results(tlsb_handle *h, void *cuda_stream, double *chi2_out, int64_t *row_out,
                     double *depth_out, int64_t *t0_index_out)
{
    if (!h || !chi2_out || !row_out || !depth_out) return fail(TLSB_ERR_ARG, "tlsb_get_results: NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const size_t P = (size_t)h->P;
    if (P == 0) return 0;
    if (!h->out.p) return fail(TLSB_ERR_STATE, "tlsb_get_results: no search has written the handle's buffer");
    std::vector<long long> packed(P + 1);
    for (int attempt = 0; attempt < 2; ++attempt) {
        CUDA_TRY(cudaMemcpyAsync(chi2_out, h->out.p, P * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(depth_out, h->out.as<double>() + P, P * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(packed.data(), h->out.as<double>() + 2 * P, (P + 1) * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (packed[P] == 0 || attempt == 1) break;
        // the device plan was not sure about some periods' limits: settle those on the host and search
        // again only the ones whose admissible widths really differ
        const int64_t before = h->repairs + h->fallbacks;
        int rc = resolve_records(h, s, h->out.p, packed[P]);
        if (rc) return rc;
        if (h->repairs + h->fallbacks == before) break;  // every flagged limit was right: results stand
    }
    for (size_t p = 0; p < P; ++p) {
        row_out[p] = (int64_t)(uint32_t)(packed[p] & 0xffffffffLL);
        if (t0_index_out) t0_index_out[p] = (int64_t)(int32_t)(packed[p] >> 32);
    }
    return 0;
}

int64_t tlsb_last_launch_count(const tlsb_handle *h) { return h ? h->launches : 0; }
int32_t tlsb_last_path_resident(const tlsb_handle *h) { return h && h->layout.resident ? 1 : 0; }
int32_t tlsb_last_path(const tlsb_handle *h) { return !h ? 0 : h->layout.resident ? 1 : h->layout.tiled ? 2 : 3; }
int32_t tlsb_last_chunk(const tlsb_handle *h) { return h ? h->layout.chunk : 0; }
int32_t tlsb_last_block(const tlsb_handle *h) { return h ? h->layout.kb : 0; }
int32_t tlsb_last_tiled_widths(const tlsb_handle *h) { return h ? (h->layout.tiled ? h->layout.n_tiled : h->nU) : 0; }

int tlsb_last_sort_info(tlsb_handle *h, int32_t *segment_capacity, int32_t *n_segments, int64_t *global_sort_periods)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_last_sort_info: NULL handle");
    if (segment_capacity) *segment_capacity = h->layout.seg_cap;
    if (n_segments) *n_segments = h->layout.n_seg;
    if (global_sort_periods) {
        CUDA_TRY(cudaSetDevice(h->device));
        int v = 0;
        CUDA_TRY(cudaMemcpy(&v, h->counter.as<int>() + 4, 4, cudaMemcpyDeviceToHost));  // synchronises
        *global_sort_periods = v;
    }
    return 0;
}
int64_t tlsb_plan_fallback_count(const tlsb_handle *h) { return h ? h->fallbacks : 0; }
int64_t tlsb_plan_repair_count(const tlsb_handle *h) { return h ? h->repairs : 0; }

int tlsb_resolve_plan(tlsb_handle *h, void *cuda_stream, void *records_dev)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_resolve_plan: NULL handle");
    if (!h->have_lc || !h->have_tp || !h->have_periods || h->P == 0)
        return fail(TLSB_ERR_STATE, "tlsb_resolve_plan: nothing has been searched");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (!records_dev) records_dev = h->out.p;
    if (!records_dev) return fail(TLSB_ERR_STATE, "tlsb_resolve_plan: no record buffer");
    long long count = 0;
    CUDA_TRY(cudaMemcpyAsync(&count, reinterpret_cast<double *>(records_dev) + 3 * (size_t)h->P, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (count == 0) return 0;
    return resolve_records(h, s, records_dev, count);
}

int tlsb_last_layout(const tlsb_handle *h, int32_t *threads, int32_t *ctas_per_sm, int32_t *queue_capacity,
                     int64_t *smem_bytes)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_last_layout: NULL handle");
    if (threads) *threads = h->layout.threads;
    if (ctas_per_sm) *ctas_per_sm = h->layout.ctas_per_sm;
    if (queue_capacity) *queue_capacity = h->layout.qcap;
    if (smem_bytes) *smem_bytes = (int64_t)h->layout.smem;
    return 0;
}

double tlsb_last_search_kernel_ms(tlsb_handle *h)
{
    if (!h || !h->timed) return 0.0;
    cudaSetDevice(h->device);
    float ms = 0.f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess || cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) {
        cudaGetLastError();
        return 0.0;
    }
    return (double)ms;
}

// The one-shot entry point keeps one handle per device alive between calls (device buffers,
// events), so that a second search of similar size pays no cudaMalloc/cudaFree.
static std::mutex g_pool_mutex;
static std::vector<std::pair<int, tlsb_handle *>> g_pool;  // (device, idle handle)

static tlsb_handle *pool_take(int device)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (size_t k = 0; k < g_pool.size(); ++k)
        if (g_pool[k].first == device) {
            tlsb_handle *h = g_pool[k].second;
            g_pool.erase(g_pool.begin() + (long)k);
            return h;
        }
    return nullptr;
}

static void pool_give(tlsb_handle *h)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool.emplace_back(h->device, h);
}

static int search_on_device(int device, const tlsb_lightcurve *lc, const double *periods, int64_t nP,
                            const tlsb_templates *tp, const tlsb_params *prm, double *chi2, int64_t *row,
                            double *depth, int64_t *t0, std::string *err)
{
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        g_error = "no CUDA device available (this library has no CPU fallback)";
        if (err) *err = g_error;
        return TLSB_ERR_CUDA;
    }
    tlsb_handle *h = pool_take(device);
    int rc = h ? 0 : tlsb_create(&h, device);
    if (!rc) h->defer_sync = true;  // every input buffer outlives this call: one synchronisation, at the end
    if (!rc) rc = tlsb_set_lightcurve(h, lc);
    if (!rc) rc = tlsb_set_templates(h, tp, prm);
    if (!rc) rc = tlsb_set_periods(h, periods, nP);
    if (!rc) rc = tlsb_search_async(h, nullptr, nullptr);
    if (!rc) rc = tlsb_get_results(h, nullptr, chi2, row, depth, t0);
    if (h) h->defer_sync = false;
    if (rc && h) cudaStreamSynchronize(nullptr);
    if (rc && err) *err = g_error;
    if (rc)
        tlsb_destroy(h);  // do not recycle a handle that failed
    else
        pool_give(h);
    return rc;
}

int tlsb_search_periods(const tlsb_lightcurve *lc, const double *periods, int64_t n_periods,
                        const tlsb_templates *tp, const tlsb_params *prm, const tlsb_exec *ex,
                        double *chi2_out, int64_t *row_out, double *depth_out, int64_t *t0_index_out)
{
    if (!lc || !tp || !prm || (!periods && n_periods > 0) || !chi2_out || !row_out || !depth_out)
        return fail(TLSB_ERR_ARG, "tlsb_search_periods: NULL argument");
    std::vector<int> devs;
    if (ex && ex->devices && ex->n_devices > 0) devs.assign(ex->devices, ex->devices + ex->n_devices);
    if (devs.size() <= 1) {
        std::string err;
        int rc = search_on_device(devs.empty() ? -1 : devs[0], lc, periods, n_periods, tp, prm, chi2_out,
                                  row_out, depth_out, t0_index_out, &err);
        if (rc) g_error = err;
        return rc;
    }
    // several GPUs from one process: deal the periods round-robin, one host thread per GPU
    const int G = (int)devs.size();
    std::vector<std::vector<double>> per(G);
    std::vector<std::vector<int64_t>> where(G);
    for (int64_t p = 0; p < n_periods; ++p) {
        per[p % G].push_back(periods[p]);
        where[p % G].push_back(p);
    }
    std::vector<int> rcs(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> pool;
    for (int g = 0; g < G; ++g) {
        pool.emplace_back([&, g]() {
            const size_t n = per[g].size();
            std::vector<double> c(n), d(n);
            std::vector<int64_t> r(n), t0(n);
            rcs[g] = search_on_device(devs[g], lc, per[g].data(), (int64_t)n, tp, prm, c.data(), r.data(),
                                      d.data(), t0.data(), &errs[g]);
            if (rcs[g]) return;
            for (size_t k = 0; k < n; ++k) {
                const int64_t p = where[g][k];
                chi2_out[p] = c[k];
                row_out[p] = r[k];
                depth_out[p] = d[k];
                if (t0_index_out) t0_index_out[p] = t0[k];
            }
        });
    }
    for (auto &th : pool) th.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return fail(rcs[g], errs[g]);
    return 0;
}

}  // extern "C"

// ---- final_T0_fit -------------------------------------------------------------------------
namespace {

template <typename K>
cudaError_t launch_t0(K kernel, const T0Args &a, int grid, int threads, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

int run_t0_fit(tlsb_handle *h, cudaStream_t s, const double *model_in, int64_t dur, double period,
               const double *trials, int64_t n_trials, double *residuals_out, int64_t *best_index_out)
{
    if (!h->have_lc) return fail(TLSB_ERR_STATE, "tlsb_final_t0_fit: set the light curve first");
    if (!model_in || !trials || !residuals_out) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit: NULL argument");
    const int N = h->N;
    if (dur < 1 || dur > N) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit: need 1 <= dur <= n");
    if (n_trials < 1 || n_trials > (int64_t)1 << 30) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit: bad trial count");
    if (!(period > 0)) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit: period must be positive");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = upload(h->t0_trials, trials, sizeof(double) * (size_t)n_trials, s))) return rc;
    if ((rc = upload(h->t0_model, model_in, sizeof(double) * (size_t)dur, s))) return rc;
    if (h->t0_resid.ensure(sizeof(double) * (size_t)n_trials)) return fail(TLSB_ERR_ALLOC, "device allocation failed");

    T0Args a{};
    const size_t cur_off = (size_t)h->cur * (size_t)N;
    a.t = h->t.as<double>() + (h->shared_t ? 0 : cur_off); a.y = h->y.as<double>() + cur_off; a.N = N;
    a.trials = h->t0_trials.as<double>(); a.n_trials = (int)n_trials;
    a.model = h->t0_model.as<double>(); a.dur = (int)dur; a.shift = (int)(dur / 2) + 1;  // stats.py:186
    a.period = period;
    a.residuals = h->t0_resid.as<double>();
    a.counter = h->counter.as<int>() + 2;

    // layout: sort keys + sorted flux (+ ids, histogram) in shared memory when they fit
    const size_t n_even = ((size_t)N + 1) & ~(size_t)1;
    bool resident = false;
    int threads = 256, per_sm = 2;
    size_t smem = 0;
    if (N < 65536) {
        const int tries[2][2] = {{256, 2}, {512, 1}};
        for (const auto &t : tries) {
            const size_t tail = (size_t)(t[0] / 32 + 2) * 8 + 32;
            const size_t bytes = 2 * n_even * 8 + (size_t)((N + 2) & ~1) * 4 + align16((size_t)N * 2) + tail;
            if (bytes > h->max_smem || (bytes + 1024) * (size_t)t[1] > h->smem_per_sm) continue;
            resident = true; threads = t[0]; per_sm = t[1]; smem = bytes; a.NB = N;
            break;
        }
    }
    const int grid = (int)std::min<int64_t>(n_trials, (int64_t)h->num_sms * per_sm);
    if (!resident) {
        threads = 256; per_sm = 2;
        const size_t tail = (size_t)(threads / 32 + 2) * 8 + 32;
        const size_t per_cta = std::min(h->max_smem, h->smem_per_sm / 2 - 1024);
        a.NB = (int)std::min<size_t>((size_t)N, (per_cta - tail - 64) / 4 - 2);
        smem = align16((size_t)(a.NB + 1) * 4) + tail;
        a.scratch_per_cta = (2 * n_even * 8 + (size_t)N * 4 + 255) & ~(size_t)255;
        if (h->scratch.ensure(a.scratch_per_cta * (size_t)grid)) return fail(TLSB_ERR_ALLOC, "device allocation failed (scratch)");
        a.scratch = h->scratch.as<unsigned char>();
    }
    h->t0_resident = resident;
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    if (resident && threads == 256) CUDA_TRY(launch_t0(tlsb_t0fit_kernel<256, true>, a, grid, 256, smem, s));
    else if (resident) CUDA_TRY(launch_t0(tlsb_t0fit_kernel<512, true>, a, grid, 512, smem, s));
    else CUDA_TRY(launch_t0(tlsb_t0fit_kernel<256, false>, a, grid, 256, smem, s));
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    CUDA_TRY(cudaMemcpyAsync(residuals_out, h->t0_resid.p, sizeof(double) * (size_t)n_trials, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->t0_ms = ms; else cudaGetLastError();
    h->timed = false;
    if (best_index_out) {  // stats.py:200-202: strict '<' from +inf, so the first minimum wins and NaN never does
        int64_t best = -1;
        double lowest = INFINITY;
        for (int64_t k = 0; k < n_trials; ++k)
            if (residuals_out[k] < lowest) { lowest = residuals_out[k]; best = k; }
        *best_index_out = best;
    }
    return 0;
}

}  // namespace

extern "C" {

int tlsb_final_t0_fit(tlsb_handle *h, void *cuda_stream, const double *model_in, int64_t dur, double period,
                      const double *trials, int64_t n_trials, double *residuals_out, int64_t *best_index_out)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit: NULL handle");
    return run_t0_fit(h, reinterpret_cast<cudaStream_t>(cuda_stream), model_in, dur, period, trials, n_trials,
                      residuals_out, best_index_out);
}

double tlsb_last_t0_fit_ms(const tlsb_handle *h) { return h ? h->t0_ms : 0.0; }

int tlsb_final_t0_fit_lc(const tlsb_lightcurve *lc, int32_t device, const double *model_in, int64_t dur,
                         double period, const double *trials, int64_t n_trials, double *residuals_out,
                         int64_t *best_index_out)
{
    if (!lc) return fail(TLSB_ERR_ARG, "tlsb_final_t0_fit_lc: NULL light curve");
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        return fail(TLSB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    tlsb_handle *h = pool_take(device);
    int rc = h ? 0 : tlsb_create(&h, device);
    if (!rc) rc = tlsb_set_lightcurve(h, lc);
    if (!rc) rc = run_t0_fit(h, nullptr, model_in, dur, period, trials, n_trials, residuals_out, best_index_out);
    if (rc) {
        std::string keep = g_error;
        tlsb_destroy(h);
        g_error = keep;
    } else
        pool_give(h);
    return rc;
}

}  // extern "C"

// ---- batch pipeline: every resident light curve through plan/search, then spectra, one sync ----
namespace {

__global__ void tlsb_gather_rows_kernel(const double *__restrict__ records, size_t record_stride,
                                        const int *__restrict__ order, double *__restrict__ out, int P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t c = blockIdx.y;
    if (k < P) out[c * P + k] = records[c * record_stride + order[k]];  // chi2 plane, ascending period
}

}  // namespace

extern "C" int tlsb_search_batch(tlsb_handle *h, void *cuda_stream, int64_t median_window, double *chi2_out,
                                 int64_t *row_out, double *depth_out, int64_t *t0_index_out, double *power_out,
                                 double *SDE_raw_out, double *SDE_out, int64_t *best_period_index_out)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_search_batch: NULL handle");
    if (!h->have_lc || !h->have_tp || !h->have_periods)
        return fail(TLSB_ERR_STATE, "tlsb_search_batch: light curves, templates and periods must be set first");
    if (!SDE_raw_out || !SDE_out) return fail(TLSB_ERR_ARG, "tlsb_search_batch: NULL argument");
    if (median_window < 1 || median_window > 24000) return fail(TLSB_ERR_ARG, "tlsb_search_batch: median window must be in 1..24000");
    if (h->P < 1) return fail(TLSB_ERR_ARG, "tlsb_search_batch: no periods");
    if (h->M > h->N) return fail(TLSB_ERR_ARG, "widest template is longer than the light curve");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    const size_t B = (size_t)h->n_curves, P = (size_t)h->P, stride = 3 * P + 1;
    if (h->brec.ensure(B * stride * 8) || h->bchi.ensure(B * P * 8) || h->bSR.ensure(B * P * 8) ||
        h->bpr.ensure(B * P * 8) || h->bpw.ensure(B * P * 8) || h->bscal.ensure(B * 32) || h->bamax.ensure(B * 8))
        return fail(TLSB_ERR_ALLOC, "device allocation failed (batch buffers)");
    if (!h->asc_valid) {  // main.py:190-196: the spectra consume chi2 in ascending-period order
        const std::vector<double> &per = h->h_periods;
        h->h_asc_order.resize(P);
        std::iota(h->h_asc_order.begin(), h->h_asc_order.end(), 0);
        std::stable_sort(h->h_asc_order.begin(), h->h_asc_order.end(), [&](int x, int y) { return per[x] < per[y]; });
        int rc0 = upload(h->asc_order, h->h_asc_order.data(), sizeof(int) * P, s);
        if (rc0) return rc0;
        h->asc_valid = true;
    }
    double *rec = h->brec.as<double>();
    std::vector<long long> status(B);
    int64_t launches = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool exact = attempt == 1 || h->plan_mode == 1;
        int rc, n_plans = 0;
        for (size_t c = 0; c < B; ++c) {
            if ((rc = tlsb_select_lightcurve(h, (int64_t)c))) return rc;
            if ((rc = enqueue_search(h, s, rec + c * stride, exact))) return rc;
            launches += h->launches;
            n_plans += h->launches == 2 ? 1 : 0;
            // the device plan of this launch serves the following curves while span/periods/bank stay the same
            if (!exact && h->plan_mode == 0) h->dev_plan_valid = true;
        }
        // one strided copy of the B status words
        CUDA_TRY(cudaMemcpy2DAsync(status.data(), 8, rec + 3 * P, stride * 8, 8, B, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        long long flagged = 0;
        for (size_t c = 0; c < B; ++c) flagged = std::max(flagged, status[c]);
        if (flagged == 0 || exact) break;
        // Some T14 limit was too close to an integer for the device pow().  One shared plan: settle the
        // flagged periods on the host and search again only those whose admissible widths differ, for
        // every curve.  Several plans in the batch (different spans): the exact host plan, everything again.
        h->dev_plan_valid = false;
        bool same_span = true;  // every plan of this batch is the same plan
        for (size_t c = 1; c < B; ++c) same_span = same_span && h->c_span[c] == h->c_span[0];
        if ((n_plans == 1 || same_span) && status[0] == flagged) {
            std::vector<int> changed;
            rc = find_changed_periods(h, s, flagged, &changed);
            if (rc < 0) return rc;
            if (rc == 0) {
                for (size_t c = 0; c < B && !changed.empty(); ++c) {
                    if ((rc = tlsb_select_lightcurve(h, (int64_t)c))) return rc;
                    if ((rc = enqueue_search(h, s, rec + c * stride, false, h->unsure.as<int>(), (int)changed.size()))) return rc;
                    launches += h->launches;
                }
                h->repairs += (int64_t)changed.size();
                CUDA_TRY(cudaStreamSynchronize(s));
                break;
            }
        }
        h->fallbacks += 1;
    }
    h->dev_plan_valid = false;  // do not carry the shortcut outside the batch
    dim3 grid((unsigned)((P + 255) / 256), (unsigned)B);
    tlsb_gather_rows_kernel<<<grid, 256, 0, s>>>(rec, stride, h->asc_order.as<int>(), h->bchi.as<double>(), (int)P);
    CUDA_TRY(cudaGetLastError());
    int rc = tlsb::spectra_device(h->bchi.as<double>(), (int64_t)P, (int64_t)B, median_window, h->bSR.as<double>(),
                                  h->bpr.as<double>(), h->bpw.as<double>(), h->bscal.as<double>(),
                                  h->bamax.as<long long>(), s);
    if (rc) return rc;
    launches += 1 + (P > 2 * (size_t)median_window ? 3 : 2);
    h->launches = launches;
    std::vector<double> scal(B * 4);
    std::vector<long long> amax(B), packed;
    CUDA_TRY(cudaMemcpyAsync(scal.data(), h->bscal.p, B * 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(amax.data(), h->bamax.p, B * 8, cudaMemcpyDeviceToHost, s));
    if (chi2_out) CUDA_TRY(cudaMemcpy2DAsync(chi2_out, P * 8, rec, stride * 8, P * 8, B, cudaMemcpyDeviceToHost, s));
    if (depth_out) CUDA_TRY(cudaMemcpy2DAsync(depth_out, P * 8, rec + P, stride * 8, P * 8, B, cudaMemcpyDeviceToHost, s));
    if (row_out || t0_index_out) {
        packed.resize(B * P);
        CUDA_TRY(cudaMemcpy2DAsync(packed.data(), P * 8, rec + 2 * P, stride * 8, P * 8, B, cudaMemcpyDeviceToHost, s));
    }
    if (power_out) CUDA_TRY(cudaMemcpyAsync(power_out, h->bpw.p, B * P * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (size_t k = 0; k < packed.size(); ++k) {
        if (row_out) row_out[k] = (int64_t)(uint32_t)(packed[k] & 0xffffffffLL);
        if (t0_index_out) t0_index_out[k] = (int64_t)(int32_t)(packed[k] >> 32);
    }
    for (size_t c = 0; c < B; ++c) {
        SDE_raw_out[c] = scal[4 * c + 0];
        SDE_out[c] = scal[4 * c + 1];
        if (best_period_index_out) best_period_index_out[c] = h->h_asc_order[(size_t)amax[c]];
    }
    return 0;
}
