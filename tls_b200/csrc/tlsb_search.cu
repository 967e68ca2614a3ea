// tlsb_search.cu — the TLS period/duration/T0 grid search as sm_100a CUDA + its C ABI.
//
// What one trial period costs in the reference (core.py:96-188, per period):
//   fold (core.py:15-18) -> stable argsort (core.py:120) -> gathers (:121-123) -> patch (:126-132)
//   -> per admissible duration W: running_mean (helpers.py:70-73), out_of_transit_residuals
//   (core.py:79-93), lowest_residuals_in_this_duration (core.py:28-76) -> min over durations.
//
// Here ONE persistent CTA handles one period at a time, entirely on chip when the folded
// curve fits shared memory ("resident" path) or through a per-CTA global scratch that stays
// in L2 ("streaming" path, any N):
//
//   A. fold in fp64 with the reciprocal-multiply form numba emits, bucket-rank the phases
//      (histogram -> scan -> scatter -> rank inside the bucket by (phase, index): a stable
//      sort), gather d = 1-y and w = 1/dy^2 straight to their sorted slots, wrap the first M
//      samples to the end, block-scan d into cumulative sums, block-reduce T = sum w d^2.
//   B. warp-autonomous sweep over the admissible widths.  With d = 1-y, D = mean*overshoot and
//      q_j = (1-signal_j)/SIGNAL_DEPTH the reference's statistic is algebraically
//          chi2_i(W) = T + D^2 * sum_j q_j^2 w_{i+j} - 2 D * sum_j q_j (w d)_{i+j}
//                        - sum_{k=L..W-1} (w d^2)_{i+k}
//      (the sum over the window of w d^2 cancels against out_of_transit_residuals and the edge
//      correction).  Each warp evaluates the gate mean_i > transit_depth_min
//      from two cumulative-sum reads for 32 candidate offsets, ballot-compacts the survivors
//      into a small shared-memory queue and runs the tap loop only on full warps of survivors.
//   C. lexicographic (chi2, width order, offset) block arg-min = the reference's strict-<
//      tie rules (core.py:71, :183), sentinel N / +inf handling (core.py:46, :139-140).
//
// No tensor cores: there is no dense contraction here (per-offset depth, gate and stride).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/tlsb200.h"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBlock = 5;             // R: consecutive T0 candidates one lane carries through the tap loop
                                      // (odd: neighbouring lanes sit R*stride doubles apart in shared memory)
constexpr int kTile = 32 * kBlock;    // candidates one warp gates per scheduler grab
constexpr int kQueue = 4096;          // CTA-wide survivor queue (blocks of kBlock candidates)
constexpr int kQueueStop = kQueue - kWarps * 32;  // gating pauses here: every warp can still add a tile
constexpr int kScanItems = 5;         // items per thread per scan tile (odd: conflict-free in smem)
constexpr unsigned kFull = 0xffffffffu;
constexpr double kSignalDepth = 0.5;  // tls_constants.py:71

thread_local std::string g_error;

int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}

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
    int pad;              // readable slack behind the patched arrays (kBlock * max stride)
    // periods
    const double *periods;
    const int *ulo;       // [P] admissible unique-width index range [ulo, uhi)
    const int *uhi;
    const int *order;     // [P] processing order (most expensive first)
    int P;
    double depth_min;
    // outputs: three planes of P 8-byte words
    double *out_chi2;
    double *out_depth;
    long long *out_packed;
    // scheduling
    double w0;            // the common weight 1/dy^2 when every dy is the same (dy=None), else unused
    int *counter;         // [2] next period, finished CTAs
    // streaming path scratch
    unsigned char *scratch;
    size_t scratch_per_cta;
    int NB;               // number of phase buckets
};

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
template <typename T>
__device__ void block_inclusive_scan(T *data, int n, T *warp_tot /* [kWarps+1] shared */)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    T carry = T(0);
    for (int base = 0; base < n; base += kThreads * kScanItems) {
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
            T wv = (lane < kWarps) ? warp_tot[lane] : T(0);
            T wi = wv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                T o = __shfl_up_sync(kFull, wi, off);
                if (lane >= off) wi += o;
            }
            if (lane < kWarps) warp_tot[lane] = wi - wv;  // exclusive warp offsets
            if (lane == kWarps - 1) warp_tot[kWarps] = wi; // tile total
        }
        __syncthreads();
        T excl = __shfl_up_sync(kFull, incl, 1);  // exclusive prefix of this thread inside its warp
        if (lane == 0) excl = T(0);
        const T offset = carry + warp_tot[wid] + excl;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (first + k < n) data[first + k] = v[k] + offset;
        carry += warp_tot[kWarps];
        __syncthreads();
    }
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
// of kBlock (loads of a whole group can be in flight together): templates are zero padded in
// tq and the patched arrays have slack behind them, so ramp-in/ramp-out need no predicates.
// kUniformW: all weights equal (dy=None) -> only B = sum q_j (w d)_{i+j} is accumulated.
template <bool kUnit, bool kUniformW>
__device__ __forceinline__ void tap_block(const WidthRec &wr, const double *__restrict__ tq,
                                          const double *__restrict__ w, const double *__restrict__ wd,
                                          int i0, double (&A)[kBlock], double (&B)[kBlock])
{
    const int L = wr.L, X = kUnit ? 1 : wr.X;
#pragma unroll
    for (int r = 0; r < kBlock; ++r) { A[r] = 0.0; B[r] = 0.0; }
    const int nb = X < L ? X : L;
    for (int b = 0; b < nb; ++b) {
        const int groups = ((L - b + X - 1) / X + 2 * kBlock - 2) / kBlock;  // ceil((taps + kBlock-1) / kBlock)
        const double *__restrict__ qp = tq + wr.q + b;
        const double *__restrict__ wp = w + i0 + b;
        const double *__restrict__ wdp = wd + i0 + b;
        double qw[kBlock], pw[kBlock];  // circular: the value loaded at step m lives in slot m % kBlock
#pragma unroll
        for (int r = 0; r < kBlock; ++r) { qw[r] = 0.0; pw[r] = 0.0; }
#pragma unroll 1
        for (int g = 0; g < groups; ++g) {
            double qk[kBlock], wv[kBlock], wdv[kBlock];
#pragma unroll
            for (int mm = 0; mm < kBlock; ++mm) {
                qk[mm] = __ldg(qp + mm * X);
                wdv[mm] = wdp[mm * X];
                if (!kUniformW) wv[mm] = wp[mm * X];
            }
#pragma unroll
            for (int mm = 0; mm < kBlock; ++mm) {
                qw[mm] = qk[mm];
                if (!kUniformW) pw[mm] = qk[mm] * qk[mm];
#pragma unroll
                for (int r = 0; r < kBlock; ++r) {
                    const int slot = (mm - r + kBlock) % kBlock;  // loaded r steps ago
                    B[r] = fma(qw[slot], wdv[mm], B[r]);
                    if (!kUniformW) A[r] = fma(pw[slot], wv[mm], A[r]);
                }
            }
            qp += kBlock * X;
            wp += kBlock * X;
            wdp += kBlock * X;
        }
    }
}

// Samples L..W-1 of a window are in neither the in-transit nor the out-of-transit sum when a
// template was trimmed to L < W (rare: L == W for the limb-darkened templates).
__device__ __noinline__ double untouched_tail(const double *w, const double *wd, int from, int to)
{
    double rest = 0.0;
#pragma unroll 1
    for (int k = from; k < to; ++k) rest += wd[k] * wd[k] / w[k];  // w d^2
    return rest;
}

template <bool kResident, bool kUniformW>
__global__ void __launch_bounds__(kThreads, 1) tlsb_search_kernel(const __grid_constant__ SearchArgs a)
{
    using idx_t = typename std::conditional<kResident, unsigned short, unsigned int>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, M = a.M, NM = N + M, NB = a.NB, nU = a.nU;
    const int NMP = NM + a.pad;

    // ---- carve memory ------------------------------------------------------------------
    // cs  : (NM+1) doubles  cumulative sums of d (cs[0] = 0)
    // w   : NMP doubles
    // U   : union { wd : NMP doubles } / { skey : N doubles, sid : N idx, slot : N idx [, hist] }
    // hist: NB+1 ints  (inside U on the resident path, in shared memory on the streaming path)
    const size_t cs_elems = (size_t)(NM + 2) & ~(size_t)1;
    double *cs, *w, *wd, *skey;
    idx_t *sid, *slot;
    int *hist;
    unsigned char *tail;
    if (kResident) {
        cs = reinterpret_cast<double *>(smem_raw);
        w = cs + cs_elems;
        unsigned char *U = reinterpret_cast<unsigned char *>(w + NMP);
        wd = reinterpret_cast<double *>(U);
        skey = reinterpret_cast<double *>(U);
        hist = reinterpret_cast<int *>(skey + N);
        sid = reinterpret_cast<idx_t *>(hist + NB + 1);
        slot = sid + N;
        size_t sort_bytes = (size_t)N * 8 + (size_t)(NB + 1) * 4 + (size_t)N * 2 * sizeof(idx_t);
        size_t u_bytes = sort_bytes > (size_t)NMP * 8 ? sort_bytes : (size_t)NMP * 8;
        tail = U + ((u_bytes + 15) & ~(size_t)15);
    } else {
        unsigned char *g = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
        cs = reinterpret_cast<double *>(g);
        w = cs + cs_elems;
        unsigned char *U = reinterpret_cast<unsigned char *>(w + NMP);
        wd = reinterpret_cast<double *>(U);
        skey = reinterpret_cast<double *>(U);
        sid = reinterpret_cast<idx_t *>(skey + N);
        slot = sid + N;
        hist = reinterpret_cast<int *>(smem_raw);
        tail = smem_raw + (((size_t)(NB + 1) * 4 + 15) & ~(size_t)15);
    }
    WidthRec *rec = reinterpret_cast<WidthRec *>(tail);                       // [nU]
    double *red_d = reinterpret_cast<double *>(rec + nU);                     // [2*kWarps + 2]
    int *red_i = reinterpret_cast<int *>(red_d + 2 * kWarps + 2);             // [2*kWarps]
    int2 *queue = reinterpret_cast<int2 *>(red_i + 2 * kWarps);               // [kQueue]
    int *s_next = reinterpret_cast<int *>(queue + kQueue);  // [4] period slot, tile counter, queue fill, queue head

    for (int k = tid; k < nU * (int)(sizeof(WidthRec) / 4); k += kThreads)
        reinterpret_cast<int *>(rec)[k] = reinterpret_cast<const int *>(a.rec)[k];

    const unsigned lt_mask = (1u << lane) - 1u;
    const double depth_min = a.depth_min;

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
        for (int b = tid; b <= NB; b += kThreads) hist[b] = 0;
        __syncthreads();
        for (int k = tid; k < N; k += kThreads) {
            const double ph = fold_phase(a.t[k], r);
            slot[k] = (idx_t)atomicAdd(&hist[bucket_of(ph, NB)], 1);
        }
        __syncthreads();
        // exclusive scan: shift by one so hist[b] = number of keys in buckets < b
        block_inclusive_scan<int>(hist, NB, reinterpret_cast<int *>(red_d));
        // hist now inclusive; convert on the fly below: base(b) = b ? hist[b-1] : 0
        for (int k = tid; k < N; k += kThreads) {
            const double ph = fold_phase(a.t[k], r);
            const int b = bucket_of(ph, NB);
            const int pos = (b ? hist[b - 1] : 0) + (int)slot[k];
            skey[pos] = ph;
            sid[pos] = (idx_t)k;
        }
        __syncthreads();
        for (int q = tid; q < N; q += kThreads) {
            const double key = skey[q];
            const int id = (int)sid[q];
            const int b = bucket_of(key, NB);
            const int lo = b ? hist[b - 1] : 0, hi = hist[b];
            int rank = lo;
            for (int s = lo; s < hi; ++s) {
                const double ks = skey[s];
                const int is = (int)sid[s];
                rank += (ks < key) || (ks == key && is < id);
            }
            cs[rank + 1] = a.dval[id];   // d, scanned in place below
            w[rank] = a.wval[id];
        }
        __syncthreads();  // sort scratch is dead from here; wd may overwrite it
        // wrap the first M samples to the end (core.py:126-132), build w*d, reduce T = sum w d^2
        double tpart = 0.0;
        for (int k = tid; k < NMP; k += kThreads) {
            if (k < NM) {
                const int src = k < N ? k : k - N;
                const double d = cs[src + 1], wv = w[src];
                const double x = wv * d;
                if (k >= N) { cs[k + 1] = d; w[k] = wv; }
                wd[k] = x;
                if (k < N) tpart = fma(x, d, tpart);
            } else {  // slack read (never used) by partially valid candidate blocks
                w[k] = 0.0;
                wd[k] = 0.0;
            }
        }
        if (tid == 0) cs[0] = 0.0;
#pragma unroll
        for (int off = 16; off; off >>= 1) tpart += __shfl_xor_sync(kFull, tpart, off);
        if (lane == 0) red_d[kWarps + 1 + wid] = tpart;
        __syncthreads();
        block_inclusive_scan<double>(cs + 1, NM, red_d);
        double T = 0.0;
        for (int k = 0; k < kWarps; ++k) T += red_d[kWarps + 1 + k];  // fixed order: deterministic

        // ---- B. gate + survivor compaction + tap loop ----------------------------------------
        // B1: warps grab tiles of kTile candidates of one width from a CTA-wide counter (wide
        //     widths first), gate them from two cumulative-sum reads per candidate and append
        //     the surviving blocks to a CTA-wide queue (one reservation per tile, so a tile's
        //     survivors stay together and the queue is nearly sorted by width).
        // B2: warps grab 32 consecutive queue entries - almost always one width, so template
        //     loads broadcast and the lanes run in lockstep - and run the tap loop.
        // The queue is bounded; B1/B2 alternate until all tiles are gated.
        Best best;
        best.chi2 = (double)N;  // core.py:46: a model must beat N to count
        best.D = 0.0;
        best.u = -1;  // "no model yet": loses every tie, so a candidate must be strictly below N
        best.i = -1;

        const int tile_base = rec[uhi - 1].cum;
        const int tile_end = rec[ulo].cum + rec[ulo].tiles;
        int cur_u = uhi - 1;
        for (;;) {
            // B1
            for (;;) {
                int g = tile_end;
                if (lane == 0 && *(volatile int *)&s_next[2] < kQueueStop) g = atomicAdd(&s_next[1], 1) + tile_base;
                g = __shfl_sync(kFull, g, 0);
                if (g >= tile_end) break;
                while (g >= rec[cur_u].cum + rec[cur_u].tiles) --cur_u;
                const int u = cur_u;
                const int W = rec[u].W, X = rec[u].X, ncand = rec[u].ncand;
                const double invW = rec[u].invW;
                const int c0 = (g - rec[u].cum) * kTile + lane * kBlock;
                int mask = 0;
#pragma unroll
                for (int rr = 0; rr < kBlock; ++rr) {
                    const int c = c0 + rr;
                    if (c < ncand) {
                        const int i = c * X;
                        const double mean = (cs[i + W] - cs[i]) * invW;
                        if (mean > depth_min) mask |= 1 << rr;  // core.py:58 (the stride is built into c)
                    }
                }
                const unsigned m = __ballot_sync(kFull, mask != 0);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_next[2], __popc(m));
                    base = __shfl_sync(kFull, base, 0);
                    if (mask) queue[base + __popc(m & lt_mask)] = make_int2(c0, u | (mask << 16));
                }
            }
            __syncthreads();
            const int qfill = s_next[2];
            const bool done = s_next[1] + tile_base >= tile_end;  // every tile has been handed out
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
                    if (wr.X == 1)
                        tap_block<true, kUniformW>(wr, a.tq, w, wd, i0, A, B);
                    else
                        tap_block<false, kUniformW>(wr, a.tq, w, wd, i0, A, B);
#pragma unroll
                    for (int rr = 0; rr < kBlock; ++rr) {
                        if (mask & (1 << rr)) {
                            const int i = i0 + rr * wr.X;
                            const double mean = (cs[i + wr.W] - cs[i]) * wr.invW;
                            const double D = mean * wr.os;
                            const double Aq = kUniformW ? a.w0 * wr.sq2 : A[rr];
                            double chi = T + D * (D * Aq - 2.0 * B[rr]);
                            if (wr.L < wr.W) chi -= untouched_tail(w, wd, i + wr.L, i + wr.W);
                            if (better(chi, u, i, best)) { best.chi2 = chi; best.D = D; best.u = u; best.i = i; }
                        }
                    }
                }
            }
            if (done) break;
            __syncthreads();  // everyone has left B2 before the queue is reused
            if (tid == 0) { s_next[2] = 0; s_next[3] = 0; }
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
        __syncthreads();  // everyone is done reading red_d (T) before it is reused
        if (lane == 0) {
            red_d[wid] = best.chi2;
            red_d[kWarps + wid] = best.D;
            red_i[wid] = best.u;
            red_i[kWarps + wid] = best.i;
        }
        __syncthreads();
        if (wid == 0) {
            Best b2;
            b2.chi2 = (double)N; b2.D = 0.0; b2.u = -1; b2.i = -1;
            if (lane < kWarps) {
                b2.chi2 = red_d[lane];
                b2.D = red_d[kWarps + lane];
                b2.u = red_i[lane];
                b2.i = red_i[kWarps + lane];
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

// d = 1 - y, w = 1/dy^2 (core.py:127 computes 1/dy**2 the same way), once per light curve.
__global__ void tlsb_prepare_kernel(const double *__restrict__ y, const double *__restrict__ dy,
                                    double *__restrict__ dval, double *__restrict__ wval, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
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

// tls_constants.py:20-25,78 and grid.py:9-32 (T14)
double t14_fraction(double R_s, double M_s, double P, bool small)
{
    const double G = 6.673e-11, R_sun = 695508000.0, R_jup = 69911000.0, M_sun = 1.989e30;
    const double Ps = P * 86400.0, R = R_sun * R_s, Ms = M_sun * M_s;
    const double cube = std::pow((4 * Ps) / (M_PI * G * Ms), 1.0 / 3);
    const double t14 = small ? R * cube : (R + 2 * R_jup) * cube;
    const double frac = t14 / Ps;
    return frac > 0.12 ? 0.12 : frac;
}

}  // namespace

struct tlsb_handle {
    int device = 0;
    int num_sms = 0;
    size_t max_smem = 0;
    // light curve
    int N = 0;
    double span = 0.0;
    bool uniform_w = false;  // every dy identical (dy=None -> std(y) everywhere, validate.py:39-40)
    double w0 = 0.0;         // 1/dy^2 in that case
    DevBuf t, y, dy, dval, wval;
    bool have_lc = false;
    // templates
    tlsb_params prm{};
    int nU = 0, M = 0, pad = 0;
    std::vector<WidthRec> recs;   // unique widths, ascending
    DevBuf tq, d_rec;
    bool have_tp = false;
    // periods
    int P = 0;
    std::vector<double> h_periods;
    DevBuf periods, ulo, uhi, order;
    bool have_periods = false;
    bool periods_stale = true;  // admissible ranges depend on light curve + templates + params
    // outputs / scheduling / scratch
    DevBuf out, counter, scratch;
    // bookkeeping
    int64_t launches = 0;
    bool resident = false;
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

// admissible unique-width range per period (core.py:143-156) + processing order
int refresh_periods(tlsb_handle *h)
{
    const int P = h->P, N = h->N;
    // candidates and scheduler tiles per width (depend on N + M), wide -> narrow prefix
    int cum = 0;
    for (int u = h->nU - 1; u >= 0; --u) {
        WidthRec &wr = h->recs[u];
        wr.ncand = (N + h->M - wr.W) / wr.X + 1;  // offsets i = c*X, i in [0, N+M-W]
        wr.tiles = (wr.ncand + kTile - 1) / kTile;
        wr.cum = cum;
        cum += wr.tiles;
    }
    {
        int rc0;
        if ((rc0 = upload(h->d_rec, h->recs.data(), sizeof(WidthRec) * (size_t)h->nU))) return rc0;
    }
    std::vector<int> lo(P), hi(P), order(P);
    for (int p = 0; p < P; ++p) {
        const double period = h->h_periods[p];
        const double dmax = t14_fraction(h->prm.R_star_max, h->prm.M_star_max, period, false);
        const double dmin = t14_fraction(h->prm.R_star_min, h->prm.M_star_min, period, true);
        const double naive = h->span / period;
        const double corr = (naive + 1) / naive;
        const double wmin_f = std::floor(dmin * (double)N);
        const double wmax_f = std::ceil(dmax * (double)N * corr);
        int a = 0;
        while (a < h->nU && (double)h->recs[a].W < wmin_f) ++a;
        int b = h->nU;
        while (b > a && (double)h->recs[b - 1].W > wmax_f) --b;
        if (!(wmax_f >= wmin_f)) b = a;  // NaN / empty
        lo[p] = a;
        hi[p] = b;
    }
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
    h->periods_stale = false;
    return 0;
}

size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

size_t tail_bytes(int nU)
{
    return (size_t)nU * sizeof(WidthRec) + (2 * kWarps + 2) * 8 + 2 * kWarps * 4 + (size_t)kQueue * 8 + 32;
}

size_t resident_smem_bytes(int N, int M, int pad, int NB, int nU)
{
    const size_t NM = (size_t)N + M, NMP = NM + pad;
    const size_t cs = ((NM + 2) & ~(size_t)1) * 8, w = NMP * 8;
    const size_t sort_bytes = (size_t)N * 8 + (size_t)(NB + 1) * 4 + (size_t)N * 2 * 2;
    const size_t u = std::max(sort_bytes, NMP * 8);
    return cs + w + align16(u) + tail_bytes(nU);
}

size_t streaming_scratch_bytes(int N, int M, int pad)
{
    const size_t NM = (size_t)N + M, NMP = NM + pad;
    const size_t cs = ((NM + 2) & ~(size_t)1) * 8, w = NMP * 8;
    const size_t sort_bytes = (size_t)N * 8 + (size_t)N * 2 * 4;
    const size_t u = std::max(sort_bytes, NMP * 8);
    return (cs + w + align16(u) + 255) & ~(size_t)255;
}

}  // namespace

extern "C" {

const char *tlsb_last_error(void) { return g_error.c_str(); }
const char *tlsb_version(void) { return "tlsb200 0.1 (sm_100a)"; }

int32_t tlsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
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
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    if (h->counter.ensure(16)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    CUDA_TRY(cudaMemset(h->counter.p, 0, 16));
    *out = h;
    return 0;
}

int tlsb_destroy(tlsb_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (DevBuf *b : {&h->t, &h->y, &h->dy, &h->dval, &h->wval, &h->tq, &h->d_rec, &h->periods, &h->ulo, &h->uhi,
                      &h->order, &h->out, &h->counter, &h->scratch})
        b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}

int tlsb_set_lightcurve(tlsb_handle *h, const tlsb_lightcurve *lc)
{
    if (!h || !lc || !lc->t || !lc->y || !lc->dy) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurve: NULL argument");
    if (lc->n < 3 || lc->n > (int64_t)1 << 28) return fail(TLSB_ERR_ARG, "tlsb_set_lightcurve: need 3 <= n <= 2^28 samples");
    CUDA_TRY(cudaSetDevice(h->device));
    const int n = (int)lc->n;
    const size_t bytes = sizeof(double) * (size_t)n;
    int rc;
    if ((rc = upload(h->t, lc->t, bytes))) return rc;
    if ((rc = upload(h->y, lc->y, bytes))) return rc;
    if ((rc = upload(h->dy, lc->dy, bytes))) return rc;
    if (h->dval.ensure(bytes) || h->wval.ensure(bytes)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    tlsb_prepare_kernel<<<(n + 255) / 256, 256>>>(h->y.as<double>(), h->dy.as<double>(),
                                                  h->dval.as<double>(), h->wval.as<double>(), n);
    CUDA_TRY(cudaGetLastError());
    double tmin = lc->t[0], tmax = lc->t[0];  // core.py:148: max(t) - min(t)
    bool uniform = true;
    for (int k = 1; k < n; ++k) {
        tmin = std::min(tmin, lc->t[k]);
        tmax = std::max(tmax, lc->t[k]);
        uniform = uniform && lc->dy[k] == lc->dy[0];
    }
    h->uniform_w = uniform;
    h->w0 = 1.0 / (lc->dy[0] * lc->dy[0]);
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->N = n;
    h->span = tmax - tmin;
    h->have_lc = true;
    h->periods_stale = true;
    return 0;
}

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
        wr.ncand = 0;  // needs N: filled by refresh_periods
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
        for (int j = 0; j < xth * 2 * kBlock; ++j) tq.push_back(0.0);  // ramp-out of tap_block
    }
    int M = recs[nU - 1].W;  // core.py:114-116
    if (M % 2 != 0) M += 1;
    int rc;
    if ((rc = upload(h->tq, tq.data(), tq.size() * 8))) return rc;
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->recs.swap(recs);
    h->pad = 2 * kBlock * xmax;
    h->nU = nU;
    h->M = M;
    h->prm = *prm;
    h->have_tp = true;
    h->periods_stale = true;
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
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    h->have_periods = true;
    h->periods_stale = true;
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
    int rc;
    if (h->periods_stale && (rc = refresh_periods(h))) return rc;
    if (!records_dev) {
        if (h->out.ensure((size_t)h->P * 24)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
        records_dev = h->out.p;
    }

    SearchArgs a{};
    a.t = h->t.as<double>(); a.dval = h->dval.as<double>(); a.wval = h->wval.as<double>(); a.N = h->N;
    a.tq = h->tq.as<double>(); a.rec = h->d_rec.as<WidthRec>(); a.nU = h->nU; a.M = h->M; a.pad = h->pad;
    a.periods = h->periods.as<double>(); a.ulo = h->ulo.as<int>(); a.uhi = h->uhi.as<int>();
    a.order = h->order.as<int>(); a.P = h->P; a.depth_min = h->prm.transit_depth_min;
    a.out_chi2 = reinterpret_cast<double *>(records_dev);
    a.out_depth = a.out_chi2 + h->P;
    a.out_packed = reinterpret_cast<long long *>(a.out_depth + h->P);
    a.counter = h->counter.as<int>();

    const int grid = std::min(h->P, h->num_sms);
    const size_t need = resident_smem_bytes(h->N, h->M, h->pad, h->N, h->nU);
    const bool resident = h->N < 65536 && need <= h->max_smem;
    h->resident = resident;
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    a.w0 = h->w0;
    auto launch = [&](auto kernel, size_t smem) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<grid, kThreads, smem, s>>>(a);
        return cudaGetLastError();
    };
    if (resident) {
        a.NB = h->N;
        if (h->uniform_w)
            CUDA_TRY(launch(tlsb_search_kernel<true, true>, need));
        else
            CUDA_TRY(launch(tlsb_search_kernel<true, false>, need));
    } else {
        if (tail_bytes(h->nU) + 4096 > h->max_smem) return fail(TLSB_ERR_ARG, "too many distinct template widths for shared memory");
        const size_t budget = h->max_smem - tail_bytes(h->nU) - 64;
        a.NB = (int)std::min<size_t>((size_t)h->N, budget / 4 - 2);
        a.scratch_per_cta = streaming_scratch_bytes(h->N, h->M, h->pad);
        if (h->scratch.ensure(a.scratch_per_cta * (size_t)grid)) return fail(TLSB_ERR_ALLOC, "device allocation failed (scratch)");
        a.scratch = h->scratch.as<unsigned char>();
        const size_t smem = align16((size_t)(a.NB + 1) * 4) + tail_bytes(h->nU);
        if (h->uniform_w)
            CUDA_TRY(launch(tlsb_search_kernel<false, true>, smem));
        else
            CUDA_TRY(launch(tlsb_search_kernel<false, false>, smem));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev1, s));
    h->launches = 1;
    h->timed = true;
    return 0;
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
    std::vector<long long> packed(P);
    CUDA_TRY(cudaMemcpyAsync(chi2_out, h->out.p, P * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(depth_out, h->out.as<double>() + P, P * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(packed.data(), h->out.as<double>() + 2 * P, P * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (size_t p = 0; p < P; ++p) {
        row_out[p] = (int64_t)(uint32_t)(packed[p] & 0xffffffffLL);
        if (t0_index_out) t0_index_out[p] = (int64_t)(int32_t)(packed[p] >> 32);
    }
    return 0;
}

int64_t tlsb_last_launch_count(const tlsb_handle *h) { return h ? h->launches : 0; }
int32_t tlsb_last_path_resident(const tlsb_handle *h) { return h && h->resident ? 1 : 0; }

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
        if (err) *err = "no CUDA device available (this library has no CPU fallback)";
        g_error = *err;
        return TLSB_ERR_CUDA;
    }
    tlsb_handle *h = pool_take(device);
    int rc = h ? 0 : tlsb_create(&h, device);
    if (!rc) rc = tlsb_set_lightcurve(h, lc);
    if (!rc) rc = tlsb_set_templates(h, tp, prm);
    if (!rc) rc = tlsb_set_periods(h, periods, nP);
    if (!rc) rc = tlsb_search_async(h, nullptr, nullptr);
    if (!rc) rc = tlsb_get_results(h, nullptr, chi2, row, depth, t0);
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
