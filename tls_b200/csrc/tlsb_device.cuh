// tlsb_device.cuh — the device-side building blocks shared by the search kernels (resident / streaming:
// tlsb_resident.cu, tiled: tlsb_tiled.cu) and the T0-fit kernel (tlsb_aux_kernels.cu): fold, stable
// bucket-rank sort, scans, gate, register-blocked tap loop, lexicographic arg-min, bulk-copy helpers.
// Header-only, anonymous namespace: each translation unit gets its own copy.
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
//        Filter layouts (everything but the streaming path): the gate reads detrended fp32
//        cumulative sums (a superset passes: Gate32), B2 accumulates the correlation(s) in fp32
//        with rigorous error bounds (tap_block32 / tap_block32w), a screen and exact bounds
//        discard what cannot be the minimum, and the few candidates left are evaluated in fp64
//        by whole warps (eval_exact_warp): results are bit-identical to an all-fp64 evaluation.
//     C. lexicographic (chi2, width order, offset) block arg-min = the reference's strict-<
//        tie rules (core.py:71, :183), sentinel N / +inf handling (core.py:46, :139-140).
//
// No tensor cores: there is no dense contraction here (per-offset depth, gate and stride).
#ifndef TLSB_DEVICE_CUH
#define TLSB_DEVICE_CUH

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <type_traits>

#include "../../include/tlsb200.h"
#include "tlsb_internal.h"

namespace {
using namespace tlsb;

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
template <int kT, typename T, int kScanItems = tlsb::kScanItems>
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

// The same scan over 16-bit counters packed two per word (entry 2k in the low half of word k): every prefix must stay
// below 65536 (the callers sort fewer than 65536 keys), so no half ever carries into its neighbour.
template <int kT, int kScanItems>
__device__ void block_inclusive_scan_u16x2(unsigned *data, int nwords, int *warp_tot /* [kT/32+1] shared */)
{
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;
    for (int base = 0; base < nwords; base += kT * kScanItems) {
        const int first = base + tid * kScanItems;
        unsigned v[kScanItems];
        int run = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const unsigned x = (first + k < nwords) ? data[first + k] : 0u;
            run += (int)(x & 0xffffu);
            const unsigned lo = (unsigned)run;
            run += (int)(x >> 16);
            v[k] = lo | ((unsigned)run << 16);
        }
        int incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(kFull, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int wv = (lane < kW) ? warp_tot[lane] : 0;
            int wi = wv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(kFull, wi, off);
                if (lane >= off) wi += o;
            }
            if (lane < kW) warp_tot[lane] = wi - wv;
            if (lane == kW - 1) warp_tot[kW] = wi;
        }
        __syncthreads();
        int excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 0;
        const unsigned offset = (unsigned)(carry + warp_tot[wid] + excl) * 0x10001u;  // the same offset in both halves
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (first + k < nwords) data[first + k] = v[k] + offset;
        carry += warp_tot[kW];
        __syncthreads();
    }
}

// Second half of the sort: skey/sid hold the keys grouped by bucket (any order inside a bucket); rank every key
// inside its bucket by (phase, index) = numpy's stable mergesort order, and gather src1 (src2) to the sorted
// slots of dst1 (dst2).  Bucket b spans [H[b-1], H[b]) with kShift = 0 (H[-1] = 0), [H[b], H[b+1]) with
// kShift = 1.  Ends WITHOUT a barrier.
template <int kT, typename idx_t, bool kTwo, int kU, int kShift, bool kKeepIds = false, bool kH16 = false>
__device__ __forceinline__ void rank_gather(int N, int NB, const int *H, const double *skey, const idx_t *sid,
                                            const double *__restrict__ src1, const double *__restrict__ src2,
                                            double *dst1, double *dst2, idx_t *sid_sorted = nullptr)
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
            if (kH16) {
                const unsigned short *H16 = reinterpret_cast<const unsigned short *>(H);
                lo[u] = (bk + kShift) ? (int)H16[bk - 1 + kShift] : 0;
                hi[u] = (int)H16[bk + kShift];
            } else {
                lo[u] = (bk + kShift) ? H[bk - 1 + kShift] : 0;
                hi[u] = H[bk + kShift];
            }
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
                if (kKeepIds) sid_sorted[rank[u]] = (idx_t)id[u];
            }
        }
    }
}

// Phase fold + stable sort of one trial (core.py:15-18 + :120-123, or stats.py:172-176 with an
// epoch): histogram of NB phase buckets -> block scan -> scatter (any order inside a bucket) ->
// rank inside the bucket by (phase, index) = numpy's stable mergesort order; src1 (and src2) are
// gathered to their sorted slots in dst1 (dst2).  dst1 doubles as the store of the unsorted
// phases until the ranking step; skey/sid/H are scratch.  Ends WITHOUT a barrier.
// kH16: the histogram is kept as 16-bit counters, two per word (fewer than 65536 keys), so that twice as many buckets fit
// the same shared memory: with NB = 2 N most keys are alone in their bucket and the ranking loops are short.
template <int kT, typename idx_t, bool kTwo, bool kEpoch, int kU = 4, int kHScanItems = tlsb::kScanItems, bool kKeepIds = false,
          bool kH16 = false>
__device__ __forceinline__ void fold_sort_gather(const double *__restrict__ t, double T0, double r, int N, int NB,
                                                 int *H, double *skey, idx_t *sid,
                                                 const double *__restrict__ src1, const double *__restrict__ src2,
                                                 double *dst1, double *dst2, int *scan_scratch,
                                                 idx_t *sid_sorted = nullptr)
{
    // kU independent load chains per thread (the streaming layouts sort in L2/HBM)
    const int tid = threadIdx.x;
    unsigned *Hw = reinterpret_cast<unsigned *>(H);
    const int nwords = (NB + 2) / 2;  // kH16: entries 0..NB, two per word
    if (kH16) {
        for (int b = tid; b < nwords; b += kT) Hw[b] = 0u;
    } else {
        for (int b = tid; b <= NB; b += kT) H[b] = 0;
    }
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
                const int e = bucket_of(ph, NB) + 1;
                if (kH16) atomicAdd(&Hw[e >> 1], 1u << ((e & 1) * 16));
                else atomicAdd(&H[e], 1);
            }
        }
    }
    __syncthreads();
    // inclusive scan of H[0..NB] (H[0] = 0): H[b] = number of keys in buckets < b
    if (kH16) block_inclusive_scan_u16x2<kT, kHScanItems>(Hw, nwords, scan_scratch);
    else block_inclusive_scan<kT, int, kHScanItems>(H, NB + 1, scan_scratch);
    for (int k0 = tid; k0 < N; k0 += kT * kU) {
        double ph[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) ph[u] = (k0 + u * kT < N) ? ph_unsorted[k0 + u * kT] : 0.0;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int k = k0 + u * kT;
            if (k < N) {
                int pos;  // any order inside the bucket
                if (kH16) {
                    const int e = bucket_of(ph[u], NB), sh = (e & 1) * 16;
                    pos = (int)((atomicAdd(&Hw[e >> 1], 1u << sh) >> sh) & 0xffffu);
                } else {
                    pos = atomicAdd(&H[bucket_of(ph[u], NB)], 1);
                }
                skey[pos] = ph[u];
                sid[pos] = (idx_t)k;
            }
        }
    }
    __syncthreads();  // now H[b] = end of bucket b; the unsorted phases are dead
    rank_gather<kT, idx_t, kTwo, kU, 0, kKeepIds, kH16>(N, NB, H, skey, sid, src1, src2, dst1, dst2, sid_sorted);
}

// After the sort: cs1[0..N) holds the sorted d = 1 - y (cs1 = cs + 1).  Wrap the first M samples to
// the end (core.py:126-132), then ONE pass turns d into its inclusive cumulative sum in place
// (helpers.py:70-73), writes wd = w * d and returns this thread's share of T = sum_{k<N} w d^2.
// With begin > 0 the pass resumes at position `begin` with the running sum `carry` (the samples
// there already hold their d; nothing is wrapped).
// kWd64 / kWd32: write the products w*d in fp64 (wd) and / or rounded to fp32 (wd32, the filter pass's samples).
// kCs32: also write the DETRENDED cumulative sums rounded to fp32, cs32_1[e] = fl32(cs1[e] - (e + 1) mu) (cs32_1 = cs32 + 1:
// the fp32 gate and screen read these, tlsb_device.cuh "fp32 gate"), and return the largest |cs32| written in *cmax32.
// kGatherW (unequal weights, filter layouts): the weights are not kept per folded position in fp64; they are gathered
// from the light curve through the sorted sample ids (wsrc[sid[k]]) and written rounded to fp32 (w32).
template <int kT, bool kUniformW, int kScanItems = tlsb::kScanItems, bool kWd64 = true, bool kWd32 = false, bool kCs32 = false,
          bool kGatherW = false, typename sid_t = unsigned short, bool kW32 = kGatherW>
__device__ __forceinline__ double wrap_weight_scan(double *cs1, double *w, double *wd, double w0, int N, int NM,
                                                   int NMP, double *warp_tot /* [kT/32 + 1] shared */,
                                                   int begin = 0, double carry = 0.0, float *wd32 = nullptr,
                                                   float *cs32_1 = nullptr, double mu = 0.0, float *cmax32 = nullptr,
                                                   const double *__restrict__ wsrc = nullptr, const sid_t *sid = nullptr,
                                                   float *w32 = nullptr)
{
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int k = N + tid; k < NMP; k += kT) {
        if (k < NM) {
            if (begin == 0) {
                cs1[k] = cs1[k - N];
                if (!kUniformW && !kGatherW) w[k] = w[k - N];
            }
        } else {  // slack read (never used) by the unguarded tap groups
            if (kWd64) wd[k] = 0.0;
            if (kWd32) wd32[k] = 0.f;
            if (!kUniformW && !kGatherW) w[k] = 0.0;
            if (kW32) w32[k] = 0.f;
        }
    }
    __syncthreads();
    double tpart = 0.0;
    float cmax = 0.f;
    for (int base = begin; base < NM; base += kT * kScanItems) {
        const int first = base + tid * kScanItems;
        double v[kScanItems];
        double run = 0.0;
        double dv[kScanItems], wv[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {  // every load in flight before the arithmetic
            const int e = first + k < NM ? first + k : NM - 1;
            dv[k] = cs1[e];
            if (kGatherW) wv[k] = __ldg(wsrc + sid[e < N ? e : e - N]);
            else wv[k] = kUniformW ? w0 : w[e];
        }
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int e = first + k;
            double d = 0.0;
            if (e < NM) {
                d = dv[k];
                const double x = wv[k] * d;
                if (kWd64) wd[e] = x;
                if (kWd32) wd32[e] = (float)x;
                if (kW32) w32[e] = (float)wv[k];
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
            if (first + k < NM) {
                const double c = v[k] + offset;
                cs1[first + k] = c;
                if (kCs32) {
                    const float c32 = (float)fma(-(double)(first + k + 1), mu, c);
                    cs32_1[first + k] = c32;
                    cmax = fmaxf(cmax, fabsf(c32));
                }
            }
        carry += warp_tot[kW];
        __syncthreads();
    }
    if (kCs32) *cmax32 = cmax;
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

// ------------------------------------------------------------------------------------------
// fp32 gate (equal-weights paths).  The gate of core.py:58 is mean_i = (cs[i+W] - cs[i]) / W > transit_depth_min, two
// 8-byte shared-memory reads and three fp64 operations per candidate - for EVERY offset of every admissible width,
// most of which fail on quiet data.  The filter paths therefore keep the cumulative sums as DETRENDED fp32 values,
// cs32[k] = fl32(cs[k] - k mu) (mu = the light curve's mean of d: without it cs would grow linearly for flux that is not
// normalised to 1 and fp32 would resolve nothing), and test diff32 = cs32[i+W] - cs32[i] against a threshold lowered by
// a rigorous bound on |diff32 - ((cs[i+W] - cs[i]) - W mu)|:
//     2^-24 (|cs32[i+W]| + |cs32[i]|) (rounding cs to fp32) + 2^-24 |diff32| (the subtraction) <= 2^-22 Cmax,
//     plus the fp64 roundings of cs[k] - k mu (<= 2^-52 (Cmax + NM |mu|)),     Cmax = max_k |cs32[k]| of this period.
// Every candidate the exact gate passes also passes this one (a SUPERSET; NaN passes); the exact fp64 gate is applied
// where it matters, in bound_one, before a candidate can lower the threshold or become a finalist.  Half the
// shared-memory bytes, one FADD + one FSETP instead of DADD + DMUL + DSETP per candidate.
struct Gate32 {
    double mu;         // detrending slope
    double err;        // bound on |diff32 - (diff64 - W mu)| for this period
    double depth_min;  // core.py:58
    __device__ __forceinline__ void set_err(float cmax, int NM)
    {
        err = 2.5e-7 * (double)cmax + 1e-13 * ((double)cmax + (double)NM * fabs(mu));
    }
    // a candidate passes iff !(diff32 <= thr(W))
    __device__ __forceinline__ float thr(int W) const
    {
        const double w = (double)W;
        return __double2float_rd(w * (depth_min - mu) - err - 1e-13 * (fabs(w * depth_min) + fabs(w * mu)));
    }
};

template <int kBlock, bool kUnit>
__device__ __forceinline__ int gate_block32(const float *cs32, int c0, int c_end, int W, int X, float thr)
{
    const int Xs = kUnit ? 1 : X;
    int mask = 0;
    if (c0 + kBlock <= c_end) {  // straight line: all loads in flight, then the compares
        const float *__restrict__ lo = cs32 + (size_t)c0 * Xs;
        const float *__restrict__ hi = lo + W;
        float diff[kBlock];
#pragma unroll
        for (int rr = 0; rr < kBlock; ++rr) diff[rr] = hi[rr * Xs] - lo[rr * Xs];
#pragma unroll
        for (int rr = 0; rr < kBlock; ++rr) mask |= (diff[rr] <= thr ? 0 : 1) << rr;
    } else {  // the last, partial block of this width (or nothing)
        for (int rr = 0; rr < kBlock; ++rr) {
            const int c = c0 + rr;
            if (c < c_end) {
                const int i = c * Xs;
                if (!(cs32[i + W] - cs32[i] <= thr)) mask |= 1 << rr;
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

// ------------------------------------------------------------------------------------------
// fp32 filter pass (equal weights).  chi2 of a candidate is T + D (D Aq - 2 B) with B = sum_j q_j (w d)_{i+j}; B is
// the only O(L) piece.  The filter computes B in fp32 (128 FMA/clk/SM instead of 64, half the shared-memory
// bytes per sample) together with a RIGOROUS bound on |B32 - B64|, keeps the smallest upper bound U on chi2 seen so
// far, and passes on ("finalist") only candidates whose lower bound does not exceed U.  Finalists are then evaluated
// in fp64 exactly as before (eval_exact), and the lexicographic arg-min runs over them alone.  Because the true
// minimum m satisfies lo(c*) <= chi2(c*) = m <= U at all times, the arg-min candidate (and every candidate that
// ties with it) is always a finalist: results are bit-identical with the filter on or off.
//
// The bound: inputs rounded to fp32 (relative 2^-24 each) and an FMA chain of L terms give
//   |B32 - sum q_j wd_j| <= ((1 + u)^2 (1 + gamma_L) - 1) sum |q_j| |wd_j|,  u = 2^-24, gamma_L = L u / (1 - L u),
// which is below (L + 4) u S for L u < 1e-3; the fp64 chain adds L 2^-53 S.  With S <= max|wd| sum_j |q_j| the
// host stores eb = (L + 8) 2^-24 sum_j |q_j| (1 + 1e-6) per width, and the kernel multiplies by max|w d| of the light curve.
// Zero padded taps add exact zeros.  NaN / inf anywhere make the comparison fail safe (the candidate is a finalist).
// ------------------------------------------------------------------------------------------
// One pass of the fp32 tap loop over a residue class (V = 1) or a pair of adjacent classes (V = 2, even strides:
// samples are then aligned float2).  Step m multiplies the sample vector at sp + m*X with the template vectors of
// steps m, m-1, ..., m-kBlock+1 (one per candidate).  The class's template values are CONTIGUOUS in tq32
// (residue-class major, tlsb_internal.h), so 8 / V steps' worth arrive per float4 broadcast load.  Steps run in
// unguarded groups of kG = 8; the template values of two consecutive groups live in two register sets that swap
// roles every group: the set of the previous group is refilled with the NEXT group's values as soon as its entries
// are dead (entry k is last used at step k-2), so there is no window copy, and samples are loaded one group ahead.
template <int kBlock, int V, bool kUnit>
__device__ __forceinline__ void tap_pass32(const float *__restrict__ qp, const float *__restrict__ sp, int X, int groups,
                                           float (&B)[kBlock])
{
    constexpr int kG = kGroup32;
    static_assert(kBlock + 1 <= kG, "entry k of the previous group must be dead after step k-2");
    const int Xs = kUnit ? 1 : X;
    float qa[kG][V], qb[kG][V];  // template values of the even / odd groups
    float sa[kG][V], sb[kG][V];  // samples, one group ahead
    auto load_q = [&](float (&dst)[kG][V], const float *__restrict__ p, int from) {  // entries [from, from + 4)
        if (V == 2) {
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p + 2 * from));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p + 2 * from + 4));
            dst[from][0] = v0.x; dst[from][V - 1] = v0.y; dst[from + 1][0] = v0.z; dst[from + 1][V - 1] = v0.w;
            dst[from + 2][0] = v1.x; dst[from + 2][V - 1] = v1.y; dst[from + 3][0] = v1.z; dst[from + 3][V - 1] = v1.w;
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + from));
            dst[from][0] = v.x; dst[from + 1][0] = v.y; dst[from + 2][0] = v.z; dst[from + 3][0] = v.w;
        }
    };
    auto load_s = [&](float (&dst)[kG][V], const float *__restrict__ p) {
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            if (V == 2) {
                const float2 v = *reinterpret_cast<const float2 *>(p + k * Xs);
                dst[k][0] = v.x; dst[k][V - 1] = v.y;
            } else {
                dst[k][0] = p[k * Xs];
            }
        }
    };
    // one group: cur = this group's template values, prev = the previous group's, refilled with the next group's
    auto group = [&](float (&cur)[kG][V], float (&prev)[kG][V], const float (&sv)[kG][V], const float *__restrict__ qnext) {
#pragma unroll
        for (int mm = 0; mm < kG; ++mm) {
#pragma unroll
            for (int r = 0; r < kBlock; ++r) {
                const int slot = mm - r;
#pragma unroll
                for (int v = 0; v < V; ++v) B[r] = fmaf(slot >= 0 ? cur[slot][v] : prev[slot + kG][v], sv[mm][v], B[r]);
            }
            if (mm == 2) load_q(prev, qnext, 0);  // prev[0..3] are dead after step 1
            if (mm == 6) load_q(prev, qnext, 4);  // prev[4..7] are dead after step 5
        }
    };
#pragma unroll
    for (int k = 0; k < kG; ++k)
#pragma unroll
        for (int v = 0; v < V; ++v) qb[k][v] = 0.f;  // "group -1": nothing before the first tap
    load_q(qa, qp, 0);
    load_q(qa, qp, 4);
    load_s(sa, sp);
#pragma unroll 1
    for (int g = 0; g < groups; g += 2) {
        load_s(sb, sp + kG * Xs);
        group(qa, qb, sa, qp + kG * V);   // qb <- template values of group g + 1
        if (g + 1 >= groups) break;
        load_s(sa, sp + 2 * kG * Xs);
        group(qb, qa, sb, qp + 2 * kG * V);  // qa <- template values of group g + 2
        qp += 2 * kG * V;
        sp += 2 * kG * Xs;
    }
}

// The fp32 correlation B[r] = sum_j q_j (w d)_{i0 + r X + j} of one block of kBlock candidates of width record wr.
// With the stride X the taps split into residue classes j = X a + b; odd strides run one pass per class (neighbouring
// lanes sit kBlock*X floats apart: odd, so the 32 banks are conflict free), even strides one pass per PAIR of classes
// with 8-byte sample loads (candidates start at multiples of X, so the pairs are aligned).  Every lane of a warp
// walks the classes in the same order: template loads are broadcasts.
template <int kBlock, bool kUnit>
__device__ __forceinline__ void tap_block32(const WidthRec &wr, const float *__restrict__ tq32,
                                            const float *__restrict__ wd32, int c0, float (&B)[kBlock])
{
    constexpr int kG = kGroup32;
    const int L = wr.L, X = kUnit ? 1 : wr.X;
#pragma unroll
    for (int r = 0; r < kBlock; ++r) B[r] = 0.f;
    const float *__restrict__ qp = tq32 + wr.q32;
    const float *__restrict__ sp = wd32 + c0 * X;
    const int groups = ((L + X - 1) / X + kBlock - 1 + kG - 1) / kG;  // steps: taps of the widest class + kBlock - 1
    if (kUnit) {
        tap_pass32<kBlock, 1, true>(qp, sp, 1, groups, B);
    } else if ((X & 1) == 0) {
        // Neighbouring lanes sit kBlock * X / 2 eight-byte units apart: conflict free on the 16 eight-byte banks of a
        // half-warp when X / 2 is odd.  When X / 2 = 2^e * odd, lanes 16 / g blocks apart (g = min(2^e, 16)) share a
        // bank; they start at DIFFERENT pair classes (one class = one unit further), rotated by the block index, so
        // the half-warp is conflict free again and the template loads touch g addresses instead of one.
        const int npc = X / 2;
        const int g = min(npc & -npc, 16);
        int pc = ((c0 / kBlock) / (16 / g)) & (g - 1);
        for (int t = 0; t < npc && 2 * t < L; ++t) {
            tap_pass32<kBlock, 2, false>(qp + 2 * pc * wr.astride, sp + 2 * pc, X, groups, B);
            if (++pc == npc) pc = 0;
        }
    } else {
        for (int b = 0; b < X && b < L; ++b) tap_pass32<kBlock, 1, false>(qp + b * wr.astride, sp + b, X, groups, B);
    }
}

// ---- unequal weights (per-point dy): the filter pass accumulates TWO correlations in fp32,
//   B = sum_j q_j (w d)_{i+j}   and   A = sum_j q_j^2 w_{i+j},
// from wd32 = fl32(w d) and w32 = fl32(w).  One residue class per pass (V = 1); kQ2 = 2: the class's template values sit
// at every second float (the pair-interleaved layout of even strides) and are fetched with scalar broadcast loads.
// The squares q^2 are formed in fp32 from the fp32 template values when they enter the register window (three roundings
// against the exact q^2 instead of one: the bound on |A32 - A64| uses L + 12 instead of L + 8 units).
template <int kBlock, bool kUnit, int kQ2>
__device__ __forceinline__ void tap_pass32w(const float *__restrict__ qp, const float *__restrict__ sp, const float *__restrict__ wp,
                                            int X, int groups, float (&A)[kBlock], float (&B)[kBlock])
{
    constexpr int kG = kGroup32;
    static_assert(kBlock + 1 <= kG, "entry k of the previous group must be dead after step k-2");
    const int Xs = kUnit ? 1 : X;
    float qa[kG], qb[kG], pa[kG], pb[kG];  // template values and their squares of the even / odd groups
    float sa[kG], sb[kG], wa[kG], wb[kG];  // samples w d and w, one group ahead
    auto load_q = [&](float (&dq)[kG], float (&dp)[kG], const float *__restrict__ p, int from) {  // entries [from, from + 4)
        if (kQ2 == 1) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + from));
            dq[from] = v.x; dq[from + 1] = v.y; dq[from + 2] = v.z; dq[from + 3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dq[from + k] = __ldg(p + 2 * (from + k));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dp[from + k] = dq[from + k] * dq[from + k];
    };
    auto load_s = [&](float (&ds)[kG], float (&dw)[kG], const float *__restrict__ p, const float *__restrict__ pw) {
#pragma unroll
        for (int k = 0; k < kG; ++k) {
            ds[k] = p[k * Xs];
            dw[k] = pw[k * Xs];
        }
    };
    auto group = [&](float (&cq)[kG], float (&cp)[kG], float (&pq)[kG], float (&pp)[kG], const float (&sv)[kG], const float (&wv)[kG],
                     const float *__restrict__ qnext) {
#pragma unroll
        for (int mm = 0; mm < kG; ++mm) {
#pragma unroll
            for (int r = 0; r < kBlock; ++r) {
                const int slot = mm - r;
                B[r] = fmaf(slot >= 0 ? cq[slot] : pq[slot + kG], sv[mm], B[r]);
                A[r] = fmaf(slot >= 0 ? cp[slot] : pp[slot + kG], wv[mm], A[r]);
            }
            // prev[j] is last read at step j - kG + kBlock - 1: entries 0..3 are dead after step 1, entries 4..7 after step 5
            if (mm == 2) load_q(pq, pp, qnext, 0);
            if (mm == 6) load_q(pq, pp, qnext, 4);
        }
    };
#pragma unroll
    for (int k = 0; k < kG; ++k) { qb[k] = 0.f; pb[k] = 0.f; }  // "group -1": nothing before the first tap
    load_q(qa, pa, qp, 0);
    load_q(qa, pa, qp, 4);
    load_s(sa, wa, sp, wp);
#pragma unroll 1
    for (int g = 0; g < groups; g += 2) {
        load_s(sb, wb, sp + kG * Xs, wp + kG * Xs);
        group(qa, pa, qb, pb, sa, wa, qp + kG * kQ2);   // qb <- template values of group g + 1
        if (g + 1 >= groups) break;
        load_s(sa, wa, sp + 2 * kG * Xs, wp + 2 * kG * Xs);
        group(qb, pb, qa, pa, sb, wb, qp + 2 * kG * kQ2);  // qa <- template values of group g + 2
        qp += 2 * kG * kQ2;
        sp += 2 * kG * Xs;
        wp += 2 * kG * Xs;
    }
}

template <int kBlock, bool kUnit>
__device__ __forceinline__ void tap_block32w(const WidthRec &wr, const float *__restrict__ tq32, const float *__restrict__ wd32,
                                             const float *__restrict__ w32, int c0, float (&A)[kBlock], float (&B)[kBlock])
{
    constexpr int kG = kGroup32;
    const int L = wr.L, X = kUnit ? 1 : wr.X;
#pragma unroll
    for (int r = 0; r < kBlock; ++r) { A[r] = 0.f; B[r] = 0.f; }
    const float *__restrict__ qp = tq32 + wr.q32;
    const float *__restrict__ sp = wd32 + c0 * X;
    const float *__restrict__ wp = w32 + c0 * X;
    const int groups = ((L + X - 1) / X + kBlock - 1 + kG - 1) / kG;
    if (kUnit) {
        tap_pass32w<kBlock, true, 1>(qp, sp, wp, 1, groups, A, B);
    } else if ((X & 1) == 0) {  // pair-interleaved template layout: class b = 2 c + v at ((c A + a) 2 + v)
        for (int b = 0; b < X && b < L; ++b)
            tap_pass32w<kBlock, false, 2>(qp + 2 * (b >> 1) * wr.astride + (b & 1), sp + b, wp + b, X, groups, A, B);
    } else {
        for (int b = 0; b < X && b < L; ++b) tap_pass32w<kBlock, false, 1>(qp + b * wr.astride, sp + b, wp + b, X, groups, A, B);
    }
}

__device__ __forceinline__ double chi2_value(double T, double D, double Aq, double B)
{
    return __fma_rn(D, __fma_rn(D, Aq, -2.0 * B), T);  // T + D (D Aq - 2 B), core.py:67-70 after the algebra of DESIGN.md §3
}

// What the exact evaluation of a finalist reads: cumulative sums, the fp64 products w*d.  kGather = 1 (resident kernel):
// rebuilt from the sorted sample ids in shared memory (u16), w0 * d[id], the same product a folded fp64 array would
// hold; kGather = 2 (tiled kernel): the same through 32-bit ids in the CTA's global scratch; kGather = 0: fp64 arrays.
template <int kGather>
struct ExactView {
    const double *cs;      // cumulative sums, indexable by the global offset
    const double *wd;      // !kGather: w*d per folded position
    const double *dval;    // kGather: 1 - y per sample
    const unsigned short *sid;  // kGather: sample id per folded position < N
    const double *tq;
    double w0, T;
    int N;
    const double *wval;    // kGather, unequal weights: 1 / dy^2 per sample
    const double *w;       // !kGather, unequal weights: w per folded position
    const unsigned *sid32; // kGather == 2: 32-bit ids (tiled kernel: the sorted ids live in the CTA's global scratch)
    __device__ __forceinline__ int id_at(int k) const
    {
        const int kk = k < N ? k : k - N;
        return kGather == 2 ? (int)__ldcg(sid32 + kk) : (int)sid[kk];
    }
    __device__ __forceinline__ double wdv(int k) const
    {
        if (kGather) return __dmul_rn(w0, __ldg(dval + id_at(k)));
        return wd[k];
    }
    // unequal weights: the weight and the product w*d of folded position k (the same product the fp64 arrays would hold)
    __device__ __forceinline__ void wpair(int k, double &wk, double &wdk) const
    {
        if (kGather) {
            const int id = id_at(k);
            wk = __ldg(wval + id);
            wdk = __dmul_rn(wk, __ldg(dval + id));
        } else {
            wk = w[k];
            wdk = wd[k];
        }
    }
};

// Exact fp64 chi2 of candidate c0 + rr of width record u, by one WARP (all 32 lanes must call with the same
// arguments): lane l accumulates the taps j = l, l + 32, ... in ascending order, a fixed butterfly adds the 32
// partial sums (every lane ends with the same bits), and every lane offers the candidate to its running best.
// This is the ONLY place the equal-weights paths evaluate chi2 in fp64, so results do not depend on what the
// filter let through.
// kUW = false (unequal weights): the quadratic term A = sum_j q_j^2 w_{i+j} is accumulated the same way.
template <int kGather, bool kUW = true>
__device__ __noinline__ void eval_exact_warp(const ExactView<kGather> &v, const WidthRec *rec, int c0, int rr, int u,
                                             Best &best)
{
    const int lane = threadIdx.x & 31;
    const WidthRec wr = rec[u];
    const int L = wr.L, i = (c0 + rr) * wr.X;
    const double *__restrict__ q = v.tq + wr.q;
    double B = 0.0, A = 0.0;
    if (kUW) {
#pragma unroll 4
        for (int j = lane; j < L; j += 32) B = fma(__ldg(q + j), v.wdv(i + j), B);
    } else {
#pragma unroll 2
        for (int j = lane; j < L; j += 32) {
            const double qv = __ldg(q + j);
            double wk, wdk;
            v.wpair(i + j, wk, wdk);
            B = fma(qv, wdk, B);
            A = fma(__dmul_rn(qv, qv), wk, A);
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        B += __shfl_xor_sync(kFull, B, off);
        if (!kUW) A += __shfl_xor_sync(kFull, A, off);
    }
    const double mean = (__ldcg(v.cs + i + wr.W) - __ldcg(v.cs + i)) * wr.invW;  // global memory, written by this CTA
    const double D = mean * wr.os;
    double chi = chi2_value(v.T, D, kUW ? v.w0 * wr.sq2 : A, B);
    if (L < wr.W) {  // samples L..W-1 are in neither sum
        double rest = 0.0;
        for (int k = i + L + lane; k < i + wr.W; k += 32) {
            if (kUW) {
                const double x = v.wdv(k);
                rest += x * x / v.w0;
            } else {
                double wk, wdk;
                v.wpair(k, wk, wdk);
                rest += wdk * wdk / wk;
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) rest += __shfl_xor_sync(kFull, rest, off);
        chi -= rest;
    }
    if (better(chi, u, i, best)) { best.chi2 = chi; best.D = D; best.u = u; best.i = i; }
}

// CTA-wide state of the filter in shared memory
struct FilterShared {
    unsigned long long U;  // bits of the smallest upper bound on chi2 so far (positive doubles order like integers)
    int fq_fill;           // finalists queued in this round (may exceed the capacity: see kRedoFlag)
    int pad_;
};

// Per-lane view of the threshold: U and the reduction a candidate has to reach, G = T - U, rounded down to fp32
struct Threshold {
    double U;
    float G32;
    __device__ __forceinline__ void set(double u, double T)
    {
        U = u;
        G32 = __double2float_rd(T - u);
    }
    __device__ __forceinline__ void refresh(const FilterShared *fs, double T)
    {
        const double Ush = __longlong_as_double((long long)*(volatile const unsigned long long *)&fs->U);
        if (Ush < U) set(Ush, T);
    }
};

// After tap_block32, cheap screen in fp32.  chi2 = T - R with the reduction R = D (2 B - D Aq); a candidate can only
// matter if R + (error bounds) >= G = T - U.  R is evaluated in fp32 from D = (diff32 + W mu) c1, c1 = fl32(os / W):
// against the exact R(B32) the roundings cost less than 16 u relative to the magnitude of its terms (D: 3 roundings,
// Aq: 1, three operations), covered by 2e-6 |D| (2 |B| + |D| Aq); the ABSOLUTE error of diff32 + W mu (Gate32::err plus
// the rounding of W mu) moves D by at most dD and R by at most dD (2 |B| + 2 |D| Aq + dD Aq); 2 |D| EB bounds
// |R(B32) - R(B64)| (see tap_block32) and slopT the fp64 evaluation roundings.  Returns the candidates that are NOT
// ruled out (they get exact fp64 bounds next).  NaN anywhere fails safe.
template <int kBlock, bool kUnit>
__device__ __forceinline__ int block_screen(const WidthRec &wr, const float *cs32, double w0, int c0, int mask,
                                            const float (&B)[kBlock], float G32, float EB2f, float slopTf, float Wmu32,
                                            float dD)
{
    float lo[kBlock], hi[kBlock];
    const int X = kUnit ? 1 : wr.X;
    const float *__restrict__ p = cs32 + c0 * X;
    const float *__restrict__ ph = p + wr.W;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        lo[rr] = p[rr * X];
        hi[rr] = ph[rr * X];
    }
    const float c1 = (float)(wr.invW * wr.os), Aq32 = (float)(w0 * wr.sq2);
    const float dDA = dD * Aq32;
    int keep = 0;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        const float Df = ((hi[rr] - lo[rr]) + Wmu32) * c1, aD = fabsf(Df);
        const float R = Df * fmaf(-Df, Aq32, 2.f * B[rr]);
        const float t = fmaf(aD, Aq32, 2.f * fabsf(B[rr]));  // |D| Aq + 2 |B|
        float m = fmaf(aD * t, 2e-6f, fmaf(aD, EB2f, slopTf));
        m = fmaf(dD, fmaf(aD, Aq32, t) + dDA, m);  // dD (2 |B| + 2 |D| Aq + dD Aq)
        keep |= (R + m < G32 ? 0 : 1) << rr;
    }
    if (wr.L < wr.W) keep = -1;  // the untouched tail adds to R: no screen for trimmed templates (rare)
    return keep & mask;
}

// Exact fp64 bounds of ONE candidate the screen did not rule out (rare: kept out of line, scalars only, so that the
// hot loop keeps its registers and its instruction-cache footprint).  First the EXACT gate of core.py:58 on the fp64
// cumulative sums (the fp32 gate let a superset through): a candidate that fails it returns +inf and touches nothing.
// Otherwise returns the lower bound of chi2 rounded down to fp32 and lowers the CTA-wide threshold to the upper bound
// if that is smaller.
//   cs64 = the fp64 cumulative sums (global memory, written by this CTA before the last barrier), i = window start,
//   B = the fp32 correlation, EB = the bound on |B32 - B64|;
//   n_tail > 0: the template was trimmed (L < W) and wd32 + k_tail are the n_tail window samples it does not cover.
__device__ __noinline__ float bound_one(const double *cs64, int i, int W, double depth_min, float B, double invW, double os,
                                        double Aq, double EB, double T, double w0, const float *wd32, int k_tail, int n_tail,
                                        FilterShared *fs)
{
    const double diff = __ldcg(cs64 + i + W) - __ldcg(cs64 + i);
    if (!(diff * invW > depth_min)) return INFINITY;  // helpers.py:70-73 + core.py:58, as gate_block evaluates it
    const double D = diff * invW * os;
    const double Bd = (double)B;
    const double chi = chi2_value(T, D, Aq, Bd);
    // |chi32 - chi64| <= 2 |D| EB + the roundings of two fp64 evaluations of the same expression
    const double E = 2.0 * fabs(D) * EB + 1e-14 * (fabs(T) + fabs(D) * (fabs(D) * Aq + 2.0 * fabs(Bd)));
    double l = chi - E, h = chi + E;
    if (n_tail > 0) {  // the untouched tail only lowers chi2: [tail (1 - e), tail (1 + e)] from the fp32 samples
        float rest = 0.f;
        for (int k = k_tail; k < k_tail + n_tail; ++k) rest = fmaf(wd32[k], wd32[k], rest);
        const double tail = (double)rest / w0, e = (double)(n_tail + 4) * 1.2e-7;
        l -= tail * (1.0 + e);
        h -= tail * (1.0 - e);
    }
    if (h > 0.0) atomicMin(&fs->U, (unsigned long long)__double_as_longlong(h));  // NaN compares false: no update
    return __double2float_rd(l);
}

// Unequal weights: the screen with both fp32 correlations, R = D (2 B - D A).  Besides the terms of block_screen the margin
// carries |D|^2 EA for |A32 - A64| <= EA.
template <int kBlock, bool kUnit>
__device__ __forceinline__ int block_screen_w(const WidthRec &wr, const float *cs32, int c0, int mask, const float (&A)[kBlock],
                                              const float (&B)[kBlock], float G32, float EB2f, float EAf, float slopTf, float Wmu32,
                                              float dD)
{
    float lo[kBlock], hi[kBlock];
    const int X = kUnit ? 1 : wr.X;
    const float *__restrict__ p = cs32 + c0 * X;
    const float *__restrict__ ph = p + wr.W;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        lo[rr] = p[rr * X];
        hi[rr] = ph[rr * X];
    }
    const float c1 = (float)(wr.invW * wr.os);
    int keep = 0;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        const float Df = ((hi[rr] - lo[rr]) + Wmu32) * c1, aD = fabsf(Df), aA = fabsf(A[rr]);
        const float R = Df * fmaf(-Df, A[rr], 2.f * B[rr]);
        const float t = fmaf(aD, aA, 2.f * fabsf(B[rr]));  // |D| |A| + 2 |B|
        float m = fmaf(aD * t, 2e-6f, fmaf(aD, fmaf(aD, EAf, EB2f), slopTf));
        m = fmaf(dD, fmaf(aD, aA, t) + dD * aA, m);  // dD (2 |B| + 2 |D| |A| + dD |A|)
        keep |= (R + m < G32 ? 0 : 1) << rr;
    }
    if (wr.L < wr.W) keep = -1;  // the untouched tail adds to R: no screen for trimmed templates (rare)
    return keep & mask;
}

// bound_one for unequal weights: chi2 = T + D (D A - 2 B) from the fp32 correlations, |A32 - A64| <= EA, |B32 - B64| <= EB;
// the untouched tail of a trimmed template is sum (w d)^2 / w over the fp32 samples.
__device__ __noinline__ float bound_one_w(const double *cs64, int i, int W, double depth_min, float A, float B, double invW,
                                          double os, double EA, double EB, double T, const float *wd32, const float *w32,
                                          int k_tail, int n_tail, FilterShared *fs)
{
    const double diff = __ldcg(cs64 + i + W) - __ldcg(cs64 + i);
    if (!(diff * invW > depth_min)) return INFINITY;  // the exact gate (helpers.py:70-73 + core.py:58)
    const double D = diff * invW * os;
    const double Ad = (double)A, Bd = (double)B;
    const double chi = chi2_value(T, D, Ad, Bd);
    const double E = 2.0 * fabs(D) * EB + D * D * EA + 1e-14 * (fabs(T) + fabs(D) * (fabs(D) * fabs(Ad) + 2.0 * fabs(Bd)));
    double l = chi - E, h = chi + E;
    if (n_tail > 0) {
        float rest = 0.f;
        for (int k = k_tail; k < k_tail + n_tail; ++k) rest += wd32[k] * wd32[k] / w32[k];
        const double tail = (double)rest, e = (double)(n_tail + 8) * 2.4e-7;
        l -= tail * (1.0 + e);
        h -= tail * (1.0 - e);
    }
    if (h > 0.0) atomicMin(&fs->U, (unsigned long long)__double_as_longlong(h));  // NaN compares false: no update
    return __double2float_rd(l);
}

// What the cold paths of a sweep need, gathered once per sweep (lives in local memory)
template <int kGather>
struct FilterCtx {
    ExactView<kGather> view;
    const WidthRec *rec;
    FilterShared *fs;
    int2 *fq;
    float *fq_lo;
    int fq_cap;
    unsigned long long *stats;
    Gate32 g32;  // the fp32 gate's parameters of this period (mu, err, depth_min)
    double ea_scale;  // unequal weights: max w (the scale of the bound on |A32 - A64|)
};

// Finalists of one batch (bit rr of `fin`, lower bounds rounded down to fp32) go to the finalist queue once more
// checked against the threshold as it is NOW; the bound travels with them and is checked a last time before the
// exact evaluation.  All 32 lanes must call.  Queue full (rare): the warp evaluates the leftovers on the spot.
template <int kBlock, int kGather, bool kUW = true>
__device__ __noinline__ void warp_push(const FilterCtx<kGather> &cx, int fin, int c0, int u, float l0, float l1, float l2,
                                       float l3, float l4, float l5, float l6, Best &best)
{
    static_assert(kBlock <= 7, "warp_push carries seven bounds");
    const float lo[7] = {l0, l1, l2, l3, l4, l5, l6};
    const double U = __longlong_as_double((long long)*(volatile unsigned long long *)&cx.fs->U);
    int over = 0;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) {
        if (((fin >> rr) & 1) && !((double)lo[rr] > U)) {
            const int slot = atomicAdd(&cx.fs->fq_fill, 1);
            if (slot < cx.fq_cap) {
                cx.fq[slot] = make_int2(c0, u | (rr << 16));
                cx.fq_lo[slot] = lo[rr];
            } else {
                over |= 1 << rr;
            }
        }
    }
    unsigned any = __ballot_sync(kFull, over != 0);
    while (any) {
        const int src = __ffs(any) - 1;
        const int oc0 = __shfl_sync(kFull, c0, src), ou = __shfl_sync(kFull, u, src), oo = __shfl_sync(kFull, over, src);
        for (int rr = 0; rr < kBlock; ++rr)
            if ((oo >> rr) & 1) {
                eval_exact_warp<kGather, kUW>(cx.view, cx.rec, oc0, rr, ou, best);
                if (cx.stats && (threadIdx.x & 31) == 0) atomicAdd(cx.stats + 2, 1ull);
            }
        any &= any - 1;
    }
}

// The queued finalists, one per warp at a time (call after a barrier, all threads of the CTA): each is checked
// against the current threshold once more (most were queued while it was still settling), evaluated in fp64, and
// its exact chi2 tightens the threshold for the rest.
template <int kT, int kGather, bool kUW = true>
__device__ __forceinline__ void drain_finalists(const FilterCtx<kGather> &cx, Best &best)
{
    constexpr int kW = kT / 32;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nf = min(cx.fs->fq_fill, cx.fq_cap);
    for (int k = wid; k < nf; k += kW) {
        const double U = __longlong_as_double((long long)*(volatile unsigned long long *)&cx.fs->U);
        if ((double)cx.fq_lo[k] > U) continue;
        const int2 e = cx.fq[k];
        const double before = best.chi2;
        eval_exact_warp<kGather, kUW>(cx.view, cx.rec, e.x, e.y >> 16, e.y & 0xffff, best);
        if (lane == 0 && best.chi2 < before && best.chi2 > 0.0)
            atomicMin(&cx.fs->U, (unsigned long long)__double_as_longlong(best.chi2));
        if (cx.stats && lane == 0) atomicAdd(cx.stats + 1, 1ull);
    }
}

// Finalists of a warp's previous batch: pushed one batch late, so that their bounds meet a threshold every active
// warp has contributed to (most candidates qualify while the threshold is still at its start value).
template <int kBlock>
struct Pending {
    int fin, c0, u;
    float lo[kBlock];
    __device__ __forceinline__ void clear()
    {
        fin = 0; c0 = 0; u = 0;
#pragma unroll
        for (int rr = 0; rr < kBlock; ++rr) lo[rr] = 0.f;
    }
    __device__ __forceinline__ float at(int k) const { return k < kBlock ? lo[k < kBlock ? k : 0] : 0.f; }
    template <int kGather, bool kUW = true>
    __device__ __forceinline__ void push(const FilterCtx<kGather> &cx, Best &best)  // all 32 lanes must call
    {
        if (__any_sync(kFull, fin != 0))
            warp_push<kBlock, kGather, kUW>(cx, fin, c0, u, at(0), at(1), at(2), at(3), at(4), at(5), at(6), best);
        fin = 0;
    }
};

// One batch of survivor blocks (one per lane, `have`): fp32 correlation, fp32 screen, exact gate + exact bounds for
// what the screen lets through; then the PREVIOUS batch's finalists are pushed and this batch's become pending.
template <int kBlock, int kGather, bool kUW = true>
__device__ __forceinline__ void tap_batch(bool have, int2 e, const FilterCtx<kGather> &cx, const float *cs32, const float *wd32,
                                          const float *__restrict__ tq32, double w0, double T, double eb_scale, float slopTf,
                                          Threshold &th, Pending<kBlock> &pend, Best &best, const float *w32 = nullptr)
{
    int fin = 0;
    float lo_now[kBlock];
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) lo_now[rr] = 0.f;
    if (have) {
        const int u = e.y & 0xffff, mask = e.y >> 16;
        const WidthRec wr = cx.rec[u];
        const double EB = wr.eb * eb_scale;
        const float EB2f = __double2float_ru(2.000001 * EB);
        // the screen's D comes from diff32 + W mu: absolute error of that sum (Gate32::err + the rounding of W mu), as an error of D
        const double wmu = (double)wr.W * cx.g32.mu;
        const float Wmu32 = (float)wmu;
        const float dD = __double2float_ru((cx.g32.err + 6.1e-8 * fabs(wmu)) * (wr.invW * wr.os) * 1.00001);
        float B[kBlock];
        float A[kBlock];  // unequal weights only
        double EA = 0.0;
        th.refresh(cx.fs, T);
        int keep;
        if constexpr (kUW) {
            if (wr.X == 1) {
                tap_block32<kBlock, true>(wr, tq32, wd32, e.x, B);
                keep = block_screen<kBlock, true>(wr, cs32, w0, e.x, mask, B, th.G32, EB2f, slopTf, Wmu32, dD);
            } else {
                tap_block32<kBlock, false>(wr, tq32, wd32, e.x, B);
                keep = block_screen<kBlock, false>(wr, cs32, w0, e.x, mask, B, th.G32, EB2f, slopTf, Wmu32, dD);
            }
        } else {
            // |A32 - A64| <= EA: the same chain bound as for B with q^2 (three fp32 roundings: L + 12 units) and max w
            EA = (double)(wr.L + 12) * 5.9604644775390625e-08 * wr.sq2 * (1.0 + 1e-6) * cx.ea_scale;
            const float EAf = __double2float_ru(1.000001 * EA);
            if (wr.X == 1) {
                tap_block32w<kBlock, true>(wr, tq32, wd32, w32, e.x, A, B);
                keep = block_screen_w<kBlock, true>(wr, cs32, e.x, mask, A, B, th.G32, EB2f, EAf, slopTf, Wmu32, dD);
            } else {
                tap_block32w<kBlock, false>(wr, tq32, wd32, w32, e.x, A, B);
                keep = block_screen_w<kBlock, false>(wr, cs32, e.x, mask, A, B, th.G32, EB2f, EAf, slopTf, Wmu32, dD);
            }
        }
        if (keep) {  // rare
            const double Aq = w0 * wr.sq2;
#pragma unroll
            for (int rr = 0; rr < kBlock; ++rr)
                if ((keep >> rr) & 1) {
                    if constexpr (kUW)
                        lo_now[rr] = bound_one(cx.view.cs, (e.x + rr) * wr.X, wr.W, cx.g32.depth_min, B[rr], wr.invW, wr.os, Aq, EB, T,
                                               w0, wd32, (e.x + rr) * wr.X + wr.L, wr.W - wr.L, cx.fs);
                    else
                        lo_now[rr] = bound_one_w(cx.view.cs, (e.x + rr) * wr.X, wr.W, cx.g32.depth_min, A[rr], B[rr], wr.invW, wr.os,
                                                 EA, EB, T, wd32, w32, (e.x + rr) * wr.X + wr.L, wr.W - wr.L, cx.fs);
                }
            th.refresh(cx.fs, T);
#pragma unroll
            for (int rr = 0; rr < kBlock; ++rr) fin |= (((keep >> rr) & 1) && !((double)lo_now[rr] > th.U) ? 1 : 0) << rr;
        }
        if (cx.stats) atomicAdd(cx.stats, (unsigned long long)__popc(mask));
    }
    pend.template push<kGather, kUW>(cx, best);  // the previous batch's finalists, against the threshold as it is now
    pend.fin = fin;
    pend.c0 = e.x;
    pend.u = e.y & 0xffff;
#pragma unroll
    for (int rr = 0; rr < kBlock; ++rr) pend.lo[rr] = lo_now[rr];
}

// Scheduler words of one sweep in shared memory (all zero before a sweep starts)
struct SweepShared {
    int tile_next;   // next gate tile to hand out
    int tiles_done;  // tiles whose survivors are in the ring
    int q_tail;      // ring entries reserved by gating warps
    int q_head;      // ring entries taken by tap warps
    int q_done;      // ring entries read and cleared (their slots may be reused)
    int pad_[3];
};

// Phases B1 + B2 of one sweep (equal weights) without a barrier between them: every warp switches roles as the
// work demands.  A warp that finds a batch of 32 survivor blocks in the ring takes it (fp32 correlation, screen,
// exact bounds, finalists); otherwise, if gate tiles are left and the ring has room, it gates the next tile and
// appends the surviving blocks; it leaves when all tiles are gated and the ring is empty.  The ring is a
// power-of-two array of int2 entries, (first candidate, width index | mask << 16); a slot is valid when its second
// word is non-zero, readers clear it.  Tiles are numbered wide -> narrow: tile ids [0, t_tiles[ub-1]) belong to width
// ub-1 and cover its candidates [t_lo, t_hi), and so on downwards.  Finalists are pushed one batch late, so that
// their bounds meet a threshold every active warp has contributed to.  Ends with the finalist queue drained
// (contains barriers: all threads of the CTA must call).  cs / wd32 must be indexable by global offsets.
template <int kT, int kBlock, int kGather, bool kUW = true>
__device__ __forceinline__ void sweep_filter(SweepShared *ss, int2 *queue, int qmask, int tile_end, int ub, const int *t_lo,
                                             const int *t_hi, const int *t_tiles, const WidthRec *rec, const float *cs32,
                                             const float *wd32, const float *__restrict__ tq32, double w0, double T,
                                             const Gate32 &g32, double eb_scale, FilterShared *fs, int2 *fq, float *fq_lo,
                                             int fq_cap, const ExactView<kGather> &view, Best &best,
                                             unsigned long long *stats, const float *w32 = nullptr, double ea_scale = 0.0)
{
    constexpr int kW = kT / 32;
    constexpr int kTile = tile_size(kBlock);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int qroom = qmask + 1 - kW * 32 * kSub;  // gating pauses above this fill: every warp can still add one tile
    FilterCtx<kGather> cx;
    cx.view = view; cx.rec = rec; cx.fs = fs; cx.fq = fq; cx.fq_lo = fq_lo; cx.fq_cap = fq_cap; cx.stats = stats; cx.g32 = g32;
    cx.ea_scale = ea_scale;
    Threshold th;
    th.set(INFINITY, T);
    th.refresh(fs, T);
    const float slopTf = __double2float_ru(4e-14 * fabs(T));
    int cur_u = ub - 1, u_begin = 0, u_end = tile_end > 0 ? t_tiles[cur_u] : 0;
    Pending<kBlock> pend;  // the previous batch's finalists, not pushed yet
    pend.clear();
    int idle = 0;
    bool gating = true;  // the role this warp prefers; it keeps a role while there is work for it (instruction cache)
    for (;;) {
        const int head = *(volatile int *)&ss->q_head, tail = *(volatile int *)&ss->q_tail;
        const int avail = tail - head;
        const bool gate_done = *(volatile int *)&ss->tiles_done >= tile_end;
        const bool can_tap = avail >= 32 || (gate_done && avail > 0);
        const bool can_gate = !gate_done && *(volatile int *)&ss->tile_next < tile_end &&
                              tail - *(volatile int *)&ss->q_done <= qroom;
        if (gating) {
            if (!can_gate || avail >= 32 * kW) gating = false;  // a batch for every warp is waiting
        } else if (!can_tap) {
            gating = true;
        }
        const bool do_tap = can_tap && (!gating || !can_gate);
        const bool do_gate = !do_tap && can_gate;
        if (do_tap) {
            // ---- B2: one batch of survivor blocks ----
            const int n = min(32, avail);
            int ok = 0;
            if (lane == 0) ok = atomicCAS(&ss->q_head, head, head + n) == head;
            if (!__shfl_sync(kFull, ok, 0)) continue;
            idle = 0;
            const bool have = lane < n;
            int2 e = make_int2(0, 0);
            if (have) {
                volatile unsigned long long *slot = reinterpret_cast<volatile unsigned long long *>(queue + ((head + lane) & qmask));
                int spins = 0;
                for (;;) {  // reserved before we took it, written in a moment
                    const unsigned long long raw = *slot;
                    e.x = (int)(unsigned)(raw & 0xffffffffull);
                    e.y = (int)(unsigned)(raw >> 32);
                    if (e.y != 0) break;
                    if (++spins > (1 << 24)) __trap();
                }
                *slot = 0ull;
            }
            __syncwarp();
            if (lane == 0) atomicAdd(&ss->q_done, n);
            tap_batch<kBlock, kGather, kUW>(have, e, cx, cs32, wd32, tq32, w0, T, eb_scale, slopTf, th, pend, best, w32);
            continue;
        }
        if (do_gate) {
            // ---- B1: gate one tile ----
            int g = 0;
            if (lane == 0) g = atomicAdd(&ss->tile_next, 1);
            g = __shfl_sync(kFull, g, 0);
            if (g >= tile_end) continue;
            idle = 0;
            while (g >= u_end) {
                --cur_u;
                u_begin = u_end;
                u_end = u_begin + t_tiles[cur_u];
            }
            const int u = cur_u;
            const int W = rec[u].W, X = rec[u].X, c_end = t_hi[u];
            const float thr = g32.thr(W);
            const int c_tile = t_lo[u] + (g - u_begin) * kTile + lane * kBlock;
            int masks[kSub];
            unsigned votes[kSub];
            int total = 0;
            if (X == 1) {
#pragma unroll
                for (int sb = 0; sb < kSub; ++sb)
                    masks[sb] = gate_block32<kBlock, true>(cs32, c_tile + sb * 32 * kBlock, c_end, W, 1, thr);
            } else {
#pragma unroll
                for (int sb = 0; sb < kSub; ++sb)
                    masks[sb] = gate_block32<kBlock, false>(cs32, c_tile + sb * 32 * kBlock, c_end, W, X, thr);
            }
#pragma unroll
            for (int sb = 0; sb < kSub; ++sb) {
                votes[sb] = __ballot_sync(kFull, masks[sb] != 0);
                total += __popc(votes[sb]);
            }
            if (total) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&ss->q_tail, total);
                base = __shfl_sync(kFull, base, 0);
#pragma unroll
                for (int sb = 0; sb < kSub; ++sb) {
                    if (masks[sb])
                        queue[(base + __popc(votes[sb] & lt_mask)) & qmask] =
                            make_int2(c_tile + sb * 32 * kBlock, u | (masks[sb] << 16));
                    base += __popc(votes[sb]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&ss->tiles_done, 1);
            }
            continue;
        }
        if (gate_done && avail == 0) break;
        if (++idle > (1 << 24)) __trap();  // never seen; a hang here would take the GPU with it
        __nanosleep(20);
    }
    pend.template push<kGather, kUW>(cx, best);
    __syncthreads();  // every finalist of this sweep is in the queue
    drain_finalists<kT, kGather, kUW>(cx, best);
}

// The same work with the classic schedule, for layouts with two CTAs per SM (the other CTA fills this one's barrier
// waits, and short batches make the ring's bookkeeping the larger cost): one ROUND of the survivor queue, filled by
// the gate before a barrier; warps take batches of 32 entries until the queue is empty, then drain the finalists.
template <int kT, int kBlock, int kGather, bool kUW = true>
__device__ __forceinline__ void filter_round(const int2 *queue, int qfill, int *q_head, const WidthRec *rec, const float *cs32,
                                             const float *wd32, const float *__restrict__ tq32, double w0, double T,
                                             const Gate32 &g32, double eb_scale, FilterShared *fs, int2 *fq, float *fq_lo,
                                             int fq_cap, const ExactView<kGather> &view, Best &best, unsigned long long *stats,
                                             const float *w32 = nullptr, double ea_scale = 0.0)
{
    const int lane = threadIdx.x & 31;
    FilterCtx<kGather> cx;
    cx.view = view; cx.rec = rec; cx.fs = fs; cx.fq = fq; cx.fq_lo = fq_lo; cx.fq_cap = fq_cap; cx.stats = stats; cx.g32 = g32;
    cx.ea_scale = ea_scale;
    Threshold th;
    th.set(INFINITY, T);
    th.refresh(fs, T);
    const float slopTf = __double2float_ru(4e-14 * fabs(T));
    Pending<kBlock> pend;
    pend.clear();
    for (;;) {
        int h = 0;
        if (lane == 0) h = atomicAdd(q_head, 32);
        h = __shfl_sync(kFull, h, 0);
        if (h >= qfill) break;
        const bool have = h + lane < qfill;
        const int2 e = have ? queue[h + lane] : make_int2(0, 0);
        tap_batch<kBlock, kGather, kUW>(have, e, cx, cs32, wd32, tq32, w0, T, eb_scale, slopTf, th, pend, best, w32);
    }
    pend.template push<kGather, kUW>(cx, best);
    __syncthreads();  // every finalist of this round is in the queue
    drain_finalists<kT, kGather, kUW>(cx, best);
}

// max |x_k| over k < n, the same value in every thread (all threads must call; scratch: kT/32 doubles, shared)
template <int kT>
__device__ double block_max_abs(const double *__restrict__ x, int n, double *scratch)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double m = 0.0;
    for (int k = tid; k < n; k += kT) m = fmax(m, fabs(__ldg(x + k)));
#pragma unroll
    for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(kFull, m, off));
    __syncthreads();
    if (lane == 0) scratch[wid] = m;
    __syncthreads();
    double r = 0.0;
    for (int k = 0; k < kT / 32; ++k) r = fmax(r, scratch[k]);
    __syncthreads();
    return r;
}

// max |x_k y_k| and max |y_k| over k < n, the same values in every thread (unequal weights: the scales of the filter's bounds)
template <int kT>
__device__ void block_max_abs2(const double *__restrict__ x, const double *__restrict__ y, int n, double *scratch, double &mxy, double &my)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double a = 0.0, b = 0.0;
    for (int k = tid; k < n; k += kT) {
        const double yk = __ldg(y + k);
        a = fmax(a, fabs(__ldg(x + k) * yk));
        b = fmax(b, fabs(yk));
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        a = fmax(a, __shfl_xor_sync(kFull, a, off));
        b = fmax(b, __shfl_xor_sync(kFull, b, off));
    }
    __syncthreads();
    if (lane == 0) { scratch[wid] = a; scratch[kT / 32 + wid] = b; }
    __syncthreads();
    mxy = 0.0; my = 0.0;
    for (int k = 0; k < kT / 32; ++k) { mxy = fmax(mxy, scratch[k]); my = fmax(my, scratch[kT / 32 + k]); }
    __syncthreads();
}

// mean of x_k over k < n, the same value in every thread (all threads must call; scratch: kT/32 doubles, shared).  Only used
// as the detrending slope of the fp32 cumulative sums: any value is CORRECT there, a good one keeps them small.
template <int kT>
__device__ double block_mean(const double *__restrict__ x, int n, double *scratch)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double m = 0.0;
    for (int k = tid; k < n; k += kT) m += __ldg(x + k);
#pragma unroll
    for (int off = 16; off; off >>= 1) m += __shfl_xor_sync(kFull, m, off);
    __syncthreads();
    if (lane == 0) scratch[wid] = m;
    __syncthreads();
    double r = 0.0;
    for (int k = 0; k < kT / 32; ++k) r += scratch[k];
    __syncthreads();
    r /= (double)n;
    return (r == r && fabs(r) < 1e300) ? r : 0.0;  // NaN / inf in the data: no detrending
}

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

}  // namespace

#endif  // TLSB_DEVICE_CUH
