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
