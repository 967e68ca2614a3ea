// tlsb_tiled.cu — tlsb_search_tiled_kernel: light curves too long for shared memory.
// ------------------------------------------------------------------------------------------
// Tiled path (light curves too long for the resident path): phase A sorts the fold on chip one
// phase segment at a time and writes the folded arrays to a per-CTA global scratch; phase B walks
// the folded curve in POSITION CHUNKS.  One elected thread stages cs32[a0, a0+C) and
// wd32[a0, a0+C) (and w32 with per-point dy; cs, w, w*d in fp64 for the all-fp64 kernels) of the
// chunk into shared memory with 1-D bulk async copies (TMA, cp.async.bulk -> mbarrier
// complete_tx), then the same fp32 gate / survivor ring / fp32 filter pass as the resident kernel
// runs from shared memory for every candidate block that STARTS inside [a0, a0+TP),
// TP = C - (widest admissible window + the tap loop's overshoot); finalists are evaluated in fp64
// from the scratch (cumulative sums) and the light curve (w, d through the sorted sample ids).
// Each folded sample is read from L2 once per chunk instead of twice per admissible width.
// ------------------------------------------------------------------------------------------
#include "tlsb_device.cuh"

namespace {

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
#ifndef TLSB_STCS
#define TLSB_STCS 1  // streaming stores for the fp64 arrays of the filter layouts
#endif
template <int kT, bool kUniformW, bool kFilt>
__device__ __forceinline__ bool sort_on_chip(const SearchArgs &a, double r, unsigned char *area, int *cnt,
                                             unsigned *gid, double *cs1, double *w, double *wd, float *wd32,
                                             int nmp_even, double *red_d, double &tpart_out, float *cs32_1, double mu,
                                             float &cmax_out, float *w32, unsigned *gsid)
{
#ifndef TLSB_SEG_RANK_U
#define TLSB_SEG_RANK_U 4
#endif
#ifndef TLSB_SEG_PART_U
#define TLSB_SEG_PART_U 4
#endif
    constexpr int kU = TLSB_SEG_RANK_U;    // independent chains of the rank / gather loop
    constexpr int kUP = TLSB_SEG_PART_U;   // ... of the partition pass
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = a.N, M = a.M, NM = N + M, S = a.seg_cap, ns = a.n_seg;
    const unsigned lt_mask = (1u << lane) - 1u;
    double *val_s = reinterpret_cast<double *>(area);            // [S] sorted d of the segment
    double *wv_s = val_s + S;                                    // [S] sorted w (unequal weights only)
    double *skey_s = kUniformW ? wv_s : wv_s + S;                // [S] keys, bucket order
    // fine-bucket histogram: 2 S buckets as 16-bit counters packed two per word (S < 32768 keys per segment), S + 2 words
    unsigned *Hw = reinterpret_cast<unsigned *>(skey_s + S);
    const unsigned short *H16 = reinterpret_cast<const unsigned short *>(Hw);
    unsigned *sid_s = Hw + S + 2;                                // [S] sample ids, bucket order
    const int S2 = 2 * S;
    const double dns = (double)ns, dS = (double)S2;

    for (int j = tid; j <= ns; j += kT) cnt[j] = 0;
    __syncthreads();
    // ---- partition: one pass over t ------------------------------------------------------------
    double tnext[kUP];  // the time stamps of the next trip are in flight while this one is folded and appended
#pragma unroll
    for (int u = 0; u < kUP; ++u) tnext[u] = (wid * 32 + u * kT + lane < N) ? __ldcs(a.t + wid * 32 + u * kT + lane) : 0.0;
    for (int kb = wid * 32; kb < N; kb += kT * kUP) {  // warp-uniform bounds: every lane reaches the match
        double tv[kUP];
#pragma unroll
        for (int u = 0; u < kUP; ++u) {
            tv[u] = tnext[u];
            const int kn = kb + kT * kUP + u * kT + lane;
            tnext[u] = kn < N ? __ldcs(a.t + kn) : 0.0;
        }
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
                    gid[(size_t)sg * S + slot] = (unsigned)k;  // the list holds ids only: the phase is folded again from t[id]
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
        return fb < 0 ? 0 : (fb < S2 - 1 ? fb : S2 - 1);
    };
    double tpart = 0.0, carry = 0.0;
    float cmax = 0.f;
    int off = 0;
    for (int j = 0; j < ns; ++j) {
        const int nj = cnt[j];
        if (nj == 0) continue;
        const unsigned *li = gid + (size_t)j * S;
        for (int b = tid; b <= S; b += kT) Hw[b] = 0u;  // entries 0 .. 2 S + 1
        if (j + 1 < ns) {  // the next segment's lists were written a while ago and may have left L2: fetch them back now
            const int nn = cnt[j + 1];
            const char *pi = reinterpret_cast<const char *>(li + S);
            for (int b = tid * 128; b < nn * 4; b += kT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pi + b));
        }
        __syncthreads();
        // histogram: every key of this thread stays in registers together with its bucket and its
        // arrival index inside the bucket, so the scatter below needs neither a reload nor atomics
        double ph[kSegPerThread];
        unsigned id[kSegPerThread], where[kSegPerThread];
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            const int q = tid + i * kT;
            id[i] = q < nj ? li[q] : 0u;
        }
        // the same fold as the partition pass, bit for bit (time stamps from L2: 4 instead of 12 bytes per key went through
        // the lists - cfg-2: 4.91 -> 3.58 MB of DRAM traffic per period and 2 % less time)
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) ph[i] = __ldg(a.t + id[i]);
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) ph[i] = fold_phase(ph[i], r);
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            if (tid + i * kT < nj) {
                const int fb = fine(ph[i], j), e = fb + 1, sh = (e & 1) * 16;
                where[i] = (((atomicAdd(&Hw[e >> 1], 1u << sh) >> sh) & 0xffffu) << 16) | (unsigned)fb;
            }
        }
        __syncthreads();
        block_inclusive_scan_u16x2<kT, kSegScanItems>(Hw, S + 1, reinterpret_cast<int *>(red_d));  // H16[b] = keys in buckets < b
#pragma unroll
        for (int i = 0; i < kSegPerThread; ++i) {
            if (tid + i * kT < nj) {
                const int pos = (int)H16[where[i] & 0xffffu] + (int)(where[i] >> 16);
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
                lo[u] = (int)H16[fb];
                hi[u] = q0 + u * kT < nj ? (int)H16[fb + 1] : lo[u];
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
                    if (kFilt) __stcs(gsid + off + rank[u], sidq[u]);  // the exact evaluations gather w, d through these
                }
            }
        }
        __syncthreads();
        for (int q = tid; q < nj; q += kT) {  // emit: wd, T, w, and the samples the wrap repeats
            const int pos = off + q;
            const double d = val_s[q];
            const double wv = kUniformW ? a.w0 : wv_s[q];
            const double x = wv * d;
            if (!kFilt) wd[pos] = x;          // (filter layouts: finalists rebuild w*d from the sorted ids)
            if (kFilt) wd32[pos] = (float)x;  // the filter pass's samples
            tpart = fma(x, d, tpart);
            if (!kUniformW && !kFilt) w[pos] = wv;
            if (kFilt && !kUniformW) w32[pos] = (float)wv;
            if (pos < M) {
                cs1[N + pos] = d;
                if (!kUniformW) w[N + pos] = wv;
            }
        }
        __syncthreads();
        block_inclusive_scan<kT, double, kSegScanItems>(val_s, nj, red_d);
        for (int q = tid; q < nj; q += kT) {
            const double c = carry + val_s[q];
            // filter layouts: the fp64 sums are read again only by the rare exact evaluations - streaming stores keep them
            // from pushing the fp32 arrays and the segment lists (re-read within the period) out of L2
            if (kFilt && TLSB_STCS) __stcs(cs1 + off + q, c); else cs1[off + q] = c;
            if (kFilt) {  // the detrended fp32 copy the gate and the screen read (tlsb_device.cuh: fp32 gate)
                const float c32 = (float)fma(-(double)(off + q + 1), mu, c);
                cs32_1[off + q] = c32;
                cmax = fmaxf(cmax, fabsf(c32));
            }
        }
        carry += val_s[nj - 1];
        off += nj;
        __syncthreads();
    }
    // positions N .. NM-1 hold the wrapped d: continue the cumulative sum, weight them, zero the slack
    float cmax2 = 0.f;
    wrap_weight_scan<kT, kUniformW, kSegScanItems, !kFilt, kFilt, kFilt, false, unsigned short, (kFilt && !kUniformW)>(
        cs1, w, wd, a.w0, N, NM, nmp_even, red_d, N, carry, wd32, cs32_1, mu, &cmax2, nullptr, nullptr, w32);
    tpart_out = tpart;
    cmax_out = fmaxf(cmax, cmax2);
    return true;
}

// kFilt: fp32 gate + fp32 filter pass from chunks of cs32 / wd32 (/ w32 with unequal weights), exact fp64 evaluation of
// the finalists from the CTA's scratch; without it (unequal weights, TLSB_WFILTER=0) chunks of cs / w / w*d in fp64.
template <int kT, bool kUniformW, int kBlock, bool kFilt = kUniformW>
__global__ void __launch_bounds__(kT, (kT <= 256 ? 2 : 1)) tlsb_search_tiled_kernel(const __grid_constant__ SearchArgs a)
{
    static_assert(kFilt || !kUniformW, "equal weights always take the filter");
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
    const size_t nmp4 = ((size_t)NMP + 3) & ~(size_t)3;
    const size_t cs4 = ((size_t)NM + 2 + 3) & ~(size_t)3;
    // equal weights: the detrended cumulative sums and the products w*d rounded to fp32 (what the chunks stage)
    float *cs32 = reinterpret_cast<float *>(wd + nmp_even);
    float *wd32 = cs32 + cs4;
    float *w32 = wd32 + nmp4;  // unequal weights only
    unsigned *sid = !kFilt ? reinterpret_cast<unsigned *>(wd + nmp_even)
                           : reinterpret_cast<unsigned *>(kUniformW ? wd32 + nmp4 : w32 + nmp4);
    double *skey = wd;  // the sort keys borrow the wd area; the sorted d go straight to cs[1..N]

    // ---- shared: queue | finalist queue | chunk of cs | chunk of w*d in fp32 (equal weights) or chunks of w and w*d
    //              | records, tables, scratch ----
    int2 *queue = reinterpret_cast<int2 *>(smem_raw);
    int2 *fq = queue + a.qcap;
    float *fq_lo = reinterpret_cast<float *>(fq + a.fq_cap);
    // filter layouts: chunks of cs32 and wd32 (8 bytes per folded sample; 12 with w32 for unequal weights); else chunks of
    // cs, w, w*d in fp64 (24 bytes)
    double *cs_s = reinterpret_cast<double *>(fq_lo + a.fq_cap);
    double *w_s = cs_s + C;
    double *wd_s = kUniformW ? w_s : w_s + C;
    float *cs32_s = reinterpret_cast<float *>(cs_s);
    float *wd32_s = cs32_s + C;
    float *w32_s = wd32_s + C;
    int *H = reinterpret_cast<int *>(cs_s);  // phase A only: the histogram borrows the chunk area
    WidthRec *rec = kFilt ? reinterpret_cast<WidthRec *>(kUniformW ? wd32_s + C : w32_s + C)
                          : reinterpret_cast<WidthRec *>(wd_s + C);  // [nU]
    double *red_d = reinterpret_cast<double *>(rec + nU);                     // [2*kW + 2]
    FilterShared *fs = reinterpret_cast<FilterShared *>(red_d + 2 * kW + 2);
    SweepShared *ss = reinterpret_cast<SweepShared *>(fs + 1);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(ss + 1);
    int *red_i = reinterpret_cast<int *>(bar + 1);                            // [2*kW]
    int *s_next = red_i + 2 * kW;  // [8] period slot, "tiles left" flag, queue fill, queue head, chunk tiles
    int *ch_lo = s_next + 8;       // [nU] first candidate of the chunk, per width
    int *ch_hi = ch_lo + nU;       // [nU] one past the last
    int *ch_tiles = ch_hi + nU;    // [nU]
    int *seg_cnt = ch_tiles + nU;  // [kMaxSegments + 1] on-chip sort: keys per phase segment
    // segment lists of the on-chip sort, behind the arrays above
    // (at least N entries: a period whose sort falls back to the global scratch keeps its sorted ids here)
    unsigned *gid = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(sid) + (((size_t)N * 4 + 15) & ~(size_t)15));

    for (int k = tid; k < nU * (int)(sizeof(WidthRec) / 4); k += kT)
        reinterpret_cast<int *>(rec)[k] = reinterpret_cast<const int *>(a.rec)[k];
    if (kFilt)  // the survivor ring: a slot is valid when it is non-zero, readers clear it
        for (int k = tid; k < a.qcap; k += kT) reinterpret_cast<unsigned long long *>(queue)[k] = 0ull;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned parity = 0;

    const unsigned lt_mask = (1u << lane) - 1u;
    const double depth_min = a.depth_min;
    const int qstop = a.qcap - kW * 32 * kSub;
    double eb_scale = 0.0, ea_scale = 0.0;  // filter pass: scales of the error bounds = max |w d| (and max w) over the light curve
    Gate32 g32;
    g32.mu = 0.0; g32.err = 0.0; g32.depth_min = depth_min;
    if (kFilt) {
        if (kUniformW) eb_scale = a.w0 * block_max_abs<kT>(a.dval, N, red_d);
        else block_max_abs2<kT>(a.dval, a.wval, N, red_d, eb_scale, ea_scale);
        if (!a.filter) eb_scale = ea_scale = INFINITY;
        g32.mu = block_mean<kT>(a.dval, N, red_d);
    }

    for (;;) {
        if (tid == 0) {
            s_next[0] = atomicAdd(a.counter, 1);
            if (kFilt) {
                fs->U = (unsigned long long)__double_as_longlong((double)N);  // core.py:46: a model must beat N to count
                fs->fq_fill = 0;
            }
        }
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
        float cmax = 0.f;
        if (tid == 0) {
            cs[0] = 0.0;
            if (kFilt) cs32[0] = 0.f;
        }
        bool on_chip = false;
        if (a.seg_cap > 0)
            on_chip = sort_on_chip<kT, kUniformW, kFilt>(a, r, reinterpret_cast<unsigned char *>(cs_s), seg_cnt, gid,
                                                         cs + 1, w, wd, wd32, (int)nmp_even, red_d, tpart, cs32 + 1, g32.mu, cmax, w32, sid);
        if (!on_chip) {  // clustered phases (or no room for segments): sort in the global scratch
            if (tid == 0 && a.seg_cap > 0) atomicAdd(a.counter + 4, 1);
            // (filter layouts: the sorted ids go to the segment lists' memory, which this period does not use)
            fold_sort_gather<kT, unsigned, !kUniformW, false, 8, kScanItems, kFilt>(a.t, 0.0, r, N, NB, H, skey, sid, a.dval, a.wval,
                                                                                    cs + 1, w, reinterpret_cast<int *>(red_d), gid);
            __syncthreads();
            tpart = wrap_weight_scan<kT, kUniformW, kScanItems, !kFilt, kFilt, kFilt, false, unsigned short, (kFilt && !kUniformW)>(
                cs + 1, w, wd, a.w0, N, NM, (int)nmp_even, red_d, 0, 0.0, wd32, cs32 + 1, g32.mu, &cmax, nullptr, nullptr, w32);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            tpart += __shfl_xor_sync(kFull, tpart, off);
            cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, off));
        }
        if (lane == 0) {
            red_d[kW + 1 + wid] = tpart;
            red_i[wid] = __float_as_int(cmax);
        }
        __syncthreads();
        double T = 0.0;
        for (int k = 0; k < kW; ++k) T += red_d[kW + 1 + k];
        if (kFilt) {
            float cm = 0.f;
            for (int k = 0; k < kW; ++k) cm = fmaxf(cm, __int_as_float(red_i[k]));
            g32.set_err(cm, NM);
        }
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
        // finalists: fp64 cumulative sums from the scratch, w and w*d rebuilt from the light curve through the sorted ids
        ExactView<2> view;
        view.cs = cs; view.wd = nullptr; view.dval = a.dval; view.sid = nullptr; view.tq = a.tq; view.w0 = a.w0; view.T = T; view.N = N;
        view.wval = a.wval; view.w = nullptr; view.sid32 = on_chip ? sid : gid;
        auto sweep = [&](const double *csb, const double *wb, const double *wdb, const float *cs32b, const float *wd32b,
                         const float *w32b, int ub) {
            const int tile_end = s_next[4];
            if constexpr (kFilt) {  // barrier-free gate + filter sweep (tlsb_device.cuh)
                sweep_filter<kT, kBlock, 2, kUniformW>(ss, queue, a.qcap - 1, tile_end, ub, ch_lo, ch_hi, ch_tiles, rec, cs32b, wd32b,
                                                           a.tq32, a.w0, T, g32, eb_scale, fs, fq, fq_lo, a.fq_cap, view, best, a.stats,
                                                           w32b, ea_scale);
                return;
            }
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
                {
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
                }
                if (!more) break;
                __syncthreads();
                if (tid == 0) { s_next[1] = 0; s_next[2] = 0; s_next[3] = 0; }
                __syncthreads();
            }
        };

        if (ulo < uT) {
            const int TP = (C - window_need(rec[uT - 1].W, rec[uT - 1].X, kBlock)) & ~3;  // 16-byte aligned fp32 chunks
            const int i_last = NM - rec[ulo].W;  // the narrowest admissible width has the most offsets
            for (int a0 = 0; a0 <= i_last; a0 += TP) {
                fence_proxy_async();
                __syncthreads();  // phase A / the previous chunk are done with the staging area and the tables
                if (tid == 0) {
                    const int len_cs = min(C, (int)cs_elems - a0);
                    if (kFilt) {
                        const int len_c32 = min(C, (int)cs4 - a0), len_32 = min(C, (int)nmp4 - a0);
                        mbar_expect_tx(bar, 4u * (unsigned)(len_c32 + (kUniformW ? 1 : 2) * len_32));
                        bulk_copy_g2s(cs32_s, cs32 + a0, 4u * (unsigned)len_c32, bar);
                        bulk_copy_g2s(wd32_s, wd32 + a0, 4u * (unsigned)len_32, bar);
                        if (!kUniformW) bulk_copy_g2s(w32_s, w32 + a0, 4u * (unsigned)len_32, bar);
                        fs->fq_fill = 0;
                        *ss = SweepShared{};
                    } else {
                        const int len_wd = min(C, (int)nmp_even - a0);
                        mbar_expect_tx(bar, 8u * (unsigned)(len_cs + 2 * len_wd));
                        bulk_copy_g2s(cs_s, cs + a0, 8u * (unsigned)len_cs, bar);
                        bulk_copy_g2s(w_s, w + a0, 8u * (unsigned)len_wd, bar);
                        bulk_copy_g2s(wd_s, wd + a0, 8u * (unsigned)len_wd, bar);
                    }
                    s_next[1] = 0;
                    s_next[2] = 0;
                    s_next[3] = 0;
                }
                build_tables(a0, TP, ulo, uT);
                __syncthreads();
                mbar_wait(bar, parity);
                parity ^= 1u;
                sweep(cs_s - a0, w_s - a0, wd_s - a0, cs32_s - a0, wd32_s - a0, w32_s - a0, uT);
            }
        }
        if (uT < uhi) {  // the widest widths: gate and taps read the scratch through L1/L2
            __syncthreads();  // the last chunk's sweep is done with the queue and the tables
            if (tid == 0) {
                s_next[1] = 0; s_next[2] = 0; s_next[3] = 0;
                if (kFilt) { fs->fq_fill = 0; *ss = SweepShared{}; }
            }
            build_tables(0, 1 << 30, max(ulo, uT), uhi);
            __syncthreads();
            sweep(cs, w, wd, cs32, wd32, w32, uhi);
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

}  // namespace

namespace tlsb {

#define TLSB_GO(K)                                                                                        \
    do {                                                                                                  \
        cudaError_t e_ = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e_ != cudaSuccess) return e_;                                                                 \
        K<<<grid, threads, smem, s>>>(a);                                                                 \
        return cudaGetLastError();                                                                        \
    } while (0)

cudaError_t launch_search_tiled(const SearchArgs &a, int threads, bool uniform_w, int kb, int grid, size_t smem,
                                cudaStream_t s)
{
    const bool uni = uniform_w;
    if (threads == 256) {
        if (uni && kb == 7) TLSB_GO((tlsb_search_tiled_kernel<256, true, 7>));
        else if (uni) TLSB_GO((tlsb_search_tiled_kernel<256, true, 5>));
        else if (a.fq_cap > 0 && kb == 7) TLSB_GO((tlsb_search_tiled_kernel<256, false, 7, true>));
        else if (a.fq_cap > 0) TLSB_GO((tlsb_search_tiled_kernel<256, false, 5, true>));
        else TLSB_GO((tlsb_search_tiled_kernel<256, false, 5>));
    } else {
        if (uni && kb == 7) TLSB_GO((tlsb_search_tiled_kernel<512, true, 7>));
        else if (uni) TLSB_GO((tlsb_search_tiled_kernel<512, true, 5>));
        else if (a.fq_cap > 0 && kb == 7) TLSB_GO((tlsb_search_tiled_kernel<512, false, 7, true>));
        else if (a.fq_cap > 0) TLSB_GO((tlsb_search_tiled_kernel<512, false, 5, true>));
        else TLSB_GO((tlsb_search_tiled_kernel<512, false, 5>));
    }
}
#undef TLSB_GO

}  // namespace tlsb
