// tlsb_resident.cu — tlsb_search_kernel: the period/duration/T0 search with the folded light curve resident in
// shared memory (or, kResident = false, streamed from a per-CTA global scratch that stays in L2).
// Reference: core.py:96-188 per period; see tlsb_device.cuh for the phases.
#include "tlsb_device.cuh"

namespace {

// kFilt: the fp32 gate + fp32 filter pass with exact fp64 evaluation of the finalists (resident layouts; with unequal
// weights two fp32 correlations).  Without it (streaming layout; unequal weights when the filter layout does not fit)
// every gate survivor is evaluated in fp64 by tap_block / block_min.
template <int kT, bool kResident, bool kUniformW, int kBlock, bool kFilt = (kResident && kUniformW)>
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
    // kFilter (resident, equal weights): X | sorted sample ids (u16) | tail | area.  X holds the sorted d, scanned in place
    // into the fp64 cumulative sums, which are then copied to this CTA's global scratch (only the rare exact evaluations
    // read them, from L2) - after that X is the survivor queue.  The area holds the sort's keys, histogram and
    // bucket-ordered ids first and then what the search reads: cs32 (detrended cumulative sums in fp32: the gate and the
    // screen), wd32 (the products w*d rounded to fp32: the taps) and the finalist queue.
    constexpr bool kFilter = kFilt;
    static_assert(!kFilt || kResident, "the filter layouts keep the folded curve in shared memory");
    const size_t cs_elems = (size_t)(NM + 2) & ~(size_t)1;
    double *cs, *w, *wd;
    idx_t *sid;
    int *H;
    int2 *queue;
    unsigned char *tail;
    float *wd32 = nullptr, *cs32 = nullptr, *w32 = nullptr;
    double *cs64g = nullptr;
    idx_t *sid_sorted = nullptr;
    int2 *fq = nullptr;
    float *fq_lo = nullptr;
    double *skey;
    if (kFilter) {
        cs = reinterpret_cast<double *>(smem_raw);
        w = wd = nullptr;
        queue = reinterpret_cast<int2 *>(smem_raw);
        // unequal weights: the finalist queue follows the survivor queue inside X (the area holds a third fp32 array, w32)
        const size_t qbytes = (size_t)a.qcap * 8 + (kUniformW ? 0 : (size_t)a.fq_cap * 12);
        const size_t xbytes = ((cs_elems * 8 > qbytes ? cs_elems * 8 : qbytes) + 15) & ~(size_t)15;
        sid_sorted = reinterpret_cast<idx_t *>(smem_raw + xbytes);
        tail = reinterpret_cast<unsigned char *>(sid_sorted) + (((size_t)N * sizeof(idx_t) + 15) & ~(size_t)15);
        unsigned char *area = tail + filter_tail_bytes(nU, kT);
        skey = reinterpret_cast<double *>(area);
        H = reinterpret_cast<int *>(skey + N);                    // 16-bit counters, two per word: NB + 1 entries
        sid = reinterpret_cast<idx_t *>(H + (NB + 2) / 2);
        cs32 = reinterpret_cast<float *>(area);
        wd32 = cs32 + (((size_t)NM + 2 + 3) & ~(size_t)3);
        if (kUniformW) {
            fq = reinterpret_cast<int2 *>(wd32 + (((size_t)NMP + 3) & ~(size_t)3));
        } else {
            w32 = wd32 + (((size_t)NMP + 3) & ~(size_t)3);
            fq = queue + a.qcap;
        }
        fq_lo = reinterpret_cast<float *>(fq + a.fq_cap);
        cs64g = reinterpret_cast<double *>(a.scratch + (size_t)blockIdx.x * a.scratch_per_cta);
    } else if (kResident) {
        cs = reinterpret_cast<double *>(smem_raw);
        w = cs + cs_elems;
        wd = kUniformW ? w : w + NMP;
        queue = reinterpret_cast<int2 *>(wd + NMP);
        H = reinterpret_cast<int *>(queue);
        sid = reinterpret_cast<idx_t *>(H + NB + 1);
        tail = reinterpret_cast<unsigned char *>(queue + a.qcap);
        skey = wd;  // the sort keys borrow the wd area; the sorted d go straight to cs[1..N]
    } else {
        unsigned char *g = a.scratch + (size_t)blockIdx.x * a.scratch_per_cta;
        cs = reinterpret_cast<double *>(g);
        w = cs + cs_elems;
        wd = kUniformW ? w : w + NMP;
        sid = reinterpret_cast<idx_t *>(wd + NMP);
        queue = reinterpret_cast<int2 *>(smem_raw);
        H = reinterpret_cast<int *>(queue + a.qcap);
        tail = smem_raw + (size_t)a.qcap * 8 + (((size_t)(NB + 1) * 4 + 15) & ~(size_t)15);
        skey = wd;
    }
    WidthRec *rec = reinterpret_cast<WidthRec *>(tail);                       // [nU]
    double *red_d = reinterpret_cast<double *>(rec + nU);                     // [2*kW + 2]
    FilterShared *fs = reinterpret_cast<FilterShared *>(red_d + 2 * kW + 2);  // (kFilter)
    SweepShared *ss = reinterpret_cast<SweepShared *>(fs + 1);                // (kFilter)
    int *red_i = reinterpret_cast<int *>(ss + 1);                             // [2*kW]
    int *s_next = red_i + 2 * kW;  // [4] period slot, "tiles left" flag, queue fill, queue head
    int *t_lo = s_next + 4;        // (kFilter) [nU] per width: first candidate, one past the last, gate tiles
    int *t_hi = t_lo + nU;
    int *t_tiles = t_hi + nU;

    for (int k = tid; k < nU * (int)(sizeof(WidthRec) / 4); k += kT)
        reinterpret_cast<int *>(rec)[k] = reinterpret_cast<const int *>(a.rec)[k];
    if (kFilter)
        for (int u = tid; u < nU; u += kT) {
            t_lo[u] = 0;
            t_hi[u] = a.rec[u].ncand;
            t_tiles[u] = a.rec[u].tiles;
        }

    const unsigned lt_mask = (1u << lane) - 1u;
    const double depth_min = a.depth_min;
    const int qstop = a.qcap - kW * 32 * kSub;  // gating pauses here: every warp can still add one tile
    // filter pass: scale of the error bound = max |w d| over the light curve (period independent)
    double eb_scale = 0.0, ea_scale = 0.0;  // max |w d| and (unequal weights) max w over the light curve
    Gate32 g32;
    g32.mu = 0.0; g32.err = 0.0; g32.depth_min = depth_min;
    if (kFilter) {
        if (kUniformW) {
            eb_scale = a.w0 * block_max_abs<kT>(a.dval, N, red_d);
        } else {
            block_max_abs2<kT>(a.dval, a.wval, N, red_d, eb_scale, ea_scale);
        }
        if (!a.filter) eb_scale = ea_scale = INFINITY;
        g32.mu = block_mean<kT>(a.dval, N, red_d);
    }

    for (;;) {
        if (tid == 0) {
            s_next[0] = atomicAdd(a.counter, 1);
            s_next[1] = 0;
            s_next[2] = 0;
            s_next[3] = 0;
            if (kFilter) {
                fs->U = (unsigned long long)__double_as_longlong((double)N);  // core.py:46: a model must beat N to count
                fs->fq_fill = 0;
                *ss = SweepShared{};
            }
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
        fold_sort_gather<kT, idx_t, (!kUniformW && !kFilter), false, (kResident ? kResSortU : 4), (kResident ? kResHScanItems : kScanItems), kFilter, kFilter>(
            a.t, 0.0, r, N, NB, H, skey, sid, a.dval, a.wval, cs + 1, w, reinterpret_cast<int *>(red_d), sid_sorted);
        if (tid == 0) cs[0] = 0.0;
        __syncthreads();  // the sorted d sit in cs[1..N]; the keys (in the wd area) are dead
        float cmax = 0.f;
        if (kFilter && tid == 0) cs32[0] = 0.f;
        double tpart = wrap_weight_scan<kT, kUniformW, (kResident ? kResScanItems : kScanItems), !kFilter, kFilter, kFilter,
                                        (kFilter && !kUniformW), idx_t>(
            cs + 1, w, wd, a.w0, N, NM, NMP, red_d, 0, 0.0, wd32, cs32 + 1, g32.mu, &cmax, a.wval, sid_sorted, w32);
        if (kFilter) {  // the fp64 cumulative sums leave the SM: only bound_one / eval_exact_warp read them again (L2)
            for (int k = tid; k <= NM; k += kT) __stcg(cs64g + k, cs[k]);
#pragma unroll
            for (int off = 16; off; off >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(kFull, cmax, off));
            if (lane == 0) red_i[wid] = __float_as_int(cmax);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) tpart += __shfl_xor_sync(kFull, tpart, off);
        if (lane == 0) red_d[kW + 1 + wid] = tpart;
        __syncthreads();  // X (the fp64 cumulative sums) is dead from here on: it becomes the survivor queue
        double T = 0.0;
        for (int k = 0; k < kW; ++k) T += red_d[kW + 1 + k];  // T = sum w d^2 over the unpatched curve; fixed order
        if (kFilter) {
            float cm = 0.f;
            for (int k = 0; k < kW; ++k) cm = fmaxf(cm, __int_as_float(red_i[k]));
            g32.set_err(cm, NM);
            if (kT > 256) {  // the survivor ring: a slot is valid when it is non-zero (a barrier follows inside the sweep set-up)
                for (int k = tid; k < a.qcap; k += kT) reinterpret_cast<unsigned long long *>(queue)[k] = 0ull;
                __syncthreads();
            }
        }

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

        constexpr bool kDynamic = kFilter && kT > 256;  // one big CTA per SM: nothing else would fill a barrier wait
        ExactView<1> view;
        if (kFilter) {
            view.cs = cs64g; view.wd = nullptr; view.dval = a.dval; view.sid = reinterpret_cast<const unsigned short *>(sid_sorted);
            view.tq = a.tq; view.w0 = a.w0; view.T = T; view.N = N; view.wval = a.wval; view.w = nullptr; view.sid32 = nullptr;
        }
        if constexpr (kDynamic) {
            // B1 + B2 as one barrier-free sweep: warps switch between gating tiles and taking batches of survivors
            // (fp32 correlation, screen, bounds); finalists are evaluated in fp64 by whole warps at the end
            const int tile_total = rec[ulo].cum + rec[ulo].tiles - rec[uhi - 1].cum;
            sweep_filter<kT, kBlock, 1, kUniformW>(ss, queue, a.qcap - 1, tile_total, uhi, t_lo, t_hi, t_tiles, rec, cs32, wd32, a.tq32,
                                                      a.w0, T, g32, eb_scale, fs, fq, fq_lo, a.fq_cap, view, best, a.stats, w32, ea_scale);
        } else {
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
                    const int c_tile = (g - u_begin) * kTile + lane * kBlock;
                    int masks[kSub];
                    unsigned votes[kSub];
                    int total = 0;
                    if constexpr (kFilter) {  // fp32 gate on the detrended cumulative sums: a superset (bound_one settles it)
                        const float thr = g32.thr(W);
                        if (X == 1) {
    #pragma unroll
                            for (int sb = 0; sb < kSub; ++sb)
                                masks[sb] = gate_block32<kBlock, true>(cs32, c_tile + sb * 32 * kBlock, ncand, W, 1, thr);
                        } else {
    #pragma unroll
                            for (int sb = 0; sb < kSub; ++sb)
                                masks[sb] = gate_block32<kBlock, false>(cs32, c_tile + sb * 32 * kBlock, ncand, W, X, thr);
                        }
                    } else {
                        const double invW = rec[u].invW;
                        if (X == 1) {
    #pragma unroll
                            for (int sb = 0; sb < kSub; ++sb)
                                masks[sb] = gate_block<kBlock, true>(cs, c_tile + sb * 32 * kBlock, ncand, W, 1, invW, depth_min);
                        } else {
    #pragma unroll
                            for (int sb = 0; sb < kSub; ++sb)
                                masks[sb] = gate_block<kBlock, false>(cs, c_tile + sb * 32 * kBlock, ncand, W, X, invW, depth_min);
                        }
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
                if constexpr (kFilter) {
                    filter_round<kT, kBlock, 1, kUniformW>(queue, qfill, &s_next[3], rec, cs32, wd32, a.tq32, a.w0, T, g32, eb_scale, fs,
                                                              fq, fq_lo, a.fq_cap, view, best, a.stats, w32, ea_scale);
                } else
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
                if (tid == 0) { s_next[1] = 0; s_next[2] = 0; s_next[3] = 0; if (kFilter) fs->fq_fill = 0; }
                __syncthreads();
            }
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
        // Filter paths: the barrier in front of drain_finalists already separates every read of red_d / red_i (T, the
        // largest |cs32|) from the writes below.
        if (!kFilter) __syncthreads();  // everyone is done reading red_d (T) and the queue before they are reused
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
        // no barrier here: the one at the top of the loop (after thread 0 has taken the next period) is the first point at
        // which red_d / red_i / the scheduler words are written again by anyone but warp 0 itself
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

}  // namespace

namespace tlsb {

#define TLSB_GO(K)                                                                                        \
    do {                                                                                                  \
        cudaError_t e_ = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e_ != cudaSuccess) return e_;                                                                 \
        K<<<grid, threads, smem, s>>>(a);                                                                 \
        return cudaGetLastError();                                                                        \
    } while (0)

cudaError_t launch_search_resident(const SearchArgs &a, int threads, bool resident, bool uniform_w, int kb, int grid,
                                   size_t smem, cudaStream_t s)
{
    const bool uni = uniform_w;
    if (resident) {
        if (threads == 256) {
            if (uni && kb == 7) TLSB_GO((tlsb_search_kernel<256, true, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_kernel<256, true, true, 5>));
            else if (a.fq_cap > 0 && kb == 7) TLSB_GO((tlsb_search_kernel<256, true, false, 7, true>));
            else if (a.fq_cap > 0) TLSB_GO((tlsb_search_kernel<256, true, false, 5, true>));
            else TLSB_GO((tlsb_search_kernel<256, true, false, 5>));
        } else {
            if (uni && kb == 7) TLSB_GO((tlsb_search_kernel<512, true, true, 7>));
            else if (uni) TLSB_GO((tlsb_search_kernel<512, true, true, 5>));
            else if (a.fq_cap > 0 && kb == 7) TLSB_GO((tlsb_search_kernel<512, true, false, 7, true>));
            else if (a.fq_cap > 0) TLSB_GO((tlsb_search_kernel<512, true, false, 5, true>));
            else TLSB_GO((tlsb_search_kernel<512, true, false, 5>));
        }
    } else {
        if (uni && kb == 7) TLSB_GO((tlsb_search_kernel<256, false, true, 7>));
        else if (uni) TLSB_GO((tlsb_search_kernel<256, false, true, 5>));
        else TLSB_GO((tlsb_search_kernel<256, false, false, 5>));
    }
}
#undef TLSB_GO

}  // namespace tlsb
