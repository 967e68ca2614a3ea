// tlsb_aux_kernels.cu — the small kernels around the search: the per-period plan (admissible widths +
// processing order), the per-light-curve preparation (d = 1 - y, w = 1/dy^2), final_T0_fit
// (stats.py:135-204) and the ascending-period row gather of the batch pipeline.
#include "tlsb_device.cuh"

namespace {

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

// ------------------------------------------------------------------------------------------
// final_T0_fit (stats.py:135-204): at the best period, every trial epoch Tx folds the light
// curve with fold(t, period, Tx) (core.py:9-12), sorts it stably (stats.py:173), rolls the
// sorted flux by dur/2+1 (stats.py:186-190), and sums the weighted residuals against the
// in-transit model (first dur samples) and against 1 (the rest).  The reference overwrites
// its weights with a SECOND roll of the rolled flux (stats.py:191), so the weight of slot k is
// 1 / flux_sorted[k - 2*shift]^2 and dy plays no part; that quirk is kept.
// One CTA per trial epoch, persistent; same fold + bucket-rank sort as the search kernel.
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

__global__ void tlsb_gather_rows_kernel(const double *__restrict__ records, size_t record_stride,
                                        const int *__restrict__ order, double *__restrict__ out, int P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t c = blockIdx.y;
    if (k < P) out[c * P + k] = records[c * record_stride + order[k]];  // chi2 plane, ascending period
}

// Multi-GPU (main.py:190-196 wants one array per quantity in the job's period order): the all-gathered buffer is
// rank-major, rank r holding periods r, r + world, ... as three planes of its own length n_r plus a status word.
// One thread per period writes the final three planes; thread 0 adds up the shards' status words.
__global__ void tlsb_unshard_kernel(const long long *__restrict__ g, int n, int world, int words_per_rank,
                                    long long *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int r = k % world, j = k / world;
        const int n_r = (n - r + world - 1) / world;
        const long long *src = g + (size_t)r * words_per_rank;
        out[k] = src[j];
        out[(size_t)n + k] = src[n_r + j];
        out[2 * (size_t)n + k] = src[2 * n_r + j];
    }
    if (k == 0) {
        long long status = 0;
        for (int r = 0; r < world; ++r) status += g[(size_t)r * words_per_rank + 3 * ((n - r + world - 1) / world)];
        out[3 * (size_t)n] = status;
    }
}

}  // namespace

namespace tlsb {

cudaError_t launch_unshard(const long long *gathered, int n, int world, int words_per_rank, long long *out, cudaStream_t s)
{
    tlsb_unshard_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(gathered, n, world, words_per_rank, out);
    return cudaGetLastError();
}

cudaError_t launch_plan(const PlanArgs &a, int grid, cudaStream_t s)
{
    tlsb_plan_kernel<<<grid, kPlanThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_prepare(const double *y, const double *dy, double *dval, double *wval, size_t n, cudaStream_t s)
{
    tlsb_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, dy, dval, wval, n);
    return cudaGetLastError();
}

template <typename K>
static cudaError_t launch_t0(K kernel, const T0Args &a, int grid, int threads, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_t0fit(const T0Args &a, int threads, bool resident, int grid, size_t smem, cudaStream_t s)
{
    if (resident && threads == 256) return launch_t0(tlsb_t0fit_kernel<256, true>, a, grid, 256, smem, s);
    if (resident) return launch_t0(tlsb_t0fit_kernel<512, true>, a, grid, 512, smem, s);
    return launch_t0(tlsb_t0fit_kernel<256, false>, a, grid, 256, smem, s);
}

cudaError_t launch_gather_rows(const double *records, size_t record_stride, const int *order, double *out, int P,
                               int n_curves, cudaStream_t s)
{
    dim3 grid((unsigned)((P + 255) / 256), (unsigned)n_curves);
    tlsb_gather_rows_kernel<<<grid, 256, 0, s>>>(records, record_stride, order, out, P);
    return cudaGetLastError();
}

}  // namespace tlsb
