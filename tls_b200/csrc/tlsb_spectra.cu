// tlsb_spectra.cu — chi2[P] -> SR, power_raw, power (median-detrended SDE spectrum), SDE_raw, SDE
// on the device: stats.spectra (stats.py:105-132) + helpers.running_median (helpers.py:93-108).
//
// Three small kernels per call, batched over light curves (one chi2 row per curve):
//   pre     one CTA per curve: min(chi2), SR = min/chi2, mean and population std of SR,
//           SDE_raw = (1 - mean)/std, power_raw = (SR - mean) * SDE_raw / max(SR - mean)
//   median  one thread per period: median of its window of `win` power_raw samples by rank
//           counting from a shared-memory tile (edge windows replicated exactly as the
//           reference pads them), detrended value power_raw - median
//   post    one CTA per curve: subtract the mean, SDE = max / population std, rescale so that
//           the peak equals SDE, first arg-max
// Reductions are fp64, fixed order (deterministic).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tlsb200.h"
#include "tlsb_internal.h"

namespace {

constexpr int kT = 1024;        // threads of the per-curve kernels
constexpr int kMedT = 256;      // threads (= outputs) per CTA of the median kernel
constexpr unsigned kFull = 0xffffffffu;

struct SpectraArgs {
    const double *chi2;   // [n_curves][P]
    double *SR;           // [n_curves][P]
    double *power_raw;    // [n_curves][P]
    double *power;        // [n_curves][P]
    double *scal;         // [n_curves][4]: SDE_raw, SDE, min chi2, max of the final power
    long long *argmax;    // [n_curves]
    int P;
    int win;              // median window (samples)
    int nwin;             // number of full windows, P - win + 1
    int detrend;          // len(power_raw) > 2 * kernel (stats.py:119)
};

enum { kSum = 0, kMin = 1, kMax = 2 };

template <int kOp> __device__ __forceinline__ double combine(double a, double b)
{
    if (kOp == kSum) return a + b;
    if (kOp == kMin) return fmin(a, b);
    return fmax(a, b);
}

// all threads get the result; `red` holds kT/32 doubles
template <int kOp> __device__ double block_reduce(double v, double *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off; off >>= 1) v = combine<kOp>(v, __shfl_xor_sync(kFull, v, off));
    __syncthreads();  // red may still be read from the previous call
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = red[0];
    for (int k = 1; k < kT / 32; ++k) r = combine<kOp>(r, red[k]);
    return r;
}

__global__ void __launch_bounds__(kT) tlsb_spectra_pre_kernel(const SpectraArgs a)
{
    __shared__ double red[kT / 32];
    const int P = a.P, tid = threadIdx.x;
    const size_t row = (size_t)blockIdx.x * P;
    const double *chi2 = a.chi2 + row;
    double *SR = a.SR + row, *pr = a.power_raw + row;

    double mn = INFINITY;
    for (int i = tid; i < P; i += kT) mn = fmin(mn, chi2[i]);
    mn = block_reduce<kMin>(mn, red);
    double s = 0.0;
    for (int i = tid; i < P; i += kT) {
        const double v = mn / chi2[i];  // stats.py:106
        SR[i] = v;
        s += v;
    }
    const double mean = block_reduce<kSum>(s, red) / P;
    double q = 0.0, mx = -INFINITY;
    for (int i = tid; i < P; i += kT) {
        const double d = SR[i] - mean;
        q += d * d;
        mx = fmax(mx, d);
    }
    const double sd = sqrt(block_reduce<kSum>(q, red) / P);  // numpy.std: population
    mx = block_reduce<kMax>(mx, red);
    const double sde_raw = (1.0 - mean) / sd;  // stats.py:107
    const double scale = sde_raw / mx;         // stats.py:111
    for (int i = tid; i < P; i += kT) pr[i] = (SR[i] - mean) * scale;
    if (tid == 0) {
        a.scal[4 * blockIdx.x + 0] = sde_raw;
        a.scal[4 * blockIdx.x + 2] = mn;
    }
}

// helpers.py:93-108: median of every full window, first/last value repeated in front/behind.
__global__ void __launch_bounds__(kMedT) tlsb_spectra_median_kernel(const SpectraArgs a)
{
    extern __shared__ double tile[];  // kMedT + win samples
    const int P = a.P, win = a.win, nwin = a.nwin;
    const size_t row = (size_t)blockIdx.y * P;
    const double *pr = a.power_raw + row;
    const int front = (int)((P - nwin) * 0.5);  // helpers.py:103
    const int i0 = blockIdx.x * kMedT;
    // window of output i starts at clamp(i - front, 0, nwin-1)
    int jlo = i0 - front;
    jlo = jlo < 0 ? 0 : (jlo > nwin - 1 ? nwin - 1 : jlo);
    for (int k = threadIdx.x; k < kMedT + win; k += kMedT) tile[k] = (jlo + k < P) ? pr[jlo + k] : 0.0;
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= P) return;
    int j = i - front;
    j = j < 0 ? 0 : (j > nwin - 1 ? nwin - 1 : j);
    const double *wv = tile + (j - jlo);
    // order statistics by rank counting; ties broken by position so that ranks are unique
    const int r_hi = win / 2, r_lo = (win & 1) ? r_hi : r_hi - 1;
    double v_lo = 0.0, v_hi = 0.0;
    for (int x = 0; x < win; ++x) {
        const double v = wv[x];
        int rank = 0;
        for (int y = 0; y < win; ++y) {
            const double u = wv[y];
            rank += (u < v) || (u == v && y < x);
        }
        if (rank == r_lo) v_lo = v;
        if (rank == r_hi) v_hi = v;
    }
    const double med = (win & 1) ? v_hi : (v_lo + v_hi) * 0.5;  // numpy.median of an even count: mean of the middle two
    a.power[row + i] = pr[i] - med;  // stats.py:121
}

__global__ void __launch_bounds__(kT) tlsb_spectra_post_kernel(const SpectraArgs a)
{
    __shared__ double red[kT / 32];
    __shared__ long long first;
    const int P = a.P, tid = threadIdx.x;
    const size_t row = (size_t)blockIdx.x * P;
    double *pw = a.power + row;
    const double *pr = a.power_raw + row;
    double peak;
    if (a.detrend) {
        double s = 0.0;
        for (int i = tid; i < P; i += kT) s += pw[i];
        const double mean = block_reduce<kSum>(s, red) / P;
        double s2 = 0.0;
        for (int i = tid; i < P; i += kT) {
            const double v = pw[i] - mean;  // stats.py:124
            pw[i] = v;
            s2 += v;
        }
        const double mean2 = block_reduce<kSum>(s2, red) / P;  // numpy.std takes its own mean again
        double q = 0.0, mx = -INFINITY;
        for (int i = tid; i < P; i += kT) {
            const double v = pw[i], d = v - mean2;
            q += d * d;
            mx = fmax(mx, v);
        }
        const double sd = sqrt(block_reduce<kSum>(q, red) / P);
        mx = block_reduce<kMax>(mx, red);
        const double sde = mx / sd;      // stats.py:125: max(power / std)
        const double scale = sde / mx;   // stats.py:127
        double mx2 = -INFINITY;
        for (int i = tid; i < P; i += kT) {
            const double v = pw[i] * scale;
            pw[i] = v;
            mx2 = fmax(mx2, v);
        }
        peak = block_reduce<kMax>(mx2, red);
        if (tid == 0) a.scal[4 * blockIdx.x + 1] = sde;
    } else {  // stats.py:130-131
        double mx = -INFINITY;
        for (int i = tid; i < P; i += kT) {
            const double v = pr[i];
            pw[i] = v;
            mx = fmax(mx, v);
        }
        peak = block_reduce<kMax>(mx, red);
        if (tid == 0) a.scal[4 * blockIdx.x + 1] = a.scal[4 * blockIdx.x + 0];
    }
    // first arg-max (numpy.argmax)
    if (tid == 0) first = (long long)P;
    __syncthreads();
    long long mine = P;
    for (int i = tid; i < P; i += kT)
        if (pw[i] == peak) { mine = i; break; }
    if (mine < P) atomicMin(&first, mine);
    __syncthreads();
    if (tid == 0) {
        a.argmax[blockIdx.x] = first < P ? first : 0;  // all-NaN rows: numpy.argmax returns 0
        a.scal[4 * blockIdx.x + 3] = peak;
    }
}

struct SpectraPool {
    std::mutex mu;
    int device = -1;
    tlsb::DeviceBuffer chi2, SR, pr, pw, scal, amax;
} g_pool;

}  // namespace

namespace tlsb {

// Device-pointer entry used by the batch pipeline as well: everything stays in HBM.
int spectra_device(const double *chi2, int64_t P, int64_t n_curves, int64_t win, double *SR, double *power_raw,
                   double *power, double *scal, long long *argmax, cudaStream_t s)
{
    SpectraArgs a{};
    a.chi2 = chi2; a.SR = SR; a.power_raw = power_raw; a.power = power; a.scal = scal; a.argmax = argmax;
    a.P = (int)P; a.win = (int)win; a.nwin = (int)(P - win + 1);
    a.detrend = P > 2 * win ? 1 : 0;
    tlsb_spectra_pre_kernel<<<(unsigned)n_curves, kT, 0, s>>>(a);
    TLSB_CUDA_TRY(cudaGetLastError());
    if (a.detrend) {
        const size_t smem = (size_t)(kMedT + win) * sizeof(double);
        if (smem > 48 * 1024)
            TLSB_CUDA_TRY(cudaFuncSetAttribute(tlsb_spectra_median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)((P + kMedT - 1) / kMedT), (unsigned)n_curves);
        tlsb_spectra_median_kernel<<<grid, kMedT, smem, s>>>(a);
        TLSB_CUDA_TRY(cudaGetLastError());
    }
    tlsb_spectra_post_kernel<<<(unsigned)n_curves, kT, 0, s>>>(a);
    TLSB_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace tlsb

extern "C" int tlsb_spectra(int32_t device, const double *chi2, int64_t n_periods, int64_t n_curves,
                            int64_t median_window, double *SR_out, double *power_raw_out, double *power_out,
                            double *SDE_raw_out, double *SDE_out, int64_t *argmax_out)
{
    if (!chi2 || !SDE_raw_out || !SDE_out) return tlsb::fail(TLSB_ERR_ARG, "tlsb_spectra: NULL argument");
    if (n_periods < 1 || n_periods > (int64_t)1 << 30 || n_curves < 1 || n_curves > 65535)
        return tlsb::fail(TLSB_ERR_ARG, "tlsb_spectra: need 1 <= n_periods <= 2^30 and 1 <= n_curves <= 65535");
    if (median_window < 1 || median_window > 24000)
        return tlsb::fail(TLSB_ERR_ARG, "tlsb_spectra: median window must be in 1..24000");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return tlsb::fail(TLSB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0) TLSB_CUDA_TRY(cudaGetDevice(&device));
    if (device >= count) return tlsb::fail(TLSB_ERR_ARG, "device ordinal out of range");
    TLSB_CUDA_TRY(cudaSetDevice(device));
    std::lock_guard<std::mutex> lock(g_pool.mu);
    if (g_pool.device != device) {  // buffers belong to one device at a time
        for (tlsb::DeviceBuffer *b : {&g_pool.chi2, &g_pool.SR, &g_pool.pr, &g_pool.pw, &g_pool.scal, &g_pool.amax}) {
            if (b->p) {
                if (g_pool.device >= 0) cudaSetDevice(g_pool.device);
                cudaFree(b->p);
                b->p = nullptr;
                b->cap = 0;
            }
        }
        TLSB_CUDA_TRY(cudaSetDevice(device));
        g_pool.device = device;
    }
    const size_t n = (size_t)n_periods * (size_t)n_curves, bytes = n * sizeof(double);
    if (g_pool.chi2.ensure(bytes) || g_pool.SR.ensure(bytes) || g_pool.pr.ensure(bytes) || g_pool.pw.ensure(bytes) ||
        g_pool.scal.ensure((size_t)n_curves * 32) || g_pool.amax.ensure((size_t)n_curves * 8))
        return tlsb::fail(TLSB_ERR_ALLOC, "device allocation failed");
    cudaStream_t s = nullptr;
    TLSB_CUDA_TRY(cudaMemcpyAsync(g_pool.chi2.p, chi2, bytes, cudaMemcpyHostToDevice, s));
    int rc = tlsb::spectra_device(g_pool.chi2.as<double>(), n_periods, n_curves, median_window, g_pool.SR.as<double>(),
                                  g_pool.pr.as<double>(), g_pool.pw.as<double>(), g_pool.scal.as<double>(),
                                  g_pool.amax.as<long long>(), s);
    if (rc) return rc;
    std::vector<double> scal((size_t)n_curves * 4);
    std::vector<long long> amax((size_t)n_curves);
    if (SR_out) TLSB_CUDA_TRY(cudaMemcpyAsync(SR_out, g_pool.SR.p, bytes, cudaMemcpyDeviceToHost, s));
    if (power_raw_out) TLSB_CUDA_TRY(cudaMemcpyAsync(power_raw_out, g_pool.pr.p, bytes, cudaMemcpyDeviceToHost, s));
    if (power_out) TLSB_CUDA_TRY(cudaMemcpyAsync(power_out, g_pool.pw.p, bytes, cudaMemcpyDeviceToHost, s));
    TLSB_CUDA_TRY(cudaMemcpyAsync(scal.data(), g_pool.scal.p, scal.size() * 8, cudaMemcpyDeviceToHost, s));
    TLSB_CUDA_TRY(cudaMemcpyAsync(amax.data(), g_pool.amax.p, amax.size() * 8, cudaMemcpyDeviceToHost, s));
    TLSB_CUDA_TRY(cudaStreamSynchronize(s));
    for (int64_t c = 0; c < n_curves; ++c) {
        SDE_raw_out[c] = scal[4 * c + 0];
        SDE_out[c] = scal[4 * c + 1];
        if (argmax_out) argmax_out[c] = (int64_t)amax[c];
    }
    return 0;
}
