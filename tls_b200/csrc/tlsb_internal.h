// Internal helpers shared by the translation units of libtlsb200.so (not part of the C ABI).
#ifndef TLSB_INTERNAL_H
#define TLSB_INTERNAL_H

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace tlsb {

extern thread_local std::string g_error;  // message behind tlsb_last_error()
int fail(int code, const std::string &msg);

// Grow-only device buffer.
struct DeviceBuffer {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        cap = want;
        return 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// stats.spectra on device buffers (tlsb_spectra.cu): chi2 [n_curves][P] in ascending-period order ->
// SR, power_raw, power [n_curves][P]; scal [n_curves][4] = SDE_raw, SDE, min chi2, peak; first arg-max.
int spectra_device(const double *chi2, int64_t P, int64_t n_curves, int64_t win, double *SR, double *power_raw,
                   double *power, double *scal, long long *argmax, cudaStream_t s);


// ---- constants and argument structs shared by the kernels' translation units and the host side ----
// R = kB (template parameter of the kernels): consecutive T0 candidates one lane carries through the
// tap loop.  Odd (neighbouring lanes sit R*stride doubles apart in shared memory): 7 when all weights
// are equal (one correlation: the register window fits), 5 with unequal weights (two correlations).
constexpr int kBlockMax = 7;          // host-side slack (template padding, array slack) is sized for the largest R
#ifndef TLSB_SUB
#define TLSB_SUB 4
#endif
constexpr int kSub = TLSB_SUB;        // sub-tiles of 32 blocks a warp gates per queue reservation
__host__ __device__ constexpr int tile_size(int kb) { return 32 * kb * kSub; }  // candidates one warp gates at a time
constexpr int kGroup32 = 8;           // steps per unrolled group of the fp32 tap loop (template values arrive as two float4)
constexpr int kPadGroups = 4;         // slack (in groups of kBlock steps) behind templates and patched arrays
constexpr int kScanItems = 5;         // items per thread per scan tile (odd: conflict-free in smem)
constexpr int kResScanItems = 19;     // ... of the resident kernel: one tile covers 19 * 256 = 4864 samples (cfg-1: one tile, 3 barriers)
#ifndef TLSB_RES_HSCAN
#define TLSB_RES_HSCAN 19  // one tile for cfg-1's 4,322 histogram words, like kResScanItems (5 -> 17: 2.611 -> 2.579 ms per cfg-1 grid)
#endif
#ifndef TLSB_RES_SORT_U
#define TLSB_RES_SORT_U 3  // 2 / 3 / 4 chains: 2.092 / 2.084 / 2.103 ms per cfg-1 grid
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
    double eb;    // fp32 filter pass: |B32 - B64| <= eb * max|w d|, eb = (L + 8) * 2^-24 * sum_j |q_j| (rounded up)
    int q32;      // offset of the template in tq32 (fp32, residue-class major: see class_stride)
    int astride;  // floats per residue class in tq32 (zero padded, a multiple of 4)
};

// tq32 layout (the filter pass's templates).  Stride X splits the taps into residue classes j = X a + b.  Odd X:
// class b occupies floats [q32 + b * astride, +astride) holding q[X a + b] for a = 0, 1, ...  Even X: the PAIR of
// classes (2 c, 2 c + 1) occupies [q32 + c * 2 * astride, + 2 * astride) as interleaved float2 (q[X a + 2c],
// q[X a + 2c + 1]).  Zero padded behind the last tap; class starts are 16-byte aligned, so every lane of a warp
// fetches the template values of 8 (or 4 x 2) consecutive steps with float4 loads of the same address (broadcast).
__host__ __device__ inline int tq32_class_stride(int L, int X)
{
    return (((L + X - 1) / X + (kBlockMax - 1) + 2 * kGroup32 + 1) + 3) & ~3;
}

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
    // fp32 filter pass (equal weights): every surviving candidate gets an fp32 correlation with a rigorous error
    // bound first; only candidates whose lower bound does not exceed the best upper bound so far ("finalists") are
    // evaluated in fp64.  Results are bit-identical with the filter on or off (filter = 0: everyone is a finalist).
    const float *tq32;    // float copy of tq (same offsets)
    int filter;           // 1: filter on; 0: every candidate goes through the exact evaluation (tests)
    int fq_cap;           // capacity of the CTA-wide finalist queue
    unsigned long long *stats;  // [4] optional device counters: candidates, finalists, queue overflows (NULL: off)
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

// bytes of the small shared-memory tail (width records, reduction scratch, filter state, scheduler words), 16-aligned
__host__ __device__ inline size_t filter_tail_bytes(int nU, int threads)
{
    const int kW = threads / 32;
    // records, reduction doubles, filter state (16), sweep scheduler (32), reduction ints, period scheduler (16), tile tables
    return ((size_t)nU * sizeof(WidthRec) + (size_t)(2 * kW + 2) * 8 + 16 + 32 + (size_t)2 * kW * 4 + 16 + (size_t)nU * 12 + 15) & ~(size_t)15;
}

// what a candidate block of width record wr may read behind its start offset
__host__ __device__ inline int window_need(int W, int X, int kb) { return W + kPadGroups * kb * X + 2; }

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

// kernel launchers (one per translation unit that holds kernels)
cudaError_t launch_plan(const PlanArgs &a, int grid, cudaStream_t s);                       // tlsb_aux_kernels.cu
cudaError_t launch_prepare(const double *y, const double *dy, double *dval, double *wval, size_t n, cudaStream_t s);
cudaError_t launch_t0fit(const T0Args &a, int threads, bool resident, int grid, size_t smem, cudaStream_t s);
cudaError_t launch_gather_rows(const double *records, size_t record_stride, const int *order, double *out, int P,
                               int n_curves, cudaStream_t s);
cudaError_t launch_unshard(const long long *gathered, int n, int world, int words_per_rank, long long *out, cudaStream_t s);
// resident (folded curve in shared memory) or streaming (global scratch) layout           // tlsb_resident.cu
cudaError_t launch_search_resident(const SearchArgs &a, int threads, bool resident, bool uniform_w, int kb, int grid,
                                   size_t smem, cudaStream_t s);
cudaError_t launch_search_tiled(const SearchArgs &a, int threads, bool uniform_w, int kb, int grid, size_t smem,
                                cudaStream_t s);                                            // tlsb_tiled.cu

}  // namespace tlsb

#define TLSB_CUDA_TRY(expr)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return tlsb::fail(TLSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

#endif  // TLSB_INTERNAL_H
