// tlsb_host.cu — the host side of libtlsb200.so: the search handle (device buffers, light curves, template bank,
// period grid), the choice of the on-chip layout, the plan/search/repair sequencing and the C ABI of
// include/tlsb200.h.  The kernels live in tlsb_resident.cu, tlsb_tiled.cu, tlsb_aux_kernels.cu and
// tlsb_spectra.cu; this file launches them through the tlsb::launch_* functions of tlsb_internal.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
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

namespace tlsb {
thread_local std::string g_error;

int fail(int code, const std::string &msg)
{
    g_error = msg;
    return code;
}
}  // namespace tlsb

namespace {
using namespace tlsb;
using tlsb::fail;
using tlsb::g_error;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(TLSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

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

// Grow-only pinned host buffer (upload sources / download targets that must not be staged by the driver).
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// TLSB_MEMO=0 switches off what the handle remembers between calls with identical inputs (derived template arrays,
// device plan): every call then redoes all of its work (bench.py's end-to-end figure is measured that way).
bool memo_enabled()
{
    const char *e = std::getenv("TLSB_MEMO");
    return !(e && e[0] == '0');
}

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
    int fq_cap = 0;        // finalist queue of the fp32 filter pass (0: the layout has no filter)
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
    cudaStream_t up_stream = nullptr;  // stream the setters upload on (tlsb_set_inputs_async: the search's stream)
    bool out_is_current = false;   // the last search wrote the handle's own record buffer (tlsb_get_results reads it)
    bool defer_sync = false;       // one-shot call: the caller's buffers outlive the whole call, setters need not wait
    PinBuf h_tq;                   // host copy of tq, pinned (keeps the upload source alive without a synchronisation)
    size_t n_tq = 0, n_tq32 = 0;   // elements of h_tq / h_tq32
    // templates
    tlsb_params prm{};
    int nU = 0, M = 0, pad = 0;
    std::vector<WidthRec> recs;   // unique widths, ascending
    DevBuf tq, tq32, d_rec, filter_stats;
    PinBuf h_tq32;                // float copy of h_tq for the fp32 filter pass (residue-class major), pinned
    // what the bank was built from (tlsb_set_templates): an identical bank is re-uploaded, not re-derived
    std::vector<double> in_signal, in_overshoot;
    std::vector<int64_t> in_meta;  // offset | length | width
    // results of the one-shot call come back through ONE copy into pinned memory
    PinBuf pin_out;
    bool tq_in_flight = false;     // an asynchronous upload from h_tq / h_tq32 may not have finished (on tq_stream)
    cudaStream_t tq_stream = nullptr;
    // the device plan is a pure function of (periods, bank, N, span, stellar limits): remembered across searches once
    // its status word has been read back clean
    int64_t bank_ver = 0, periods_ver = 0;
    struct PlanKey {
        int N = -1, kb = 0;
        double span = 0, rs_min = 0, rs_max = 0, ms_min = 0, ms_max = 0;
        int64_t bank_ver = -1, periods_ver = -1;
        bool operator==(const PlanKey &o) const
        {
            return N == o.N && kb == o.kb && span == o.span && rs_min == o.rs_min && rs_max == o.rs_max && ms_min == o.ms_min &&
                   ms_max == o.ms_max && bank_ver == o.bank_ver && periods_ver == o.periods_ver;
        }
    };
    PlanKey plan_key_dev;          // key of the plan that is on the device now (ulo / uhi / order)
    bool plan_key_launched = false;  // ... and it was produced by the plan kernel in the most recent search
    bool plan_key_clean = false;     // ... and its status word has been seen clean (or repaired on the device)
    int filter_mode = 1;          // 1: fp32 filter pass on (equal weights); 0: every candidate through the exact evaluation
    bool want_stats = false;      // count candidates / finalists on the device (tlsb_last_filter_stats)
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

size_t tail_bytes(int nU, int threads) { return filter_tail_bytes(nU, threads); }

// resident layout with the fp32 filter pass (equal weights): cs | sorted ids (u16) | tail | area, the area being the
// larger of the sort's scratch (keys, histogram with N buckets, bucket-ordered ids) and the search's arrays (w*d in
// fp32, survivor queue, finalist queue); mirrors the carve in tlsb_search_kernel
size_t resident_filter_smem_bytes(int N, int M, int pad, int nU, int qcap, int fq_cap, int threads, int NB, bool uniform = true)
{
    const size_t NM = (size_t)N + M, NMP = NM + pad;
    // X: the sorted d / fp64 cumulative sums while they are built, the survivor queue afterwards (unequal weights: the
    // finalist queue too, because the area then holds a third fp32 array, w32)
    const size_t x = align16(std::max(((NM + 2) & ~(size_t)1) * 8, (size_t)qcap * 8 + (uniform ? 0 : (size_t)fq_cap * 12)));
    const size_t sort_area = (size_t)N * 8 + (((size_t)NB + 2) / 2) * 4 + (size_t)N * 2;  // 16-bit bucket counters
    const size_t nmp4 = ((NMP + 3) & ~(size_t)3) * 4;
    const size_t search_area = ((NM + 2 + 3) & ~(size_t)3) * 4 + (uniform ? nmp4 + (size_t)fq_cap * 12 : 2 * nmp4);
    return x + align16((size_t)N * 2) + tail_bytes(nU, threads) + align16(std::max(sort_area, search_area));
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
    // block size R: 7 candidates per lane in the filter layouts (one or two fp32 correlations), 5 in the all-fp64 kernels
    // of unequal weights (two fp64 correlations: registers)
    const char *kbe = std::getenv("TLSB_BLOCK");   // experiments: 5 forces R = 5 with equal weights
    const char *wke = std::getenv("TLSB_WBLOCK");  // experiments: 5 forces R = 5 in the filter layouts of unequal weights
    const int kb_w = (wke && std::atoi(wke) == 5) ? 5 : 7;
    const char *wf = std::getenv("TLSB_WFILTER");  // "0": unequal weights keep the all-fp64 kernels (experiments / A-B runs)
    const bool wfilter = !h->uniform_w && !(wf && std::atoi(wf) == 0);  // unequal weights with the fp32 gate + filter pass
    const int kb_pref = h->uniform_w ? ((kbe && std::atoi(kbe) == 5) ? 5 : 7) : 5;
    best.kb = kb_pref;
    if (N < 65536 && h->path_mode <= 1 && h->uniform_w && kb_pref == 7) {
        // equal weights: fp32 filter pass; the folded curve costs 8 (cs) + 4 (w*d in fp32) + 2 (ids) bytes per sample
        // threads, CTAs per SM, survivor queue, finalist queue, phase buckets of the sort as a divisor of N
        // (512 threads: the survivor queue is a ring, a power of two, at least twice what all warps can append at once)
        // (two CTAs per SM: the queue takes over the memory of the fp64 cumulative sums once they have been copied out,
        // so it is at least N + M + 2 entries for free)
        const int tries[7][5] = {{256, 2, 3584, 1024, 1}, {256, 2, 3072, 1024, 1}, {256, 2, 3072, 1024, 2}, {256, 2, 3072, 512, 3},
                                 {256, 2, 2560, 512, 4}, {512, 1, 8192, 2048, 1}, {512, 1, 4096, 1024, 1}};
        const int cs_elems = (N + h->M + 2) & ~1;
        for (const auto &t : tries) {
            const int NB = std::max(64, 2 * (N / t[4]));  // 16-bit counters: two buckets per sample in the space of one
            const int qcap = t[0] == 256 ? std::min(16384, std::max(t[2], cs_elems)) : t[2];
            const size_t bytes = resident_filter_smem_bytes(N, h->M, h->pad, h->nU, qcap, t[3], t[0], NB);
            if (bytes > h->max_smem) continue;
            if ((bytes + 1024) * (size_t)t[1] > h->smem_per_sm) continue;
            best.resident = true;
            best.threads = t[0];
            best.ctas_per_sm = t[1];
            best.qcap = qcap;
            best.fq_cap = t[3];
            best.NB = NB;
            best.smem = bytes;
            best.scratch_per_cta = ((size_t)cs_elems * 8 + 255) & ~(size_t)255;  // the fp64 cumulative sums (exact evaluations)
            return best;
        }
    }
    if (N < 65536 && h->path_mode <= 1 && wfilter) {
        // unequal weights: fp32 gate + two fp32 correlations; cs32, wd32, w32 and the sorted ids stay on chip
        const int tries[5][5] = {{256, 2, 3584, 832, 1}, {256, 2, 3072, 512, 1}, {256, 2, 3072, 512, 2}, {512, 1, 8192, 2048, 1},
                                 {512, 1, 4096, 1024, 1}};
        const int cs_elems = (N + h->M + 2) & ~1;
        for (const auto &t : tries) {
            const int NB = std::max(64, 2 * (N / t[4]));
            const size_t bytes = resident_filter_smem_bytes(N, h->M, h->pad, h->nU, t[2], t[3], t[0], NB, false);
            if (bytes > h->max_smem) continue;
            if ((bytes + 1024) * (size_t)t[1] > h->smem_per_sm) continue;
            best.kb = kb_w;  // R = 5 -> 7: cfg-1 at 500 ppm with dy 12.3 -> 11.0 ms (no spills at 128 registers)
            best.resident = true;
            best.threads = t[0];
            best.ctas_per_sm = t[1];
            best.qcap = t[2];
            best.fq_cap = t[3];
            best.NB = NB;
            best.smem = bytes;
            best.scratch_per_cta = ((size_t)cs_elems * 8 + 255) & ~(size_t)255;
            return best;
        }
    }
    if (N < 65536 && h->path_mode <= 1) {
        const int tries[2][2] = {{256, 2}, {512, 1}};
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
    // bytes a folded sample takes in a staged chunk: detrended cs in fp32 (4) + w*d in fp32 (4) with equal weights
    // (fp32 gate + filter pass), cs + w + w*d in fp64 otherwise
    const bool filt = h->uniform_w || wfilter;
    const int kb_t = wfilter ? kb_w : kb_pref;  // tiled layouts (R = 5 -> 7 with unequal weights: cfg-3 with dy 54.6 -> 45.5 ms)
    const size_t elem = h->uniform_w ? 8 : wfilter ? 12 : 24;
    const size_t nmp4 = (NMP + 3) & ~(size_t)3;
    const size_t cs4 = (NM + 2 + 3) & ~(size_t)3;
    const int fq_cap = filt ? 1024 : 0;
    int need_max = 0, need5 = 0;
    for (const WidthRec &wr : h->recs) {
        need_max = std::max(need_max, window_need(wr.W, wr.X, kb_t));
        need5 = std::max(need5, window_need(wr.W, wr.X, 5));
    }
    const char *force = std::getenv("TLSB_TILED");  // "0": never, "256"/"512": force that CTA size (experiments)
    const int forced = force ? std::atoi(force) : -1;
    if (forced != 0 && h->path_mode != 3) {
        // threads, CTAs per SM, queue entries (equal weights: a ring, power of two)
        const int tries[2][3] = {{256, 2, filt ? 2048 : 3072}, {512, 1, 4096}};
        for (const auto &t : tries) {
            if (forced > 0 && forced != t[0]) continue;
            size_t per_cta = std::min(h->max_smem, h->smem_per_sm / (size_t)t[1] - 1024);
            if (const char *cap = std::getenv("TLSB_SMEM_KB"))  // experiments: leave part of the SM's 256 KB to L1
                per_cta = std::min(per_cta, (size_t)std::atoi(cap) * 1024 / (size_t)t[1]);
            const size_t fixed = (size_t)t[2] * 8 + (size_t)fq_cap * 12 + tail_bytes(h->nU, t[0]) + (size_t)h->nU * 12 + 128 +
                                 (size_t)(kMaxSegments + 2) * 4;
            if (per_cta <= fixed) continue;
            long long C = (long long)(((per_cta - fixed) / elem) & ~(size_t)3);
            const bool exact_cap = h->chunk_cap < 0;  // tests: cap the chunk exactly; widths that do not fit take the L2 pass
            if (h->chunk_cap > 0) C = std::min<long long>(C, std::max<long long>(h->chunk_cap, need_max + 64) & ~3LL);
            if (exact_cap) C = std::min<long long>(C, (long long)(-h->chunk_cap) & ~3LL);
            int kb = kb_t, n_tiled = h->nU;
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
            } else if (h->chunk_cap == 0 && t[1] == 2 && forced < 0 && (TP < 2048 || 5 * TP < 4 * C)) {
                // two CTAs per SM only when nearly all of a chunk is start offsets (halo below 20 %); one big CTA with the
                // barrier-free sweep otherwise (cfg-3, halo 34 %: 22.0 ms with two CTAs, 21.2 ms with one)
                continue;
            }
            best.resident = false;
            best.tiled = true;
            best.kb = kb;
            best.threads = t[0];
            best.ctas_per_sm = t[1];
            best.qcap = t[2];
            best.chunk = (int)C;
            best.n_tiled = n_tiled;
            best.fq_cap = fq_cap;
            best.NB = (int)std::min<long long>(N, (long long)(elem * (size_t)C / 4) - 2);
            best.smem = (size_t)t[2] * 8 + (size_t)fq_cap * 12 + elem * (size_t)C + tail_bytes(h->nU, t[0]) + (size_t)h->nU * 12 + 128 +
                        (size_t)(kMaxSegments + 2) * 4;
            best.scratch_per_cta = cs + (size_t)(narr - 1) * nmp_even * 8 +
                                   (h->uniform_w ? (cs4 + nmp4) * 4 : wfilter ? (cs4 + 2 * nmp4) * 4 : 0) + align16((size_t)N * 4);
            const size_t list_base = best.scratch_per_cta;
            // on-chip sort: segments of S keys sorted in the chunk area; 1.5x head room over N / n_seg
            const char *oc = std::getenv("TLSB_ONCHIP_SORT");  // "0" disables (experiments)
            const size_t area = elem * (size_t)C;
            long long S = (long long)(((area - 64) / (h->uniform_w ? 24 : 32)) & ~(size_t)1);
            S = std::min<long long>(S, (long long)kSegPerThread * t[0]);
            const long long ns = S > 0 ? (3LL * N + 2 * S - 1) / (2 * S) : 0;
            if (!(oc && std::atoi(oc) == 0) && S >= 64 && S <= 65534 && ns >= 1 && ns <= kMaxSegments) {
                best.seg_cap = (int)S;
                best.n_seg = (int)ns;
                best.scratch_per_cta += (size_t)ns * (size_t)S * 4 + 16;  // lists of sample ids
            }
            best.scratch_per_cta = std::max(best.scratch_per_cta, list_base + align16((size_t)N * 4) + 16);  // sorted ids of a fallback sort
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

tlsb_handle::PlanKey current_plan_key(const tlsb_handle *h, int kb)
{
    tlsb_handle::PlanKey k;
    k.N = h->N; k.kb = kb; k.span = h->span;
    k.rs_min = h->prm.R_star_min; k.rs_max = h->prm.R_star_max; k.ms_min = h->prm.M_star_min; k.ms_max = h->prm.M_star_max;
    k.bank_ver = h->bank_ver; k.periods_ver = h->periods_ver;
    return k;
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
    } else if (h->plan_mode == 0 && ((h->dev_plan_valid && h->dev_plan_span == h->span) ||
                                     (h->plan_key_clean && memo_enabled() && h->plan_key_dev == current_plan_key(h, lay.kb)))) {
        CUDA_TRY(cudaMemsetAsync(status, 0, 8, s));  // same periods, bank, N and span as the plan on the device
        h->plan_key_launched = false;
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
        CUDA_TRY(launch_plan(pa, plan_grid, s));
        h->host_plan_valid = false;
        h->launches += 1;
        h->dev_plan_valid = false;  // becomes valid only once its status word has been seen clean (batch)
        h->dev_plan_span = h->span;
        h->plan_key_dev = current_plan_key(h, lay.kb);
        h->plan_key_clean = false;
        h->plan_key_launched = h->plan_mode == 0;
    }
    if (exact_plan) h->plan_key_clean = h->plan_key_launched = false;  // the host plan overwrote ulo / uhi / order

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
    a.tq32 = h->tq32.as<float>();
    a.filter = h->filter_mode;
    a.fq_cap = lay.fq_cap;
    a.stats = nullptr;
    if (h->want_stats && lay.fq_cap > 0) {
        if (h->filter_stats.ensure(32)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
        if (!only) CUDA_TRY(cudaMemsetAsync(h->filter_stats.p, 0, 32, s));
        a.stats = h->filter_stats.as<unsigned long long>();
    }
    const int grid = std::min(only ? n_only : P, h->num_sms * lay.ctas_per_sm);
    if (!lay.resident || lay.scratch_per_cta > 0) {
        if (lay.NB < 1) return fail(TLSB_ERR_ARG, "too many distinct template widths for shared memory");
        a.scratch_per_cta = lay.scratch_per_cta;
        if (h->scratch.ensure(a.scratch_per_cta * (size_t)grid)) return fail(TLSB_ERR_ALLOC, "device allocation failed (scratch)");
        a.scratch = h->scratch.as<unsigned char>();
    }
    CUDA_TRY(cudaMemsetAsync(h->counter.as<int>() + 4, 0, 4, s));  // periods whose on-chip sort overflowed
    CUDA_TRY(cudaEventRecord(h->ev0, s));
    const bool uni = h->uniform_w;
    cudaError_t le = cudaSuccess;
    if (lay.resident || !lay.tiled)
        le = launch_search_resident(a, lay.threads, lay.resident, uni, lay.kb, grid, lay.smem, s);
    else
        le = launch_search_tiled(a, lay.threads, uni, lay.kb, grid, lay.smem, s);
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
const char *tlsb_version(void) { return "tlsb200 0.3 (sm_100a)"; }

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
    if (const char *fe = std::getenv("TLSB_FILTER")) h->filter_mode = std::atoi(fe) != 0;  // experiments / tests
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
                      &h->uhi, &h->order, &h->bin_of, &h->out, &h->tq32, &h->filter_stats, &h->counter, &h->scratch, &h->t0_trials, &h->t0_model,
                      &h->t0_resid, &h->plan_bins, &h->unsure, &h->asc_order, &h->brec, &h->bchi, &h->bSR, &h->bpr, &h->bpw, &h->bscal, &h->bamax})
        b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    h->pin_out.release();
    h->h_tq.release();
    h->h_tq32.release();
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
    if ((rc = upload(h->t, t, shared_t ? sizeof(double) * (size_t)n : bytes, h->up_stream))) return rc;
    if ((rc = upload(h->y, y, bytes, h->up_stream))) return rc;
    if ((rc = upload(h->dy, dy, bytes, h->up_stream))) return rc;
    if (h->dval.ensure(bytes) || h->wval.ensure(bytes)) return fail(TLSB_ERR_ALLOC, "device allocation failed");
    CUDA_TRY(launch_prepare(h->y.as<double>(), h->dy.as<double>(), h->dval.as<double>(), h->wval.as<double>(), total, h->up_stream));
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
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(h->up_stream));
    h->n_curves = (int)n_curves;
    h->shared_t = shared_t;
    h->cur = 0;
    h->uniform_w = h->c_uniform[0] != 0;
    h->w0 = h->c_w0[0];
    if (!h->have_lc || h->N != n) h->recs_stale = true;  // candidates and tiles per width depend on N
    h->N = n;
    h->span = h->c_span[0];
    h->have_lc = true;
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
    {
        // The same bank and parameters as the handle already holds (the one-shot call re-sends them with every search):
        // the derived arrays (unique widths, q = (1 - signal) / SIGNAL_DEPTH, its fp32 residue-class-major copy, error
        // bounds) are still right; only the uploads are repeated.
        size_t total = 0;
        for (int r = 0; r < R; ++r) {
            if (tp->offset[r] < 0 || tp->length[r] < 0) return fail(TLSB_ERR_ARG, "tlsb_set_templates: negative offset or length");
            total = std::max(total, (size_t)(tp->offset[r] + tp->length[r]));
        }
        const bool same = h->have_tp && memo_enabled() && h->in_meta.size() == (size_t)3 * R && h->in_signal.size() == total &&
                          std::memcmp(&h->prm, prm, sizeof(tlsb_params)) == 0 &&
                          std::memcmp(h->in_meta.data(), tp->offset, sizeof(int64_t) * R) == 0 &&
                          std::memcmp(h->in_meta.data() + R, tp->length, sizeof(int64_t) * R) == 0 &&
                          std::memcmp(h->in_meta.data() + 2 * R, tp->width, sizeof(int64_t) * R) == 0 &&
                          std::memcmp(h->in_overshoot.data(), tp->overshoot, sizeof(double) * R) == 0 &&
                          std::memcmp(h->in_signal.data(), tp->signal, sizeof(double) * total) == 0;
        if (same) {
            int rc0;
            if ((rc0 = upload(h->tq, h->h_tq.p, h->n_tq * 8, h->up_stream))) return rc0;
            if ((rc0 = upload(h->tq32, h->h_tq32.p, h->n_tq32 * 4, h->up_stream))) return rc0;
            if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(h->up_stream));
            h->tq_in_flight = h->defer_sync;
            h->tq_stream = h->up_stream;
            return 0;
        }
        if (memo_enabled()) {
            h->in_meta.resize((size_t)3 * R);
            std::memcpy(h->in_meta.data(), tp->offset, sizeof(int64_t) * R);
            std::memcpy(h->in_meta.data() + R, tp->length, sizeof(int64_t) * R);
            std::memcpy(h->in_meta.data() + 2 * R, tp->width, sizeof(int64_t) * R);
            h->in_overshoot.assign(tp->overshoot, tp->overshoot + R);
            h->in_signal.assign(tp->signal, tp->signal + total);
        } else {
            h->in_meta.clear();
        }
        h->have_tp = false;  // until the new bank is complete
    }
    // unique widths ascending, first row with each width (core.py:113, :163-165)
    std::vector<int64_t> uniq(tp->width, tp->width + R);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    const int nU = (int)uniq.size();
    if (nU > 65535) return fail(TLSB_ERR_ARG, "tlsb_set_templates: more than 65535 distinct widths");
    std::vector<WidthRec> recs(nU);
    // pass 1: geometry of every unique width and the offsets of its templates in tq (fp64) and tq32 (fp32, class major)
    int xmax = 1;
    size_t n_tq = 0, n_tq32 = 0;
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
        n_tq = (n_tq + 3) & ~(size_t)3;  // 16-byte aligned template starts
        wr.q = (int)n_tq;
        n_tq += (size_t)L + (size_t)xth * kPadGroups * kBlockMax;  // + ramp-out and pipeline overshoot (zeros)
        wr.astride = tq32_class_stride((int)L, xth);
        wr.q32 = (int)n_tq32;
        n_tq32 += (size_t)xth * (size_t)wr.astride;  // X classes (odd X) or X/2 interleaved pairs (even X) of astride floats
    }
    int M = recs[nU - 1].W;  // core.py:114-116
    if (M % 2 != 0) M += 1;
    int rc;
    if (h->tq_in_flight) {  // an earlier asynchronous upload may still read h_tq / h_tq32
        CUDA_TRY(cudaStreamSynchronize(h->tq_stream));
        h->tq_in_flight = false;
    }
    if (h->h_tq.ensure(n_tq * 8 + 8) || h->h_tq32.ensure(n_tq32 * 4 + 4)) return fail(TLSB_ERR_ALLOC, "pinned host allocation failed");
    double *tq = h->h_tq.as<double>();
    float *tq32 = h->h_tq32.as<float>();
    std::memset(tq, 0, n_tq * 8);
    std::memset(tq32, 0, n_tq32 * 4);
    // pass 2: q_j = (1 - signal_j) / SIGNAL_DEPTH (core.py:61-68), its sums, and the fp32 residue-class-major copy of the
    // filter pass (tlsb_internal.h: tq32_class_stride)
    for (int u = 0; u < nU; ++u) {
        WidthRec &wr = recs[u];
        const int L = wr.L, X = wr.X, A = wr.astride;
        const double *sg = tp->signal + tp->offset[wr.row];
        double *q = tq + wr.q;
        double sq2 = 0.0, qabs = 0.0;
        for (int j = 0; j < L; ++j) {
            const double v = (1 - sg[j]) / kSignalDepth;
            q[j] = v;
            sq2 = std::fma(v, v, sq2);
            qabs += std::fabs(v);
        }
        wr.sq2 = sq2;
        wr.eb = (double)(L + 8) * 5.9604644775390625e-08 * qabs * (1.0 + 1e-6);  // tlsb_device.cuh: the filter's bound
        float *dst = tq32 + wr.q32;
        const int V = (X & 1) ? 1 : 2;
        if (X == 1) {
            for (int j = 0; j < L; ++j) dst[j] = (float)q[j];
        } else {
            for (int j = 0; j < L; ++j) {  // tap j = X a + V c + v lives at ((c A + a) V + v)
                const int a = j / X, b = j - a * X, c = b / V, v = b - c * V;
                dst[((size_t)c * A + a) * V + v] = (float)q[j];
            }
        }
    }
    h->n_tq = n_tq;
    h->n_tq32 = n_tq32;
    if ((rc = upload(h->tq, tq, n_tq * 8, h->up_stream))) return rc;
    if ((rc = upload(h->tq32, tq32, n_tq32 * 4, h->up_stream))) return rc;
    h->tq_in_flight = h->defer_sync;
    h->tq_stream = h->up_stream;
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(h->up_stream));
    h->recs.swap(recs);
    h->pad = kPadGroups * kBlockMax * xmax;
    h->nU = nU;
    h->M = M;
    h->prm = *prm;
    h->have_tp = true;
    h->bank_ver += 1;
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
    const bool same = h->have_periods && (size_t)n_periods == h->h_periods.size() &&
                      (n_periods == 0 || std::memcmp(periods, h->h_periods.data(), sizeof(double) * (size_t)n_periods) == 0);
    h->P = (int)n_periods;
    if (!same) h->h_periods.assign(periods, periods + n_periods);
    int rc;
    if ((rc = upload(h->periods, periods, sizeof(double) * (size_t)n_periods, h->up_stream))) return rc;
    if (!h->defer_sync) CUDA_TRY(cudaStreamSynchronize(h->up_stream));
    h->have_periods = true;
    if (same) return 0;  // the same grid again: orders and plans that depend on it stay valid
    h->periods_ver += 1;
    h->asc_valid = false;
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

int tlsb_set_inputs_async(tlsb_handle *h, void *cuda_stream, const tlsb_lightcurve *lc, const tlsb_templates *tp,
                          const tlsb_params *prm, const double *periods, int64_t n_periods)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_set_inputs_async: NULL handle");
    h->up_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    h->defer_sync = true;
    int rc = 0;
    if (lc) rc = tlsb_set_lightcurve(h, lc);
    if (!rc && tp) rc = prm ? tlsb_set_templates(h, tp, prm) : fail(TLSB_ERR_ARG, "tlsb_set_inputs_async: templates without params");
    if (!rc && periods) rc = tlsb_set_periods(h, periods, n_periods);
    h->defer_sync = false;
    h->up_stream = nullptr;
    return rc;
}

int tlsb_unshard_records(const void *gathered_dev, int64_t n_periods, int32_t world, void *out_dev, void *cuda_stream)
{
    if (!gathered_dev || !out_dev) return fail(TLSB_ERR_ARG, "tlsb_unshard_records: NULL argument");
    if (n_periods < 1 || n_periods > (int64_t)1 << 30 || world < 1 || world > 65536)
        return fail(TLSB_ERR_ARG, "tlsb_unshard_records: bad period count or world size");
    const int64_t cap = (n_periods + world - 1) / world;
    CUDA_TRY(launch_unshard(reinterpret_cast<const long long *>(gathered_dev), (int)n_periods, (int)world, (int)(3 * cap + 1),
                            reinterpret_cast<long long *>(out_dev), reinterpret_cast<cudaStream_t>(cuda_stream)));
    return 0;
}

int tlsb_set_filter(tlsb_handle *h, int32_t mode, int32_t count_stats)
{
    if (!h || mode < 0 || mode > 1) return fail(TLSB_ERR_ARG, "tlsb_set_filter: mode must be 0 or 1");
    h->filter_mode = mode;
    h->want_stats = count_stats != 0;
    return 0;
}

int tlsb_last_filter_stats(tlsb_handle *h, int64_t *candidates, int64_t *finalists, int64_t *overflows)
{
    if (!h) return fail(TLSB_ERR_ARG, "tlsb_last_filter_stats: NULL handle");
    unsigned long long v[4] = {0, 0, 0, 0};
    if (h->filter_stats.p && h->want_stats) {
        CUDA_TRY(cudaSetDevice(h->device));
        CUDA_TRY(cudaMemcpy(v, h->filter_stats.p, 32, cudaMemcpyDeviceToHost));  // synchronises
    }
    if (candidates) *candidates = (int64_t)v[0];
    if (finalists) *finalists = (int64_t)(v[1] + v[2]);
    if (overflows) *overflows = (int64_t)v[2];
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
    h->out_is_current = !records_dev;
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
    if (!h->out.p || !h->out_is_current)
        return fail(TLSB_ERR_STATE, "tlsb_get_results: the most recent search did not write the handle's own buffer "
                                    "(it was given records_dev; read that buffer instead)");
    // one device -> host copy of the three planes + status word into pinned staging memory (the caller's arrays may
    // be pageable: three staged copies cost more than one pinned copy and a host-side unpack)
    const size_t words = 3 * P + 1;
    if (h->pin_out.ensure(words * 8)) return fail(TLSB_ERR_ALLOC, "pinned host allocation failed");
    const double *stage = h->pin_out.as<double>();
    const long long *packed = reinterpret_cast<const long long *>(stage + 2 * P);
    for (int attempt = 0; attempt < 2; ++attempt) {
        CUDA_TRY(cudaMemcpyAsync(h->pin_out.p, h->out.p, words * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (s == h->tq_stream) h->tq_in_flight = false;
        if (packed[P] == 0 && attempt == 0 && h->plan_key_launched) h->plan_key_clean = true;  // the device plan stands
        if (packed[P] == 0 || attempt == 1) break;
        // the device plan was not sure about some periods' limits: settle those on the host and search
        // again only the ones whose admissible widths really differ
        const int64_t before = h->repairs + h->fallbacks;
        int rc = resolve_records(h, s, h->out.p, packed[P]);
        if (rc) return rc;
        if (h->repairs + h->fallbacks == before) break;  // every flagged limit was right: results stand
    }
    std::memcpy(chi2_out, stage, P * 8);
    std::memcpy(depth_out, stage + P, P * 8);
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
    // TLSB_TRACE=1: host microseconds per stage of the one-shot call on stderr (experiments; nothing is synchronised)
    static const bool trace = std::getenv("TLSB_TRACE") != nullptr;
    using clk = std::chrono::steady_clock;
    clk::time_point tp0 = clk::now();
    double us[6] = {0, 0, 0, 0, 0, 0};
    int stage = 0;
    auto mark = [&]() {
        if (!trace) return;
        const clk::time_point now = clk::now();
        us[stage++] = std::chrono::duration<double, std::micro>(now - tp0).count();
        tp0 = now;
    };
    tlsb_handle *h = pool_take(device);
    int rc = h ? 0 : tlsb_create(&h, device);
    mark();
    if (!rc) h->defer_sync = true;  // every input buffer outlives this call: one synchronisation, at the end
    if (!rc) rc = tlsb_set_lightcurve(h, lc);
    mark();
    if (!rc) rc = tlsb_set_templates(h, tp, prm);
    mark();
    if (!rc) rc = tlsb_set_periods(h, periods, nP);
    mark();
    if (!rc) rc = tlsb_search_async(h, nullptr, nullptr);
    mark();
    if (!rc) rc = tlsb_get_results(h, nullptr, chi2, row, depth, t0);
    mark();
    if (trace)
        std::fprintf(stderr, "tlsb trace (host us): handle %.1f  lightcurve %.1f  templates %.1f  periods %.1f  enqueue %.1f  results+wait %.1f\n",
                     us[0], us[1], us[2], us[3], us[4], us[5]);
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
    CUDA_TRY(launch_t0fit(a, threads, resident, grid, smem, s));
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
    CUDA_TRY(launch_gather_rows(rec, stride, h->asc_order.as<int>(), h->bchi.as<double>(), (int)P, (int)B, s));
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
