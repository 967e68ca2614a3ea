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

}  // namespace tlsb

#define TLSB_CUDA_TRY(expr)                                                                          \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return tlsb::fail(TLSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

#endif  // TLSB_INTERNAL_H
