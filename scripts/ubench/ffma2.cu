// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 throughput on sm_100a, in the access pattern of the tap loop
// (7 accumulators per lane, one sample reused by all, template values from a register window).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float &d0, float &d1, float a0, float a1, float b0, float b1)
{
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float *out, const float *in, int iters)
{
    float B[8];
    float q[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) B[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) q[i] = in[(threadIdx.x + i) & 255];
    float s0 = in[threadIdx.x & 63], s1 = in[(threadIdx.x + 7) & 63];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int mm = 0; mm < 8; ++mm) {
            const float s = (mm & 1) ? s1 : s0;
            if (MODE == 0) {
#pragma unroll
                for (int r = 0; r < 8; ++r) B[r] = fmaf(q[(mm - r + 16) & 15], s, B[r]);
            } else {
#pragma unroll
                for (int r = 0; r < 8; r += 2) fma2(B[r], B[r + 1], q[(mm - r + 16) & 15], q[(mm - r - 1 + 16) & 15], s, s);
            }
        }
        s0 += 1e-9f; s1 -= 1e-9f;
#pragma unroll
        for (int i = 0; i < 16; ++i) q[i] += 1e-9f;
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += B[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, 148 * 2 * 256 * 4);
    cudaMalloc(&in, 1024);
    cudaMemset(in, 0, 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 256>>>(out, in, iters); else k<1><<<148 * 2, 256>>>(out, in, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * 2 * 256 * (double)iters * 64;
            printf("mode %d (%s): %.3f ms  %.1f GFMA/s  = %.1f FMA/clk/SM at 1.965 GHz\n", mode, mode ? "fma.rn.f32x2" : "FFMA", ms, fma / ms * 1e-6,
                   fma / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
