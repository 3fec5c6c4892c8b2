// peaks.cu -- measured issue-rate ceilings of the SM integer pipes (bench support, not the product).
//
// MEASURED_PEAKS.json has no integer peak, so the short-string and long-Levenshtein
// kernels -- which ncu shows bound by the ALU pipe, not by HBM -- are quoted against a measured one.
// Each kernel runs `iters` rounds of 8 independent dependency chains per thread (enough ILP to hide
// the 4-cycle ALU latency at 32 resident warps per SM) and the host converts the CUDA-event time to
// warp instructions per clock per SM with the SM clock sampled during the run.
//   mode 0: LOP3 only (alu pipe)      mode 2: IMAD only (fma pipe)
//   mode 3: LOP3 + IMAD interleaved 1:1 (both pipes)
// (mode 1, plain adds, is not reported: ptxas folds two dependent adds into one IADD3)
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B1u;
    const uint32_t k1 = seed | 1u, k2 = seed ^ 0x5bd1e995u;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                } else if (MODE == 1) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(k1));
                } else if (MODE == 2) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                } else {
                    if (i & 1)
                        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                    else
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;  // keeps the chains alive
}

// runs one mode; returns milliseconds for `reps` launches and the warp instructions of ONE launch
extern "C" __attribute__((visibility("default"))) int int_peak_run(int mode, int iters, int reps, double* ms,
                                                                   double* warp_instructions, int* sm_count) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas = sms * 4;  // 4 x 256 threads = 32 warps per SM
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, (size_t)ctas * 256 * 4) != cudaSuccess) return 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto launch = [&]() {
        switch (mode) {
            case 0: int_peak_kernel<0><<<ctas, 256>>>(d, iters, 12345u); break;
            case 1: int_peak_kernel<1><<<ctas, 256>>>(d, iters, 12345u); break;
            case 2: int_peak_kernel<2><<<ctas, 256>>>(d, iters, 12345u); break;
            default: int_peak_kernel<3><<<ctas, 256>>>(d, iters, 12345u); break;
        }
    };
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) launch();
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return 3;
    float f = 0;
    cudaEventElapsedTime(&f, e0, e1);
    *ms = f;
    *warp_instructions = (double)ctas * 8 /*warps*/ * iters * 16.0 * 8.0;
    *sm_count = sms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}
