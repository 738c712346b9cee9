// Register-resident FP64 FMA throughput probe: the denominator of the FP64-pipe roofline.
// 8 independent DFMA chains per thread, 256 threads per CTA, 8 CTAs per SM; no memory traffic.
#include <cuda_runtime.h>

#include "engine_host.h"

namespace tob {

__global__ void __launch_bounds__(256) dfma_probe(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

int measure_fp64_peak(int device, double* tflops) {
    if (cudaSetDevice(device) != cudaSuccess) return TO_ERR_CUDA;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double* out = nullptr;
    if (cudaMalloc(&out, 64) != cudaSuccess) return TO_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = prop.multiProcessorCount * 8, iters = 1 << 16;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        dfma_probe<<<grid, 256>>>(out, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8.0 * (double)iters * 256.0 * grid;
        double t = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return cudaGetLastError() == cudaSuccess ? 0 : TO_ERR_CUDA;
}

}  // namespace tob
