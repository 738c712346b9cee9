// LDS.64 throughput in the shape of the CTA Riccati step 1: 256 threads, each loads 13 + 13 doubles from shared memory, then a 13-FMA chain
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, int active_warps) {
    __shared__ double AB[13 * 18 + 16];
    __shared__ double S[13 * 14 + 16];
    for (int e = threadIdx.x; e < 13 * 18 + 16; e += blockDim.x) AB[e] = 1.0 + e * 1e-6;
    for (int e = threadIdx.x; e < 13 * 14 + 16; e += blockDim.x) S[e] = 1.0 - e * 1e-6;
    __syncthreads();
    const int t = threadIdx.x;
    const int i = t % 13, j = (t / 13) % 13;
    volatile double* vAB = AB;
    volatile double* vS = S;
    double tot = 0.0;
    long long t0 = 0, t1 = 0;
    if ((t >> 5) < active_warps) {
        t0 = clock64();
        for (int it = 0; it < iters; it++) {
            double x[13], y[13];
            if (MODE == 0) {            // the step-1 pattern: x strided rows, y contiguous column
#pragma unroll
                for (int l = 0; l < 13; l++) { x[l] = vAB[l * 18 + i]; y[l] = vS[j * 14 + l]; }
            } else if (MODE == 1) {     // all lanes the same address (broadcast)
#pragma unroll
                for (int l = 0; l < 13; l++) { x[l] = vAB[l * 18]; y[l] = vS[l]; }
            } else if (MODE == 2) {     // x only
#pragma unroll
                for (int l = 0; l < 13; l++) { x[l] = vAB[l * 18 + i]; y[l] = 1.0; }
            } else {                    // y only
#pragma unroll
                for (int l = 0; l < 13; l++) { x[l] = 1.0; y[l] = vS[j * 14 + l]; }
            }
            double acc = tot;
#pragma unroll
            for (int l = 0; l < 13; l++) acc = fma(x[l], y[l], acc);
            tot = acc * 1e-3;
        }
        t1 = clock64();
    }
    out[t] = tot;
    if (t == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char* name) {
    double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
    for (int w : {1, 2, 4, 8}) {
        k<MODE><<<1, 256>>>(out, cyc, 1000, w);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %d warps: %.1f cycles per (26 LDS.64 + 13 DFMA) step\n", name, w, c / 1000.0);
    }
}
int main() {
    run<0>("step-1 pattern (13 rows x 3 columns)");
    run<1>("broadcast");
    run<2>("x only (13 consecutive doubles per warp)");
    run<3>("y only (3 addresses per warp)");
    return 0;
}
