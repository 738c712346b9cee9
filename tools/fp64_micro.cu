// FP64 micro-benchmarks that size the engine's kernels (not part of the product library):
//   1. DFMA dependent-chain latency (1 warp, 1 chain)
//   2. DFMA throughput per SM as a function of resident warps and independent chains per thread
//   3. FP64 division and sqrt throughput / latency
//   4. broadcast LDS.64 / LDS.128 + DFMA streams (the backward-pass inner loop shape)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_micro tools/fp64_micro.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double seed, long long* cycles) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = seed + threadIdx.x + i;
    const double m = 0.999999, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void ddiv_kernel(double* out, int iters, double seed, long long* cycles, int mode) {
    double a = seed + threadIdx.x, b = 1.0000001;
    double a2 = a + 1.0, a3 = a + 2.0, a4 = a + 3.0;
    long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; it++) a = a / b;  // dependent
    } else if (mode == 1) {
        for (int it = 0; it < iters; it++) { a = a / b; a2 = a2 / b; a3 = a3 / b; a4 = a4 / b; }
    } else {
        for (int it = 0; it < iters; it++) a = sqrt(a) + 1.5;
    }
    long long t1 = clock64();
    if (a + a2 + a3 + a4 == 12345.678) out[0] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// every lane reads the same shared-memory stream (broadcast) and FMAs it into NACC accumulators
template <int NACC, bool VEC>
__global__ void lds_fma_kernel(double* out, int iters, double seed, long long* cycles) {
    __shared__ __align__(16) double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = 0.0;
    const double s = seed + 1e-12 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const double* p = sm + ((it * 16) & 511);
        if (VEC) {
#pragma unroll
            for (int i = 0; i < NACC; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(p + i);
                acc[i] = fma(v.x, s, acc[i]);
                if (i + 1 < NACC) acc[i + 1] = fma(v.y, s, acc[i + 1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NACC; i++) acc[i] = fma(p[i], s, acc[i]);
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) r += acc[i];
    if (r == 12345.678) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

static double run_ms(void (*launch)(int, int), int grid, int block) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(grid, block);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(grid, block);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

static double* g_out;
static long long* g_cyc;
static int g_iters;

template <int ILP> void l_dfma(int g, int b) { dfma_kernel<ILP><<<g, b>>>(g_out, g_iters, 1.0, g_cyc); }
template <int N, bool V> void l_lds(int g, int b) { lds_fma_kernel<N, V><<<g, b>>>(g_out, g_iters, 1.0, g_cyc); }
static int g_mode;
void l_div(int g, int b) { ddiv_kernel<<<g, b>>>(g_out, g_iters, 3.0, g_cyc, g_mode); }

long long cycles() {
    long long c;
    cudaMemcpy(&c, g_cyc, 8, cudaMemcpyDeviceToHost);
    return c;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    cudaMalloc(&g_out, 64);
    cudaMalloc(&g_cyc, 64);
    printf("device %s, %d SMs, clock %.0f MHz\n", prop.name, sms, prop.clockRate / 1e3);
    g_iters = 1 << 14;
    // 1. latency
    run_ms(l_dfma<1>, 1, 32);
    printf("DFMA dependent chain, 1 warp: %.2f cycles/op\n", (double)cycles() / g_iters);
    run_ms(l_dfma<2>, 1, 32);
    printf("DFMA 2 chains, 1 warp: %.2f cycles/op-pair\n", (double)cycles() / g_iters);
    run_ms(l_dfma<4>, 1, 32);
    printf("DFMA 4 chains, 1 warp: %.2f cycles per 4 ops\n", (double)cycles() / g_iters);
    run_ms(l_dfma<8>, 1, 32);
    printf("DFMA 8 chains, 1 warp: %.2f cycles per 8 ops\n", (double)cycles() / g_iters);
    // 2. throughput vs warps/SM and ILP (one block per SM, `w` warps)
    printf("\nDFMA throughput (TFLOP/s, FMA=2) : rows = warps per SM, cols = chains per thread 1,2,4,8\n");
    for (int w : {1, 2, 4, 8, 12, 16, 24, 32}) {
        printf("  %2d warps/SM:", w);
        double fl;
        double ms;
        ms = run_ms(l_dfma<1>, sms, 32 * w); fl = 2.0 * 1 * g_iters * 32.0 * w * sms; printf(" %6.2f", fl / ms / 1e9);
        ms = run_ms(l_dfma<2>, sms, 32 * w); fl = 2.0 * 2 * g_iters * 32.0 * w * sms; printf(" %6.2f", fl / ms / 1e9);
        ms = run_ms(l_dfma<4>, sms, 32 * w); fl = 2.0 * 4 * g_iters * 32.0 * w * sms; printf(" %6.2f", fl / ms / 1e9);
        ms = run_ms(l_dfma<8>, sms, 32 * w); fl = 2.0 * 8 * g_iters * 32.0 * w * sms; printf(" %6.2f", fl / ms / 1e9);
        printf("\n");
    }
    // 3. division / sqrt
    g_iters = 1 << 12;
    g_mode = 0; run_ms(l_div, 1, 32);
    printf("\nDDIV dependent, 1 warp: %.1f cycles/op\n", (double)cycles() / g_iters);
    g_mode = 1; run_ms(l_div, 1, 32);
    printf("DDIV 4 independent, 1 warp: %.1f cycles per 4 ops\n", (double)cycles() / g_iters);
    g_mode = 2; run_ms(l_div, 1, 32);
    printf("DSQRT(+add) dependent, 1 warp: %.1f cycles/op\n", (double)cycles() / g_iters);
    for (int w : {4, 16, 32}) {
        g_mode = 1;
        double ms = run_ms(l_div, sms, 32 * w);
        printf("DDIV throughput %2d warps/SM: %.2f Gdiv/s\n", w, 4.0 * g_iters * 32.0 * w * sms / ms / 1e6);
    }
    // 4. broadcast LDS + DFMA
    g_iters = 1 << 13;
    printf("\nbroadcast LDS + DFMA (13 accumulators): TFLOP/s at warps/SM = 4, 8, 16, 32 ; LDS.64 then LDS.128\n");
    for (int w : {4, 8, 16, 32}) {
        double ms = run_ms(l_lds<13, false>, sms, 32 * w);
        double fl = 2.0 * 13 * g_iters * 32.0 * w * sms;
        double ms2 = run_ms(l_lds<14, true>, sms, 32 * w);
        double fl2 = 2.0 * 14 * g_iters * 32.0 * w * sms;
        printf("  %2d warps/SM: LDS.64 %6.2f   LDS.128 %6.2f\n", w, fl / ms / 1e9, fl2 / ms2 / 1e9);
    }
    run_ms(l_lds<13, false>, 1, 32);
    printf("1 warp LDS.64+DFMA x13: %.1f cycles per 13-FMA step\n", (double)cycles() / g_iters);
    run_ms(l_lds<14, true>, 1, 32);
    printf("1 warp LDS.128+DFMA x14: %.1f cycles per 14-FMA step\n", (double)cycles() / g_iters);
    return 0;
}
