// FP64 peak of the device: dependent-free DFMA chains and separately rounded DMUL+DADD chains (what the
// solver kernels issue under -fmad=false).  Prints one JSON object; profiles/fp64_peak.json keeps the
// B200 result that bench.py reads for the FP64 bound of the roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64peak fp64peak.cu && ./fp64peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = a + threadIdx.x * 1e-9 + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) x[j] = __fma_rn(x[j], b, a);                       // 1 DFMA  = 2 flop
            else x[j] = __dadd_rn(__dmul_rn(x[j], b), a);                     // DMUL+DADD = 2 flop, 2 issues
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += x[j];
    if (s == 123.456) out[0] = s;
}

// FP64 tensor core: mma.sync m8n8k4 (DMMA), 8 independent accumulator tiles per warp; 8*8*4*2 = 512 flop per warp-wide mma
__global__ void __launch_bounds__(256) kmma(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = a + threadIdx.x * 1e-9 + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

static double run_mma(int sms) {
    double *d;
    cudaMalloc(&d, 8);
    const int iters = 1 << 13, blocks = sms * 8;
    kmma<<<blocks, 256>>>(d, 16, 1.0, 1e-3);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kmma<<<blocks, 256>>>(d, iters, 1.0, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    return 512.0 * 8 * (double)iters * 8.0 * blocks / (best * 1e-3) / 1e12;     // 8 warps per CTA
}

template <int MODE>
static double run(int sms) {
    double *d;
    cudaMalloc(&d, 8);
    const int iters = 1 << 14, blocks = sms * 8;
    k<MODE><<<blocks, 256>>>(d, 16, 1.0, 0.999999);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, iters, 1.0, 0.999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(d);
    return 2.0 * 8 * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double fma = run<0>(p.multiProcessorCount), ma = run<1>(p.multiProcessorCount), mma = run_mma(p.multiProcessorCount);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %.0f, \"dfma_tflops\": %.2f, \"dmul_dadd_tflops\": %.2f, "
           "\"dmma_m8n8k4_tflops\": %.2f, "
           "\"how\": \"8 independent chains (accumulator tiles) per thread (warp), 256 threads x 8 CTAs per SM, best of 5, CUDA events\"}\n",
           p.name, p.multiProcessorCount, clk / 1e3, fma, ma, mma);
    return 0;
}
