// Granularity of prefetch.global.L2 on B200: how many bytes does ONE prefetched address bring into L2?
// Pass A touches one address every X bytes of a 64 MB buffer (prefetch, or an 8-byte load with an .L2::NNNB
// hint); pass B then reads the whole buffer and is timed: L2-resident data stream at L2 speed, the rest at HBM
// speed.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pfgran pfgran.cu && ./pfgran
#include <cstdio>
#include <cuda_runtime.h>
__global__ void flush(double *p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.0; }
template <int MODE> __global__ void touch(const char *p, size_t bytes, int stride) {
    for (size_t o = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * (size_t)stride; o < bytes; o += (size_t)gridDim.x * blockDim.x * stride) {
        if (MODE == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        if (MODE == 1) { double v; asm volatile("ld.global.L2::256B.f64 %0, [%1];" : "=d"(v) : "l"(p + o)); if (v == 123.456) printf("x"); }
        if (MODE == 2) { double v; asm volatile("ld.global.L2::128B.f64 %0, [%1];" : "=d"(v) : "l"(p + o)); if (v == 123.456) printf("x"); }
        if (MODE == 3) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p + o)); if (v == 123.456) printf("x"); }
    }
}
__global__ void readall(const double2 *p, size_t n, double *out) {
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double2 v = p[i]; s += v.x + v.y; }
    if (s == 123.456) *out = s;
}
int main() {
    const size_t bytes = 64ull << 20, big = 1ull << 30;
    char *buf; double *fl, *out;
    cudaMalloc(&buf, bytes); cudaMalloc(&fl, big); cudaMalloc(&out, 8);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"prefetch.global.L2", "ld .L2::256B", "ld .L2::128B", "ld plain"};
    for (int mode = -1; mode < 4; mode++)
        for (int stride = 32; stride <= 512; stride *= 2) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                flush<<<1184, 256>>>(fl, big / 8);
                cudaDeviceSynchronize();
                if (mode == 0) touch<0><<<1184, 256>>>(buf, bytes, stride);
                if (mode == 1) touch<1><<<1184, 256>>>(buf, bytes, stride);
                if (mode == 2) touch<2><<<1184, 256>>>(buf, bytes, stride);
                if (mode == 3) touch<3><<<1184, 256>>>(buf, bytes, stride);
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                readall<<<1184, 256>>>((const double2 *)buf, bytes / 16, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("%-20s stride %4d B: read pass %.3f ms = %.0f GB/s\n", mode < 0 ? "(no touch)" : names[mode], stride, best, bytes / best / 1e6);
            if (mode < 0) break;
        }
    return 0;
}
