// Microbenchmark 2: per-warp load throughput vs number of independent loads per chunk (U), load
// width and warps per SM, with addresses computed arithmetically (no dependent offset loads).
#include <cstdio>
#include <cuda_runtime.h>

template <int U, typename T>
__global__ void __launch_bounds__(32) chunks(const T *__restrict__ m, int nchunks, unsigned slots, size_t tile_elems,
                                             double *out) {
    const T *mb = m + (size_t)blockIdx.x * tile_elems + threadIdx.x;
    double acc = 0;
    unsigned s = blockIdx.x * 7919u;
    for (int c = 0; c < nchunks; c++) {
        T v[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            s = (s + 2654435761u) ;
            v[j] = mb[(size_t)((s >> 8) % slots) * 32];
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            if constexpr (sizeof(T) == 8) acc += *(const double *)&v[j];
            else acc += ((const double *)&v[j])[0] + ((const double *)&v[j])[1];
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int U, typename T>
void run(const void *m, int nchunks, unsigned slots, size_t tile_bytes, int wps, double *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int nw = 148 * wps;
    chunks<U, T><<<nw, 32>>>((const T *)m, 4, slots, tile_bytes / sizeof(T), out);
    cudaEventRecord(e0);
    chunks<U, T><<<nw, 32>>>((const T *)m, nchunks, slots, tile_bytes / sizeof(T), out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%3d %4d %4d %9.3f ms %8.1f ns/chunk %7.1f ns/load %7.3f TB/s\n", (int)sizeof(T), U, wps, ms,
           1e6 * ms / nchunks, 1e6 * ms / nchunks / U, (double)nw * nchunks * U * 32 * sizeof(T) / ms / 1e9);
}

int main() {
    const unsigned slots = 38567 / 2;
    const size_t tile_bytes = (size_t)38567 * 256;
    void *m;
    if (cudaMalloc(&m, tile_bytes * 148 * 16) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(m, 0, tile_bytes * 148 * 16);
    double *out;
    cudaMalloc(&out, 8);
    printf("bytes/lane U warps/SM\n");
    for (int wps : {1, 4, 8, 16}) {
        run<4, double>(m, 2000, slots, tile_bytes, wps, out);
        run<8, double>(m, 2000, slots, tile_bytes, wps, out);
        run<16, double>(m, 2000, slots, tile_bytes, wps, out);
        run<32, double>(m, 2000, slots, tile_bytes, wps, out);
        run<64, double>(m, 1000, slots, tile_bytes, wps, out);
        run<96, double>(m, 1000, slots, tile_bytes, wps, out);
        run<16, double2>(m, 2000, slots, tile_bytes, wps, out);
        run<48, double2>(m, 1000, slots, tile_bytes, wps, out);
    }
    return 0;
}
