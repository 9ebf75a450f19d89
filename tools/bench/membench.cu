// Microbenchmark behind DESIGN.md's memory model: what HBM bandwidth does the access pattern
// of the compiled solver reach?  One instance per lane, every operand access of a warp is one
// contiguous 32*W-byte segment at a (pseudo-random or sequential) slot of the warp's private
// tile; U independent loads are issued back to back (one "chunk"), then consumed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int U, typename T>
__global__ void __launch_bounds__(32) chunks(const T *__restrict__ m, const unsigned *__restrict__ offs, int nchunks,
                                             size_t tile_elems, double *out) {
    const T *mb = m + (size_t)blockIdx.x * tile_elems + threadIdx.x;
    double acc = 0;
    for (int c = 0; c < nchunks; c++) {
        T v[U];
#pragma unroll
        for (int j = 0; j < U; j++) v[j] = mb[(size_t)__ldg(offs + (size_t)c * U + j) * 32];
#pragma unroll
        for (int j = 0; j < U; j++) {
            if constexpr (sizeof(T) == 8) acc += *(const double *)&v[j];
            else acc += ((const double *)&v[j])[0] + ((const double *)&v[j])[1];
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

template <int U, typename T>
float run(const void *m, const unsigned *offs, int nchunks, size_t tile_elems, int nwarps, double *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    chunks<U, T><<<nwarps, 32>>>((const T *)m, offs, 4, tile_elems, out);
    cudaEventRecord(e0);
    chunks<U, T><<<nwarps, 32>>>((const T *)m, offs, nchunks, tile_elems, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    const size_t slots = 38567;                  // quadcopter scratchbook
    const int max_warps = 148 * 32;
    const size_t tile_bytes = slots * 256;
    void *m;
    if (cudaMalloc(&m, tile_bytes * max_warps) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(m, 0, tile_bytes * max_warps);
    double *out;
    cudaMalloc(&out, 8);
    const int nchunks = 1500, UMAX = 128;
    std::vector<unsigned> rnd(nchunks * UMAX), seq(nchunks * UMAX);
    srand(1);
    for (size_t i = 0; i < rnd.size(); i++) { rnd[i] = rand() % (slots / 2); seq[i] = (unsigned)(i % (slots / 2)); }
    unsigned *d_rnd, *d_seq;
    cudaMalloc(&d_rnd, rnd.size() * 4); cudaMalloc(&d_seq, seq.size() * 4);
    cudaMemcpy(d_rnd, rnd.data(), rnd.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_seq, seq.data(), seq.size() * 4, cudaMemcpyHostToDevice);
    printf("pattern  bytes/lane  U  warps/SM   ms     TB/s\n");
    for (int pat = 0; pat < 2; pat++) {
        const unsigned *offs = pat ? d_seq : d_rnd;
        for (int wps : {1, 2, 4, 8, 16, 32}) {
            int nw = 148 * wps;
#define RUN(U, T, W)                                                                                      \
    {                                                                                                     \
        float ms = run<U, T>(m, offs, nchunks, tile_bytes / sizeof(T), nw, out);                          \
        printf("%s %3d %4d %4d %9.3f %7.3f\n", pat ? "seq" : "rnd", W, U, wps, ms,                        \
               (double)nw * nchunks * U * 32 * W / ms / 1e9);                                             \
    }
            RUN(32, double, 8) RUN(64, double, 8) RUN(96, double, 8)
            if (wps <= 8) { RUN(32, double2, 16) RUN(48, double2, 16) }
        }
    }
    return 0;
}
