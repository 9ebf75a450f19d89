// Dense path of the batched interior-point backend: box-constrained dense least squares
//
//     minimise ||A x - b||^2 / N   subject to   lo <= x <= hi,        A: N x n dense per instance
//
// (BASELINE.json config 5: N=4000, n=256, batch 4096; the same problem class as examples/sls.m:136-166).  The
// reference generates scalar straight-line code for every entry of the KKT factorisation, which does not scale
// to a dense 256x256 Schur block (its instruction table is capped, instructionsTableUTHash.c:171), so this
// path keeps the reference's ALGORITHM -- ipmPD_CStimesLambda.m:86-290 driven by ipmPD_CSsolver.c:132-920 --
// and replaces the scalar instruction list by dense kernels:
//
//   tcd_gram_kernel   once per set of (A,b): H0 = (2/N) A'A, g0 = (2/N) A'b, c0 = b'b/N  (the parameter-only
//                     dependency groups of the generated solver) -- FP64 tensor-core SYRK, mma.sync m8n8k4 (DMMA)
//   tcd_solve_kernel  the whole Newton loop, ONE CTA PER INSTANCE: iterate and multipliers in registers
//                     (thread i owns x_i and the two bound constraints on it), vectors in shared memory, block
//                     reductions for the norms / gap / step bounds, and per iteration the LDL' of the KKT matrix:
//                     the 2n inequality pivots are eliminated in closed form (they are diagonal), the dense
//                     n x n Schur block is factored by a blocked right-looking LDL' whose trailing updates are
//                     DMMA tiles (panels in shared memory, the matrix in an L2-resident per-CTA buffer).
//
// Results agree with the CPU oracle (oracle/dense_ls.c: the reference driver around hand-written dense group
// functions) to rounding: identical status, iterations within +-1, solution within 1e-8 relative
// (tests/test_dense.py).  Not bit-identical: tensor-core tiles and block reductions add in another order.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/tenscalc_b200.h"

#define TCD_THREADS 256
#define TCD_NB 16            // panel width of the blocked LDL' (the serial work per factorisation grows with it)
#define TCD_LDS 20           // row stride (doubles) of the panel buffers: = 4 mod 16 -> conflict-free DMMA fragment loads
#ifndef TCD_CTAS
#define TCD_CTAS 2           // resident CTAs per SM of the solve kernel (register budget 65536/(256*TCD_CTAS))
#endif
#define TCD_SB 32            // block width of the triangular solves (one warp per diagonal block)
#define TCD_KC 32            // rows of A per shared-memory stage of the Gram kernel
#define TCD_GLDS 36          // row stride (doubles) of a Gram stage, = 4 mod 16

#define TCD_CUDA(x)                             \
    do {                                        \
        cudaError_t e_ = (x);                   \
        if (e_ != cudaSuccess) return (int)e_;  \
    } while (0)

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core.  Fragment layout (PTX ISA, mma.m8n8k4 .f64): lane l holds
// A[l/4][l%4], B[l%4][l/4] and C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void tcd_dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------------------
// Gram kernel: grid (batch, ceil(n/64)); CTA = 8 warps, warp w owns the 8 rows i0+8w.. of the 64-row block and
// accumulates the 8x8 tiles (i, j>=block start) plus one extra tile column holding b (-> A'b).  A is streamed in
// stages of TCD_KC rows through shared memory, stored transposed [column][k] so that the global reads (a column's
// 32 consecutive rows = 256 B) and the fragment loads are both conflict-free.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tcd_cp_async8(double *smem_dst, const double *gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 8 : 0;               // src-size 0: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}

// body of the Gram kernel for the block row that starts at tile column TC0 (= 8 * blockIdx.y): with TC0 a
// compile-time constant every fragment address is an immediate offset from two per-lane base pointers and the
// tile loop carries no predicate -- two instructions (LDS.64 + DMMA) per tile instead of eleven
template <int NT, int TC0>
__device__ __forceinline__ void tcd_gram_body(const double *__restrict__ A, const double *__restrict__ b, int64_t N,
                                              double *__restrict__ H0, double *__restrict__ g0, double *__restrict__ c0,
                                              double *sm) {
    constexpr int n = NT * 8;
    constexpr int STAGE = (n + 8) * TCD_GLDS;       // doubles per shared-memory stage
    constexpr int j0 = TC0 * 8;                     // only the columns >= j0 are needed by this block
    const int64_t inst = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tr = TC0 + warp;                      // tile row of this warp
    const double *Ai = A + (size_t)inst * (size_t)N * n;
    const double *bi = b + (size_t)inst * (size_t)N;
    double acc[NT + 1 - TC0][2];
#pragma unroll
    for (int t = 0; t <= NT - TC0; t++) acc[t][0] = acc[t][1] = 0.0;
    double bsq = 0.0;
    const bool active = tr < NT;
    for (int e = threadIdx.x; e < 2 * 7 * TCD_GLDS; e += TCD_THREADS)          // zero padding of the b tile, both stages
        sm[(e / (7 * TCD_GLDS)) * STAGE + (n + 1) * TCD_GLDS + e % (7 * TCD_GLDS)] = 0.0;
    // cp.async double buffering: stage s+1 (rows k0+32.. of the needed columns, 256 contiguous bytes per warp and
    // column, written transposed [column][k]) is in flight while the tensor cores work on stage s.  Warp w fetches
    // the columns j0+w, j0+w+8, ..; warp 0 also fetches b (column n of the stage).
    const double *src0 = Ai + lane + (size_t)N * (j0 + warp);
    double *dst0 = sm + (j0 + warp) * TCD_GLDS + lane;
    auto issue = [&](int64_t k0, int buf) {
        const bool ok = k0 + lane < N;
        const double *src = src0 + k0;
        double *dst = dst0 + buf * STAGE;
#pragma unroll 4
        for (int j = j0 + warp; j < n; j += TCD_THREADS / 32, src += (size_t)N * (TCD_THREADS / 32), dst += (TCD_THREADS / 32) * TCD_GLDS)
            tcd_cp_async8(dst, src, ok);
        if (warp == 0) tcd_cp_async8(sm + buf * STAGE + n * TCD_GLDS + lane, bi + k0 + lane, ok);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0, 0);
    int buf = 0;
    const int fa = (tr * 8 + (lane >> 2)) * TCD_GLDS + (lane & 3);     // A fragment: rows of this warp's tile row
    const int fb = (j0 + (lane >> 2)) * TCD_GLDS + (lane & 3);        // B fragment of tile column TC0; +8*TCD_GLDS per tile
    for (int64_t k0 = 0; k0 < N; k0 += TCD_KC, buf ^= 1) {
        if (k0 + TCD_KC < N) {
            issue(k0 + TCD_KC, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                           // stage `buf` is complete and visible
        const double *st = sm + buf * STAGE;
        if (TC0 == 0 && warp == 0) { const double v = st[n * TCD_GLDS + lane]; bsq += v * v; }
        if (active) {
            const double *pa = st + fa, *pb = st + fb;
#pragma unroll
            for (int kk = 0; kk < TCD_KC / 4; kk++) {
                const double a = pa[4 * kk];
#pragma unroll
                for (int t = 0; t <= NT - TC0; t++)
                    tcd_dmma(acc[t][0], acc[t][1], a, pb[t * 8 * TCD_GLDS + 4 * kk]);
            }
        }
        __syncthreads();                           // everybody is done with `buf` before the next issue overwrites it
    }
    const double scale = 2.0 / (double)N;
    if (active) {
        const int i = tr * 8 + (lane >> 2);
        double *Hi = H0 + (size_t)inst * n * n;
#pragma unroll
        for (int t = 0; t < NT - TC0; t++) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int j = (TC0 + t) * 8 + 2 * (lane & 3) + q;
                if (j >= i) {                      // the upper entry is authoritative: exactly symmetric H0
                    const double v = scale * acc[t][q];
                    Hi[i + (size_t)n * j] = v;
                    Hi[j + (size_t)n * i] = v;
                }
            }
        }
        if ((lane & 3) == 0) g0[(size_t)inst * n + i] = scale * acc[NT - TC0][0];
    }
    if (TC0 == 0 && warp == 0) {
        for (int o = 16; o; o >>= 1) bsq += __shfl_xor_sync(0xffffffffu, bsq, o);
        if (lane == 0) c0[inst] = bsq / (double)N;
    }
}

template <int NT>   // NT = n/8 tile columns
__global__ void __launch_bounds__(TCD_THREADS) tcd_gram_kernel(const double *__restrict__ A, const double *__restrict__ b,
                                                              int64_t N, double *__restrict__ H0,
                                                              double *__restrict__ g0, double *__restrict__ c0) {
    extern __shared__ double sm[];                 // [2][(n + 8)][TCD_GLDS]
    // upper triangle by 64-row blocks: block row y needs the tile columns >= 8 y (uniform per CTA)
    if (blockIdx.y == 0) tcd_gram_body<NT, 0>(A, b, N, H0, g0, c0, sm);
    if constexpr (NT > 8) { if (blockIdx.y == 1) tcd_gram_body<NT, 8>(A, b, N, H0, g0, c0, sm); }
    if constexpr (NT > 16) {
        if (blockIdx.y == 2) tcd_gram_body<NT, 16>(A, b, N, H0, g0, c0, sm);
        if (blockIdx.y == 3) tcd_gram_body<NT, 24>(A, b, N, H0, g0, c0, sm);
    }
}

// ------------------------------------------------------------------------------------------------------------
// solve kernel
// ------------------------------------------------------------------------------------------------------------
struct tcd_args {
    int64_t batch;
    int n;
    double lo, hi;
    // the #define table of lib/cmex2optimizeCS.m:303-331 (defaults of parameters4all.m)
    double gradTolerance, addEye2HessianUtolerance, desiredDualityGap, alphaMin, alphaMax, muFactorAggressive,
        muFactorConservative;
    const double *H0, *g0, *c0, *x0;
    double *x, *J, *work;         // work: [gridDim.x][n*n]
    int32_t *status, *iter;
    double mu0, aeU0;
    int maxIter;
    unsigned long long *flops;    // DMMA flops issued (measurement)
};

struct tcd_smem {
    double red[TCD_THREADS / 32];
    double vec[TCD_THREADS];      // vector operand of the matrix-vector products / right-hand side of the solves
    double diag[TCD_NB * (TCD_NB + 1)];
    double dd[TCD_NB], rd[TCD_NB];             // D of the current diagonal block and 1/D
    double Ls[(TCD_THREADS - 8) * TCD_LDS];   // L21 of the current panel, [row][TCD_LDS]
    double dW[TCD_NB];                        // -D of that panel (the A operand of its trailing update is L21 * -D)
};

enum { RED_SUM = 0, RED_MIN, RED_MAX };

template <int OP> __device__ __forceinline__ double tcd_op(double a, double b) {
    if (OP == RED_SUM) return a + b;
    if (OP == RED_MIN) return a < b ? a : b;
    return a > b ? a : b;
}

// block-wide reduction; every thread receives the same value (so all control decisions are CTA-uniform)
template <int OP> __device__ __forceinline__ double tcd_reduce(double v, tcd_smem &S) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = tcd_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) S.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = S.red[0];
#pragma unroll
    for (int w = 1; w < TCD_THREADS / 32; w++) r = tcd_op<OP>(r, S.red[w]);
    return r;
}

// y_i = sum_j M[i + n j] v_j  (M symmetric, column-major: consecutive threads read consecutive addresses)
__device__ __forceinline__ double tcd_matvec(const double *__restrict__ M, int n, double v, tcd_smem &S) {
    const int i = threadIdx.x;
    __syncthreads();
    S.vec[i] = v;
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (i < n) {
#pragma unroll 2
        for (int j = 0; j < n; j += 8) {           // n is a multiple of 8; eight loads in flight per thread
            const double m0 = M[i + (size_t)n * j], m1 = M[i + (size_t)n * (j + 1)], m2 = M[i + (size_t)n * (j + 2)],
                         m3 = M[i + (size_t)n * (j + 3)], m4 = M[i + (size_t)n * (j + 4)], m5 = M[i + (size_t)n * (j + 5)],
                         m6 = M[i + (size_t)n * (j + 6)], m7 = M[i + (size_t)n * (j + 7)];
            a0 = fma(m0, S.vec[j], a0); a1 = fma(m1, S.vec[j + 1], a1); a2 = fma(m2, S.vec[j + 2], a2); a3 = fma(m3, S.vec[j + 3], a3);
            a0 = fma(m4, S.vec[j + 4], a0); a1 = fma(m5, S.vec[j + 5], a1); a2 = fma(m6, S.vec[j + 6], a2); a3 = fma(m7, S.vec[j + 7], a3);
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// Tiles (ti, tj), tj0 <= tj <= min(ti, tj1-1), of the trailing update A22 -= (L21 D) L21' of one panel: 8x8 DMMA tiles.
// The warps `wid` of `nw` take whole tile rows (long and short rows alternate, so every warp gets about the same
// number of tiles), keep the row's A fragments in registers and walk the row four tiles at a time with the loads of
// the C tiles issued first.
__device__ __forceinline__ void tcd_trailing(double *__restrict__ W22, int n, int mt, int tj0, int tj1, int wid, int nw,
                                             const tcd_smem &S, unsigned long long &flops) {
    const int lane = threadIdx.x & 31;
    const int rows = mt - tj0;                                    // tile rows tj0 .. mt-1 have tiles in the range
    // per-lane bases: the C fragment of tile (ti, tj) is W22[8 ti + lane/4 + n (8 tj + 2 (lane%4) + {0,1})], the B
    // fragment of tile column tj and step kk is Ls[(8 tj + lane/4) * TCD_LDS + 4 kk + lane%4] (indices fit an int)
    const int cb = (lane >> 2) + n * (2 * (lane & 3));
    const double *lb = S.Ls + (lane >> 2) * TCD_LDS + (lane & 3);
    const int n8 = 8 * n;
    unsigned long long done = 0;
    for (int q = wid; q < rows; q += nw) {
        const int ti = tj0 + ((q & 1) ? (q >> 1) : (rows - 1 - (q >> 1)));
        const int tend = min(ti + 1, tj1);
        double af[TCD_NB / 4];
#pragma unroll
        for (int kk = 0; kk < TCD_NB / 4; kk++) af[kk] = lb[ti * 8 * TCD_LDS + 4 * kk] * S.dW[4 * kk + (lane & 3)];
        double *pc = W22 + cb + 8 * ti + tj0 * n8;                // tile (ti, tj0); + n8 per tile column
        const double *pl = lb + tj0 * 8 * TCD_LDS;
        int t0 = tj0;
        for (; t0 + 4 <= tend; t0 += 4, pc += 4 * n8, pl += 4 * 8 * TCD_LDS) {     // four tiles: eight loads in flight
            double c[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) { c[u][0] = pc[u * n8]; c[u][1] = pc[u * n8 + n]; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
#pragma unroll
                for (int kk = 0; kk < TCD_NB / 4; kk++) tcd_dmma(c[u][0], c[u][1], af[kk], pl[u * 8 * TCD_LDS + 4 * kk]);
                pc[u * n8] = c[u][0];
                pc[u * n8 + n] = c[u][1];
            }
            done += 4;
        }
        for (; t0 < tend; t0++, pc += n8, pl += 8 * TCD_LDS) {                    // the row's last one to three tiles
            double c0 = pc[0], c1 = pc[n];
#pragma unroll
            for (int kk = 0; kk < TCD_NB / 4; kk++) tcd_dmma(c0, c1, af[kk], pl[4 * kk]);
            pc[0] = c0;
            pc[n] = c1;
            done += 1;
        }
    }
    if (lane == 0) flops += done * (2ull * 8 * 8 * TCD_NB);
}

// Blocked right-looking LDL' of the symmetric matrix whose LOWER triangle is in W (n x n, column-major), in place:
// D on the diagonal, L strictly below it and, mirrored, strictly above it (the backward substitution then reads
// rows of L' with consecutive threads on consecutive addresses).  Recipe per column as sparsity_ldl.m:144-236
// (v_i = L(col,i) D(i), D(col) = A(col,col) - sum L(col,i) v_i, L(row,col) = (A(row,col) - sum L(row,i) v_i)/D(col),
// no pivoting); the trailing update A22 -= (L21 D) L21' runs on the FP64 tensor cores.
// Look-ahead: of a panel's trailing update only the tiles in the NEXT panel's columns are applied at once; the rest
// is applied by warps 1..7 while warp 0 runs the (serial) factorisation of the next diagonal block.
__device__ __noinline__ void tcd_ldl(double *__restrict__ W, int n, tcd_smem &S, unsigned long long &flops) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int prev_mt = 0;                       // tile rows of the previous panel's trailing matrix (0: nothing pending)
    double *prev_W22 = W;
    for (int k0 = 0; k0 < n; k0 += TCD_NB) {
        const int nb = min(TCD_NB, n - k0), m = n - k0 - nb;
        __syncthreads();
        for (int e = tid; e < nb * nb; e += TCD_THREADS) {             // diagonal block -> shared memory
            const int r = e % nb, c = e / nb;
            if (c <= r) S.diag[r * (TCD_NB + 1) + c] = W[(k0 + r) + (size_t)n * (k0 + c)];
        }
        __syncthreads();
        if (warp == 0) {
            // right-looking elimination of the nb x nb block in shared memory: lane r owns row r; column j, which
            // every lane needs at step j, is a broadcast read
            for (int j = 0; j < nb; j++) {
                const double dj = S.diag[j * (TCD_NB + 1) + j];        // D(j): row j is final
                const double rj = 1.0 / dj;                             // one reciprocal per column; L = A * (1/D)
                double l = 0.0;
                if (lane > j && lane < nb) {
                    l = S.diag[lane * (TCD_NB + 1) + j] * rj;           // L(r,j)
                    S.diag[lane * (TCD_NB + 1) + j] = l;
                }
                if (lane == j) S.rd[j] = rj;
                __syncwarp();
                if (lane > j && lane < nb) {
                    const double vr = -(l * dj);                        // -v_j = -L(r,j) D(j)
                    double *row = S.diag + lane * (TCD_NB + 1);
                    int c = j + 1;
                    for (; c + 3 <= lane; c += 4) {                     // A(r,c) -= L(r,j) D(j) L(c,j), four at a time
                        const double l0 = S.diag[c * (TCD_NB + 1) + j], l1 = S.diag[(c + 1) * (TCD_NB + 1) + j],
                                     l2 = S.diag[(c + 2) * (TCD_NB + 1) + j], l3 = S.diag[(c + 3) * (TCD_NB + 1) + j];
                        const double r0 = row[c], r1 = row[c + 1], r2 = row[c + 2], r3 = row[c + 3];
                        row[c] = fma(vr, l0, r0); row[c + 1] = fma(vr, l1, r1); row[c + 2] = fma(vr, l2, r2); row[c + 3] = fma(vr, l3, r3);
                    }
                    for (; c <= lane; c++) row[c] = fma(vr, S.diag[c * (TCD_NB + 1) + j], row[c]);
                }
                __syncwarp();
            }
            if (lane < nb) S.dd[lane] = S.diag[lane * (TCD_NB + 1) + lane];
        } else if (prev_mt > TCD_NB / 8) {
            // meanwhile: the previous panel's trailing tiles outside the current panel's columns
            tcd_trailing(prev_W22, n, prev_mt, TCD_NB / 8, prev_mt, warp - 1, TCD_THREADS / 32 - 1, S, flops);
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += TCD_THREADS) {             // factor -> W (+ mirror); V = L D -> S.diag
            const int r = e % nb, c = e / nb;
            if (r == c) W[(k0 + r) + (size_t)n * (k0 + c)] = S.dd[r];
            else if (r > c) {
                const double l = S.diag[r * (TCD_NB + 1) + c];
                W[(k0 + r) + (size_t)n * (k0 + c)] = l;
                W[(k0 + c) + (size_t)n * (k0 + r)] = l;
                S.diag[r * (TCD_NB + 1) + c] = l * S.dd[c];
            }
        }
        __syncthreads();
        prev_mt = 0;
        if (m == 0) break;
        // ---- panel: one thread per row below the block; the row stays in registers (right-looking: the updates of
        // one step are independent of each other) ----
        if (tid < m) {
            const int row = k0 + nb + tid;
            double a[TCD_NB];
#pragma unroll
            for (int j = 0; j < TCD_NB; j++) a[j] = j < nb ? W[row + (size_t)n * (k0 + j)] : 0.0;
#pragma unroll
            for (int j = 0; j < TCD_NB; j++) {
                if (j < nb) {
                    const double l = a[j] * S.rd[j];                    // L(row,j) = A(row,j) * (1/D(j))
                    a[j] = l;
#pragma unroll
                    for (int c = j + 1; c < TCD_NB; c++)
                        if (c < nb) a[c] = fma(-l, S.diag[c * (TCD_NB + 1) + j], a[c]);   // A(row,c) -= L(row,j) (L(c,j) D(j))
                }
            }
#pragma unroll
            for (int j = 0; j < TCD_NB; j++) {
                const double l = j < nb ? a[j] : 0.0;
                S.Ls[tid * TCD_LDS + j] = l;
                if (j < nb) {
                    W[row + (size_t)n * (k0 + j)] = l;
                    W[(k0 + j) + (size_t)n * row] = l;
                }
            }
        }
        if (tid < TCD_NB) S.dW[tid] = tid < nb ? -S.dd[tid] : 0.0;       // minus: the update is a subtraction
        __syncthreads();
        // ---- trailing update, urgent part: the tiles in the next panel's columns (all warps) ----
        prev_mt = m / 8;
        prev_W22 = W + (size_t)(k0 + nb) * n + (k0 + nb);
        tcd_trailing(prev_W22, n, prev_mt, 0, TCD_NB / 8, warp, TCD_THREADS / 32, S, flops);
    }
    __syncthreads();
}

// x = (L D L') \ b with the factor of tcd_ldl; thread i holds b_i and receives x_i
__device__ __noinline__ double tcd_ldl_solve(const double *__restrict__ W, int n, double bi, tcd_smem &S) {
    const int i = threadIdx.x, lane = i & 31;
    // forward substitution, panel by panel: the warp that owns the panel solves the nb x nb unit-lower block with
    // shuffles (its rows of the block are loaded into registers first: no memory access inside the dependent
    // chain), then every row below subtracts the panel's contribution
    double y = bi;
    for (int k0 = 0; k0 < n; k0 += TCD_SB) {
        const int nb = min(TCD_SB, n - k0);
        if (i >= k0 && i < k0 + TCD_SB) {                   // one whole warp (k0 is a multiple of 32)
            double lrow[TCD_SB];
#pragma unroll
            for (int j = 0; j < TCD_SB; j++) lrow[j] = (j < lane && lane < nb) ? W[i + (size_t)n * (k0 + j)] : 0.0;
#pragma unroll
            for (int j = 0; j < TCD_SB; j++) {
                const double yj = __shfl_sync(0xffffffffu, y, j);
                y -= lrow[j] * yj;                          // lrow[j] = 0 for j >= lane
            }
            if (lane < nb) S.vec[i] = y;
        }
        __syncthreads();
        if (i >= k0 + nb && i < n) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int j = 0; j < nb; j += 4) {
                a0 += W[i + (size_t)n * (k0 + j)] * S.vec[k0 + j];
                a1 += W[i + (size_t)n * (k0 + j + 1)] * S.vec[k0 + j + 1];
                a2 += W[i + (size_t)n * (k0 + j + 2)] * S.vec[k0 + j + 2];
                a3 += W[i + (size_t)n * (k0 + j + 3)] * S.vec[k0 + j + 3];
            }
            y -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    if (i < n) y = y / W[i + (size_t)n * i];
    // backward substitution with L' (mirrored into the upper triangle), last panel first
    for (int k0 = ((n - 1) / TCD_SB) * TCD_SB; k0 >= 0; k0 -= TCD_SB) {
        const int nb = min(TCD_SB, n - k0);
        if (i >= k0 && i < k0 + TCD_SB) {
            double lrow[TCD_SB];
#pragma unroll
            for (int j = 0; j < TCD_SB; j++) lrow[j] = (j > lane && j < nb) ? W[i + (size_t)n * (k0 + j)] : 0.0;
#pragma unroll
            for (int j = TCD_SB - 1; j >= 0; j--) {
                const double xj = __shfl_sync(0xffffffffu, y, j);
                y -= lrow[j] * xj;                          // lrow[j] = 0 for j <= lane
            }
            if (lane < nb) S.vec[i] = y;
        }
        __syncthreads();
        if (i < k0) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int j = 0; j < nb; j += 4) {
                a0 += W[i + (size_t)n * (k0 + j)] * S.vec[k0 + j];
                a1 += W[i + (size_t)n * (k0 + j + 1)] * S.vec[k0 + j + 1];
                a2 += W[i + (size_t)n * (k0 + j + 2)] * S.vec[k0 + j + 2];
                a3 += W[i + (size_t)n * (k0 + j + 3)] * S.vec[k0 + j + 3];
            }
            y -= (a0 + a1) + (a2 + a3);
        }
        __syncthreads();
    }
    return y;
}

#define TCD_MIN(A, B) (((A) < (B)) ? (A) : (B))
#define TCD_MAX(A, B) (((A) > (B)) ? (A) : (B))

// ipmPD_CSsolver() (lib/ipmPD_CSsolver.c:132-920) for the option set cmex2optimizeCS generates for this problem
// (nG=0, nF=2n, skipAffine=1, adjustAddEye2Hessian=1, useInertia=0, scaleCost=0, scaleInequalities=1,
// coupledAlphas=1), one instance per CTA; CTAs take instances from the grid-strided range.
__global__ void __launch_bounds__(TCD_THREADS, TCD_CTAS) tcd_solve_kernel(tcd_args P) {
    extern __shared__ double smraw[];
    tcd_smem &S = *reinterpret_cast<tcd_smem *>(smraw);
    const int n = P.n, i = threadIdx.x;
    const bool on = i < n;
    double *W = P.work + (size_t)blockIdx.x * n * n;
    unsigned long long flops = 0;
    const double BIG = DBL_MAX;
    for (int64_t inst = blockIdx.x; inst < P.batch; inst += gridDim.x) {
        const double *H0 = P.H0 + (size_t)inst * n * n;
        const double g0 = on ? P.g0[(size_t)inst * n + i] : 0.0;
        // ---- state of thread i: x_i and the constraints x_i >= lo (k=i), x_i <= hi (k=n+i) ----
        double x = on ? P.x0[(size_t)inst * n + i] : 0.0;                      // initPrimal__
        double s1 = on ? fabs(1.0 / (x - P.lo)) : 1.0, s2 = on ? fabs(1.0 / (P.hi - x)) : 1.0;   // scaleIneq__ (:86-89)
        double mu = P.mu0, aeU = P.aeU0, aeU_set = P.aeU0;   // driver variable / value stored by setAddEye2HessianU__
        double F1 = s1 * (x - P.lo), F2 = s2 * (P.hi - x);
        double l1 = mu / F1, l2 = mu / F2;                                     // initDualIneq__ (:198)
        double dU = 0, d1 = 0, d2 = 0, Lfu = 0, hdu = 0;
        bool dx_valid = false, updU = false;
        int status = 0, iter = 0;
        const double gapTol = P.desiredDualityGap;
        const double muMin = gapTol / (2 * n) / 2;
        double alphaPrimal = 0, alphaDualIneq = 0, alphaDualEq = 0, alphaMax_ = P.alphaMax;
        double norminf_grad = 0, gap = 0, ineq = 0, ineq1 = 0, dual = 0;

        // Newton direction (ipmPD_CStimesLambda.m:235-256): closed-form elimination of the 2n inequality pivots,
        // dense LDL' of the Schur block, three solves; also hdu = H0*dU for the curvature / direction error
        auto eval_dx = [&]() {
            if (dx_valid) return;
            const double DF1 = -(F1 * l1), DF2 = -(F2 * l2);
            const double LF1 = (-(l1 * s1)) / DF1, LF2 = (l2 * s2) / DF2;
            __syncthreads();
            if (on) {                                                           // W = H0 (coalesced copy, 8 loads in flight)
#pragma unroll 2
                for (int j = 0; j < n; j += 8) {
                    double t[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) t[u] = H0[i + (size_t)n * (j + u)];
#pragma unroll
                    for (int u = 0; u < 8; u++) W[i + (size_t)n * (j + u)] = t[u];
                }
            }
            if (on) W[i + (size_t)n * i] = (H0[i + (size_t)n * i] + aeU_set) - (LF1 * (LF1 * DF1) + LF2 * (LF2 * DF2));
            tcd_ldl(W, n, S, flops);
            const double yF1 = l1 * F1 - mu, yF2 = l2 * F2 - mu;
            const double bU = -Lfu - (LF1 * yF1 + LF2 * yF2);
            dU = tcd_ldl_solve(W, n, on ? bU : 0.0, S);
            d1 = yF1 / DF1 - LF1 * dU;
            d2 = yF2 / DF2 - LF2 * dU;
            hdu = tcd_matvec(H0, n, dU, S);
            dx_valid = true;
        };

        while (true) {
            iter++;
            if (iter > P.maxIter) { status = 8; break; }                       // :308-313
            // ---- exit conditions (:323-381) ----
            const double hx = tcd_matvec(H0, n, x, S);
            Lfu = (hx - g0) - s1 * l1 + s2 * l2;                                // Lf_u = f_u - F_u' lambda
            norminf_grad = tcd_reduce<RED_MAX>(on ? fabs(Lfu) : 0.0, S);
            if (isnan(norminf_grad)) { status = 4; break; }
            F1 = s1 * (x - P.lo); F2 = s2 * (P.hi - x);
            gap = tcd_reduce<RED_SUM>(on ? l1 * F1 + l2 * F2 : 0.0, S);
            ineq = tcd_reduce<RED_MIN>(on ? TCD_MIN(F1, F2) : BIG, S);
            dual = tcd_reduce<RED_MIN>(on ? TCD_MIN(l1, l2) : BIG, S);
            if (ineq <= 0) { status = 1; break; }
            if (dual <= 0) { status = 2; break; }
            if (norminf_grad <= P.gradTolerance && gap <= gapTol && aeU <= P.addEye2HessianUtolerance) { status = 0; break; }
            // ---- regularisation control (:393-548, useInertia=0) ----
            if (updU) { aeU_set = aeU; dx_valid = false; updU = false; }        // delayed from the last iteration
            eval_dx();
            const double curvature = tcd_reduce<RED_SUM>(on ? dU * (hdu + aeU_set * dU) : 0.0, S);
            double r = on ? fabs((hdu + aeU_set * dU) + (-(l1 * s1)) * d1 + (l2 * s2) * d2 + Lfu) : 0.0;
            r = TCD_MAX(r, on ? fabs(-(l1 * s1) * dU - (F1 * l1) * d1 - (l1 * F1 - mu)) : 0.0);
            r = TCD_MAX(r, on ? fabs((l2 * s2) * dU - (F2 * l2) * d2 - (l2 * F2 - mu)) : 0.0);
            const double derr = tcd_reduce<RED_MAX>(r, S);
            if (curvature > 0) {                                                // :415-456: stored at the next iteration
                if (aeU > 1e-20 && derr < 1e-6) { aeU = TCD_MAX(.75 * aeU, 1e-20); updU = true; }
                if (aeU < 1e2 && derr > 1e-6) { aeU = TCD_MIN(10 * aeU, 1e2); updU = true; }
            } else if (curvature < 0) {                                         // :462-520: stored now, single pass
                if (aeU < 1e2) { aeU = TCD_MIN(10 * TCD_MAX(aeU, 1e-20), 1e2); aeU_set = aeU; dx_valid = false; }
            }
            // ---- combined search direction (:679-772) ----
            eval_dx();
            double ap = BIG, ad = BIG;
            if (on) {
                const double q1 = s1 * dU, q2 = -(s2 * dU);
                if (q1 < 0) ap = TCD_MIN(ap, -F1 / q1);
                if (q2 < 0) ap = TCD_MIN(ap, -F2 / q2);
                if (d1 < 0) ad = TCD_MIN(ad, -1.0 / d1);
                if (d2 < 0) ad = TCD_MIN(ad, -1.0 / d2);
            }
            alphaPrimal = tcd_reduce<RED_MIN>(ap, S);
            alphaDualIneq = tcd_reduce<RED_MIN>(ad, S);
            if (alphaDualIneq < alphaPrimal) alphaPrimal = alphaDualIneq;       // coupledAlphas
            const double stepback = .99;
            alphaPrimal *= stepback;
            alphaMax_ = (alphaPrimal < P.alphaMax) ? alphaPrimal : P.alphaMax;
            auto minF = [&](double a) {                                         // getMinF_s__ (:276-277)
                const double xn = x + a * dU;
                return tcd_reduce<RED_MIN>(on ? TCD_MIN(s1 * (xn - P.lo), s2 * (P.hi - xn)) : BIG, S);
            };
            bool nan_exit = false;
            if (alphaMax_ >= P.alphaMin) {
                alphaPrimal = alphaMax_ / stepback;                             // try max
                ineq = minF(alphaPrimal);
                if (isnan(ineq)) nan_exit = true;
                else {
                    if (ineq > 0) { alphaPrimal *= stepback; ineq1 = minF(alphaPrimal); }
                    if (ineq <= 0 || ineq1 <= ineq / 10) {
                        alphaPrimal = P.alphaMin / stepback;                    // try min
                        ineq = minF(alphaPrimal);
                        if (ineq > 0) {
                            for (alphaPrimal = alphaMax_ * .95; alphaPrimal >= P.alphaMin; alphaPrimal /= 2) {
                                ineq = minF(alphaPrimal);
                                if (ineq > 0) {
                                    alphaPrimal *= stepback;
                                    ineq1 = minF(alphaPrimal);
                                    if (ineq1 > ineq / 10) break;
                                }
                            }
                            if (alphaPrimal < P.alphaMin) alphaPrimal = 0;
                        } else alphaPrimal = 0;
                    }
                }
            } else alphaPrimal = 0;
            if (nan_exit) { status = 4; break; }
            alphaDualEq = alphaPrimal;
            alphaDualIneq = alphaPrimal;
            // updatePrimalDual__ (:261-281)
            x = x + alphaPrimal * dU;
            l1 = l1 * (1.0 + alphaDualIneq * d1);
            l2 = l2 * (1.0 + alphaDualIneq * d2);
            dx_valid = false;
            // ---- barrier parameter (:782-859) ----
            {
                const bool th_grad = norminf_grad < TCD_MAX(1e-6, 1e0 * P.gradTolerance);
                if (alphaPrimal > alphaMax_ / 2 && th_grad) {
                    mu *= TCD_MIN(P.muFactorAggressive, sqrt(mu));
                    if (mu < muMin) mu = muMin;
                } else if (alphaPrimal < .1) {
                    mu *= 1.1;
                    if (mu > P.mu0) mu = P.mu0;
                    F1 = s1 * (x - P.lo); F2 = s2 * (P.hi - x);                 // initDualIneq__ (:816)
                    l1 = mu / F1; l2 = mu / F2;
                } else if (alphaPrimal > .99) {
                    mu *= P.muFactorConservative;
                    if (mu < muMin) mu = muMin;
                }
                if (alphaPrimal < P.alphaMin && alphaDualIneq < P.alphaMin && alphaDualEq < P.alphaMin) {
                    mu /= (P.muFactorConservative * P.muFactorConservative);
                    if (mu < muMin) mu = muMin;
                }
            }
        }
        if (status == 8) {                                                      // :885-920
            const double hx = tcd_matvec(H0, n, x, S);
            Lfu = (hx - g0) - s1 * l1 + s2 * l2;
            norminf_grad = tcd_reduce<RED_MAX>(on ? fabs(Lfu) : 0.0, S);
            if (norminf_grad > P.gradTolerance) status |= 16;
            F1 = s1 * (x - P.lo); F2 = s2 * (P.hi - x);
            gap = tcd_reduce<RED_SUM>(on ? l1 * F1 + l2 * F2 : 0.0, S);
            if (gap > gapTol) status |= 64;
            if (mu > muMin) status |= 128;
            if (alphaPrimal <= P.alphaMin && alphaDualIneq < P.alphaMin && alphaDualEq < P.alphaMin) status |= 1792;
            else if (alphaPrimal <= .1 && alphaDualIneq < .1 && alphaDualEq < .1) status |= 1536;
            else if (alphaPrimal <= .5 && alphaDualIneq < .5 && alphaDualEq < .5) status |= 1024;
            if (aeU > P.addEye2HessianUtolerance) status |= 2048;
        }
        // getOutputs: J = ||Ax-b||^2/N = x'(H0 x)/2 - g0'x + b'b/N, and x
        {
            const double hx = tcd_matvec(H0, n, x, S);
            const double q = tcd_reduce<RED_SUM>(on ? x * (.5 * hx - g0) : 0.0, S);
            if (on) P.x[(size_t)inst * n + i] = x;
            if (i == 0) {
                P.J[inst] = q + P.c0[inst];
                P.status[inst] = status;
                P.iter[inst] = iter;
            }
        }
        __syncthreads();
    }
    if ((i & 31) == 0 && flops) atomicAdd(P.flops, flops);
}

// ------------------------------------------------------------------------------------------------------------
// host side / C ABI
// ------------------------------------------------------------------------------------------------------------
struct tc_dense {
    int device = 0, sms = 0;
    int64_t N = 0, n = 0, batch = 0;
    double lo = 0, hi = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    std::vector<void *> allocs;
    const double *A = nullptr, *b = nullptr;       // device pointers (own copy or borrowed)
    double *own_A = nullptr, *own_b = nullptr, *stage = nullptr;
    double *H0 = nullptr, *g0 = nullptr, *c0 = nullptr, *x0 = nullptr, *x = nullptr, *J = nullptr, *work = nullptr;
    int32_t *status = nullptr, *iter = nullptr;
    unsigned long long *flops = nullptr;
    bool gram_dirty = true;
    int grid = 0;
    int64_t launches = 0;
    float gram_ms = 0.f, solve_ms = 0.f;
};

template <typename T> static int tcd_alloc(tc_dense *s, T **p, size_t count) {
    void *d = nullptr;
    TCD_CUDA(cudaMalloc(&d, count * sizeof(T)));
    s->allocs.push_back(d);
    *p = (T *)d;
    return 0;
}

template <int NT> static int tcd_launch_gram(tc_dense *s) {
    const size_t smem = 2 * (size_t)(NT * 8 + 8) * TCD_GLDS * sizeof(double);
    TCD_CUDA(cudaFuncSetAttribute(tcd_gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)s->batch, (unsigned)((NT + 7) / 8));
    tcd_gram_kernel<NT><<<grid, TCD_THREADS, smem, s->stream>>>(s->A, s->b, s->N, s->H0, s->g0, s->c0);
    return (int)cudaGetLastError();
}

extern "C" {

int tc_dense_create(tc_dense **out, int64_t N, int64_t n, int64_t batch, double lo, double hi, int device) {
    if (!out || N <= 0 || batch <= 0 || !(lo < hi)) return TC_EINVAL;
    if (!(n == 8 || n == 16 || n == 32 || n == 64 || n == 128 || n == 256)) return TC_EINVAL;   // n/8 DMMA tile columns, n <= CTA size
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return TC_ENOGPU;
    tc_dense *s = new (std::nothrow) tc_dense();
    if (!s) return TC_EINVAL;
    s->device = device; s->N = N; s->n = n; s->batch = batch; s->lo = lo; s->hi = hi;
    int rc = 0;
#define TCD_TRY(x) do { rc = (x); if (rc) { tc_dense_destroy(s); return rc; } } while (0)
    TCD_TRY((int)cudaSetDevice(device));
    TCD_TRY((int)cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device));
    TCD_TRY((int)cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TCD_TRY((int)cudaEventCreate(&s->ev0));
    TCD_TRY((int)cudaEventCreate(&s->ev1));
    TCD_TRY((int)cudaEventCreate(&s->ev2));
    s->grid = (int)(batch < TCD_CTAS * s->sms ? batch : TCD_CTAS * s->sms);   // several CTAs per SM: one's serial sections overlap the others'
    TCD_TRY(tcd_alloc(s, &s->H0, (size_t)batch * n * n));
    TCD_TRY(tcd_alloc(s, &s->g0, (size_t)batch * n));
    TCD_TRY(tcd_alloc(s, &s->c0, (size_t)batch));
    TCD_TRY(tcd_alloc(s, &s->x0, (size_t)batch * n));
    TCD_TRY(tcd_alloc(s, &s->x, (size_t)batch * n));
    TCD_TRY(tcd_alloc(s, &s->J, (size_t)batch));
    TCD_TRY(tcd_alloc(s, &s->work, (size_t)s->grid * n * n));
    TCD_TRY(tcd_alloc(s, &s->status, (size_t)batch));
    TCD_TRY(tcd_alloc(s, &s->iter, (size_t)batch));
    TCD_TRY(tcd_alloc(s, &s->flops, 1));
    TCD_TRY((int)cudaMemsetAsync(s->flops, 0, sizeof(unsigned long long), s->stream));
    TCD_TRY((int)cudaMemsetAsync(s->x0, 0, (size_t)batch * n * sizeof(double), s->stream));
    TCD_TRY((int)cudaFuncSetAttribute(tcd_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tcd_smem)));
#undef TCD_TRY
    *out = s;
    return 0;
}

int tc_dense_destroy(tc_dense *s) {
    if (!s) return 0;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void *p : s->allocs) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

// which: 0 = A (double[batch][N*n], each instance column-major), 1 = b (double[batch][N]), 2 = x (double[batch][n])
static int tcd_set(tc_dense *s, int which, const double *v, int on_device) {
    if (!s || !v || which < 0 || which > 2) return TC_EINVAL;
    TCD_CUDA(cudaSetDevice(s->device));
    const size_t count = (size_t)s->batch * (which == 0 ? (size_t)s->N * s->n : which == 1 ? (size_t)s->N : (size_t)s->n);
    if (which == 2) {
        TCD_CUDA(cudaMemcpyAsync(s->x0, v, count * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s->stream));
        return 0;
    }
    s->gram_dirty = true;
    if (on_device) {                      // borrowed: must stay valid until the next solve has been enqueued
        (which == 0 ? s->A : s->b) = v;
        return 0;
    }
    double **own = which == 0 ? &s->own_A : &s->own_b;
    if (!*own) {
        int rc = tcd_alloc(s, own, count);
        if (rc) return rc;
    }
    TCD_CUDA(cudaMemcpyAsync(*own, v, count * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    (which == 0 ? s->A : s->b) = *own;
    return 0;
}

int tc_dense_set_A(tc_dense *s, const double *A, int on_device) { return tcd_set(s, 0, A, on_device); }
int tc_dense_set_b(tc_dense *s, const double *b, int on_device) { return tcd_set(s, 1, b, on_device); }
int tc_dense_set_x(tc_dense *s, const double *x, int on_device) { return tcd_set(s, 2, x, on_device); }

int tc_dense_solve_async(tc_dense *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                         const double *addEye2Hessian) {
    if (!s || !mu0 || !maxIter || !addEye2Hessian) return TC_EINVAL;
    (void)saveIter;                       // allowSave=0: the reference ignores it too (ipmPD_CSsolver.c:562-564)
    if (!s->A || !s->b) return TC_EINVAL;
    TCD_CUDA(cudaSetDevice(s->device));
    TCD_CUDA(cudaEventRecord(s->ev0, s->stream));
    if (s->gram_dirty) {                  // the parameter-only groups: evaluated lazily, once per set of A / b
        int rc = 0;
        switch (s->n / 8) {
        case 1: rc = tcd_launch_gram<1>(s); break;
        case 2: rc = tcd_launch_gram<2>(s); break;
        case 4: rc = tcd_launch_gram<4>(s); break;
        case 8: rc = tcd_launch_gram<8>(s); break;
        case 16: rc = tcd_launch_gram<16>(s); break;
        case 32: rc = tcd_launch_gram<32>(s); break;
        default: return TC_EINVAL;
        }
        if (rc) return rc;
        s->launches++;
        s->gram_dirty = false;
    }
    TCD_CUDA(cudaEventRecord(s->ev1, s->stream));
    tcd_args P{};
    P.batch = s->batch; P.n = (int)s->n; P.lo = s->lo; P.hi = s->hi;
    P.gradTolerance = 1e-4; P.addEye2HessianUtolerance = 1e-6; P.desiredDualityGap = 1e-5; P.alphaMin = 1e-7;
    P.alphaMax = 1.0; P.muFactorAggressive = .2; P.muFactorConservative = .95;
    P.H0 = s->H0; P.g0 = s->g0; P.c0 = s->c0; P.x0 = s->x0; P.x = s->x; P.J = s->J; P.work = s->work;
    P.status = s->status; P.iter = s->iter; P.mu0 = *mu0; P.aeU0 = addEye2Hessian[0]; P.maxIter = *maxIter;
    P.flops = s->flops;
    tcd_solve_kernel<<<s->grid, TCD_THREADS, sizeof(tcd_smem), s->stream>>>(P);
    s->launches++;
    TCD_CUDA(cudaGetLastError());
    TCD_CUDA(cudaEventRecord(s->ev2, s->stream));
    return 0;
}

int tc_dense_solve(tc_dense *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                   const double *addEye2Hessian, int32_t *status, int32_t *iter, double *time) {
    if (!status || !iter) return TC_EINVAL;
    int rc = tc_dense_solve_async(s, mu0, maxIter, saveIter, addEye2Hessian);
    if (rc) return rc;
    TCD_CUDA(cudaMemcpyAsync(status, s->status, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    TCD_CUDA(cudaMemcpyAsync(iter, s->iter, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    TCD_CUDA(cudaStreamSynchronize(s->stream));
    TCD_CUDA(cudaEventElapsedTime(&s->gram_ms, s->ev0, s->ev1));
    TCD_CUDA(cudaEventElapsedTime(&s->solve_ms, s->ev1, s->ev2));
    if (time) *time = 1e-3 * (double)(s->gram_ms + s->solve_ms);
    return 0;
}

// <class>_getOutputs: J double[batch], x double[batch][n]; status/iter int32[batch] (may be NULL)
int tc_dense_get_outputs(tc_dense *s, double *J, double *x, int32_t *status, int32_t *iter, int on_device) {
    if (!s || !J || !x) return TC_EINVAL;
    TCD_CUDA(cudaSetDevice(s->device));
    const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    TCD_CUDA(cudaMemcpyAsync(J, s->J, s->batch * sizeof(double), k, s->stream));
    TCD_CUDA(cudaMemcpyAsync(x, s->x, (size_t)s->batch * s->n * sizeof(double), k, s->stream));
    if (status) TCD_CUDA(cudaMemcpyAsync(status, s->status, s->batch * sizeof(int32_t), k, s->stream));
    if (iter) TCD_CUDA(cudaMemcpyAsync(iter, s->iter, s->batch * sizeof(int32_t), k, s->stream));
    if (!on_device) TCD_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int tc_dense_synchronize(tc_dense *s) {
    if (!s) return TC_EINVAL;
    TCD_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

void *tc_dense_stream(tc_dense *s) { return s ? (void *)s->stream : nullptr; }

// measurement hooks: CUDA-event times of the last solve (Gram kernel, Newton-loop kernel), kernel launches so
// far, and the FP64 tensor-core flops the LDL' trailing updates have issued since creation
int tc_dense_times_ms(tc_dense *s, float *gram_ms, float *solve_ms) {
    if (!s || !gram_ms || !solve_ms) return TC_EINVAL;
    TCD_CUDA(cudaEventSynchronize(s->ev2));
    TCD_CUDA(cudaEventElapsedTime(gram_ms, s->ev0, s->ev1));
    TCD_CUDA(cudaEventElapsedTime(solve_ms, s->ev1, s->ev2));
    return 0;
}

int tc_dense_counters(tc_dense *s, int64_t *launches, int64_t *dmma_flops) {
    if (!s || !launches || !dmma_flops) return TC_EINVAL;
    unsigned long long h = 0;
    TCD_CUDA(cudaSetDevice(s->device));
    TCD_CUDA(cudaMemcpyAsync(&h, s->flops, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    TCD_CUDA(cudaStreamSynchronize(s->stream));
    *launches = s->launches;
    *dmma_flops = (int64_t)h;
    return 0;
}

}  // extern "C"
