// Tape interpreter: straight-line code of one dependency group, one instance per lane.
//
// Replaces the per-problem C functions `classname_groupK()` that the reference emits with
// writeCinstructionsC (lib/csparse/instructionsTableUTHash.c:950-1539).  The 32 lanes of a
// warp walk the SAME tape (no divergence inside a group) and each lane applies it to its own
// instance.  The scratchbook is tiled per warp and lane-interleaved (tc_image.h), so a
// warp-wide operand access is one contiguous 256-byte segment.
//
// Execution model (tape format: tenscalc_b200/tape.py).  The tape is a sequence of 128-bit
// cells of an accumulator machine (registers acc, p; up to two memory operands per cell).
// Cells are processed in bundles of TC_BUNDLE: while bundle b executes, the 2*TC_BUNDLE
// scratchbook/constant operands of bundle b+1 are already in flight (branch-free, predicated
// loads issued back to back), which is what hides HBM latency at modest occupancy.  Operands
// produced by one of the last TC_RING instructions never come from the scratchbook -- the
// lowering routes them through the per-lane RESULT RING in shared memory -- so fetching a
// bundle early can never read a stale value, and short-lived temporaries never touch HBM.
//
// Arithmetic contract (parity with the CPU oracle): every instruction is evaluated in the C
// expression order of the reference emitter, with separately rounded multiplies and adds
// (__dmul_rn/__dadd_rn are never contracted into DFMA), IEEE division and square root.
#pragma once
#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

#include "tc_image.h"
#include "tc_math.cuh"

#define TC_FULL 0xFFFFFFFFu

// Uniform control decisions.  By default a warp is autonomous and decides by warp vote.  With
// TC_LOCKSTEP=W (compiled back end) the W warps of a CTA walk the driver and the chunk functions
// TOGETHER (CTA-wide votes, one barrier per chunk), so that the instruction stream of the
// straight-line code -- megabytes of SASS per Newton step, far beyond the instruction caches --
// is fetched once per CTA instead of once per warp.
#if defined(TC_LOCKSTEP) && TC_LOCKSTEP > 1
#define TC_ANY(x) (__syncthreads_or((int)(x)) != 0)
#else
#define TC_ANY(x) (__any_sync(TC_FULL, (x)) != 0)
#endif

#ifndef TC_TILES
#define TC_TILES 1         // instances per lane = scratchbook tiles per warp (compiled back end: emit_cuda.py)
#endif
// doubles between two slots of one instance: 32, or 64 when the two instances of a lane are interleaved
// (emitter option pair, tc_device_image::ilv = 1)
#if defined(TC_PAIR) || defined(TC_WIDE)
#define TC_SLOT_STRIDE (2 * TC_TILE)
#else
#define TC_SLOT_STRIDE TC_TILE
#endif

struct tc_lane {
    double *mb[TC_TILES];  // &m[tile][0][lane]; slot k of the instance in tile t is mb[t][k*32]
    const double *consts;
    const uint4 *tape;
    double *sm;            // shared memory of the CTA: result ring [TC_RING][TC_BLOCK], then constants
    unsigned lane;
    uint64_t st[TC_TILES][TC_GW];   // groupStatus[] of every instance, one bit per group
};

#ifdef TC_COMPILED_GROUPS
// Compiled back end: the straight-line code of the groups and of the copy functions is CUDA C
// generated per problem by tenscalc_b200/emit_cuda.py and linked into the same module.
// run dependency group g on the instances with pred set; called by every thread that took the
// vote (warp, or CTA under TC_LOCKSTEP: then every chunk ends in a CTA barrier)
// `mask` bit t: run it on the instance of tile t (tiles of a warp are adjacent: m + t*S*32)
extern __device__ void tcm_group(int g, double *m, unsigned mask);
extern __device__ bool tcm_copy(int fn, double *m, unsigned mask);   // copy function fn; false: no compiled body

__device__ __forceinline__ void tc_exec_group(tc_lane &c, const uint32_t *, int g, const bool (&pred)[TC_TILES]) {
    unsigned mask = 0;
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) mask |= pred[t] ? (1u << t) : 0u;
    tcm_group(g, c.mb[0], mask);
    __syncwarp();
}
#else
__device__ __forceinline__ double tc_unary(int op, double x) {
    switch (op) {   // C forms: instructionsTableUTHash.c:1347-1539
    case I_inv: return 1.0 / x;
    case I_minus_inv_sqr: return -1.0 / __dmul_rn(x, x);
    case I_2_inv_cube: return 2.0 / __dmul_rn(__dmul_rn(x, x), x);
    case I_exp: return exp(x);
    case I_log: return log(x);
    case I_cos: return cos(x);
    case I_minus_cos: return -cos(x);
    case I_sin: return sin(x);
    case I_minus_sin: return -sin(x);
    case I_round: return round(x);
    case I_ceil: return ceil(x);
    case I_floor: return floor(x);
    case I_abs: return fabs(x);
    case I_sign: return (x >= 0) ? 1.0 : -1.0;
    case I_sqrt: return sqrt(x);
    case I_Dsqrt: return __dmul_rn(.5, tc_pow_m05(x));
    case I_DDsqrt: return __dmul_rn(-.25, tc_pow_m15(x));
    case I_sqr: return __dmul_rn(x, x);
    case I_2times: return __dmul_rn(2.0, x);
    case I_cube: return __dmul_rn(__dmul_rn(x, x), x);
    case I_3sqr: return __dmul_rn(__dmul_rn(3.0, x), x);
    case I_6times: return __dmul_rn(6.0, x);
    case I_atan: return atan(x);
    case I_Datan: return 1.0 / __dadd_rn(1.0, __dmul_rn(x, x));
    case I_DDatan: {
        double d = __dadd_rn(1.0, __dmul_rn(x, x));
        return (__dmul_rn(-2.0, x) / d) / d;
    }
    case I_srelu: return log(__dadd_rn(1.0, exp(x)));
    case I_dsrelu: return 1.0 / __dadd_rn(1.0, exp(-x));
    case I_ddsrelu: return 1.0 / __dadd_rn(__dadd_rn(2.0, exp(-x)), exp(x));
    case I_relu: return (x > 0) ? x : 0.0;
    case I_heaviside: return (x > 0) ? 1.0 : 0.0;
    }
    // not an opcode the interpreter knows (I_componentwise: user C text, compiled modules only): NaN, so that
    // a handle that runs such a tape without its module fails loudly (status 4) instead of computing something else
    return __longlong_as_double(0x7ff8000000000000ll);
}

// every cell kind except OP_MADD, out of line so that the hot path stays small
__device__ __noinline__ double2 tc_cell_other(unsigned w0, double acc, double p, double x, double y) {
    const unsigned op = w0 & 0x7fu;
    const double sx = (w0 & TC_F_NEGA) ? -x : x, sy = (w0 & TC_F_NEGB) ? -y : y;
    switch (op) {
    case OP_NOP: break;
    case OP_SUM2F: acc = __dadd_rn(sx, sy); break;                          // +a-b+c...    :964-968
    case OP_SUM2: acc = __dadd_rn(__dadd_rn(acc, sx), sy); break;
    case OP_SUM1F: acc = sx; break;
    case OP_SUM1: acc = __dadd_rn(acc, sx); break;
    case OP_PSET: p = __dmul_rn(x, y); break;                               // products of >2 factors :1007-1029
    case OP_PMUL1: p = __dmul_rn(p, x); break;
    case OP_PMUL2: p = __dmul_rn(__dmul_rn(p, x), y); break;
    case OP_P1SET: acc = __dmul_rn(p, x); break;
    case OP_P1ADD: acc = __dadd_rn(acc, __dmul_rn(p, x)); break;
    case OP_P2SET: acc = __dmul_rn(__dmul_rn(p, x), y); break;
    case OP_P2ADD: acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(p, x), y)); break;
    case OP_RSUBDIV: acc = __dadd_rn(x, -acc) / y; break;                   // (a0-(dot))/d :1081-1098
    case OP_RSUB: acc = __dadd_rn(x, -acc); break;                          // a0-(dot)     :1128-1144
    case OP_NEGDIV: acc = (-acc) / x; break;                                // -(dot)/d     :1106-1121
    case OP_DIV: acc = x / y; break;
    case OP_SQ1F: acc = __dmul_rn(x, x); break;                             // a*a+b*b+...  :1201-1210
    case OP_SQ2F: acc = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)); break;
    case OP_SQ1: acc = __dadd_rn(acc, __dmul_rn(x, x)); break;
    case OP_SQ2: acc = __dadd_rn(__dadd_rn(acc, __dmul_rn(x, x)), __dmul_rn(y, y)); break;
    case OP_ABS1F: acc = fabs(x); break;                                    // |a|+|b|+...  :1217-1225
    case OP_ABS2F: acc = __dadd_rn(fabs(x), fabs(y)); break;
    case OP_ABS1: acc = __dadd_rn(acc, fabs(x)); break;
    case OP_ABS2: acc = __dadd_rn(__dadd_rn(acc, fabs(x)), fabs(y)); break;
    // sequential compare chains (:1228-1313): a NaN in a non-first operand is dropped
    case OP_MINF1: acc = x; break;
    case OP_MINF2: acc = x; if (acc > y) acc = y; break;
    case OP_MIN2: if (w0 & TC_F_INIT) acc = 0.0; if (acc > x) acc = x; if (acc > y) acc = y; break;
    case OP_MIN1: if (w0 & TC_F_INIT) acc = 0.0; if (acc > x) acc = x; break;
    case OP_MAXF1: acc = x; break;
    case OP_MAXF2: acc = x; if (acc < y) acc = y; break;
    case OP_MAX2: if (w0 & TC_F_INIT) acc = 0.0; if (acc < x) acc = x; if (acc < y) acc = y; break;
    case OP_MAX1: if (w0 & TC_F_INIT) acc = 0.0; if (acc < x) acc = x; break;
    case OP_AMAXF1: acc = fabs(x); break;
    case OP_AMAXF2: acc = fabs(x); if (acc < fabs(y)) acc = fabs(y); break;
    case OP_AMAX2: if (acc < fabs(x)) acc = fabs(x); if (acc < fabs(y)) acc = fabs(y); break;
    case OP_AMAX1: if (acc < fabs(x)) acc = fabs(x); break;
    case OP_CLP:                                                            // :1315-1331
        if (w0 & TC_F_INIT) acc = DBL_MAX;
        if (y < 0) {
            double q = (-x) / y;
            if (acc > q) acc = q;
        }
        break;
    default: acc = tc_unary((int)op - OP_UNARY, x); break;
    }
    return make_double2(acc, p);
}

struct tc_bundle {                 // one bundle in flight: cells and their early (scratchbook) operands
    uint4 c[TC_BUNDLE];
    double x[TC_BUNDLE], y[TC_BUNDLE];
};

// Look-ahead fetch: scratchbook operands only, one predicated load each, no branches -- the
// 2*TC_BUNDLE loads of a bundle issue back to back and stay in flight while the previous
// bundle executes.
__device__ __forceinline__ void tc_fetch(tc_bundle &B, const uint4 *cells, const char *mb, bool pred) {
#pragma unroll
    for (int j = 0; j < TC_BUNDLE; j++) B.c[j] = __ldg(cells + j);
#pragma unroll
    for (int j = 0; j < TC_BUNDLE; j++) {
        const unsigned w0 = B.c[j].x;
        double x = 0.0, y = 0.0;
        if (pred && !(w0 & TC_F_ALATE)) x = *(const double *)(mb + B.c[j].y);
        if (pred && !(w0 & TC_F_BLATE)) y = *(const double *)(mb + B.c[j].z);
        B.x[j] = x;
        B.y[j] = y;
    }
}

// late operand: shared memory.  Word = byte offset | 1 for per-thread data (result ring),
// even for CTA-wide data (constants); tid8m1 = 8*threadIdx.x-1.
__device__ __forceinline__ double tc_late(const double *sm, unsigned w, unsigned tid8m1) {
    return sm[(w + (w & 1u) * tid8m1) >> 3];
}

__device__ __forceinline__ void tc_exec(const tc_bundle &B, char *mb, double *sm, unsigned tid, bool pred,
                                        double &acc, double &p) {
    const unsigned tid8m1 = 8u * tid - 1u;
#pragma unroll
    for (int j = 0; j < TC_BUNDLE; j++) {
        const unsigned w0 = B.c[j].x;
        double x = B.x[j], y = B.y[j];
        if (w0 & TC_F_ALATE) x = tc_late(sm, B.c[j].y, tid8m1);
        if (w0 & TC_F_BLATE) y = tc_late(sm, B.c[j].z, tid8m1);
        if ((w0 & 0x7fu) == OP_MADD) {                    // acc = [acc] +- x*y   (a*b+c*d+... :1007-1029)
            double t = __dmul_rn(x, y);
            if (w0 & TC_F_NEGA) t = -t;
            acc = (w0 & TC_F_FIRST) ? t : __dadd_rn(acc, t);
        } else {
            const double2 r = tc_cell_other(w0, acc, p, x, y);
            acc = r.x;
            p = r.y;
        }
        if (w0 & TC_F_END) {                              // warp-uniform
            if (w0 & TC_F_NEG) acc = -acc;                // -(dot)               :1151-1164
            sm[((w0 >> 16) & 0xffu) * TC_BLOCK + tid] = acc;
            if ((w0 & TC_F_STORE) && pred) *(double *)(mb + B.c[j].w) = acc;
        }
    }
}

// Execute dependency group g for the lanes with pred set.  Must be called by all 32 lanes.
__device__ __noinline__ void tc_exec_group1(tc_lane &c, const uint32_t *group_start, int g, bool pred) {
    const uint32_t first = __ldg(group_start + g);
    const uint4 *cells = c.tape + first;
    const int nb = (int)((__ldg(group_start + g + 1) - first) / TC_BUNDLE);
    char *const mb = (char *)c.mb[0];
    double *const sm = c.sm;
    const unsigned tid = threadIdx.x;
    double acc = 0.0, p = 0.0;
    tc_bundle B0, B1;              // double buffer: one bundle executes while the next is in flight
    tc_fetch(B0, cells, mb, pred);
    for (int b = 0; b < nb; b += 2) {
        const bool more1 = b + 1 < nb, more2 = b + 2 < nb;    // warp-uniform
        if (more1) tc_fetch(B1, cells + (size_t)(b + 1) * TC_BUNDLE, mb, pred);
        tc_exec(B0, mb, sm, tid, pred, acc, p);
        if (more1) {
            if (more2) tc_fetch(B0, cells + (size_t)(b + 2) * TC_BUNDLE, mb, pred);
            tc_exec(B1, mb, sm, tid, pred, acc, p);
        }
    }
}

static_assert(TC_TILES == 1, "the tape interpreter runs one instance per lane");
__device__ __forceinline__ void tc_exec_group(tc_lane &c, const uint32_t *group_start, int g, const bool (&pred)[1]) {
    tc_exec_group1(c, group_start, g, pred[0]);
}

#endif  // TC_COMPILED_GROUPS

// ---- lazy group evaluation and set/get/copy wrappers (writeCfunctionpergroup.m:513-852) ------
// io[] word of a get/copy source: scratchbook slot, or constant-pool index when TC_CONST_BIT is set
__device__ __forceinline__ double tc_ldop(const tc_lane &c, int t, uint32_t w, bool pred) {
    const uint32_t idx = w & TC_SLOT_MASK;
    if (w & TC_CONST_BIT) return __ldg(c.consts + idx);                 // warp-uniform branch
    return pred ? c.mb[t][(size_t)idx * TC_SLOT_STRIDE] : 0.0;
}

__device__ __forceinline__ void tc_run_parents(tc_lane &c, const tc_device_image &P, const tc_func *f,
                                               const bool (&pred)[TC_TILES]) {
#pragma unroll
    for (int w = 0; w < TC_GW; w++) {
        uint64_t bits = f->parents[w];          // warp-uniform
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            bool need[TC_TILES], any = false;
            unsigned n = 0;
#pragma unroll
            for (int t = 0; t < TC_TILES; t++) {
                need[t] = pred[t] && !((c.st[t][w] >> b) & 1ull);
                any = any || need[t];
                n += __popc(__ballot_sync(TC_FULL, need[t]));
            }
            if (TC_ANY(any)) {
                tc_exec_group(c, P.group_start, 64 * w + b, need);
                // static per-group flop/byte tables x these counts = algorithmic work (the
                // reference's profiling counters, writeCfunctionpergroup.m:149-177,481-492)
                if (c.lane == 0 && n) {
                    atomicAdd(P.group_count + 64 * w + b, (unsigned long long)n);
                    atomicAdd(P.group_count + TC_MAXGROUPS + 64 * w + b, 1ull);
                }
            }
#pragma unroll
            for (int t = 0; t < TC_TILES; t++)
                if (need[t]) c.st[t][w] |= 1ull << b;
        }
    }
}

__device__ __forceinline__ void tc_invalidate(tc_lane &c, const tc_func *f, const bool (&pred)[TC_TILES]) {
#pragma unroll
    for (int t = 0; t < TC_TILES; t++)
        if (pred[t]) {
#pragma unroll
            for (int w = 0; w < TC_GW; w++) c.st[t][w] &= ~f->children[w];
        }
}

__device__ __forceinline__ void tc_set_scalar(tc_lane &c, const tc_device_image &P, int fn,
                                              const double (&value)[TC_TILES], const bool (&pred)[TC_TILES]) {
    const tc_func *f = P.funcs + fn;
    uint32_t slot = __ldg(P.io + f->io_off);
#pragma unroll
    for (int t = 0; t < TC_TILES; t++)
        if (pred[t]) c.mb[t][(size_t)slot * TC_SLOT_STRIDE] = value[t];
    tc_invalidate(c, f, pred);
}

__device__ __forceinline__ void tc_get(tc_lane &c, const tc_device_image &P, int fn, const bool (&pred)[TC_TILES]) {
    tc_run_parents(c, P, P.funcs + fn, pred);
}

__device__ __forceinline__ void tc_value(const tc_lane &c, const tc_device_image &P, int fn, int k,
                                         const bool (&pred)[TC_TILES], double (&out)[TC_TILES]) {
    uint32_t w = __ldg(P.io + P.funcs[fn].io_off + k);
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) out[t] = (w == TC_NO_SLOT) ? 0.0 : tc_ldop(c, t, w, pred[t]);
}

__device__ __noinline__ void tc_copy(tc_lane &c, const tc_device_image &P, int fn, const bool (&pred)[TC_TILES]) {
    const tc_func *f = P.funcs + fn;
    tc_run_parents(c, P, f, pred);
#ifdef TC_COMPILED_GROUPS
    {
        unsigned mask = 0;
#pragma unroll
        for (int t = 0; t < TC_TILES; t++) mask |= pred[t] ? (1u << t) : 0u;
        bool done = tcm_copy(fn, c.mb[0], mask);
        __syncwarp();
        if (__all_sync(TC_FULL, done)) {
            tc_invalidate(c, f, pred);
            return;
        }
    }
#endif
    const uint32_t *io = P.io + f->io_off;
    const uint32_t cnt = f->io_cnt;
    for (uint32_t k = 0; k < cnt; k += 2) {
        uint32_t d = __ldg(io + k), s = __ldg(io + k + 1);
#pragma unroll
        for (int t = 0; t < TC_TILES; t++) {
            double v = tc_ldop(c, t, s, pred[t]);
            if (pred[t]) c.mb[t][(size_t)d * TC_SLOT_STRIDE] = v;
        }
    }
    tc_invalidate(c, f, pred);
}
