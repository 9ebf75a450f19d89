// Tape interpreter: straight-line code of one dependency group, one instance per lane.
//
// Replaces the per-problem C functions `classname_groupK()` that the reference emits with
// writeCinstructionsC (lib/csparse/instructionsTableUTHash.c:950-1539).  The 32 lanes of a
// warp walk the SAME tape (so there is no divergence inside a group) and each lane applies
// it to its own instance.  The scratchbook is instance-interleaved, m[slot][instance], so a
// warp-wide operand load is one contiguous 256-byte segment.
//
// Arithmetic contract (parity with the CPU oracle): every instruction is evaluated in the
// C expression order of the reference emitter, with separately rounded multiplies and adds
// (__dmul_rn/__dadd_rn are never contracted into DFMA), IEEE division and square root.
#pragma once
#include <cfloat>
#include <cmath>
#include <cuda_runtime.h>

#include "tc_image.h"

#define TC_FULL 0xFFFFFFFFu
#define TC_CHUNK 8

struct tc_lane {
    double *mb;          // &m[0][instance]
    int64_t stride;
    const double *consts;
    const uint32_t *tape;
    unsigned lane;
    uint64_t st[2];      // groupStatus[] of this instance, one bit per group
};

// The tape is fetched cooperatively: 32 consecutive words per warp load, then broadcast
// word by word with a shuffle -- one LSU wavefront per 32 tape words instead of one per word.
struct tc_tape_reader {
    const uint32_t *base;
    uint32_t pc, win;
    unsigned lane;
    __device__ __forceinline__ void seek(uint32_t p) {
        pc = p;
        win = __ldg(base + (p & ~31u) + lane);
    }
    __device__ __forceinline__ uint32_t next() {
        uint32_t w = __shfl_sync(TC_FULL, win, pc & 31u);
        pc++;
        if ((pc & 31u) == 0) win = __ldg(base + pc + lane);
        return w;
    }
};

__device__ __forceinline__ double tc_ldop(const tc_lane &c, uint32_t w, bool pred) {
    uint32_t idx = w & TC_SLOT_MASK;
    double v;
    if (w & TC_CONST_BIT) v = __ldg(c.consts + idx);                 // warp-uniform branch
    else v = pred ? c.mb[(size_t)idx * (size_t)c.stride] : 0.0;
    return (w & TC_NEG_BIT) ? -v : v;
}

template <int N>
__device__ __forceinline__ void tc_load_chunk(const tc_lane &c, tc_tape_reader &tr, double (&v)[N], int cnt,
                                              bool pred) {
    uint32_t w[N];
#pragma unroll
    for (int j = 0; j < N; j++)
        if (j < cnt) w[j] = tr.next();
#pragma unroll
    for (int j = 0; j < N; j++)
        if (j < cnt) v[j] = tc_ldop(c, w[j], pred);
}

// sum_k x_{2k}*x_{2k+1}, left to right, over `npairs` pairs read from the tape
__device__ __forceinline__ double tc_dot_pairs(const tc_lane &c, tc_tape_reader &tr, int npairs, bool pred) {
    double acc = 0.0;
    bool first = true;
    for (int k = 0; k < npairs; k += TC_CHUNK / 2) {
        int cnt = min(TC_CHUNK / 2, npairs - k);
        double v[TC_CHUNK];
        tc_load_chunk<TC_CHUNK>(c, tr, v, 2 * cnt, pred);
#pragma unroll
        for (int j = 0; j < TC_CHUNK / 2; j++)
            if (j < cnt) {
                double p = __dmul_rn(v[2 * j], v[2 * j + 1]);
                acc = first ? p : __dadd_rn(acc, p);
                first = false;
            }
    }
    return acc;
}

template <int M>
__device__ __forceinline__ double tc_sumprod_m(const tc_lane &c, tc_tape_reader &tr, int nprod, bool pred) {
    double acc = 0.0;
    bool first = true;
    for (int k = 0; k < nprod; k++) {
        double v[M];
        tc_load_chunk<M>(c, tr, v, M, pred);
        double p = v[0];
#pragma unroll
        for (int j = 1; j < M; j++) p = __dmul_rn(p, v[j]);
        acc = first ? p : __dadd_rn(acc, p);
        first = false;
    }
    return acc;
}

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
    case I_Dsqrt: return __dmul_rn(.5, pow(x, -.5));
    case I_DDsqrt: return __dmul_rn(-.25, pow(x, -1.5));
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
    return x;
}

// Execute dependency group g for the lanes with pred set.  Must be called by all 32 lanes.
__device__ __noinline__ void tc_exec_group(tc_lane &c, const uint32_t *group_start, int g, bool pred) {
    tc_tape_reader tr;
    tr.base = c.tape;
    tr.lane = c.lane;
    tr.seek(__ldg(group_start + g));
    const uint32_t end = __ldg(group_start + g + 1);
    while (tr.pc < end) {
        const uint32_t hdr = tr.next();
        const int op = hdr & 0xff;
        int n = hdr >> 8;
        const uint32_t r = tr.next();
        double acc = 0.0;
        switch (op) {
        case I_sum: {                                       // +a-b+c...       :964-968
            bool first = true;
            for (int k = 0; k < n; k += TC_CHUNK) {
                int cnt = min(TC_CHUNK, n - k);
                double v[TC_CHUNK];
                tc_load_chunk<TC_CHUNK>(c, tr, v, cnt, pred);
#pragma unroll
                for (int j = 0; j < TC_CHUNK; j++)
                    if (j < cnt) {
                        acc = first ? v[j] : __dadd_rn(acc, v[j]);
                        first = false;
                    }
            }
        } break;
        case I_sumprod: {                                   // a*b+c*d+...     :1007-1029
            const int m = (int)tr.next();
            if (m == 2) acc = tc_dot_pairs(c, tr, n / 2, pred);
            else if (m == 3) acc = tc_sumprod_m<3>(c, tr, n / 3, pred);
            else if (m == 4) acc = tc_sumprod_m<4>(c, tr, n / 4, pred);
            else {
                bool first = true;
                for (int k = 0; k < n; k += m) {
                    double p = tc_ldop(c, tr.next(), pred);
                    for (int j = 1; j < m; j++) p = __dmul_rn(p, tc_ldop(c, tr.next(), pred));
                    acc = first ? p : __dadd_rn(acc, p);
                    first = false;
                }
            }
        } break;
        case I_plus_minus_dot_div: {                        // (a0-(..))/d     :1081-1098
            double a0 = tc_ldop(c, tr.next(), pred);
            double d = tc_ldop(c, tr.next(), pred);
            acc = __dadd_rn(a0, -tc_dot_pairs(c, tr, (n - 2) / 2, pred)) / d;
        } break;
        case I_minus_dot_div: {                             // -(..)/d         :1106-1121
            double d = tc_ldop(c, tr.next(), pred);
            acc = (-tc_dot_pairs(c, tr, (n - 1) / 2, pred)) / d;
        } break;
        case I_plus_minus_dot: {                            // a0-(..)         :1128-1144
            double a0 = tc_ldop(c, tr.next(), pred);
            acc = __dadd_rn(a0, -tc_dot_pairs(c, tr, (n - 1) / 2, pred));
        } break;
        case I_minus_dot:                                   // -(..)           :1151-1164
            acc = -tc_dot_pairs(c, tr, n / 2, pred);
            break;
        case I_div: {
            double a = tc_ldop(c, tr.next(), pred);
            double b = tc_ldop(c, tr.next(), pred);
            acc = a / b;
        } break;
        case I_plus_sqr:                                    // a*a+b*b+...     :1201-1210
        case I_plus_abs:                                    // |a|+|b|+...     :1217-1225
        case I_min:
        case I_max:
        case I_min0:
        case I_max0:
        case I_max_abs: {                                   // sequential compare chains :1228-1313
            bool first = !(op == I_min0 || op == I_max0);   // the ...0 forms start from 0
            for (int k = 0; k < n; k += TC_CHUNK) {
                int cnt = min(TC_CHUNK, n - k);
                double v[TC_CHUNK];
                tc_load_chunk<TC_CHUNK>(c, tr, v, cnt, pred);
#pragma unroll
                for (int j = 0; j < TC_CHUNK; j++)
                    if (j < cnt) {
                        double x = v[j];
                        if (op == I_plus_sqr) x = __dmul_rn(x, x);
                        else if (op == I_plus_abs || op == I_max_abs) x = fabs(x);
                        if (first) acc = x;
                        else if (op == I_plus_sqr || op == I_plus_abs) acc = __dadd_rn(acc, x);
                        else if (op == I_min || op == I_min0) { if (acc > x) acc = x; }
                        else { if (acc < x) acc = x; }      // NaN in a non-first operand is dropped
                        first = false;
                    }
            }
        } break;
        case I_clp: {                                       // :1315-1331
            acc = DBL_MAX;
            for (int k = 0; k < n; k += TC_CHUNK) {
                int cnt = min(TC_CHUNK, n - k);
                double v[TC_CHUNK];
                tc_load_chunk<TC_CHUNK>(c, tr, v, cnt, pred);
#pragma unroll
                for (int j = 0; j < TC_CHUNK; j += 2)
                    if (j < cnt) {
                        if (v[j + 1] < 0) {
                            double x = (-v[j]) / v[j + 1];
                            if (acc > x) acc = x;
                        }
                    }
            }
        } break;
        default:
            acc = tc_unary(op, tc_ldop(c, tr.next(), pred));
        }
        if (pred) c.mb[(size_t)r * (size_t)c.stride] = acc;
    }
}

// ---- lazy group evaluation and set/get/copy wrappers (writeCfunctionpergroup.m:513-852) ------
__device__ __forceinline__ void tc_run_parents(tc_lane &c, const tc_device_image &P, const tc_func *f, bool pred) {
#pragma unroll
    for (int w = 0; w < 2; w++) {
        uint64_t bits = f->parents[w];          // warp-uniform
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            bool need = pred && !((c.st[w] >> b) & 1ull);
            const unsigned vote = __ballot_sync(TC_FULL, need);
            if (vote) {
                tc_exec_group(c, P.group_start, 64 * w + b, need);
                // static per-group flop/byte tables x these counts = algorithmic work (the
                // reference's profiling counters, writeCfunctionpergroup.m:149-177,481-492)
                if (c.lane == 0) atomicAdd(P.group_count + 64 * w + b, (unsigned long long)__popc(vote));
            }
            if (need) c.st[w] |= 1ull << b;
        }
    }
}

__device__ __forceinline__ void tc_invalidate(tc_lane &c, const tc_func *f, bool pred) {
    if (pred) {
        c.st[0] &= ~f->children[0];
        c.st[1] &= ~f->children[1];
    }
}

__device__ __forceinline__ void tc_set_scalar(tc_lane &c, const tc_device_image &P, int fn, double value, bool pred) {
    const tc_func *f = P.funcs + fn;
    uint32_t slot = __ldg(P.io + f->io_off);
    if (pred) c.mb[(size_t)slot * (size_t)c.stride] = value;
    tc_invalidate(c, f, pred);
}

__device__ __forceinline__ void tc_get(tc_lane &c, const tc_device_image &P, int fn, bool pred) {
    tc_run_parents(c, P, P.funcs + fn, pred);
}

__device__ __forceinline__ double tc_value(const tc_lane &c, const tc_device_image &P, int fn, int k, bool pred) {
    uint32_t w = __ldg(P.io + P.funcs[fn].io_off + k);
    return (w == TC_NO_SLOT) ? 0.0 : tc_ldop(c, w, pred);
}

__device__ __noinline__ void tc_copy(tc_lane &c, const tc_device_image &P, int fn, bool pred) {
    const tc_func *f = P.funcs + fn;
    tc_run_parents(c, P, f, pred);
    const uint32_t *io = P.io + f->io_off;
    const uint32_t cnt = f->io_cnt;
    for (uint32_t k = 0; k < cnt; k += 2) {
        uint32_t d = __ldg(io + k), s = __ldg(io + k + 1);
        double v = tc_ldop(c, s, pred);
        if (pred) c.mb[(size_t)d * (size_t)c.stride] = v;
    }
    tc_invalidate(c, f, pred);
}
