// Solve / lazy-evaluation kernels shared by the two back ends: the tape interpreter built into
// libtenscalc_b200.so and the per-problem compiled modules (TC_COMPILED_GROUPS, emit_cuda.py).
#pragma once
#ifdef TC_DRIVER_LANES
#include "tc_driver_lanes.cuh"   // every lane advances through the solver loop on its own (emitter option driver="lanes")
#else
#include "tc_driver.cuh"
#endif

#ifdef TC_COMPILED_GROUPS
#ifndef TC_LOCKSTEP
#define TC_LOCKSTEP 1      // warps per CTA that walk the code together (see TC_ANY in tc_vm.cuh)
#endif
#define TC_KBLOCK (32 * TC_LOCKSTEP)   // a finished CTA frees its slot for the next tile(s) of instances
#define tc_solve_kernel tcm_solve_kernel   // distinct kernel names per back end (profiles, symbol tables)
#define tc_eval_kernel tcm_eval_kernel
#ifndef TC_KMINBLOCKS
#define TC_KMINBLOCKS (8 / TC_LOCKSTEP)   // 255 registers per thread: 8 warps per SM
#endif
#else
#define TC_KBLOCK TC_BLOCK
#define TC_KMINBLOCKS 4
#endif

// warp w of the grid owns the TC_TILES adjacent tiles w*TC_TILES ..; lane l carries instance
// (w*TC_TILES + t)*32 + l of every tile t
__device__ __forceinline__ int64_t tc_inst(int t) {
    const int64_t warp = ((int64_t)blockIdx.x * TC_KBLOCK + threadIdx.x) >> 5;
    return (warp * TC_TILES + t) * 32 + (threadIdx.x & 31);
}

__device__ __forceinline__ void tc_lane_init(tc_lane &c, const tc_device_image &P) {
    // result ring [slot][thread] (conflict-free), then the constant pool (broadcast reads)
#ifdef TC_COMPILED_GROUPS
    c.sm = nullptr;
#else
    __shared__ double sm[TC_RING_DOUBLES + TC_MAXCONST];
    for (int k = threadIdx.x; k < P.n_consts; k += TC_BLOCK) sm[TC_RING_DOUBLES + k] = P.consts[k];
    __syncthreads();
    c.sm = sm;
#endif
    c.consts = P.consts;
    c.tape = P.tape;
    c.lane = threadIdx.x & 31u;
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) {
        const int64_t inst = tc_inst(t);
        c.mb[t] = &TC_M(P, 0, inst);
#pragma unroll
        for (int w = 0; w < TC_GW; w++) c.st[t][w] = P.group_status[w * P.stride + inst];
    }
}

__device__ __forceinline__ void tc_lane_finish(const tc_lane &c, const tc_device_image &P) {
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) {
        const int64_t inst = tc_inst(t);
#pragma unroll
        for (int w = 0; w < TC_GW; w++) P.group_status[w * P.stride + inst] = c.st[t][w];
    }
}

#ifdef TC_DRIVER_QUEUE
// whole solve of ipmPD_CSsolver() for every instance of the batch by a grid of RESIDENT warps that
// pull instances from a queue (tc_driver_lanes.cuh); one launch per tc_solve
__global__ void __launch_bounds__(TC_KBLOCK, TC_KMINBLOCKS) tc_solve_kernel(tc_device_image P, double mu0, int maxIter,
                                                            double aeU, double aeEq,
                                                            int32_t *__restrict__ status,
                                                            int32_t *__restrict__ iter, int, int) {
    tc_lane c;
    c.sm = nullptr;
    c.consts = P.consts;
    c.tape = P.tape;
    c.lane = threadIdx.x & 31u;
    c.mb[0] = P.m_work + ((size_t)blockIdx.x * (size_t)P.n_slots) * TC_TILE + c.lane;
#pragma unroll
    for (int w = 0; w < TC_GW; w++) c.st[0][w] = 0;
    tc_queue_solve(c, P, mu0, maxIter, aeU, aeEq, status, iter);
}
#else
// ipmPD_CSsolver() for every instance of the batch.  phase_iters == 0: the whole solve in one launch.  Otherwise
// one PHASE of a phased solve (tc_driver.cuh, tc_abi.cu): at most phase_iters Newton iterations for the
// instances in the columns [0, *P.n_live); results of the instances that finish go to status/iter at their
// ORIGINAL index (P.orig: compaction moves instances between columns)
__global__ void __launch_bounds__(TC_KBLOCK, TC_KMINBLOCKS) tc_solve_kernel(tc_device_image P, double mu0, int maxIter,
                                                            double aeU, double aeEq,
                                                            int32_t *__restrict__ status,
                                                            int32_t *__restrict__ iter, int phase_first,
                                                            int phase_iters) {
    const int64_t limit = phase_iters ? (int64_t)*P.n_live : P.batch;
    if ((((int64_t)blockIdx.x * TC_KBLOCK) * TC_TILES) >= limit) return;     // no unfinished instance in this CTA's columns
    tc_lane c;
    tc_lane_init(c, P);
    bool valid[TC_TILES], done[TC_TILES];
    int st[TC_TILES], it[TC_TILES];
    int64_t pos[TC_TILES];
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) {
        pos[t] = tc_inst(t);
        valid[t] = pos[t] < limit;
    }
#if defined(TC_DRIVER_LANES)
    tc_ipm_solve(c, P, valid, mu0, maxIter, aeU, aeEq, st, it);
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) done[t] = valid[t];
#else
    tc_ipm_solve(c, P, valid, mu0, maxIter, aeU, aeEq, st, it, done, pos, phase_first, phase_iters);
#endif
    tc_lane_finish(c, P);
#pragma unroll
    for (int t = 0; t < TC_TILES; t++)
        if (done[t]) {
            const int64_t o = phase_iters ? (int64_t)P.orig[pos[t]] : pos[t];
            status[o] = st[t];
            iter[o] = it[t];
        }
}

#endif  // TC_DRIVER_QUEUE

// lazy evaluation of the parents of one get function (classname_getOutputs)
__global__ void __launch_bounds__(TC_KBLOCK, TC_KMINBLOCKS) tc_eval_kernel(tc_device_image P, int fn) {
    tc_lane c;
    tc_lane_init(c, P);
    bool valid[TC_TILES];
#pragma unroll
    for (int t = 0; t < TC_TILES; t++) valid[t] = tc_inst(t) < P.batch;
    tc_get(c, P, fn, valid);
    tc_lane_finish(c, P);
}
