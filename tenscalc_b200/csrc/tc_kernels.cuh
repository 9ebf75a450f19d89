// Solve / lazy-evaluation kernels shared by the two back ends: the tape interpreter built into
// libtenscalc_b200.so and the per-problem compiled modules (TC_COMPILED_GROUPS, emit_cuda.py).
#pragma once
#include "tc_driver.cuh"

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

__device__ __forceinline__ void tc_lane_init(tc_lane &c, const tc_device_image &P, int64_t inst) {
    // result ring [slot][thread] (conflict-free), then the constant pool (broadcast reads)
#ifdef TC_COMPILED_GROUPS
    c.sm = nullptr;
#else
    __shared__ double sm[TC_RING_DOUBLES + TC_MAXCONST];
    for (int k = threadIdx.x; k < P.n_consts; k += TC_BLOCK) sm[TC_RING_DOUBLES + k] = P.consts[k];
    __syncthreads();
    c.sm = sm;
#endif
    c.mb = &TC_M(P, 0, inst);
    c.consts = P.consts;
    c.tape = P.tape;
    c.lane = threadIdx.x & 31u;
    c.st[0] = P.group_status[inst];
    c.st[1] = P.group_status[P.stride + inst];
}

__device__ __forceinline__ void tc_lane_finish(const tc_lane &c, const tc_device_image &P, int64_t inst) {
    P.group_status[inst] = c.st[0];
    P.group_status[P.stride + inst] = c.st[1];
}

// whole solve of ipmPD_CSsolver() for every instance of the batch; one launch per tc_solve
__global__ void __launch_bounds__(TC_KBLOCK, TC_KMINBLOCKS) tc_solve_kernel(tc_device_image P, double mu0, int maxIter,
                                                            double aeU, double aeEq,
                                                            int32_t *__restrict__ status,
                                                            int32_t *__restrict__ iter) {
    const int64_t inst = (int64_t)blockIdx.x * TC_KBLOCK + threadIdx.x;
    const bool valid = inst < P.batch;
    tc_lane c;
    tc_lane_init(c, P, inst);
    int st, it;
    tc_ipm_solve(c, P, valid, mu0, maxIter, aeU, aeEq, st, it);
    tc_lane_finish(c, P, inst);
    if (valid) {
        status[inst] = st;
        iter[inst] = it;
    }
}

// lazy evaluation of the parents of one get function (classname_getOutputs)
__global__ void __launch_bounds__(TC_KBLOCK, TC_KMINBLOCKS) tc_eval_kernel(tc_device_image P, int fn) {
    const int64_t inst = (int64_t)blockIdx.x * TC_KBLOCK + threadIdx.x;
    tc_lane c;
    tc_lane_init(c, P, inst);
    tc_get(c, P, fn, inst < P.batch);
    tc_lane_finish(c, P, inst);
}

