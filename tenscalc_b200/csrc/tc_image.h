// Problem image shared by the host loader and the device kernels.
// Written by tenscalc_b200/tape.py (serialize); see that file for the tape encoding.
#pragma once
#include <stdint.h>

#define TC_MAGIC 0x3030324243540001LL
#define TC_CONST_BIT 0x40000000u
#define TC_NEG_BIT 0x80000000u
#define TC_SLOT_MASK 0x3FFFFFFFu
#define TC_NO_SLOT 0xFFFFFFFFu

// driver function ids -- MUST match FN in tenscalc_b200/tape.py.  Names are the reference's
// (lib/ipmPD_CStimesLambda.m declares them, lib/ipmPD_CSsolver.c:28-51 calls them).
enum tc_fn {
    FN_initPrimal = 0, FN_scaleIneq, FN_scaleCost, FN_getScale4Cost, FN_setMu, FN_initDualIneq,
    FN_initDualEqX, FN_setAddEye2HessianU, FN_setAddEye2HessianEq, FN_getNorminf_Grad,
    FN_getNorminf_G, FN_getGapMinFMinLambda, FN_getCurvature, FN_getDirectionError,
    FN_getHessInertia, FN_getMaxAlphas_s, FN_setAlphaPrimal, FN_getMinF_s, FN_setAlphaDualEq,
    FN_setAlphaDualIneq, FN_updatePrimalDual, FN_getJ, FN_getOutputs, FN_USER0
};

// opcodes = reference enum (lib/csparse/instructionsTableTypes.h:2-56)
enum tc_op {
    I_set = 1, I_load, I_sum, I_sumprod, I_inv, I_minus_inv_sqr, I_2_inv_cube, I_div, I_minus_dot,
    I_minus_dot_div, I_plus_minus_dot, I_plus_sqr, I_plus_abs, I_plus_minus_dot_div, I_min, I_min0,
    I_max, I_max0, I_max_abs, I_clp, I_exp, I_log, I_cos, I_minus_cos, I_sin, I_minus_sin, I_round,
    I_ceil, I_floor, I_abs, I_sign, I_sqrt, I_Dsqrt, I_DDsqrt, I_sqr, I_2times, I_2, I_cube, I_3sqr,
    I_6times, I_atan, I_Datan, I_DDatan, I_srelu, I_dsrelu, I_ddsrelu, I_relu, I_heaviside
};

struct tc_func {            // one set / get / copy function
    uint64_t parents[2];    // dependency groups to (re)evaluate before reading
    uint64_t children[2];   // dependency groups invalidated by writing
    uint32_t io_off;        // into io[]: set = dst slots, get = src words, copy = (dst,src) pairs
    uint32_t io_cnt;
    uint64_t kind;
};

struct tc_defines {         // the #define table of lib/cmex2optimizeCS.m:303-331
    int32_t nU, nG, nF;
    int32_t scaleCost, scaleInequalities, coupledAlphas, setAddEye2Hessian, adjustAddEye2Hessian,
        useInertia, skipAffine, delta;
    double gradTolerance, addEye2HessianUtolerance, equalTolerance, desiredDualityGap, alphaMin,
        alphaMax, muFactorAggressive, muFactorConservative;
};

struct tc_device_image {    // device pointers, passed by value to the kernels
    const uint32_t *tape;
    const uint32_t *group_start;
    const double *consts;
    const tc_func *funcs;
    const uint32_t *io;
    double *m;              // scratchbook, [n_slots][stride], instance-interleaved
    uint64_t *group_status; // [2][stride] lazy-evaluation flags per instance
    unsigned long long *group_count; // [n_groups] instance-executions of every group (profiling)
    int64_t stride;         // padded batch (multiple of 32)
    int64_t batch;
    int32_t n_groups;
    tc_defines def;
};
