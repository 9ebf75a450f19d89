// Problem image shared by the host loader and the device kernels.
// Written by tenscalc_b200/tape.py (serialize); see that file for the tape encoding.
#pragma once
#include <stdint.h>
#include <vector_types.h>

#define TC_MAGIC 0x3030324243540006LL
#define TC_CONST_BIT 0x40000000u   // io[] words (get/copy sources): constant-pool index
#define TC_SLOT_MASK 0x3FFFFFFFu
#define TC_NO_SLOT 0xFFFFFFFFu
#define TC_TILE 32                 // instances per scratchbook tile = lanes per warp
#define TC_BLOCK 128               // threads per CTA = instances per CTA -- MUST match tape.py BLOCK
#define TC_BUNDLE 4                // cells fetched together               -- MUST match tape.py BUNDLE
#define TC_RING 32                 // results in the per-lane ring         -- MUST match tape.py RING
#define TC_MAXCONST 1024           // constant pool in shared memory       -- MUST match tape.py MAX_CONSTS
#define TC_RING_DOUBLES (TC_RING * TC_BLOCK)
#define TC_GW 4                    // 64-bit words of a dependency-group mask                -- MUST match tape.py GW
#define TC_MAXGROUPS (64 * TC_GW)  // dependency groups per problem (groupStatus[] of the reference, one bit each)

// tape cell = uint4 {op|flags|ring slot<<16, operand A, operand B, result offset}; see tape.py
#define TC_F_ALATE (1u << 7)
#define TC_F_BLATE (1u << 8)
#define TC_F_STORE (1u << 9)
#define TC_F_END (1u << 10)
#define TC_F_NEG (1u << 11)
#define TC_F_NEGA (1u << 12)
#define TC_F_NEGB (1u << 13)
#define TC_F_INIT (1u << 14)
#define TC_F_FIRST (1u << 15)

enum tc_cellop {   // MUST match the OP_* list of tenscalc_b200/tape.py
    OP_MADD = 0, OP_SUM1F, OP_SUM2F, OP_SUM1, OP_SUM2, OP_NOP, OP_RESERVED, OP_PSET, OP_PMUL1, OP_PMUL2,
    OP_P1SET, OP_P1ADD, OP_P2SET, OP_P2ADD, OP_RSUB, OP_RSUBDIV, OP_NEGDIV, OP_DIV, OP_SQ1F, OP_SQ2F,
    OP_SQ1, OP_SQ2, OP_ABS1F, OP_ABS2F, OP_ABS1, OP_ABS2, OP_MINF1, OP_MINF2, OP_MIN1, OP_MIN2,
    OP_MAXF1, OP_MAXF2, OP_MAX1, OP_MAX2, OP_AMAXF1, OP_AMAXF2, OP_AMAX1, OP_AMAX2, OP_CLP, OP_PACC,
    OP_UNARY = 64   // + reference opcode of the unary function
};

// driver function ids -- MUST match FN in tenscalc_b200/tape.py.  Names are the reference's
// (lib/ipmPD_CStimesLambda.m declares them, lib/ipmPD_CSsolver.c:28-51 calls them).
enum tc_fn {
    FN_initPrimal = 0, FN_scaleIneq, FN_scaleCost, FN_getScale4Cost, FN_setMu, FN_initDualIneq,
    FN_initDualEqX, FN_setAddEye2HessianU, FN_setAddEye2HessianEq, FN_getNorminf_Grad,
    FN_getNorminf_G, FN_getGapMinFMinLambda, FN_getCurvature, FN_getDirectionError,
    FN_getHessInertia, FN_getMaxAlphas_s, FN_setAlphaPrimal, FN_getMinF_s, FN_setAlphaDualEq,
    FN_setAlphaDualIneq, FN_updatePrimalDual, FN_getJ, FN_getOutputs,
    FN_getMaxAlphas_a, FN_getMinF_a, FN_getRho,      // affine-scaling (Mehrotra) step, skipAffine=0 (ipmPD_CS.m:313-315,422-425)
    FN_setAddEye2HessianD, FN_getHessDinertia, FN_getHessUinertia,   // min-max driver (ipmPDminmax_CS.m:181-192,266-272)
    FN_saveWW,               // allowSave: saveWW__ (ipmPD_CSsolver.c:51) as a copy of the KKT values into save variables
    FN_USER0
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
    uint64_t parents[TC_GW];    // dependency groups to (re)evaluate before reading
    uint64_t children[TC_GW];   // dependency groups invalidated by writing
    uint32_t io_off;        // into io[]: set = dst slots, get = src words, copy = (dst,src) pairs
    uint32_t io_cnt;
    uint64_t kind;
};

struct tc_defines {         // the #define table of lib/cmex2optimizeCS.m:303-331
    int32_t nU, nG, nF;
    int32_t scaleCost, scaleInequalities, coupledAlphas, setAddEye2Hessian, adjustAddEye2Hessian,
        useInertia, skipAffine, delta;
    int32_t driver;         // 0: ipmPD_CSsolver (lib/ipmPD_CSsolver.c), 1: ipmPDeq_CSsolver (lib/ipmPDeq_CSsolver.c),
                            // 2: ipmPDminmax_CSsolver (lib/ipmPDminmax_CSsolver.m)
    int32_t mpUdesired, mnDdesired;         // min-max: wanted inertia (ipmPDminmax_CSsolver.m:68-69)
    int32_t allowSave;                      // the solver was generated with allowSave (cmex2optimizeCS.m:327)
    double addEye2HessianDtolerance;
    double gradTolerance, addEye2HessianUtolerance, equalTolerance, desiredDualityGap, alphaMin,
        alphaMax, muFactorAggressive, muFactorConservative;
};

struct tc_device_image {    // device pointers, passed by value to the kernels
    const uint4 *tape;            // cells
    const uint32_t *group_start;  // first cell of every group (multiples of TC_BUNDLE)
    const double *consts;
    const tc_func *funcs;
    const uint32_t *io;
    double *m;              // scratchbook, [stride/32][n_slots][32]: 32 consecutive instances (one warp)
                            // share one contiguous n_slots*256-byte block, lane-interleaved, so a warp-wide
                            // operand access is one 256-byte segment and a warp's whole working set sits
                            // in a handful of 2 MB pages (TLB reach) instead of being strided over the batch
    int64_t n_slots;
    uint64_t *group_status; // [TC_GW][stride] lazy-evaluation flags per scratchbook column
    unsigned long long *group_count; // [2][TC_MAXGROUPS] instance- and warp-executions of every group (profiling)
    int64_t stride;         // padded batch (multiple of 32)
    int64_t batch;
    int32_t n_groups;
    int32_t n_consts;
    tc_defines def;
    // work-queue driver (emitter option driver="queue"; NULL/0 otherwise)
    double *m_work;               // [work_tiles][n_slots][32]: one column per RESIDENT lane
    unsigned long long *queue;    // next instance to hand out (reset before every solve)
    int64_t work_tiles;
    int32_t n_funcs;              // funcs[FN_USER0 .. n_funcs) are the user set functions
    int32_t phased;               // 1: the solve is a sequence of phase launches with compaction in between (tc_abi.cu)
    // phased execution (compiled back end, warp driver): the solver loop's locals of every scratchbook
    // column between two phase launches (tc_driver.cuh), the instance each column currently holds, and
    // the number of columns [0, n_live) that hold unfinished instances
    double *ps_d;                 // [TC_PS_ND][stride]
    int32_t *ps_i;                // [TC_PS_NI][stride]
    int32_t *orig;                // [stride] original instance whose state lives in the column, -1: none
    int32_t *n_live;              // device scalar
    double aeD0;                  // min-max driver: initial addEye2HessianD of this solve (addEye2Hessian[1])
    // log2 of the instances a lane carries INTERLEAVED (emitter option pair): 0: m[tile][slot][lane];
    // 1: m[tile pair][slot][lane][2] -- the two instances of a lane sit side by side, so one 128-bit access
    // moves slot s of both (set by tc_attach_module from the module, before anything is written)
    int32_t ilv;
    // 1: tiles of 64 instances, m[tile][slot][64] (emitter option wide): the two warps of a lockstep CTA that sit on
    // neighbouring 32-instance blocks touch the two 256-byte halves of one 512-byte segment at about the same time
    int32_t wide;
    // saveIter of this solve (ipmPD_CSsolver.c:136): with allowSave, the Newton iteration whose KKT values are kept
    // (0: the values at a successful exit; status 4 keeps them too, :330-332,700-702,877-882)
    int32_t save_iter;
    int32_t pad_;
};
#define TC_PS_ND 12
#define TC_PS_NI 2

// address of scratchbook slot `slot` of instance `inst`
#define TC_M(P, slot, inst)                                                                                        \
    ((P).m[(((size_t)((inst) >> (5 + (P).ilv + (P).wide)) * (size_t)(P).n_slots + (size_t)(slot)) * TC_TILE            \
            << ((P).ilv + (P).wide)) +                                                                                 \
           ((P).ilv ? ((size_t)((inst) & 31) << 1) + (((inst) >> 5) & 1) : (size_t)((inst) & ((32 << (P).wide) - 1)))])
