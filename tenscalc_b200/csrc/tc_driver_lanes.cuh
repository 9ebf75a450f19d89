// Device-resident primal-dual interior-point driver, LANE-DECOUPLED form (emitter option
// driver="lanes").
//
// Same control flow as tc_driver.cuh (lib/ipmPD_CSsolver.c:132-920 for the option set
// cmex2optimizeCS can emit), but every lane runs the solver loop of its own instance as a
// resumable coroutine: the sequential code below is the reference's loop almost verbatim, and
// every call of a generated get/copy function is a suspension point (TC_AWAIT) at which the lane
// states which function it needs.  The warp then repeatedly
//   1. lets every lane name the first dependency group it still misses for that function,
//   2. resumes the lanes whose function is complete,
//   3. otherwise runs one group: the one the lane furthest behind waits for -- except that the
//      dominant group (the KKT factorisation) only runs once every living lane waits for it.
// Lanes therefore drift apart in solver phase and meet again at the expensive groups: a lane that
// must re-factorise the KKT matrix within an iteration (negative curvature: the regulariser
// is raised at once, :458-520) waits for the rest of the warp to arrive at ITS next
// factorisation instead of making the whole warp walk the LDL code a second time for one lane.
// Per instance the sequence of operations -- and so every rounding -- is unchanged.
#pragma once
#include "tc_vm.cuh"

#if TC_TILES != 1
#error "the lane-decoupled driver carries one instance per lane"
#endif

#define TC_addEye2HessianUMAX 1e2
#define TC_addEye2HessianEqMAX 1e2
#define TC_addEye2HessianMIN 1e-20
#define TC_maxDirectionError 1e-6
#define TC_stepback .99
#define TC_MIN(A, B) (((A) < (B)) ? (A) : (B))
#define TC_MAX(A, B) (((A) > (B)) ? (A) : (B))

enum { TC_WANT_NONE = 0, TC_WANT_GET = 1, TC_WANT_COPY = 2 };

struct tc_lane_state {          // everything the reference keeps in locals of ipmPD_CSsolver()
    int pc, want_fn, want_kind;
    int status, iter;
    bool updU, updEq, th_grad, th_eq;
    double aeU, aeEq, mu, muMin, gapTol, alphaMax_, alphaPrimal, alphaDualEq, alphaDualIneq;
    double norminf_grad, norminf_eq, gap, ineq, ineq1, dual, curvature, derr, mp, mn;
};

#define TC_AWAIT(FN, KIND)           \
    do {                             \
        s.want_fn = (FN);            \
        s.want_kind = (KIND);        \
        s.pc = __LINE__;             \
        return;                      \
    case __LINE__:;                  \
    } while (0)
#define TC_GET(FN) TC_AWAIT(FN, TC_WANT_GET)
#define TC_COPY(FN) TC_AWAIT(FN, TC_WANT_COPY)
#define TC_VAL(FN, K) tc_lane_value(c, P, FN, K)
#define TC_SET(FN, V) tc_lane_set(c, P, FN, V)

__device__ __forceinline__ double tc_lane_value(const tc_lane &c, const tc_device_image &P, int fn, int k) {
    uint32_t w = __ldg(P.io + P.funcs[fn].io_off + k);
    if (w == TC_NO_SLOT) return 0.0;
    const uint32_t idx = w & TC_SLOT_MASK;
    return (w & TC_CONST_BIT) ? __ldg(c.consts + idx) : c.mb[0][(size_t)idx * TC_TILE];
}

__device__ __forceinline__ void tc_lane_set(tc_lane &c, const tc_device_image &P, int fn, double value) {
    const tc_func *f = P.funcs + fn;
    c.mb[0][(size_t)__ldg(P.io + f->io_off) * TC_TILE] = value;
#pragma unroll
    for (int w = 0; w < TC_GW; w++) c.st[0][w] &= ~f->children[w];
}

// Runs the instance's solver loop until its next get/copy (s.want_*), or to the end (s.pc = -1).
__device__ __noinline__ void tc_lane_resume(tc_lane_state &s, tc_lane &c, const tc_device_image &P, double mu0,
                                            int maxIter) {
    const tc_defines &D = P.def;
    const bool hasF = D.nF > 0, hasG = D.nG > 0;
    const bool adjust = D.setAddEye2Hessian && D.adjustAddEye2Hessian;
    switch (s.pc) {
    case 0:
        if (D.setAddEye2Hessian) {                                          // :150-155
            TC_SET(FN_setAddEye2HessianU, s.aeU);
            TC_SET(FN_setAddEye2HessianEq, s.aeEq);
        }
        TC_COPY(FN_initPrimal);                                             // :204
        if (D.scaleCost) {                                                  // :206-214
            TC_COPY(FN_scaleCost);
            TC_GET(FN_getScale4Cost);
            if (hasF) s.gapTol *= TC_VAL(FN_getScale4Cost, 0);
        }
        if (hasF) {                                                         // :216-230
            if (D.scaleInequalities) TC_COPY(FN_scaleIneq);
            TC_SET(FN_setMu, s.mu);
            s.muMin = s.gapTol / D.nF / 2;
            TC_COPY(FN_initDualIneq);
        }
        if (hasG) TC_COPY(FN_initDualEqX);                                  // :232-235

        while (true) {                                                      // :282
            s.iter++;
            if (s.iter > maxIter) { s.status = 8; break; }                  // :308-313

            // ---- exit conditions (:323-381) ----
            TC_GET(FN_getNorminf_Grad);
            s.norminf_grad = TC_VAL(FN_getNorminf_Grad, 0);
            if (isnan(s.norminf_grad)) { s.status = 4; break; }
            if (hasG) {
                TC_GET(FN_getNorminf_G);
                s.norminf_eq = TC_VAL(FN_getNorminf_G, 0);
            }
            if (hasF) {
                TC_GET(FN_getGapMinFMinLambda);
                s.gap = TC_VAL(FN_getGapMinFMinLambda, 0);
                s.ineq = TC_VAL(FN_getGapMinFMinLambda, 1);
                s.dual = TC_VAL(FN_getGapMinFMinLambda, 2);
                if (s.ineq <= 0) { s.status = 1; break; }
                if (s.dual <= 0) { s.status = 2; break; }
            }
            if (s.norminf_grad <= D.gradTolerance && (!hasF || s.gap <= s.gapTol) &&
                (!hasG || s.norminf_eq <= D.equalTolerance) && (!adjust || s.aeU <= D.addEye2HessianUtolerance)) {
                s.status = 0;
                break;
            }

            // ---- regularisation control (:393-548) ----
            if (adjust) {
                if (s.updU) { TC_SET(FN_setAddEye2HessianU, s.aeU); s.updU = false; }      // delayed
                if (s.updEq) { TC_SET(FN_setAddEye2HessianEq, s.aeEq); s.updEq = false; }
                TC_GET(FN_getCurvature);
                s.curvature = TC_VAL(FN_getCurvature, 0);
                if (D.useInertia && s.curvature <= 0) {                     // :408-413
                    TC_GET(FN_getHessInertia);
                    s.mp = TC_VAL(FN_getHessInertia, 0);
                    s.mn = TC_VAL(FN_getHessInertia, 1);
                }
                TC_GET(FN_getDirectionError);
                s.derr = TC_VAL(FN_getDirectionError, 0);
                if (s.curvature > 0 || (D.useInertia && s.mp == D.nU && s.mn == D.nF + D.nG)) {   // :415-456
                    if (s.aeU > TC_addEye2HessianMIN && s.derr < TC_maxDirectionError) {
                        s.aeU = TC_MAX(.75 * s.aeU, TC_addEye2HessianMIN);
                        s.updU = true;
                    }
                    if (s.aeEq > TC_addEye2HessianMIN && s.derr < TC_maxDirectionError) {
                        s.aeEq = TC_MAX(.75 * s.aeEq, TC_addEye2HessianMIN);
                        s.updEq = true;
                    }
                    if (s.aeU < TC_addEye2HessianUMAX && s.derr > TC_maxDirectionError) {
                        s.aeU = TC_MIN(10 * s.aeU, TC_addEye2HessianUMAX);
                        s.updU = true;
                    }
                    if (s.aeEq < TC_addEye2HessianEqMAX && s.derr > TC_maxDirectionError) {
                        s.aeEq = TC_MIN(10 * s.aeEq, TC_addEye2HessianEqMAX);
                        s.updEq = true;
                    }
                } else if (s.curvature < 0) {                               // :458-520, single pass, applied now
                    const bool morePos = D.useInertia && s.mp < D.nU, moreNeg = D.useInertia && s.mn < D.nF + D.nG;
                    if (!D.useInertia || morePos || moreNeg) {
                        const double fU = (!D.useInertia || morePos) ? 10 : 2;
                        const double fEq = (!D.useInertia || !morePos) ? 10 : 2;
                        if (s.aeU < TC_addEye2HessianUMAX) {
                            s.aeU = TC_MIN(fU * TC_MAX(s.aeU, TC_addEye2HessianMIN), TC_addEye2HessianUMAX);
                            TC_SET(FN_setAddEye2HessianU, s.aeU);
                        }
                        if (s.aeEq < TC_addEye2HessianEqMAX) {
                            s.aeEq = TC_MIN(fEq * TC_MAX(s.aeEq, TC_addEye2HessianMIN), TC_addEye2HessianEqMAX);
                            TC_SET(FN_setAddEye2HessianEq, s.aeEq);
                        }
                    }
                    // :522-531 re-reads curvature/inertia/direction error only to print them
                }
            }

            if (!hasF) {                                                    // :550-569
                TC_SET(FN_setAlphaPrimal, s.alphaMax_);
                if (hasG) TC_SET(FN_setAlphaDualEq, s.alphaMax_);
                TC_COPY(FN_updatePrimalDual);
                continue;
            }

            // ---- affine-scaling (Mehrotra) predictor, skipAffine==0 (:576-665) ----
            if (!D.skipAffine) {
                TC_GET(FN_getMaxAlphas_a);
                s.alphaPrimal = TC_VAL(FN_getMaxAlphas_a, 0);
                s.alphaDualIneq = TC_VAL(FN_getMaxAlphas_a, 1);
                s.alphaMax_ = (s.alphaPrimal < D.alphaMax) ? s.alphaPrimal : D.alphaMax;
                if (s.alphaDualIneq < s.alphaMax_) s.alphaMax_ = s.alphaDualIneq;
                if (s.alphaMax_ >= D.alphaMin) {
                    s.alphaPrimal = s.alphaMax_;                            // try max
                    TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                    TC_GET(FN_getMinF_a);
                    s.ineq = TC_VAL(FN_getMinF_a, 0);
                    if (s.ineq < 0) {
                        s.alphaPrimal = D.alphaMin;                         // try min
                        TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                        TC_GET(FN_getMinF_a);
                        s.ineq = TC_VAL(FN_getMinF_a, 0);
                        if (s.ineq > 0) {
                            for (s.alphaPrimal = s.alphaMax_ * .95; s.alphaPrimal >= D.alphaMin; s.alphaPrimal /= 2) {
                                TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                                TC_GET(FN_getMinF_a);
                                s.ineq = TC_VAL(FN_getMinF_a, 0);
                                if (s.ineq >= 0) break;
                            }
                            if (s.alphaPrimal < D.alphaMin) {
                                s.alphaPrimal = 0;
                                TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                            }
                        } else {
                            s.alphaPrimal = 0;
                            TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                        }
                    }
                } else {
                    s.alphaPrimal = 0;
                    TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                }
                TC_SET(FN_setAlphaDualIneq, s.alphaPrimal);
                if (s.alphaPrimal > D.alphaMax / 2 &&
                    (!hasG || s.norminf_eq < 100 * D.equalTolerance || s.norminf_eq < 1e-3)) {      // :627-660
                    TC_GET(FN_getRho);
                    s.curvature = TC_VAL(FN_getRho, 0);                     // sigma (state only: reuses a dead local)
                    if (s.curvature > 1) s.curvature = 1;
                    if (s.curvature < 0) s.curvature = 0;
                    s.curvature = (D.delta == 2) ? s.curvature * s.curvature : s.curvature * s.curvature * s.curvature;
                    s.mu = s.curvature * s.gap / D.nF;
                    if (s.mu < s.muMin) s.mu = s.muMin;
                    TC_SET(FN_setMu, s.mu);
                }
            }

            // ---- combined search direction and line search (:679-756) ----
            TC_GET(FN_getMaxAlphas_s);
            s.alphaPrimal = TC_VAL(FN_getMaxAlphas_s, 0);
            s.alphaDualIneq = TC_VAL(FN_getMaxAlphas_s, 1);
            if (D.coupledAlphas && s.alphaDualIneq < s.alphaPrimal) s.alphaPrimal = s.alphaDualIneq;
            s.alphaPrimal *= TC_stepback;
            s.alphaMax_ = (s.alphaPrimal < D.alphaMax) ? s.alphaPrimal : D.alphaMax;
            if (s.alphaMax_ >= D.alphaMin) {
                s.alphaPrimal = s.alphaMax_ / TC_stepback;                  // try the largest step :690-714
                TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                TC_GET(FN_getMinF_s);
                s.ineq = TC_VAL(FN_getMinF_s, 0);
                if (isnan(s.ineq)) { s.status = 4; break; }
                if (s.ineq > 0) {
                    s.alphaPrimal *= TC_stepback;
                    TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                    TC_GET(FN_getMinF_s);
                    s.ineq1 = TC_VAL(FN_getMinF_s, 0);                      // recheck just inside
                }
                // :704,716 as written in the reference (a NaN in the re-check keeps the step): try the smallest
                // step iff ineq<=0 or ineq1<=ineq/10.  ineq1 is only meaningful when ineq>0 (state only: no
                // local may live across an await)
                if (s.ineq <= 0 || s.ineq1 <= s.ineq / 10) {
                    s.alphaPrimal = D.alphaMin / TC_stepback;               // try the smallest step :716-749
                    TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                    TC_GET(FN_getMinF_s);
                    s.ineq = TC_VAL(FN_getMinF_s, 0);
                    if (s.ineq > 0) {
                        s.alphaPrimal = s.alphaMax_ * .95;                  // scan downwards :722-743
                        while (s.alphaPrimal >= D.alphaMin) {
                            TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                            TC_GET(FN_getMinF_s);
                            s.ineq = TC_VAL(FN_getMinF_s, 0);
                            if (s.ineq > 0) {
                                s.alphaPrimal *= TC_stepback;
                                TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                                TC_GET(FN_getMinF_s);
                                s.ineq1 = TC_VAL(FN_getMinF_s, 0);
                                if (s.ineq1 > s.ineq / 10) break;
                            }
                            s.alphaPrimal /= 2;
                        }
                        if (!(s.alphaPrimal >= D.alphaMin)) {
                            s.alphaPrimal = 0;
                            TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                        }
                    } else {
                        s.alphaPrimal = 0;
                        TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
                    }
                }
            } else {
                s.alphaPrimal = 0;                                          // :752-756
                TC_SET(FN_setAlphaPrimal, s.alphaPrimal);
            }

            if (D.coupledAlphas) {                                          // :758-767
                s.alphaDualEq = s.alphaPrimal;
                s.alphaDualIneq = s.alphaPrimal;
            } else {
                s.alphaDualIneq *= TC_stepback;
                if (s.alphaDualIneq > s.alphaMax_) s.alphaDualIneq = s.alphaMax_;
                s.alphaDualEq = s.alphaDualIneq;
            }
            if (hasG) TC_SET(FN_setAlphaDualEq, s.alphaDualEq);
            TC_SET(FN_setAlphaDualIneq, s.alphaDualIneq);
            TC_COPY(FN_updatePrimalDual);                                   // :772

            // ---- barrier parameter update, skipAffine!=0 (:782-859) ----
            s.th_grad = s.norminf_grad < TC_MAX(1e-6, 1e0 * D.gradTolerance);
            s.th_eq = !hasG || s.norminf_eq < TC_MAX(1e-5, 1e0 * D.equalTolerance);
            if (!D.skipAffine) {
            } else if (s.alphaPrimal > s.alphaMax_ / 2 && s.th_grad && s.th_eq) {
                const double r = sqrt(s.mu);   // pow(mu,.5) in the reference; sqrt is its correctly rounded value
                s.mu *= TC_MIN(D.muFactorAggressive, r);
                if (s.mu < s.muMin) s.mu = s.muMin;
                TC_SET(FN_setMu, s.mu);
            } else if (s.alphaPrimal < .1) {
                s.mu *= 1.1;
                if (s.mu > mu0) s.mu = mu0;
                TC_SET(FN_setMu, s.mu);
                TC_COPY(FN_initDualIneq);                                   // :816
            } else if (s.alphaPrimal > .99 && s.th_eq) {
                s.mu *= D.muFactorConservative;
                if (s.mu < s.muMin) s.mu = s.muMin;
                TC_SET(FN_setMu, s.mu);
            }
            if (s.alphaPrimal < D.alphaMin && s.alphaDualIneq < D.alphaMin && s.alphaDualEq < D.alphaMin) {   // :853-859
                s.mu /= (D.muFactorConservative * D.muFactorConservative);
                if (s.mu < s.muMin) s.mu = s.muMin;
                TC_SET(FN_setMu, s.mu);
            }
        }

        // ---- diagnostic bits when maxIter was reached (:885-920) ----
        if (s.status == 8) {
            TC_GET(FN_getNorminf_Grad);
            if (TC_VAL(FN_getNorminf_Grad, 0) > D.gradTolerance) s.status |= 16;
            if (hasG) {
                TC_GET(FN_getNorminf_G);
                if (TC_VAL(FN_getNorminf_G, 0) > D.equalTolerance) s.status |= 32;
            }
            if (hasF) {
                TC_GET(FN_getGapMinFMinLambda);
                if (TC_VAL(FN_getGapMinFMinLambda, 0) > s.gapTol) s.status |= 64;
                if (s.mu > s.muMin) s.status |= 128;
                if (s.alphaPrimal <= D.alphaMin && s.alphaDualIneq < D.alphaMin && s.alphaDualEq < D.alphaMin)
                    s.status |= 1792;
                else if (s.alphaPrimal <= .1 && s.alphaDualIneq < .1 && s.alphaDualEq < .1)
                    s.status |= 1536;
                else if (s.alphaPrimal <= .5 && s.alphaDualIneq < .5 && s.alphaDualEq < .5)
                    s.status |= 1024;
            }
            if (adjust && s.aeU > D.addEye2HessianUtolerance) s.status |= 2048;
        }
#ifdef TC_DRIVER_QUEUE
        TC_GET(FN_getOutputs);     // the lane's column is recycled: evaluate the outputs before it is
#endif
    }
    s.pc = -1;
    s.want_kind = TC_WANT_NONE;
}

__device__ __forceinline__ void tc_state_init(tc_lane_state &s, const tc_device_image &P, bool valid, double mu0,
                                              double aeU0, double aeEq0) {
    s.pc = valid ? 0 : -1;
    s.want_fn = 0;
    s.want_kind = TC_WANT_NONE;
    s.status = 0; s.iter = 0; s.updU = s.updEq = s.th_grad = s.th_eq = false;
    s.aeU = aeU0; s.aeEq = aeEq0; s.mu = mu0; s.muMin = 0; s.gapTol = P.def.desiredDualityGap;
    s.alphaMax_ = P.def.alphaMax; s.alphaPrimal = s.alphaDualEq = s.alphaDualIneq = 0;
    s.norminf_grad = s.norminf_eq = s.gap = s.ineq = s.ineq1 = s.dual = s.curvature = s.derr = s.mp = s.mn = 0;
}

// One scheduling round of the warp: resume the lanes whose function is complete, otherwise run
// one dependency group.  Must be called by all 32 lanes while any lane is alive (s.pc >= 0).
__device__ __forceinline__ void tc_lanes_round(tc_lane_state &s, tc_lane &c, const tc_device_image &P, double mu0,
                                               int maxIter, uint32_t big) {
    const bool alive = s.pc >= 0;
    // 1. the first group this lane still misses for the function it waits on (TC_MAXGROUPS = none)
    int front = TC_MAXGROUPS;
    if (alive) {
        const tc_func *f = P.funcs + s.want_fn;
#pragma unroll
        for (int w = TC_GW - 1; w >= 0; w--) {
            const uint64_t miss = f->parents[w] & ~c.st[0][w];
            if (miss) front = 64 * w + __ffsll((long long)miss) - 1;
        }
    }
    // 2. lanes whose function is complete go on at once: copies first (one function at a
    //    time, its code is shared by the lanes that want it), then the solver code up to the
    //    lane's next get/copy
    const bool ready = alive && front == TC_MAXGROUPS;
    if (__any_sync(TC_FULL, ready)) {
        bool copying = ready && s.want_kind == TC_WANT_COPY;
        while (__any_sync(TC_FULL, copying)) {
            const int fn = __shfl_sync(TC_FULL, s.want_fn, __ffs(__ballot_sync(TC_FULL, copying)) - 1);
            const bool mine = copying && s.want_fn == fn;
#ifdef TC_COMPILED_GROUPS
            const bool done = tcm_copy(fn, c.mb[0], mine ? 1u : 0u);
#else
            const bool done = false;
#endif
            __syncwarp();
            const tc_func *f = P.funcs + fn;
            if (!done) {                                                    // generic body (uniform)
                const uint32_t *io = P.io + f->io_off;
                for (uint32_t k = 0; k < f->io_cnt; k += 2) {
                    const uint32_t d = __ldg(io + k);
                    const double v = tc_ldop(c, 0, __ldg(io + k + 1), mine);
                    if (mine) c.mb[0][(size_t)d * TC_TILE] = v;
                }
            }
            if (mine) {
#pragma unroll
                for (int w = 0; w < TC_GW; w++) c.st[0][w] &= ~f->children[w];
                copying = false;
            }
        }
        if (ready) tc_lane_resume(s, c, P, mu0, maxIter);
        __syncwarp();
        return;
    }
    // 3. nobody can go on: run ONE group.  Cheap groups: the front group of the lane that is
    //    furthest behind (keeps the lanes together, like a warp-synchronous walk).  The
    //    dominant group (the KKT factorisation; >= half the largest group's size) is a
    //    RENDEZVOUS: it runs only when every living lane is waiting for a dominant group, so a
    //    lane that re-factorises within an iteration shares the pass with the others' next one.
    const int fg = alive ? front : 0;
    const bool held = alive && (__ldg(P.group_start + fg + 1) - __ldg(P.group_start + fg)) >= big;
    const bool some_free = __any_sync(TC_FULL, alive && !held);         // (a vote: outside any short-circuit)
    const bool cand = alive && (!held || !some_free);
    unsigned long long key = cand ? ((unsigned long long)(((unsigned)s.iter << 16) | ((unsigned)s.pc & 0xffffu)) << 5) | c.lane
                                  : ~0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(TC_FULL, key, o);
        key = other < key ? other : key;
    }
    const int g = __shfl_sync(TC_FULL, front, (int)(key & 31u));
    const bool need[1] = {front == g};
    tc_exec_group(c, P.group_start, g, need);
    if (need[0]) c.st[0][g >> 6] |= 1ull << (g & 63);
    const unsigned vote = __ballot_sync(TC_FULL, need[0]);
    if (c.lane == 0) {
        atomicAdd(P.group_count + g, (unsigned long long)__popc(vote));
        atomicAdd(P.group_count + TC_MAXGROUPS + g, 1ull);
    }
}

__device__ __forceinline__ uint32_t tc_dominant_size(const tc_device_image &P) {
    uint32_t big = 0;              // size (tape cells) that makes a group "dominant"
    for (int g = 0; g < P.n_groups; g++) big = max(big, __ldg(P.group_start + g + 1) - __ldg(P.group_start + g));
    return big / 2 + 1;
}

__device__ __forceinline__ void tc_ipm_solve(tc_lane &c, const tc_device_image &P, const bool (&valid)[1], double mu0,
                                             int maxIter, double aeU0, double aeEq0, int (&status_out)[1],
                                             int (&iter_out)[1]) {
    tc_lane_state s;
    tc_state_init(s, P, valid[0], mu0, aeU0, aeEq0);
    if (s.pc == 0) tc_lane_resume(s, c, P, mu0, maxIter);
    __syncwarp();
    const uint32_t big = tc_dominant_size(P);
    while (__any_sync(TC_FULL, s.pc >= 0)) tc_lanes_round(s, c, P, mu0, maxIter, big);
    status_out[0] = s.status;
    iter_out[0] = s.iter;
}

#ifdef TC_DRIVER_QUEUE
// WORK-QUEUE form (emitter option driver="queue"): the grid is one warp per resident slot; every
// lane works in its own column of a WORK scratchbook (P.m_work) and takes instances from a
// global counter.  Taking an instance = copying the slots of its set functions from the instance's
// column of the batch scratchbook; finishing = writing status/iter, copying the output slots back
// and marking the output groups valid there (tc_get_outputs then gathers without recomputing).
// Lanes are refilled when TC_REFILL_MIN of a warp are idle (their start-up groups then run for
// several lanes at once) or when the warp has nothing else to do.
#ifndef TC_REFILL_MIN
#define TC_REFILL_MIN 8
#endif

__device__ __forceinline__ void tc_queue_solve(tc_lane &c, const tc_device_image &P, double mu0, int maxIter,
                                               double aeU0, double aeEq0, int32_t *__restrict__ status,
                                               int32_t *__restrict__ iter) {
    tc_lane_state s;
    tc_state_init(s, P, false, mu0, aeU0, aeEq0);
    long long inst = -1;           // the instance this lane is working on
    bool exhausted = false;        // the queue had nothing left for this lane
    const uint32_t big = tc_dominant_size(P);
    const tc_func *fo = P.funcs + FN_getOutputs;
    const tc_func *fu = P.funcs + FN_updatePrimalDual;   // io = (dst,src) pairs: dst = u, nu, lambda
    while (true) {
        // a lane whose solver loop ended hands its instance back
        const bool finished = s.pc < 0 && inst >= 0;
        if (__any_sync(TC_FULL, finished)) {
            if (finished) {
                status[inst] = s.status;
                iter[inst] = s.iter;
#pragma unroll
                for (int w = 0; w < TC_GW; w++) P.group_status[w * P.stride + inst] = fo->parents[w];
            }
            for (uint32_t k = 0; k < fo->io_cnt; k++) {                     // uniform loop, per-lane targets
                const uint32_t w = __ldg(P.io + fo->io_off + k);
                if (finished && w != TC_NO_SLOT && !(w & TC_CONST_BIT))
                    TC_M(P, w, inst) = c.mb[0][(size_t)w * TC_TILE];
            }
            // the optimum itself (u, nu, lambda) goes back too: a later set_<parameter> + getOutputs without a
            // new solve re-evaluates the output groups on the batch scratchbook, like the reference does
            for (uint32_t k = 0; k < fu->io_cnt; k += 2) {
                const uint32_t w = __ldg(P.io + fu->io_off + k);
                if (finished) TC_M(P, w, inst) = c.mb[0][(size_t)w * TC_TILE];
            }
            if (finished) inst = -1;
        }
        const bool idle = s.pc < 0 && !exhausted;
        const unsigned idle_mask = __ballot_sync(TC_FULL, idle);
        const bool any_alive = __any_sync(TC_FULL, s.pc >= 0);
        if (idle_mask && (__popc(idle_mask) >= TC_REFILL_MIN || !any_alive)) {
            unsigned long long base = 0;
            if (c.lane == 0) base = atomicAdd(P.queue, (unsigned long long)__popc(idle_mask));
            base = __shfl_sync(TC_FULL, base, 0);
            const long long j = (long long)base + __popc(idle_mask & ((1u << c.lane) - 1u));
            const bool take = idle && j < P.batch;
            if (idle && !take) exhausted = true;
            for (int fn = FN_USER0; fn < P.n_funcs; fn++) {                 // <class>_set_* of the new instance
                const tc_func *f = P.funcs + fn;
                for (uint32_t k = 0; k < f->io_cnt; k++) {
                    const uint32_t slot = __ldg(P.io + f->io_off + k);
                    if (take) c.mb[0][(size_t)slot * TC_TILE] = TC_M(P, slot, j);
                }
            }
            if (take) {
                inst = j;
#pragma unroll
                for (int w = 0; w < TC_GW; w++) c.st[0][w] = 0;             // every group is stale for the new instance
                tc_state_init(s, P, true, mu0, aeU0, aeEq0);
                tc_lane_resume(s, c, P, mu0, maxIter);
            }
            __syncwarp();
            continue;
        }
        if (!any_alive) break;     // nothing running, nothing left to take
        tc_lanes_round(s, c, P, mu0, maxIter, big);
    }
}
#endif  // TC_DRIVER_QUEUE
