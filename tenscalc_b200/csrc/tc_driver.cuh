// Device-resident primal-dual interior-point driver, one instance per lane.
//
// Control flow restated from lib/ipmPD_CSsolver.c:132-920 (the template cmex2optimizeCS
// specialises with the #define table of lib/cmex2optimizeCS.m:303-331) for the option set
// that generator can emit: ipmPD_CStimesLambda => skipAffine=1, useLDL=1, large KKT matrix.
// The reference runs it on the host around ~25 generated C functions; here the whole loop
// lives in one kernel: no host round trip per Newton step.  Because all lanes of a warp
// share the tape, the driver is written in predicated form -- every lane walks the same
// sequence of set/get/copy calls and a per-lane predicate says whether the call applies to
// its instance; branches are taken on warp votes only.  Per-instance decisions (exit tests,
// regularisation control, line search, barrier update) use the same comparisons, constants
// and order as the reference, so status codes and iteration counts match it.
//
// TC_TILES (compiled back end): every lane carries TC_TILES instances (one per scratchbook tile
// of its warp), all per-instance state below is an array over t, and a dependency group is run
// for all tiles of the warp by ONE pass over its code (the chunk functions loop over the tiles):
// the instruction stream -- megabytes of straight-line SASS per Newton step, which is what
// bounds a warp -- is fetched once per TC_TILES*32 instances.
#pragma once
#include "tc_vm.cuh"

// constants of ipmPD_CSsolver.c:145-148,174
#define TC_addEye2HessianUMAX 1e2
#define TC_addEye2HessianEqMAX 1e2
#define TC_addEye2HessianMIN 1e-20
#define TC_maxDirectionError 1e-6
#define TC_stepback .99
// equilibrium driver, D.driver == 1: lib/ipmPDeq_CSsolver.c:313-316
#define TCEQ_addEye2Hessian2MAX 1e-2
#define TCEQ_maxDirectionError 1e-9

#define TC_MIN(A, B) (((A) < (B)) ? (A) : (B))
#define TC_MAX(A, B) (((A) > (B)) ? (A) : (B))

// line-search automaton, one state per `getMinF_s__` evaluation of ipmPD_CSsolver.c:690-756
enum { LS_DONE = 0, LS_TRYMAX, LS_RECHECK, LS_TRYMIN, LS_SCAN, LS_SCAN_RECHECK };
// affine-step automaton, one state per `getMinF_a__` evaluation of ipmPD_CSsolver.c:580-625
enum { AF_DONE = 0, AF_TRYMAX, AF_TRYMIN, AF_SCAN };

// PHASED EXECUTION.  One solve = a sequence of launches of this loop ("phases", tc_abi.cu): a phase runs at
// most `phase_iters` Newton iterations per instance and parks the loop's locals of the unfinished
// instances in P.ps_d / P.ps_i; between phases the host-enqueued compaction kernel moves the unfinished
// instances into the leading scratchbook columns, so every warp of the next phase is full again (a
// statically assigned warp runs until its slowest lane is done: 34.9 iterations per warp against 27.3 per
// instance on mpc_quadcopter) and the few instances that need hundreds of iterations end up together in
// one warp.  The per-instance sequence of operations -- hence every rounding -- is unchanged.
// phase_iters == 0: the whole solve in this launch (tape interpreter, lanes/queue drivers).
enum { PS_aeU = 0, PS_aeEq, PS_mu, PS_muMin, PS_gapTol, PS_alphaMax, PS_alphaPrimal, PS_alphaDualEq, PS_alphaDualIneq,
       PS_mp, PS_mn, PS_aeD };                             // TC_PS_ND doubles per column
enum { PS_iter = 0, PS_flags };                            // TC_PS_NI ints; flags: 1 running, 2 updU, 4 updEq, 8 updD
static_assert(PS_aeD + 1 == TC_PS_ND && PS_flags + 1 == TC_PS_NI, "phase-state layout");

typedef double tc_dT[TC_TILES];
typedef bool tc_bT[TC_TILES];
typedef int tc_iT[TC_TILES];
#define TC_FORT _Pragma("unroll") for (int t = 0; t < TC_TILES; t++)

__device__ __forceinline__ bool tc_any_tile(const tc_bT &x) {
    bool a = false;
    TC_FORT a = a || x[t];
    return TC_ANY(a);
}

// valid: the lane's column holds an instance this launch works on; pos: the column; done_out: the
// instance finished in this launch (status_out/iter_out are final)
__device__ __forceinline__ void tc_ipm_solve(tc_lane &c, const tc_device_image &P, const tc_bT &valid, double mu0,
                                             int maxIter, double aeU0, double aeEq0, tc_iT &status_out,
                                             tc_iT &iter_out, tc_bT &done_out, const int64_t (&pos)[TC_TILES],
                                             int phase_first, int phase_iters) {
    const tc_defines &D = P.def;
    const bool hasF = D.nF > 0, hasG = D.nG > 0;
    const bool adjust = D.setAddEye2Hessian && D.adjustAddEye2Hessian;
    // D.driver == 1: ipmPDeq_CSsolver (lib/ipmPDeq_CSsolver.c:122-802, Nash equilibria with latent variables).  Same loop
    // skeleton; what differs is marked `eq` below: only the second regulariser is adapted, from the direction error
    // alone and with up to 20 immediate retries (:333-374); thresholds of the barrier update (:622-651); alphaDualEq
    // without coupled step sizes (:611); no 2048 status bit.  FN_setAddEye2HessianU/Eq are setAddEye2Hessian1__/2__ and
    // FN_initDualEqX is initDualEq__ there (tape.py).
    const bool eq = D.driver == 1;

    // D.driver == 2: ipmPDminmax_CSsolver (lib/ipmPDminmax_CSsolver.m:1-587; its C namesake is stale in the reference,
    // so the MATLAB solver is the specification and parity with the reference is UNPINNED): a third regulariser for the
    // maximiser's block, regulariser control from two inertias and the direction error with up to 30 immediate
    // retries (:238-301), other barrier thresholds (:411-440), 2048 bit from either tolerance (:503-507).
    const bool mm = D.driver == 2;
    tc_bT run, updU, updEq, updD, pr, pr2;
    // allowSave (cmex2optimizeCS.m:327): saveWW__ -- in the reference a dump of the KKT matrix values to a file -- is a
    // copy of those values into the instance's save variables, which getOutputs returns as its last output.  The
    // events are the reference's: a failed factorisation (status 4, :330-332,700-702), iter == saveIter
    // (:561-566,671-676) and saveIter == 0 at a successful exit (:877-882).  `sv` marks the lanes of an event.
    tc_bT sv;
    TC_FORT sv[t] = false;
#define TC_SAVE_WW()                                                                                  \
    if (D.allowSave && !eq && !mm) {                                                                  \
        if (tc_any_tile(sv)) tc_copy(c, P, FN_saveWW, sv);                                            \
        TC_FORT sv[t] = false;                                                                        \
    }
    tc_iT status, iter, ls;
    tc_dT aeU, aeEq, aeD, mu, muMin, gapTol, alphaMax_, alphaPrimal, alphaDualEq, alphaDualIneq;
    tc_dT norminf_grad, norminf_eq, gap, ineq, ineq1, dual, v0, v1, v2, mp, mn;
    TC_FORT {
        run[t] = valid[t]; updU[t] = updEq[t] = updD[t] = false; status[t] = 0; iter[t] = 0; ls[t] = LS_DONE;
        aeU[t] = aeU0; aeEq[t] = aeEq0; aeD[t] = P.aeD0; mu[t] = mu0; muMin[t] = 0; gapTol[t] = D.desiredDualityGap;
        alphaMax_[t] = D.alphaMax; alphaPrimal[t] = alphaDualEq[t] = alphaDualIneq[t] = 0;
        norminf_grad[t] = norminf_eq[t] = gap[t] = ineq[t] = ineq1[t] = dual[t] = mp[t] = mn[t] = 0;
    }

    if (phase_first) {
    if (D.setAddEye2Hessian) {                                              // :150-155
        tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, run);
        if (mm) tc_set_scalar(c, P, FN_setAddEye2HessianD, aeD, run);
        tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, run);
    }
    tc_copy(c, P, FN_initPrimal, run);                                      // :204
    if (D.scaleCost) {                                                      // :206-214
        tc_copy(c, P, FN_scaleCost, run);
        tc_get(c, P, FN_getScale4Cost, run);
        tc_value(c, P, FN_getScale4Cost, 0, run, v0);
        if (hasF) TC_FORT gapTol[t] *= v0[t];
    }
    if (hasF) {                                                             // :216-230
        if (D.scaleInequalities) tc_copy(c, P, FN_scaleIneq, run);
        tc_set_scalar(c, P, FN_setMu, mu, run);
        TC_FORT muMin[t] = gapTol[t] / D.nF / 2;
        tc_copy(c, P, FN_initDualIneq, run);
    }
    if (hasG) tc_copy(c, P, FN_initDualEqX, run);                           // :232-235
    } else {                                                                // resume: the locals parked by the last phase
        TC_FORT if (valid[t]) {
            const double *d = P.ps_d + pos[t];
            const int32_t *q = P.ps_i + pos[t];
            const int64_t S = P.stride;
            aeU[t] = d[PS_aeU * S]; aeEq[t] = d[PS_aeEq * S]; mu[t] = d[PS_mu * S]; muMin[t] = d[PS_muMin * S];
            gapTol[t] = d[PS_gapTol * S]; alphaMax_[t] = d[PS_alphaMax * S]; alphaPrimal[t] = d[PS_alphaPrimal * S];
            alphaDualEq[t] = d[PS_alphaDualEq * S]; alphaDualIneq[t] = d[PS_alphaDualIneq * S];
            mp[t] = d[PS_mp * S]; mn[t] = d[PS_mn * S]; aeD[t] = d[PS_aeD * S];
            iter[t] = q[PS_iter * S];
            const int fl = q[PS_flags * S];
            run[t] = (fl & 1) != 0; updU[t] = (fl & 2) != 0; updEq[t] = (fl & 4) != 0; updD[t] = (fl & 8) != 0;
        }
    }
    tc_bT entered;
    TC_FORT entered[t] = run[t];

    for (int phase_it = 0; tc_any_tile(run) && (phase_iters == 0 || phase_it < phase_iters); phase_it++) {   // :282
        TC_FORT {
            if (run[t]) iter[t]++;
            if (run[t] && iter[t] > maxIter) { status[t] = 8; run[t] = false; }   // :308-313
        }

        // ---- exit conditions (:323-381) ----
        tc_get(c, P, FN_getNorminf_Grad, run);
        tc_value(c, P, FN_getNorminf_Grad, 0, run, v0);
        TC_FORT {
            if (run[t]) norminf_grad[t] = v0[t];
            if (run[t] && isnan(norminf_grad[t])) { status[t] = 4; run[t] = false; sv[t] = true; }
        }
        TC_SAVE_WW();                                                       // :329-332
        if (hasG) {
            tc_get(c, P, FN_getNorminf_G, run);
            tc_value(c, P, FN_getNorminf_G, 0, run, v0);
            TC_FORT if (run[t]) norminf_eq[t] = v0[t];
        }
        if (hasF) {
            tc_get(c, P, FN_getGapMinFMinLambda, run);
            tc_value(c, P, FN_getGapMinFMinLambda, 0, run, v0);
            tc_value(c, P, FN_getGapMinFMinLambda, 1, run, v1);
            tc_value(c, P, FN_getGapMinFMinLambda, 2, run, v2);
            TC_FORT {
                if (run[t]) { gap[t] = v0[t]; ineq[t] = v1[t]; dual[t] = v2[t]; }
                if (run[t] && ineq[t] <= 0) { status[t] = 1; run[t] = false; }
                if (run[t] && dual[t] <= 0) { status[t] = 2; run[t] = false; }
            }
        }
        TC_FORT if (run[t] && norminf_grad[t] <= D.gradTolerance && (!hasF || gap[t] <= gapTol[t]) &&
                    (!hasG || norminf_eq[t] <= D.equalTolerance) &&
                    (!adjust || (aeU[t] <= D.addEye2HessianUtolerance && (!mm || aeD[t] <= D.addEye2HessianDtolerance)))) {
            status[t] = 0;
            run[t] = false;
        }

        // ---- regularisation control (:393-548) ----
        if (adjust && mm) {                                                 // ipmPDminmax_CSsolver.m:225-301
            tc_bT prD;
            TC_FORT { pr[t] = run[t] && updU[t]; prD[t] = run[t] && updD[t]; pr2[t] = run[t] && updEq[t]; }
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, pr);            // delayed from last iteration
            tc_set_scalar(c, P, FN_setAddEye2HessianD, aeD, prD);
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, pr2);
            tc_bT retry;
            TC_FORT { if (run[t]) updU[t] = updD[t] = updEq[t] = false; retry[t] = run[t]; }
            for (int ii = 0; ii < 30 && tc_any_tile(retry); ii++) {
                tc_get(c, P, FN_getDirectionError, retry);
                tc_value(c, P, FN_getDirectionError, 0, retry, v0);         // derr
                tc_get(c, P, FN_getHessDinertia, retry);
                tc_value(c, P, FN_getHessDinertia, 1, retry, v1);           // mnD
                tc_get(c, P, FN_getHessUinertia, retry);
                tc_value(c, P, FN_getHessUinertia, 0, retry, v2);           // mpU
                tc_bT nowU, nowD, nowEq;
                TC_FORT {
                    nowU[t] = nowD[t] = nowEq[t] = false;
                    if (retry[t]) {
                        const double derr = v0[t], mnD = v1[t], mpU = v2[t];
                        if (mpU == D.mpUdesired && mnD == D.mnDdesired && derr <= 1e-7) {
                            if (aeEq[t] > TC_addEye2HessianMIN) { aeEq[t] = TC_MAX(.75 * aeEq[t], TC_addEye2HessianMIN); updEq[t] = true; }
                            if (aeU[t] > TC_addEye2HessianMIN) { aeU[t] = TC_MAX(.75 * aeU[t], TC_addEye2HessianMIN); updU[t] = true; }
                            if (aeD[t] > TC_addEye2HessianMIN) { aeD[t] = TC_MAX(.75 * aeD[t], TC_addEye2HessianMIN); updD[t] = true; }
                            retry[t] = false;                               // stored at the next iteration
                        } else {
                            // (:281,286 compare the update FLAG with the maximum, which always holds: restated as written)
                            if (mpU < D.mpUdesired) { aeU[t] = TC_MIN(10 * aeU[t], TC_addEye2HessianUMAX); nowU[t] = true; }
                            if (mnD < D.mnDdesired) { aeD[t] = TC_MIN(10 * aeD[t], TC_addEye2HessianUMAX); nowD[t] = true; }
                            if (derr > 1e-7 && aeEq[t] < TC_addEye2HessianEqMAX) {
                                aeEq[t] = TC_MIN(10 * aeEq[t], TC_addEye2HessianEqMAX);
                                nowEq[t] = true;
                            }
                            if (!(nowU[t] || nowD[t] || nowEq[t])) retry[t] = false;
                        }
                    }
                }
                tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, nowU);      // immediate
                tc_set_scalar(c, P, FN_setAddEye2HessianD, aeD, nowD);
                tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, nowEq);
            }
        } else if (adjust && eq) {                                          // ipmPDeq_CSsolver.c:318-374
            TC_FORT { pr[t] = run[t] && updU[t]; pr2[t] = run[t] && updEq[t]; }
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, pr);
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, pr2);
            TC_FORT if (run[t]) updU[t] = updEq[t] = false;
            tc_get(c, P, FN_getDirectionError, run);
            tc_value(c, P, FN_getDirectionError, 0, run, v1);
            tc_bT retry;
            TC_FORT {
                retry[t] = false;
                if (run[t]) {
                    if (v1[t] < TCEQ_maxDirectionError) {
                        if (aeEq[t] > TC_addEye2HessianMIN) {
                            aeEq[t] = TC_MAX(.75 * aeEq[t], TC_addEye2HessianMIN);
                            updEq[t] = true;                                // stored at the next iteration
                        }
                    } else retry[t] = true;
                }
            }
            for (int ii = 0; ii < 20; ii++) {                               // :343-372, per instance
                TC_FORT {
                    pr[t] = false;
                    if (retry[t]) {
                        if (v1[t] >= TCEQ_maxDirectionError && aeEq[t] < TCEQ_addEye2Hessian2MAX) {
                            aeEq[t] = TC_MIN(2 * TC_MAX(aeEq[t], TC_addEye2HessianMIN), TCEQ_addEye2Hessian2MAX);
                            pr[t] = true;
                        } else retry[t] = false;
                    }
                }
                if (!tc_any_tile(pr)) break;
                tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, pr);     // immediate
                tc_get(c, P, FN_getDirectionError, pr);
                tc_value(c, P, FN_getDirectionError, 0, pr, v2);
                TC_FORT if (pr[t]) v1[t] = v2[t];
            }
        } else if (adjust) {
            TC_FORT { pr[t] = run[t] && updU[t]; pr2[t] = run[t] && updEq[t]; }
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, pr);            // delayed from last iteration
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, pr2);
            TC_FORT if (run[t]) updU[t] = updEq[t] = false;
            tc_get(c, P, FN_getCurvature, run);
            tc_value(c, P, FN_getCurvature, 0, run, v0);                    // curvature
            if (D.useInertia) {                                             // :408-413
                TC_FORT pr[t] = run[t] && v0[t] <= 0;
                tc_get(c, P, FN_getHessInertia, pr);
                tc_value(c, P, FN_getHessInertia, 0, pr, v1);
                tc_value(c, P, FN_getHessInertia, 1, pr, v2);
                TC_FORT if (pr[t]) { mp[t] = v1[t]; mn[t] = v2[t]; }
            }
            tc_get(c, P, FN_getDirectionError, run);
            tc_value(c, P, FN_getDirectionError, 0, run, v1);               // direction error
            TC_FORT {
                const double curvature = v0[t], derr = v1[t];
                bool nowU = false, nowEq = false;
                // inertia of the KKT factor: nU positive, nF+nG negative pivots wanted (:164-169)
                const bool inertia_ok = D.useInertia && mp[t] == D.nU && mn[t] == D.nF + D.nG;
                const bool morePos = D.useInertia && mp[t] < D.nU, moreNeg = D.useInertia && mn[t] < D.nF + D.nG;
                const double fU = (!D.useInertia || morePos) ? 10 : 2, fEq = (!D.useInertia || !morePos) ? 10 : 2;
                if (run[t]) {
                    if (curvature > 0 || inertia_ok) {                      // :415-456
                        if (aeU[t] > TC_addEye2HessianMIN && derr < TC_maxDirectionError) {
                            aeU[t] = TC_MAX(.75 * aeU[t], TC_addEye2HessianMIN);
                            updU[t] = true;
                        }
                        if (aeEq[t] > TC_addEye2HessianMIN && derr < TC_maxDirectionError) {
                            aeEq[t] = TC_MAX(.75 * aeEq[t], TC_addEye2HessianMIN);
                            updEq[t] = true;
                        }
                        if (aeU[t] < TC_addEye2HessianUMAX && derr > TC_maxDirectionError) {
                            aeU[t] = TC_MIN(10 * aeU[t], TC_addEye2HessianUMAX);
                            updU[t] = true;
                        }
                        if (aeEq[t] < TC_addEye2HessianEqMAX && derr > TC_maxDirectionError) {
                            aeEq[t] = TC_MIN(10 * aeEq[t], TC_addEye2HessianEqMAX);
                            updEq[t] = true;
                        }
                    } else if (curvature < 0 && (!D.useInertia || morePos || moreNeg)) {   // :462-520, single pass
                        if (aeU[t] < TC_addEye2HessianUMAX) {
                            aeU[t] = TC_MIN(fU * TC_MAX(aeU[t], TC_addEye2HessianMIN), TC_addEye2HessianUMAX);
                            nowU = true;
                        }
                        if (aeEq[t] < TC_addEye2HessianEqMAX) {
                            aeEq[t] = TC_MIN(fEq * TC_MAX(aeEq[t], TC_addEye2HessianMIN), TC_addEye2HessianEqMAX);
                            nowEq = true;
                        }
                    }
                }
                pr[t] = nowU;
                pr2[t] = nowEq;
            }
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, pr);            // immediate
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, pr2);
            // The reference re-reads curvature and direction error here (:524-531) only to
            // print them; the groups they would refresh are refreshed lazily by the next get.
        }

        if (!hasF) {                                                        // :550-569
            tc_set_scalar(c, P, FN_setAlphaPrimal, alphaMax_, run);
            if (hasG) tc_set_scalar(c, P, FN_setAlphaDualEq, alphaMax_, run);
            TC_FORT sv[t] = run[t] && iter[t] == P.save_iter;               // :561-566
            TC_SAVE_WW();
            tc_copy(c, P, FN_updatePrimalDual, run);
            continue;
        }

        // ---- affine-scaling (Mehrotra) predictor, skipAffine==0 (:576-665) ----
        if (!D.skipAffine) {
            tc_get(c, P, FN_getMaxAlphas_a, run);
            tc_value(c, P, FN_getMaxAlphas_a, 0, run, v0);
            tc_value(c, P, FN_getMaxAlphas_a, 1, run, v1);
            tc_iT af;
            tc_bT azero, aev;
            TC_FORT {
                af[t] = AF_DONE;
                azero[t] = false;
                if (run[t]) {
                    alphaPrimal[t] = v0[t];
                    alphaDualIneq[t] = v1[t];
                    alphaMax_[t] = (alphaPrimal[t] < D.alphaMax) ? alphaPrimal[t] : D.alphaMax;
                    if (alphaDualIneq[t] < alphaMax_[t]) alphaMax_[t] = alphaDualIneq[t];
                    if (alphaMax_[t] >= D.alphaMin) { alphaPrimal[t] = alphaMax_[t]; af[t] = AF_TRYMAX; }   // try max
                    else { alphaPrimal[t] = 0; azero[t] = true; }
                }
            }
            while (true) {
                TC_FORT pr[t] = azero[t] || af[t] != AF_DONE;
                tc_set_scalar(c, P, FN_setAlphaPrimal, alphaPrimal, pr);
                TC_FORT { azero[t] = false; aev[t] = af[t] != AF_DONE; }
                if (!tc_any_tile(aev)) break;
                tc_get(c, P, FN_getMinF_a, aev);
                tc_value(c, P, FN_getMinF_a, 0, aev, v0);
                TC_FORT {
                    const double f = v0[t];
                    if (aev[t]) switch (af[t]) {
                    case AF_TRYMAX:
                        if (f < 0) { alphaPrimal[t] = D.alphaMin; af[t] = AF_TRYMIN; }               // try min
                        else af[t] = AF_DONE;
                        break;
                    case AF_TRYMIN:
                        if (f > 0) {                                                              // between min and max
                            alphaPrimal[t] = alphaMax_[t] * .95;
                            if (alphaPrimal[t] >= D.alphaMin) af[t] = AF_SCAN;
                            else { alphaPrimal[t] = 0; azero[t] = true; af[t] = AF_DONE; }
                        } else { alphaPrimal[t] = 0; azero[t] = true; af[t] = AF_DONE; }
                        break;
                    case AF_SCAN:
                        if (f >= 0) af[t] = AF_DONE;
                        else {
                            alphaPrimal[t] /= 2;
                            if (!(alphaPrimal[t] >= D.alphaMin)) { alphaPrimal[t] = 0; azero[t] = true; af[t] = AF_DONE; }
                        }
                        break;
                    }
                }
            }
            tc_set_scalar(c, P, FN_setAlphaDualIneq, alphaPrimal, run);
            // mu from the predicted complementarity (sigma = rho^delta), only after a long affine step (:627-660)
            TC_FORT pr[t] = run[t] && alphaPrimal[t] > D.alphaMax / 2 &&
                            (!hasG || norminf_eq[t] < 100 * D.equalTolerance || (!eq && norminf_eq[t] < 1e-3));   // eq :377
            tc_get(c, P, FN_getRho, pr);
            tc_value(c, P, FN_getRho, 0, pr, v0);
            TC_FORT if (pr[t]) {
                double sigma = v0[t];
                if (sigma > 1) sigma = 1;
                if (sigma < 0) sigma = 0;
                sigma = (D.delta == 2) ? sigma * sigma : sigma * sigma * sigma;
                mu[t] = sigma * gap[t] / D.nF;
                if (mu[t] < muMin[t]) mu[t] = muMin[t];
            }
            tc_set_scalar(c, P, FN_setMu, mu, pr);
        }

        // ---- combined search direction (:679-772) ----
        TC_FORT sv[t] = run[t] && iter[t] == P.save_iter;                   // :671-676
        TC_SAVE_WW();
        tc_get(c, P, FN_getMaxAlphas_s, run);
        tc_value(c, P, FN_getMaxAlphas_s, 0, run, v0);
        tc_value(c, P, FN_getMaxAlphas_s, 1, run, v1);
        tc_bT zero_now, ev;
        TC_FORT {
            ls[t] = LS_DONE;
            if (run[t]) {
                alphaPrimal[t] = v0[t];
                alphaDualIneq[t] = v1[t];
                if (D.coupledAlphas && alphaDualIneq[t] < alphaPrimal[t]) alphaPrimal[t] = alphaDualIneq[t];
                alphaPrimal[t] *= TC_stepback;
                alphaMax_[t] = (alphaPrimal[t] < D.alphaMax) ? alphaPrimal[t] : D.alphaMax;
                if (alphaMax_[t] >= D.alphaMin) {
                    alphaPrimal[t] = alphaMax_[t] / TC_stepback;
                    ls[t] = LS_TRYMAX;
                }
            }
            zero_now[t] = run[t] && ls[t] == LS_DONE;                       // :752-756
        }
        while (true) {
            TC_FORT {
                if (zero_now[t]) alphaPrimal[t] = 0;
                pr[t] = zero_now[t] || ls[t] != LS_DONE;
            }
            tc_set_scalar(c, P, FN_setAlphaPrimal, alphaPrimal, pr);
            TC_FORT { zero_now[t] = false; ev[t] = ls[t] != LS_DONE; }
            if (!tc_any_tile(ev)) break;
            tc_get(c, P, FN_getMinF_s, ev);
            tc_value(c, P, FN_getMinF_s, 0, ev, v0);
            TC_FORT {
                const double f = v0[t];
                if (ev[t]) switch (ls[t]) {
                case LS_TRYMAX:                                             // :693-714
                    ineq[t] = f;
                    if (isnan(ineq[t])) { status[t] = 4; run[t] = false; ls[t] = LS_DONE; sv[t] = true; }
                    else if (ineq[t] > 0) { alphaPrimal[t] *= TC_stepback; ls[t] = LS_RECHECK; }
                    else { alphaPrimal[t] = D.alphaMin / TC_stepback; ls[t] = LS_TRYMIN; }
                    break;
                case LS_RECHECK:
                    ineq1[t] = f;
                    if (ineq1[t] <= ineq[t] / 10) { alphaPrimal[t] = D.alphaMin / TC_stepback; ls[t] = LS_TRYMIN; }
                    else ls[t] = LS_DONE;
                    break;
                case LS_TRYMIN:                                             // :716-749
                    ineq[t] = f;
                    if (ineq[t] > 0) {
                        alphaPrimal[t] = alphaMax_[t] * .95;
                        if (alphaPrimal[t] >= D.alphaMin) ls[t] = LS_SCAN;
                        else { zero_now[t] = true; ls[t] = LS_DONE; }
                    } else { zero_now[t] = true; ls[t] = LS_DONE; }
                    break;
                case LS_SCAN:                                               // :722-738
                    ineq[t] = f;
                    if (ineq[t] > 0) { alphaPrimal[t] *= TC_stepback; ls[t] = LS_SCAN_RECHECK; }
                    else {
                        alphaPrimal[t] /= 2;
                        if (!(alphaPrimal[t] >= D.alphaMin)) { zero_now[t] = true; ls[t] = LS_DONE; }
                    }
                    break;
                case LS_SCAN_RECHECK:
                    ineq1[t] = f;
                    if (ineq1[t] > ineq[t] / 10) {
                        ls[t] = LS_DONE;
                        if (alphaPrimal[t] < D.alphaMin) zero_now[t] = true;    // :739-743
                    } else {
                        alphaPrimal[t] /= 2;
                        if (alphaPrimal[t] >= D.alphaMin) ls[t] = LS_SCAN;
                        else { zero_now[t] = true; ls[t] = LS_DONE; }
                    }
                    break;
                }
            }
            TC_SAVE_WW();                                                   // :699-703
        }

        TC_FORT if (run[t]) {                                               // :758-767
            if (D.coupledAlphas) {
                alphaDualEq[t] = alphaPrimal[t];
                alphaDualIneq[t] = alphaPrimal[t];
            } else {
                alphaDualIneq[t] *= TC_stepback;
                const double cap = mm ? D.alphaMax : alphaMax_[t];           // min-max: obj.alphaMax (.m:399-401)
                if (alphaDualIneq[t] > cap) alphaDualIneq[t] = cap;
                alphaDualEq[t] = eq ? D.alphaMax : alphaDualIneq[t];         // eq :611
            }
        }
        if (hasG) tc_set_scalar(c, P, FN_setAlphaDualEq, alphaDualEq, run);
        tc_set_scalar(c, P, FN_setAlphaDualIneq, alphaDualIneq, run);
        tc_copy(c, P, FN_updatePrimalDual, run);                            // :772

        // ---- barrier parameter update, skipAffine!=0 (:782-859) ----
        TC_FORT {
            bool set_mu = false, reinit_dual = false;
            if (run[t] && mm) {                                             // ipmPDminmax_CSsolver.m:411-440
                const bool th_grad = norminf_grad[t] <= TC_MAX(1e-3, 1e0 * D.gradTolerance);
                const bool th_eq = !hasG || norminf_eq[t] <= TC_MAX(1e-5, 1e0 * D.equalTolerance);
                if (alphaPrimal[t] > D.alphaMax / 2 && th_grad && th_eq) {
                    const double a = D.muFactorAggressive * mu[t], b = mu[t] * sqrt(mu[t]);   // mu^1.5
                    mu[t] = TC_MAX(muMin[t], TC_MIN(a, b));
                    set_mu = true;
                } else if (alphaPrimal[t] < .1) {
                    mu[t] = TC_MIN(mu0, 1.1 * mu[t]);
                    set_mu = true;
                    reinit_dual = true;
                } else if (alphaPrimal[t] > .99 && th_eq) {
                    mu[t] = TC_MAX(mu[t] * D.muFactorConservative, muMin[t]);
                    set_mu = true;
                }
            } else if (run[t] && D.skipAffine && eq) {                      // ipmPDeq_CSsolver.c:620-660
                const bool th_grad = norminf_grad[t] < TC_MAX(1e-3, 1e0 * D.gradTolerance);
                const bool th_eq = !hasG || norminf_eq[t] < TC_MAX(1e-5, 1e0 * D.equalTolerance);
                if (alphaPrimal[t] > D.alphaMax / 2 && th_grad && th_eq) {
                    const double r = sqrt(mu[t]);
                    mu[t] *= TC_MIN(D.muFactorAggressive, r);
                    if (mu[t] < muMin[t]) mu[t] = muMin[t];
                    set_mu = true;
                } else if (alphaPrimal[t] < .02) {
                    mu[t] *= 1.1;
                    if (mu[t] > 100 * mu0) mu[t] = 100 * mu0;
                    set_mu = true;
                    reinit_dual = true;
                } else {
                    mu[t] *= D.muFactorConservative;
                    if (mu[t] < muMin[t]) mu[t] = muMin[t];
                    set_mu = true;
                }
            } else if (run[t] && D.skipAffine) {
                const bool th_grad = norminf_grad[t] < TC_MAX(1e-6, 1e0 * D.gradTolerance);
                const bool th_eq = !hasG || norminf_eq[t] < TC_MAX(1e-5, 1e0 * D.equalTolerance);
                if (alphaPrimal[t] > alphaMax_[t] / 2 && th_grad && th_eq) {
                    const double r = sqrt(mu[t]);   // pow(mu,.5) in the reference; sqrt is its correctly rounded value
                    mu[t] *= TC_MIN(D.muFactorAggressive, r);
                    if (mu[t] < muMin[t]) mu[t] = muMin[t];
                    set_mu = true;
                } else if (alphaPrimal[t] < .1) {
                    mu[t] *= 1.1;
                    if (mu[t] > mu0) mu[t] = mu0;
                    set_mu = true;
                    reinit_dual = true;
                } else if (alphaPrimal[t] > .99 && th_eq) {
                    mu[t] *= D.muFactorConservative;
                    if (mu[t] < muMin[t]) mu[t] = muMin[t];
                    set_mu = true;
                }
            }
            pr[t] = set_mu;
            pr2[t] = reinit_dual;
        }
        tc_set_scalar(c, P, FN_setMu, mu, pr);
        tc_copy(c, P, FN_initDualIneq, pr2);                                // :816
        TC_FORT {
            pr[t] = run[t] && alphaPrimal[t] < D.alphaMin && alphaDualIneq[t] < D.alphaMin &&
                    alphaDualEq[t] < D.alphaMin;
            if (pr[t]) {                                                    // :853-859 (min-max: one division, .m:441-444)
                mu[t] /= mm ? D.muFactorConservative : (D.muFactorConservative * D.muFactorConservative);
                if (mu[t] < muMin[t]) mu[t] = muMin[t];
            }
        }
        tc_set_scalar(c, P, FN_setMu, mu, pr);
    }

    TC_FORT sv[t] = entered[t] && !run[t] && status[t] == 0 && P.save_iter == 0;   // :877-882
    TC_SAVE_WW();
#undef TC_SAVE_WW
    // ---- diagnostic bits when maxIter was reached (:885-920) ----
    tc_bT diag;
    TC_FORT diag[t] = entered[t] && !run[t] && status[t] == 8;
    if (tc_any_tile(diag)) {
        tc_get(c, P, FN_getNorminf_Grad, diag);
        tc_value(c, P, FN_getNorminf_Grad, 0, diag, v0);
        TC_FORT if (diag[t] && v0[t] > D.gradTolerance) status[t] |= 16;
        if (hasG) {
            tc_get(c, P, FN_getNorminf_G, diag);
            tc_value(c, P, FN_getNorminf_G, 0, diag, v0);
            TC_FORT if (diag[t] && v0[t] > D.equalTolerance) status[t] |= 32;
        }
        if (hasF) {
            tc_get(c, P, FN_getGapMinFMinLambda, diag);
            tc_value(c, P, FN_getGapMinFMinLambda, 0, diag, v0);
            TC_FORT if (diag[t]) {
                if (v0[t] > gapTol[t]) status[t] |= 64;
                if (mu[t] > muMin[t]) status[t] |= 128;
                const double aI = alphaDualIneq[t], aE = alphaDualEq[t];
                if (alphaPrimal[t] <= D.alphaMin && (mm ? aI <= D.alphaMin && aE <= D.alphaMin : aI < D.alphaMin && aE < D.alphaMin))
                    status[t] |= 1792;
                else if (alphaPrimal[t] <= .1 && (mm ? aI <= .1 && aE <= .1 : aI < .1 && aE < .1))
                    status[t] |= 1536;
                else if (alphaPrimal[t] <= .5 && (mm ? aI <= .5 && aE <= .5 : aI < .5 && aE < .5))
                    status[t] |= 1024;
            }
        }
        TC_FORT if (adjust && !eq && diag[t] &&
                    (aeU[t] > D.addEye2HessianUtolerance || (mm && aeD[t] > D.addEye2HessianDtolerance))) status[t] |= 2048;
    }
    TC_FORT {
        status_out[t] = status[t];
        iter_out[t] = iter[t];
        done_out[t] = entered[t] && !run[t];
    }
    if (phase_iters != 0) {            // park the locals of the instances that go on in the next phase
        TC_FORT if (entered[t]) {
            double *d = P.ps_d + pos[t];
            int32_t *q = P.ps_i + pos[t];
            const int64_t S = P.stride;
            d[PS_aeU * S] = aeU[t]; d[PS_aeEq * S] = aeEq[t]; d[PS_mu * S] = mu[t]; d[PS_muMin * S] = muMin[t];
            d[PS_gapTol * S] = gapTol[t]; d[PS_alphaMax * S] = alphaMax_[t]; d[PS_alphaPrimal * S] = alphaPrimal[t];
            d[PS_alphaDualEq * S] = alphaDualEq[t]; d[PS_alphaDualIneq * S] = alphaDualIneq[t];
            d[PS_mp * S] = mp[t]; d[PS_mn * S] = mn[t]; d[PS_aeD * S] = aeD[t];
            q[PS_iter * S] = iter[t];
            q[PS_flags * S] = (run[t] ? 1 : 0) | (updU[t] ? 2 : 0) | (updEq[t] ? 4 : 0) | (updD[t] ? 8 : 0);
        }
    }
}
