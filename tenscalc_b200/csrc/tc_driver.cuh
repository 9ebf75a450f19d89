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
#pragma once
#include "tc_vm.cuh"

// constants of ipmPD_CSsolver.c:145-148,174
#define TC_addEye2HessianUMAX 1e2
#define TC_addEye2HessianEqMAX 1e2
#define TC_addEye2HessianMIN 1e-20
#define TC_maxDirectionError 1e-6
#define TC_stepback .99

#define TC_MIN(A, B) (((A) < (B)) ? (A) : (B))
#define TC_MAX(A, B) (((A) > (B)) ? (A) : (B))

// line-search automaton, one state per `getMinF_s__` evaluation of ipmPD_CSsolver.c:690-756
enum { LS_DONE = 0, LS_TRYMAX, LS_RECHECK, LS_TRYMIN, LS_SCAN, LS_SCAN_RECHECK };

__device__ __forceinline__ void tc_ipm_solve(tc_lane &c, const tc_device_image &P, bool valid, double mu0,
                                             int maxIter, double aeU0, double aeEq0, int &status_out,
                                             int &iter_out) {
    const tc_defines &D = P.def;
    const bool hasF = D.nF > 0, hasG = D.nG > 0;
    const bool adjust = D.setAddEye2Hessian && D.adjustAddEye2Hessian;

    bool run = valid;
    int status = 0, iter = 0;
    double aeU = aeU0, aeEq = aeEq0;
    bool updU = false, updEq = false;
    double mu = mu0, muMin = 0, gapTol = D.desiredDualityGap, alphaMax_ = D.alphaMax;
    double alphaPrimal = 0, alphaDualEq = 0, alphaDualIneq = 0;
    double norminf_grad = 0, norminf_eq = 0, gap = 0, ineq = 0, ineq1 = 0, dual = 0;

    if (D.setAddEye2Hessian) {                                              // :150-155
        tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, run);
        tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, run);
    }
    tc_copy(c, P, FN_initPrimal, run);                                      // :204
    if (D.scaleCost) {                                                      // :206-214
        tc_copy(c, P, FN_scaleCost, run);
        tc_get(c, P, FN_getScale4Cost, run);
        double s4c = tc_value(c, P, FN_getScale4Cost, 0, run);
        if (hasF) gapTol *= s4c;
    }
    if (hasF) {                                                             // :216-230
        if (D.scaleInequalities) tc_copy(c, P, FN_scaleIneq, run);
        tc_set_scalar(c, P, FN_setMu, mu, run);
        muMin = gapTol / D.nF / 2;
        tc_copy(c, P, FN_initDualIneq, run);
    }
    if (hasG) tc_copy(c, P, FN_initDualEqX, run);                           // :232-235

    while (TC_ANY(run)) {                                      // :282
        if (run) iter++;
        if (run && iter > maxIter) { status = 8; run = false; }             // :308-313

        // ---- exit conditions (:323-381) ----
        tc_get(c, P, FN_getNorminf_Grad, run);
        if (run) norminf_grad = tc_value(c, P, FN_getNorminf_Grad, 0, run);
        if (run && isnan(norminf_grad)) { status = 4; run = false; }
        if (hasG) {
            tc_get(c, P, FN_getNorminf_G, run);
            if (run) norminf_eq = tc_value(c, P, FN_getNorminf_G, 0, run);
        }
        if (hasF) {
            tc_get(c, P, FN_getGapMinFMinLambda, run);
            if (run) {
                gap = tc_value(c, P, FN_getGapMinFMinLambda, 0, run);
                ineq = tc_value(c, P, FN_getGapMinFMinLambda, 1, run);
                dual = tc_value(c, P, FN_getGapMinFMinLambda, 2, run);
            }
            if (run && ineq <= 0) { status = 1; run = false; }
            if (run && dual <= 0) { status = 2; run = false; }
        }
        if (run && norminf_grad <= D.gradTolerance && (!hasF || gap <= gapTol) &&
            (!hasG || norminf_eq <= D.equalTolerance) && (!adjust || aeU <= D.addEye2HessianUtolerance)) {
            status = 0;
            run = false;
        }

        // ---- regularisation control (:393-548, useInertia==0) ----
        if (adjust) {
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, run && updU);   // delayed from last iteration
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, run && updEq);
            if (run) updU = updEq = false;
            tc_get(c, P, FN_getCurvature, run);
            double curvature = tc_value(c, P, FN_getCurvature, 0, run);
            tc_get(c, P, FN_getDirectionError, run);
            double derr = tc_value(c, P, FN_getDirectionError, 0, run);
            bool nowU = false, nowEq = false;
            if (run) {
                if (curvature > 0) {                                        // :415-456
                    if (aeU > TC_addEye2HessianMIN && derr < TC_maxDirectionError) {
                        aeU = TC_MAX(.75 * aeU, TC_addEye2HessianMIN);
                        updU = true;
                    }
                    if (aeEq > TC_addEye2HessianMIN && derr < TC_maxDirectionError) {
                        aeEq = TC_MAX(.75 * aeEq, TC_addEye2HessianMIN);
                        updEq = true;
                    }
                    if (aeU < TC_addEye2HessianUMAX && derr > TC_maxDirectionError) {
                        aeU = TC_MIN(10 * aeU, TC_addEye2HessianUMAX);
                        updU = true;
                    }
                    if (aeEq < TC_addEye2HessianEqMAX && derr > TC_maxDirectionError) {
                        aeEq = TC_MIN(10 * aeEq, TC_addEye2HessianEqMAX);
                        updEq = true;
                    }
                } else if (curvature < 0) {                                 // :462-480, single pass
                    if (aeU < TC_addEye2HessianUMAX) {
                        aeU = TC_MIN(10 * TC_MAX(aeU, TC_addEye2HessianMIN), TC_addEye2HessianUMAX);
                        nowU = true;
                    }
                    if (aeEq < TC_addEye2HessianEqMAX) {
                        aeEq = TC_MIN(10 * TC_MAX(aeEq, TC_addEye2HessianMIN), TC_addEye2HessianEqMAX);
                        nowEq = true;
                    }
                }
            }
            tc_set_scalar(c, P, FN_setAddEye2HessianU, aeU, nowU);          // immediate
            tc_set_scalar(c, P, FN_setAddEye2HessianEq, aeEq, nowEq);
            // The reference re-reads curvature and direction error here (:524-531) only to
            // print them; the groups they would refresh are refreshed lazily by the next get.
        }

        if (!hasF) {                                                        // :550-569
            tc_set_scalar(c, P, FN_setAlphaPrimal, alphaMax_, run);
            if (hasG) tc_set_scalar(c, P, FN_setAlphaDualEq, alphaMax_, run);
            tc_copy(c, P, FN_updatePrimalDual, run);
            continue;
        }

        // ---- combined search direction (:679-772) ----
        tc_get(c, P, FN_getMaxAlphas_s, run);
        int ls = LS_DONE;
        if (run) {
            alphaPrimal = tc_value(c, P, FN_getMaxAlphas_s, 0, run);
            alphaDualIneq = tc_value(c, P, FN_getMaxAlphas_s, 1, run);
            if (D.coupledAlphas && alphaDualIneq < alphaPrimal) alphaPrimal = alphaDualIneq;
            alphaPrimal *= TC_stepback;
            alphaMax_ = (alphaPrimal < D.alphaMax) ? alphaPrimal : D.alphaMax;
            if (alphaMax_ >= D.alphaMin) {
                alphaPrimal = alphaMax_ / TC_stepback;
                ls = LS_TRYMAX;
            }
        }
        bool zero_now = run && ls == LS_DONE;                               // :752-756
        while (true) {
            if (zero_now) alphaPrimal = 0;
            tc_set_scalar(c, P, FN_setAlphaPrimal, alphaPrimal, zero_now || ls != LS_DONE);
            zero_now = false;
            const bool ev = ls != LS_DONE;
            if (!TC_ANY(ev)) break;
            tc_get(c, P, FN_getMinF_s, ev);
            const double f = tc_value(c, P, FN_getMinF_s, 0, ev);
            if (ev) switch (ls) {
            case LS_TRYMAX:                                                 // :693-714
                ineq = f;
                if (isnan(ineq)) { status = 4; run = false; ls = LS_DONE; }
                else if (ineq > 0) { alphaPrimal *= TC_stepback; ls = LS_RECHECK; }
                else { alphaPrimal = D.alphaMin / TC_stepback; ls = LS_TRYMIN; }
                break;
            case LS_RECHECK:
                ineq1 = f;
                if (ineq1 <= ineq / 10) { alphaPrimal = D.alphaMin / TC_stepback; ls = LS_TRYMIN; }
                else ls = LS_DONE;
                break;
            case LS_TRYMIN:                                                 // :716-749
                ineq = f;
                if (ineq > 0) {
                    alphaPrimal = alphaMax_ * .95;
                    if (alphaPrimal >= D.alphaMin) ls = LS_SCAN;
                    else { zero_now = true; ls = LS_DONE; }
                } else { zero_now = true; ls = LS_DONE; }
                break;
            case LS_SCAN:                                                   // :722-738
                ineq = f;
                if (ineq > 0) { alphaPrimal *= TC_stepback; ls = LS_SCAN_RECHECK; }
                else {
                    alphaPrimal /= 2;
                    if (!(alphaPrimal >= D.alphaMin)) { zero_now = true; ls = LS_DONE; }
                }
                break;
            case LS_SCAN_RECHECK:
                ineq1 = f;
                if (ineq1 > ineq / 10) {
                    ls = LS_DONE;
                    if (alphaPrimal < D.alphaMin) zero_now = true;          // :739-743
                } else {
                    alphaPrimal /= 2;
                    if (alphaPrimal >= D.alphaMin) ls = LS_SCAN;
                    else { zero_now = true; ls = LS_DONE; }
                }
                break;
            }
        }

        if (run) {                                                          // :758-767
            if (D.coupledAlphas) {
                alphaDualEq = alphaPrimal;
                alphaDualIneq = alphaPrimal;
            } else {
                alphaDualIneq *= TC_stepback;
                if (alphaDualIneq > alphaMax_) alphaDualIneq = alphaMax_;
                alphaDualEq = alphaDualIneq;
            }
        }
        if (hasG) tc_set_scalar(c, P, FN_setAlphaDualEq, alphaDualEq, run);
        tc_set_scalar(c, P, FN_setAlphaDualIneq, alphaDualIneq, run);
        tc_copy(c, P, FN_updatePrimalDual, run);                            // :772

        // ---- barrier parameter update, skipAffine!=0 (:782-859) ----
        bool set_mu = false, reinit_dual = false;
        if (run) {
            const bool th_grad = norminf_grad < TC_MAX(1e-6, 1e0 * D.gradTolerance);
            const bool th_eq = !hasG || norminf_eq < TC_MAX(1e-5, 1e0 * D.equalTolerance);
            if (alphaPrimal > alphaMax_ / 2 && th_grad && th_eq) {
                const double r = sqrt(mu);   // pow(mu,.5) in the reference; sqrt is its correctly rounded value
                mu *= TC_MIN(D.muFactorAggressive, r);
                if (mu < muMin) mu = muMin;
                set_mu = true;
            } else if (alphaPrimal < .1) {
                mu *= 1.1;
                if (mu > mu0) mu = mu0;
                set_mu = true;
                reinit_dual = true;
            } else if (alphaPrimal > .99 && th_eq) {
                mu *= D.muFactorConservative;
                if (mu < muMin) mu = muMin;
                set_mu = true;
            }
        }
        tc_set_scalar(c, P, FN_setMu, mu, set_mu);
        tc_copy(c, P, FN_initDualIneq, reinit_dual);                        // :816
        bool stall = run && alphaPrimal < D.alphaMin && alphaDualIneq < D.alphaMin && alphaDualEq < D.alphaMin;
        if (stall) {                                                        // :853-859
            mu /= (D.muFactorConservative * D.muFactorConservative);
            if (mu < muMin) mu = muMin;
        }
        tc_set_scalar(c, P, FN_setMu, mu, stall);
    }

    // ---- diagnostic bits when maxIter was reached (:885-920) ----
    const bool diag = valid && status == 8;
    if (TC_ANY(diag)) {
        tc_get(c, P, FN_getNorminf_Grad, diag);
        norminf_grad = tc_value(c, P, FN_getNorminf_Grad, 0, diag);
        if (diag && norminf_grad > D.gradTolerance) status |= 16;
        if (hasG) {
            tc_get(c, P, FN_getNorminf_G, diag);
            norminf_eq = tc_value(c, P, FN_getNorminf_G, 0, diag);
            if (diag && norminf_eq > D.equalTolerance) status |= 32;
        }
        if (hasF) {
            tc_get(c, P, FN_getGapMinFMinLambda, diag);
            gap = tc_value(c, P, FN_getGapMinFMinLambda, 0, diag);
            if (diag) {
                if (gap > gapTol) status |= 64;
                if (mu > muMin) status |= 128;
                if (alphaPrimal <= D.alphaMin && alphaDualIneq < D.alphaMin && alphaDualEq < D.alphaMin)
                    status |= 1792;
                else if (alphaPrimal <= .1 && alphaDualIneq < .1 && alphaDualEq < .1)
                    status |= 1536;
                else if (alphaPrimal <= .5 && alphaDualIneq < .5 && alphaDualEq < .5)
                    status |= 1024;
            }
        }
        if (adjust && diag && aeU > D.addEye2HessianUtolerance) status |= 2048;
    }
    status_out = status;
    iter_out = iter;
}
