/* CPU ORACLE, test infrastructure only.  PARITY UNPINNED: this is a C transcription of the reference's MATLAB
 * solver lib/ipmPDminmax_CSsolver.m:1-587 (its C namesake lib/ipmPDminmax_CSsolver.c is stale -- it calls functions
 * ipmPDminmax_CS.m no longer declares, SURVEY.md Note A -- and MATLAB is not available here), so nothing of the
 * reference's own can be run against it.  What it gives the tests is a second, independent, sequential
 * implementation of the same specification to hold the device driver against (status, iterations, outputs).
 * Same calling convention as ipmPD_CSsolver(); addEye2Hessian = {U, D, Eq} (:46-60).  Line numbers: the .m file. */
#include <math.h>
#include <stdint.h>

extern void initPrimal__(void);
extern void scaleIneq__(void);
extern void scaleCost__(void);
extern void getScale4Cost__(double *);
extern void setMu__(const double *);
extern void initDualIneq__(void);
extern void initDualEq__(void);
extern void setAddEye2HessianU__(const double *);
extern void setAddEye2HessianD__(const double *);
extern void setAddEye2HessianEq__(const double *);
extern void getNorminf_Grad__(double *);
extern void getNorminf_G__(double *);
extern void getGapMinFMinLambda__(double *, double *, double *);
extern void getDirectionError__(double *);
extern void getHessDinertia__(double *, double *);
extern void getHessUinertia__(double *, double *);
extern void getMaxAlphas_s__(double *, double *);
extern void setAlphaPrimal__(const double *);
extern void setAlphaDualEq__(const double *);
extern void setAlphaDualIneq__(const double *);
extern void getMinF_s__(double *);
extern void updatePrimalDual__(void);

#define MM_MIN(A, B) (((A) < (B)) ? (A) : (B))
#define MM_MAX(A, B) (((A) > (B)) ? (A) : (B))

EXPORT void ipmPDminmax_CSsolver(double *mu0, int32_t *maxIter, int32_t *saveIter, double *addEye2Hessian,
                                 int32_t *status_out, int32_t *iter_out, double *time) {
    const double UMAX = 1e2, DMAX = 1e2, EqMAX = 1e2, MINV = 1e-20, maxDirectionError = 1e-7;     /* :37-43 */
    double aeU = addEye2Hessian[0], aeD = addEye2Hessian[1], aeEq = addEye2Hessian[2];
    int updU = 0, updD = 0, updEq = 0;
    int status = 0, iter = 0;
    double desiredGap = desiredDualityGap, mu = *mu0, muMin = 0, alphaPrimal = 0, alphaDualEq = 0, alphaDualIneq = 0;
    double norminf_grad = 0, norminf_eq = 0, gap = 0, ineq = 0, ineq1 = 0, dual = 0, derr, mpD, mnD, mpU, mnU;
    (void)saveIter; (void)DMAX;
#if setAddEye2Hessian != 0
    setAddEye2HessianU__(&aeU); setAddEye2HessianD__(&aeD); setAddEye2HessianEq__(&aeEq);          /* :61-63 */
#endif
    initPrimal__();                                                                               /* :80 */
#if scaleCost != 0
    scaleCost__();
    { double s4c; getScale4Cost__(&s4c); desiredGap = s4c * desiredDualityGap; }
#endif
#if nF > 0
#if scaleInequalities != 0
    scaleIneq__();
#endif
    setMu__(&mu);
    muMin = desiredGap / nF / 2;
    initDualIneq__();
#endif
#if nG > 0
    initDualEq__();
#endif
    while (1) {
        iter++;
        if (iter > *maxIter) { status = 8; break; }                                               /* :172-176 */
        getNorminf_Grad__(&norminf_grad);
        if (isnan(norminf_grad)) { status = 4; break; }
#if nG > 0
        getNorminf_G__(&norminf_eq);
#endif
#if nF > 0
        getGapMinFMinLambda__(&gap, &ineq, &dual);
        if (ineq <= 0) { status = 1; break; }
        if (dual <= 0) { status = 2; break; }
#endif
        if (norminf_grad <= gradTolerance && (nF == 0 || gap <= desiredGap) && (nG == 0 || norminf_eq <= equalTolerance) &&
            (!setAddEye2Hessian || !adjustAddEye2Hessian ||
             (aeU <= addEye2HessianUtolerance && aeD <= addEye2HessianDtolerance))) {              /* :211-219 */
            status = 0;
            break;
        }
#if (setAddEye2Hessian != 0) && (adjustAddEye2Hessian != 0)
        if (updU) { setAddEye2HessianU__(&aeU); updU = 0; }                                       /* :225-236 */
        if (updD) { setAddEye2HessianD__(&aeD); updD = 0; }
        if (updEq) { setAddEye2HessianEq__(&aeEq); updEq = 0; }
        for (int ii = 0; ii < 30; ii++) {                                                         /* :238-301 */
            getDirectionError__(&derr);
            getHessDinertia__(&mpD, &mnD);
            getHessUinertia__(&mpU, &mnU);
            if (mpU == mpUdesired && mnD == mnDdesired && derr <= maxDirectionError) {
                if (aeEq > MINV) { aeEq = MM_MAX(.75 * aeEq, MINV); updEq = 1; }
                if (aeU > MINV) { aeU = MM_MAX(.75 * aeU, MINV); updU = 1; }
                if (aeD > MINV) { aeD = MM_MAX(.75 * aeD, MINV); updD = 1; }
                break;
            } else {
                int change = 0;
                if (mpU < mpUdesired && (double)updU < UMAX) {        /* :281 compares the FLAG with the maximum */
                    aeU = MM_MIN(10 * aeU, UMAX);
                    setAddEye2HessianU__(&aeU);
                    change = 1;
                }
                if (mnD < mnDdesired && (double)updD < DMAX) {
                    aeD = MM_MIN(10 * aeD, DMAX);
                    setAddEye2HessianD__(&aeD);
                    change = 1;
                }
                if (derr > maxDirectionError && aeEq < EqMAX) {
                    aeEq = MM_MIN(10 * aeEq, EqMAX);
                    setAddEye2HessianEq__(&aeEq);
                    change = 1;
                }
                if (!change) break;
            }
        }
#endif
#if nF == 0
        { double a = alphaMax; setAlphaPrimal__(&a);                                              /* :318-324 */
#if nG > 0
          setAlphaDualEq__(&a);
#endif
          updatePrimalDual__(); }
#else
        getMaxAlphas_s__(&alphaPrimal, &alphaDualIneq);                                           /* :331-404 */
#if coupledAlphas != 0
        if (alphaDualIneq < alphaPrimal) alphaPrimal = alphaDualIneq;
#endif
        const double stepback = .99;
        alphaPrimal = stepback * alphaPrimal;
        const double alphaMaxL = MM_MIN(alphaPrimal, alphaMax);
        if (alphaMaxL >= alphaMin) {
            alphaPrimal = alphaMaxL / stepback;
            setAlphaPrimal__(&alphaPrimal); getMinF_s__(&ineq);
            if (isnan(ineq)) { status = 4; break; }
            if (ineq > 0) { alphaPrimal = stepback * alphaPrimal; setAlphaPrimal__(&alphaPrimal); getMinF_s__(&ineq1); }
            if (ineq <= 0 || ineq1 <= ineq / 10) {
                alphaPrimal = alphaMin / stepback;
                setAlphaPrimal__(&alphaPrimal); getMinF_s__(&ineq);
                if (ineq > 0) {
                    alphaPrimal = alphaMaxL * .95;
                    while (alphaPrimal >= alphaMin) {
                        setAlphaPrimal__(&alphaPrimal); getMinF_s__(&ineq);
                        if (ineq > 0) {
                            alphaPrimal = stepback * alphaPrimal;
                            setAlphaPrimal__(&alphaPrimal); getMinF_s__(&ineq1);
                            if (ineq1 > ineq / 10) break;
                        }
                        alphaPrimal = alphaPrimal / 2;
                    }
                    if (alphaPrimal < alphaMin) { alphaPrimal = 0; setAlphaPrimal__(&alphaPrimal); }
                } else { alphaPrimal = 0; setAlphaPrimal__(&alphaPrimal); }
            }
        } else { alphaPrimal = 0; setAlphaPrimal__(&alphaPrimal); }
#if coupledAlphas != 0
        alphaDualEq = alphaPrimal;
        alphaDualIneq = alphaPrimal;
#else
        alphaDualIneq = stepback * alphaDualIneq;
        if (alphaDualIneq > alphaMax) alphaDualIneq = alphaMax;
        alphaDualEq = alphaDualIneq;
#endif
#if nG > 0
        setAlphaDualEq__(&alphaDualEq);
#endif
        setAlphaDualIneq__(&alphaDualIneq);
        updatePrimalDual__();
        {                                                                                         /* :411-444 */
            const int th_grad = norminf_grad <= MM_MAX(1e-3, 1e0 * gradTolerance);
            const int th_eq = (nG == 0) || (norminf_eq <= MM_MAX(1e-5, 1e0 * equalTolerance));
            if (alphaPrimal > alphaMax / 2 && th_grad && th_eq) {
                mu = MM_MAX(muMin, MM_MIN(muFactorAggressive * mu, mu * sqrt(mu)));               /* mu^1.5 */
                setMu__(&mu);
            } else if (alphaPrimal < .1) {
                mu = MM_MIN(*mu0, 1.1 * mu);
                setMu__(&mu);
                initDualIneq__();
            } else if (alphaPrimal > .99 && th_eq) {
                mu = MM_MAX(mu * muFactorConservative, muMin);
                setMu__(&mu);
            }
            if (alphaPrimal < alphaMin && alphaDualIneq < alphaMin && alphaDualEq < alphaMin) {
                mu = MM_MAX(mu / muFactorConservative, muMin);
                setMu__(&mu);
            }
        }
#endif
    }
    if (status == 8) {                                                                            /* :476-508 */
        getNorminf_Grad__(&norminf_grad);
        if (norminf_grad > gradTolerance) status |= 16;
#if nG > 0
        getNorminf_G__(&norminf_eq);
        if (norminf_eq > equalTolerance) status |= 32;
#endif
#if nF > 0
        getGapMinFMinLambda__(&gap, &ineq, &dual);
        if (gap > desiredGap) status |= 64;
        if (mu > muMin) status |= 128;
        if (alphaPrimal <= alphaMin && alphaDualIneq <= alphaMin && alphaDualEq <= alphaMin) status |= 1792;
        else if (alphaPrimal <= .1 && alphaDualIneq <= .1 && alphaDualEq <= .1) status |= 1536;
        else if (alphaPrimal <= .5 && alphaDualIneq <= .5 && alphaDualEq <= .5) status |= 1024;
#endif
        if (setAddEye2Hessian && adjustAddEye2Hessian && (aeU > addEye2HessianUtolerance || aeD > addEye2HessianDtolerance))
            status |= 2048;
    }
    *status_out = status;
    *iter_out = iter;
    if (time) *time = 0;
}
