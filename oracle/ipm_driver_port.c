/* CPU ORACLE -- test infrastructure only, never linked into or called from the product path.
 *
 * Restatement ("port") of the reference's primal-dual interior-point driver
 * lib/ipmPD_CSsolver.c:132-920: skipAffine 0 or 1 (the affine-scaling predictor :576-665),
 * useInertia 0 or 1, with or without equality/inequality constraints.  It is #included at the end of the translation unit written by
 * oracle/emit_c.py, after the #define table (lib/cmex2optimizeCS.m:303-331) and the generated
 * set/get/copy functions whose names are the reference's `*__` names.
 *
 * PINNING: tests/test_oracle.py runs this driver and the reference's own unmodified driver
 * (oracle/_ref/<name>_ref_O1.so, compiled from /root/reference/lib/ipmPD_CSsolver.c) on the
 * same seeded batches and requires identical status, identical iteration counts and
 * bit-identical outputs.
 *
 * Structure differs from the reference on purpose (helper functions instead of one 800-line
 * body; the line search is a small function returning the accepted step) but every comparison,
 * constant and call order follows the cited lines.
 */
#include <math.h>
#include <stdint.h>

#define P_UMAX 1e2      /* addEye2HessianUMAX   ipmPD_CSsolver.c:145 */
#define P_EQMAX 1e2     /* addEye2HessianEqMAX  :146 */
#define P_AEMIN 1e-20   /* addEye2HessianMIN    :147 */
#define P_MAXDERR 1e-6  /* maxDirectionError    :148 */
#define P_STEPBACK .99  /* stepback             :174 */

static double p_max(double a, double b) { return a > b ? a : b; }
static double p_min(double a, double b) { return a < b ? a : b; }

#if nF > 0
static double p_minF(double alpha) { /* F(u+alpha*du) at a trial step */
    double f;
    setAlphaPrimal__(&alpha);
    getMinF_s__(&f);
    return f;
}

/* Line search of :690-756.  Returns the accepted alphaPrimal (already set through
 * setAlphaPrimal__), or NAN when min F evaluated to NaN at the first trial (status 4). */
static double p_line_search(double aMax) {
    double a, f, f1 = 0, zero = 0;
    if (aMax >= alphaMin) {
        a = aMax / P_STEPBACK;                                   /* :693 try max */
        f = p_minF(a);
        if (isnan(f)) return NAN;                                /* :696-704 */
        if (f > 0) {                                             /* :705-711 recheck */
            a *= P_STEPBACK;
            f1 = p_minF(a);
        }
        if (!(f <= 0 || f1 <= f / 10)) return a;                 /* :712 */
        f = p_minF(alphaMin / P_STEPBACK);                       /* :715-717 try min */
        if (f > 0) {
            for (a = aMax * .95; a >= alphaMin; a /= 2) {        /* :721-738 */
                f = p_minF(a);
                if (f > 0) {
                    a *= P_STEPBACK;
                    f1 = p_minF(a);
                    if (f1 > f / 10) break;
                }
            }
            if (a >= alphaMin) return a;                         /* :739-743 */
        }
    }
    setAlphaPrimal__(&zero);                                     /* :741,747,754 */
    return 0;
}
#endif

/* allowSave: the reference dumps the KKT values with saveWW__(file) at these events (:330-332,561-566,671-676,
 * 699-703,877-882); the generated saveWW__ of the oracle copies them into the save variables getOutputs returns */
#if allowSave != 0
#define P_SAVE_WW() saveWW__("")
#else
#define P_SAVE_WW() ((void)0)
#endif

EXPORT void ipmPD_CSsolver(double *mu0, int32_t *maxIter, int32_t *saveIter, double *addEye2Hessian,
                           int32_t *status, int32_t *iter, double *time) {
    (void)saveIter;
    (void)time;
    double aeU = addEye2Hessian[0], aeEq = addEye2Hessian[1];
    int updU = 0, updEq = 0;
    double grad = 0, eq = 0, aMax = alphaMax, s4c = 0;
    double mu = *mu0, muMin = 0, gap = 0, ineq = 0, dual = 0, gapTol = desiredDualityGap;
    double aP = 0, aDe = 0, aDi = 0;
    const int adjust = (setAddEye2Hessian != 0) && (adjustAddEye2Hessian != 0);
    (void)eq; (void)s4c; (void)dual; (void)ineq; (void)gap; (void)muMin; (void)mu; (void)aDe; (void)aDi; (void)aP;
    *iter = 0;
    *status = 0;

#if setAddEye2Hessian != 0
    setAddEye2HessianU__(&aeU);                                  /* :150-155 */
    setAddEye2HessianEq__(&aeEq);
#endif
    initPrimal__();                                              /* :204 */
#if scaleCost != 0
    scaleCost__();                                               /* :206-214 */
    getScale4Cost__(&s4c);
#if nF > 0
    gapTol *= s4c;
#endif
#endif
#if nF > 0
#if scaleInequalities != 0
    scaleIneq__();                                               /* :216-221 */
#endif
    setMu__(&mu);
    muMin = gapTol / nF / 2;
    initDualIneq__();                                            /* :228-230 */
#endif
#if nG > 0
    initDualEqX__();                                             /* :232-235 */
#endif

    for (;;) {                                                   /* :282 */
        if (++*iter > *maxIter) { *status = 8; break; }          /* :284,308-313 */

        getNorminf_Grad__(&grad);                                /* :323-335 */
        if (isnan(grad)) { *status = 4; P_SAVE_WW(); break; }   /* :326-334 */
#if nG > 0
        getNorminf_G__(&eq);                                     /* :337-339 */
#endif
#if nF > 0
        getGapMinFMinLambda__(&gap, &ineq, &dual);               /* :343-357 */
        if (ineq <= 0) { *status = 1; break; }
        if (dual <= 0) { *status = 2; break; }
#endif
        if (grad <= gradTolerance                                /* :362-381 */
#if nF > 0
            && gap <= gapTol
#endif
#if nG > 0
            && eq <= equalTolerance
#endif
            && (!adjust || aeU <= addEye2HessianUtolerance)) {
            *status = 0;
            break;
        }

        if (adjust) {                                            /* :393-548 */
            double curv, derr;
#if useInertia != 0
            double mp = 0, mn = 0;                               /* read only after being set (:408-419) */
#endif
            if (updU) { setAddEye2HessianU__(&aeU); updU = 0; }  /* delayed from last iteration */
            if (updEq) { setAddEye2HessianEq__(&aeEq); updEq = 0; }
            getCurvature__(&curv);
#if useInertia != 0
            if (curv <= 0) getHessInertia__(&mp, &mn);           /* :408-413 */
#endif
            getDirectionError__(&derr);
            if (curv > 0
#if useInertia != 0
                || (mp == nU && mn == nF + nG)                   /* mpDesired, mnDesired :164-169 */
#endif
                ) {                                              /* :415-456 */
                if (aeU > P_AEMIN && derr < P_MAXDERR) { aeU = p_max(.75 * aeU, P_AEMIN); updU = 1; }
                if (aeEq > P_AEMIN && derr < P_MAXDERR) { aeEq = p_max(.75 * aeEq, P_AEMIN); updEq = 1; }
                if (aeU < P_UMAX && derr > P_MAXDERR) { aeU = p_min(10 * aeU, P_UMAX); updU = 1; }
                if (aeEq < P_EQMAX && derr > P_MAXDERR) { aeEq = p_min(10 * aeEq, P_EQMAX); updEq = 1; }
            } else if (curv < 0) {                               /* :458-520: one pass, applied now */
                int change = 0;
#if useInertia == 0
                const double fU = 10, fEq = 10;
                const int act = 1;
#else
                /* not enough positive eigenvalues: U x10, Eq x2; not enough negative ones: U x2, Eq x10 */
                const int act = (mp < nU) || (mn < nF + nG);
                const double fU = (mp < nU) ? 10 : 2, fEq = (mp < nU) ? 2 : 10;
#endif
                if (act && aeU < P_UMAX) { aeU = p_min(fU * p_max(aeU, P_AEMIN), P_UMAX); setAddEye2HessianU__(&aeU); change = 1; }
                if (act && aeEq < P_EQMAX) { aeEq = p_min(fEq * p_max(aeEq, P_AEMIN), P_EQMAX); setAddEye2HessianEq__(&aeEq); change = 1; }
                if (change) {                                    /* :522-531 */
                    getCurvature__(&curv);
#if useInertia != 0
                    if (curv <= 0) getHessInertia__(&mp, &mn);
#endif
                    getDirectionError__(&derr);
                }
            }
        }

#if nF == 0
        setAlphaPrimal__(&aMax);                                 /* :550-569 */
#if nG > 0
        setAlphaDualEq__(&aMax);
#endif
        if (*iter == *saveIter) P_SAVE_WW();                     /* :561-566 */
        updatePrimalDual__();
#else
#if skipAffine == 0
        {                                                        /* affine-scaling predictor :576-665 */
            double zero = 0, sigma;
            getMaxAlphas_a__(&aP, &aDi);
            aMax = (aP < alphaMax) ? aP : alphaMax;
            if (aDi < aMax) aMax = aDi;
            if (aMax >= alphaMin) {
                aP = aMax;                                       /* try max */
                setAlphaPrimal__(&aP);
                getMinF_a__(&ineq);
                if (ineq < 0) {
                    aP = alphaMin;                               /* try min */
                    setAlphaPrimal__(&aP);
                    getMinF_a__(&ineq);
                    if (ineq > 0) {
                        for (aP = aMax * .95; aP >= alphaMin; aP /= 2) {
                            setAlphaPrimal__(&aP);
                            getMinF_a__(&ineq);
                            if (ineq >= 0) break;
                        }
                        if (aP < alphaMin) { aP = 0; setAlphaPrimal__(&zero); }
                    } else { aP = 0; setAlphaPrimal__(&zero); }
                }
            } else { aP = 0; setAlphaPrimal__(&zero); }
            setAlphaDualIneq__(&aP);
            if (aP > alphaMax / 2
#if nG > 0
                && (eq < 100 * equalTolerance || eq < 1e-3)
#endif
                ) {                                              /* :627-660 */
                getRho__(&sigma);
                if (sigma > 1) sigma = 1;
                if (sigma < 0) sigma = 0;
#if delta == 2
                sigma = sigma * sigma;
#else
                sigma = sigma * sigma * sigma;
#endif
                mu = sigma * gap / nF;
                if (mu < muMin) mu = muMin;
                setMu__(&mu);
            }
        }
#endif
        if (*iter == *saveIter) P_SAVE_WW();                     /* :671-676 */
        getMaxAlphas_s__(&aP, &aDi);                             /* :679 */
#if coupledAlphas != 0
        if (aDi < aP) aP = aDi;                                  /* :681-684 */
#endif
        aP *= P_STEPBACK;
        aMax = (aP < alphaMax) ? aP : alphaMax;                  /* :688 */
        aP = p_line_search(aMax);
        if (isnan(aP)) { *status = 4; P_SAVE_WW(); break; }      /* :696-704 */
#if coupledAlphas != 0
        aDe = aP;                                                /* :758-767 */
        aDi = aP;
#else
        aDi *= P_STEPBACK;
        if (aDi > aMax) aDi = aMax;
        aDe = aDi;
#endif
#if nG > 0
        setAlphaDualEq__(&aDe);
#endif
        setAlphaDualIneq__(&aDi);
        updatePrimalDual__();                                    /* :772 */

#if skipAffine != 0
        {                                                        /* barrier update, skipAffine!=0 :782-850 */
            int th_grad = grad < p_max(1e-6, 1e0 * gradTolerance);
            int th_eq = 1;
#if nG > 0
            th_eq = eq < p_max(1e-5, 1e0 * equalTolerance);
#endif
            if (aP > aMax / 2 && th_grad && th_eq) {
                mu *= p_min(muFactorAggressive, pow(mu, .5));
                if (mu < muMin) mu = muMin;
                setMu__(&mu);
            } else if (aP < .1) {
                mu *= 1.1;
                if (mu > *mu0) mu = *mu0;
                setMu__(&mu);
                initDualIneq__();
            } else if (aP > .99 && th_eq) {
                mu *= muFactorConservative;
                if (mu < muMin) mu = muMin;
                setMu__(&mu);
            }
        }
#endif
        if (aP < alphaMin && aDi < alphaMin && aDe < alphaMin) { /* :853-859 */
            mu /= (muFactorConservative * muFactorConservative);
            if (mu < muMin) mu = muMin;
            setMu__(&mu);
        }
#endif
    }

    if (*saveIter == 0 && *status == 0) P_SAVE_WW();             /* :877-882 */
    if (*status == 8) {                                          /* :885-920 */
        getNorminf_Grad__(&grad);
        if (grad > gradTolerance) *status |= 16;
#if nG > 0
        getNorminf_G__(&eq);
        if (eq > equalTolerance) *status |= 32;
#endif
#if nF > 0
        getGapMinFMinLambda__(&gap, &ineq, &dual);
        if (gap > gapTol) *status |= 64;
        if (mu > muMin) *status |= 128;
        if (aP <= alphaMin && aDi < alphaMin && aDe < alphaMin) *status |= 1792;
        else if (aP <= .1 && aDi < .1 && aDe < .1) *status |= 1536;
        else if (aP <= .5 && aDi < .5 && aDe < .5) *status |= 1024;
#endif
        if (adjust && aeU > addEye2HessianUtolerance) *status |= 2048;
    }
}
