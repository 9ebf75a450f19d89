/* CPU ORACLE for BASELINE config 5 (dense box-constrained least squares) -- test infrastructure only.
 *
 *     minimise  J(x) = ||A x - b||^2 / N     subject to  LO <= x <= HI        A: N x n dense, per instance
 *
 * The reference cannot GENERATE this problem at N=4000, n=256 (its instruction table is capped,
 * lib/csparse/instructionsTableUTHash.c:171), so the oracle is what SURVEY.md 8(d) prescribes:
 * the reference's OWN unmodified driver lib/ipmPD_CSsolver.c (#included below with the #define table of
 * lib/cmex2optimizeCS.m:303-331) around HAND-WRITTEN dense versions of the functions that
 * lib/ipmPD_CStimesLambda.m:86-290 declares.  Every function below cites the lines it restates; the LDL
 * follows the recipe of lib/@csparse/sparsity_ldl.m:144-236 (v_i = L(col,i) D(i); D(col) = A(col,col) - sum;
 * L(row,col) = (A(row,col) - sum)/D(col); strictly diagonal D, no pivoting) with the elimination order
 * "the 2n inequality rows first, then the n primal variables in natural order" -- the fill-free order for this
 * KKT pattern, which is also what symamd produces for it (the 2n diagonal pivots have degree 1).
 *
 * Pinned by tests/test_dense.py: built for N=100, n=8, LO=0, HI=.05 it must agree (status, iterations,
 * solution) with the reference-generated solver of examples/sls.m (oracle/_ref sls, reference emitter +
 * reference driver) on the same seeded instances.
 *
 * Build: gcc -O1 -DDN=<N> -DDn=<n> -DLO=<lo> -DHI=<hi> [-DDRIVER_PORT] -shared -fPIC dense_ls.c
 *        (oracle/dense_oracle.py); DRIVER_PORT selects oracle/ipm_driver_port.c where /root/reference is absent.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define EXPORT __attribute__((visibility("default")))

#ifndef DN
#error "define DN (rows of A), Dn (columns), LO, HI"
#endif
#define NF (2 * Dn)

/* the #define table of lib/cmex2optimizeCS.m:303-331 with the defaults of parameters4all.m (what
 * examples/sls.m:136-166 generates) */
#define nU Dn
#define nG 0
#define nF NF
#define gradTolerance 1e-4
#define addEye2HessianUtolerance 1e-6
#define equalTolerance 1e-4
#define desiredDualityGap 1e-5
#define scaleCost 0
#define scaleInequalities 1
#define scaleEqualities 0
#define alphaMin 1e-7
#define alphaMax 1.0
#define coupledAlphas 1
#define setAddEye2Hessian 1
#define adjustAddEye2Hessian 1
#define useInertia 0
#define muFactorAggressive .2
#define muFactorConservative .95
#define delta 3
#define skipAffine 1
#define smallerNewtonMatrix 0
#define useLDL 1
#define useUmfpack 0
#define allowSave 0
#define debugConvergence 0
#define profiling 0
#define verboseLevel 1

/* ---- "scratchbook": parameters, variables and what the groups compute ----------------------- */
static double *A, *bb, *H;                 /* A[N*n] column-major, b[N], H = (2/N) A'A (parameter-only group) */
static double x0_[Dn], x[Dn], lam[NF], s4[NF];
static double mu_, eU_, eEq_, aP_, aDi_;
static double F[NF], Lfu[Dn], dU[Dn], dLam[NF], Ld[Dn * Dn], Dd[Dn], DF[NF], LF[NF];
static int H_valid, dx_valid;

static void alloc_once(void) {
    if (!A) {
        A = (double *)malloc(sizeof(double) * DN * Dn);
        bb = (double *)malloc(sizeof(double) * DN);
        H = (double *)malloc(sizeof(double) * Dn * Dn);
    }
}

/* <class>_set_A / _set_b / _set_x (writeCfunctionpergroup.m:524): store and invalidate the children */
EXPORT void dense_set_A(const double *v) { alloc_once(); memcpy(A, v, sizeof(double) * DN * Dn); H_valid = dx_valid = 0; }
EXPORT void dense_set_b(const double *v) { alloc_once(); memcpy(bb, v, sizeof(double) * DN); dx_valid = 0; }
EXPORT void dense_set_x(const double *v) { memcpy(x0_, v, sizeof(x0_)); }

/* F = scale4Ineq .* [x - LO ; HI - x]  (constraints x>=LO, x<=HI; scaling ipmPD_CStimesLambda.m:86-89) */
static void eval_F(const double *xv, double *Fv) {
    for (int i = 0; i < Dn; i++) {
        Fv[i] = s4[i] * (xv[i] - (LO));
        Fv[Dn + i] = s4[Dn + i] * ((HI) - xv[i]);
    }
}

/* Lf_u = f_u - F_u' lambda, f_u = (2/N) A'(A x - b)  (:126-141,186) */
static void eval_grad(void) {
    static double *y;
    if (!y) y = (double *)malloc(sizeof(double) * DN);
    for (int k = 0; k < DN; k++) {
        double a = 0;
        for (int j = 0; j < Dn; j++) a += A[k + (size_t)DN * j] * x[j];
        y[k] = a - bb[k];
    }
    for (int i = 0; i < Dn; i++) {
        double a = 0;
        for (int k = 0; k < DN; k++) a += A[k + (size_t)DN * i] * y[k];
        Lfu[i] = (2.0 / DN) * a - s4[i] * lam[i] + s4[Dn + i] * lam[Dn + i];
    }
}

static void eval_H(void) {             /* Lf_uu = (2/N) A'A: depends on the parameter A only */
    if (H_valid) return;
    for (int i = 0; i < Dn; i++)
        for (int j = 0; j <= i; j++) {
            double a = 0;
            for (int k = 0; k < DN; k++) a += A[k + (size_t)DN * i] * A[k + (size_t)DN * j];
            H[i + Dn * j] = H[j + Dn * i] = (2.0 / DN) * a;
        }
    H_valid = 1;
}

/* Newton direction (:235-256):
 *   WW = [ H + eU I         -F_u' diag(lambda) ]    b = [ -Lf_u           ]     F_u = [ diag(s1) ; -diag(s2) ]
 *        [ -diag(lambda)F_u  -diag(F.*lambda)  ]        [ lambda.*F - mu  ]
 * LDL' with the inequality rows eliminated first (sparsity_ldl.m:144-236), then forward substitution,
 * division by D and backward substitution (sparsity_mldivide_l1.m / _d.m / _u1.m). */
static void eval_dx(void) {
    if (dx_valid) return;
    eval_H();
    eval_F(x, F);
    eval_grad();
    /* pivots of the inequality block: D(k) = -(F_k lambda_k); L(i,k) = WW(i,k)/D(k), i = k mod n */
    for (int k = 0; k < NF; k++) {
        const double fu = (k < Dn) ? s4[k] : -s4[k];
        DF[k] = -(F[k] * lam[k]);
        LF[k] = (-(lam[k] * fu)) / DF[k];
    }
    /* dense block, left-looking by columns */
    static double v[NF + Dn];
    for (int col = 0; col < Dn; col++) {
        const int k1 = col, k2 = Dn + col;
        const double v1 = LF[k1] * DF[k1], v2 = LF[k2] * DF[k2];
        for (int i = 0; i < col; i++) v[i] = Ld[col + Dn * i] * Dd[i];
        double acc = LF[k1] * v1 + LF[k2] * v2;
        for (int i = 0; i < col; i++) acc += Ld[col + Dn * i] * v[i];
        Dd[col] = (H[col + Dn * col] + eU_) - acc;
        for (int row = col + 1; row < Dn; row++) {      /* L(row,k1), L(row,k2) are structurally zero */
            double a2 = 0;
            for (int i = 0; i < col; i++) a2 += Ld[row + Dn * i] * v[i];
            Ld[row + Dn * col] = (H[row + Dn * col] - a2) / Dd[col];
        }
    }
    /* right-hand side and the three solves */
    static double yF[NF], yU[Dn];
    for (int k = 0; k < NF; k++) yF[k] = lam[k] * F[k] - mu_;
    for (int i = 0; i < Dn; i++) {
        double acc = LF[i] * yF[i] + LF[Dn + i] * yF[Dn + i];
        for (int j = 0; j < i; j++) acc += Ld[i + Dn * j] * yU[j];
        yU[i] = -Lfu[i] - acc;
    }
    for (int k = 0; k < NF; k++) yF[k] = yF[k] / DF[k];
    for (int i = 0; i < Dn; i++) yU[i] = yU[i] / Dd[i];
    for (int i = Dn - 1; i >= 0; i--) {
        double acc = 0;
        for (int j = i + 1; j < Dn; j++) acc += Ld[j + Dn * i] * dU[j];
        dU[i] = yU[i] - acc;
    }
    for (int k = 0; k < NF; k++) dLam[k] = yF[k] - LF[k] * dU[k < Dn ? k : k - Dn];
    dx_valid = 1;
}

/* ---- the functions the driver calls (names and argument lists: lib/ipmPD_CSsolver.c:28-51) ---- */
void initPrimal__(void) { memcpy(x, x0_, sizeof(x)); dx_valid = 0; }                       /* cmex2optimizeCS.m:216-218 */
void scaleIneq__(void) {                                                                  /* :86-89 */
    for (int i = 0; i < Dn; i++) {
        s4[i] = fabs(1.0 / (x[i] - (LO)));
        s4[Dn + i] = fabs(1.0 / ((HI) - x[i]));
    }
    dx_valid = 0;
}
void setMu__(const double *m) { mu_ = *m; dx_valid = 0; }                                  /* :195 */
void setAddEye2HessianU__(const double *v) { eU_ = *v; dx_valid = 0; }                     /* :107-110 */
void setAddEye2HessianEq__(const double *v) { eEq_ = *v; }                                 /* no equality rows here */
void setAlphaPrimal__(const double *a) { aP_ = *a; }                                       /* :224-229 */
void setAlphaDualEq__(const double *a) { (void)a; }
void setAlphaDualIneq__(const double *a) { aDi_ = *a; }
void initDualIneq__(void) {                                                               /* lambda = mu./F  :198 */
    eval_F(x, F);
    for (int k = 0; k < NF; k++) lam[k] = mu_ / F[k];
    dx_valid = 0;
}
void getNorminf_Grad__(double *g) {                                                       /* :186-187 */
    eval_grad();
    double m = fabs(Lfu[0]);
    for (int i = 1; i < Dn; i++) if (m < fabs(Lfu[i])) m = fabs(Lfu[i]);
    *g = m;
}
void getGapMinFMinLambda__(double *gap, double *ineq, double *dual) {                     /* :201 */
    eval_F(x, F);
    double g = 0, mf = F[0], ml = lam[0];
    for (int k = 0; k < NF; k++) {
        g += lam[k] * F[k];
        if (mf > F[k]) mf = F[k];
        if (ml > lam[k]) ml = lam[k];
    }
    *gap = g; *ineq = mf; *dual = ml;
}
void getCurvature__(double *c) {                                                          /* dU'(WW11 dU)  :266-267 */
    eval_dx();
    double acc = 0;
    for (int i = 0; i < Dn; i++) {
        double r = eU_ * dU[i];
        for (int j = 0; j < Dn; j++) r += H[i + Dn * j] * dU[j];
        acc += dU[i] * r;
    }
    *c = acc;
}
void getDirectionError__(double *derr) {                                                  /* ||WW dx - b||inf  :258 */
    eval_dx();
    double m = 0;
    for (int i = 0; i < Dn; i++) {
        double r = eU_ * dU[i];
        for (int j = 0; j < Dn; j++) r += H[i + Dn * j] * dU[j];
        r += -(lam[i] * s4[i]) * dLam[i] + (lam[Dn + i] * s4[Dn + i]) * dLam[Dn + i];
        r -= -Lfu[i];
        if (m < fabs(r)) m = fabs(r);
    }
    for (int k = 0; k < NF; k++) {
        const int i = k < Dn ? k : k - Dn;
        const double fu = (k < Dn) ? s4[k] : -s4[k];
        double r = -(lam[k] * fu) * dU[i] - (F[k] * lam[k]) * dLam[k] - (lam[k] * F[k] - mu_);
        if (m < fabs(r)) m = fabs(r);
    }
    *derr = m;
}
void getHessInertia__(double *mp, double *mn) { *mp = 0; *mn = 0; }
void getMaxAlphas_s__(double *aPrimal, double *aDual) {                                   /* clp  :273-275 */
    eval_dx();
    double ap = DBL_MAX, ad = DBL_MAX;
    for (int k = 0; k < NF; k++) {
        const int i = k < Dn ? k : k - Dn;
        const double fu = (k < Dn) ? s4[k] : -s4[k];
        const double d = fu * dU[i];
        if (d < 0) { double t = -F[k] / d; if (ap > t) ap = t; }
        if (dLam[k] < 0) { double t = -1.0 / dLam[k]; if (ad > t) ad = t; }
    }
    *aPrimal = ap; *aDual = ad;
}
void getMinF_s__(double *ineq) {                                                          /* min F(u+alphaPrimal dU)  :276-277 */
    eval_dx();
    static double xn[Dn], Fn[NF];
    for (int i = 0; i < Dn; i++) xn[i] = x[i] + aP_ * dU[i];
    eval_F(xn, Fn);
    double m = Fn[0];
    for (int k = 1; k < NF; k++) if (m > Fn[k]) m = Fn[k];
    *ineq = m;
}
void updatePrimalDual__(void) {                                                           /* :261-281 */
    eval_dx();
    for (int i = 0; i < Dn; i++) x[i] = x[i] + aP_ * dU[i];
    for (int k = 0; k < NF; k++) lam[k] = lam[k] * (1.0 + aDi_ * dLam[k]);
    dx_valid = 0;
}
void getJ__(double *J) { *J = 0; }

#ifdef DRIVER_PORT
#include "ipm_driver_port.c"
#else
#include "/root/reference/lib/ipmPD_CSsolver.c"
#endif

/* <class>_getOutputs: {J, x} */
static void get_outputs(double *y) {
    double J = 0;
    for (int k = 0; k < DN; k++) {
        double a = 0;
        for (int j = 0; j < Dn; j++) a += A[k + (size_t)DN * j] * x[j];
        a -= bb[k];
        J += a * a;
    }
    y[0] = J / DN;
    memcpy(y + 1, x, sizeof(x));
}

/* same batch runner ABI as the generated oracles (oracle/emit_c.py): P = [A(:) ; b], V = [x0], Y = [J ; x] */
EXPORT void oracle_run_batch(int64_t B, const double *P, const double *V, double mu0, int32_t maxIter, int32_t saveIter,
                             const double *addEye, int32_t *status, int32_t *iter, double *Y, double *seconds) {
    const int64_t nP = (int64_t)DN * Dn + DN, nV = Dn, nY = 1 + Dn;
    for (int64_t q = 0; q < B; q++) {
        double mu = mu0, ae[2] = {addEye[0], addEye[1]}, t;
        int32_t mi = maxIter, si = saveIter;
        dense_set_A(P + q * nP);
        dense_set_b(P + q * nP + (int64_t)DN * Dn);
        dense_set_x(V + q * nV);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        status[q] = 0;
        ipmPD_CSsolver(&mu, &mi, &si, ae, status + q, iter + q, &t);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (seconds) seconds[q] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        get_outputs(Y + q * nY);
    }
}
