/* Thin mex gateway over the C ABI of include/tenscalc_b200.h.
 *
 * Takes the place of the per-function gateways that CmexTools' createGateway generates around
 * the reference's dynamic library (lib/cmex2optimizeCS.m:403-433, template :280-294): MATLAB
 * host code keeps calling setP_* / setV_* / solve / getOutputs on a class object; the class
 * methods (bindings/matlab/tenscalc_b200_solver.m) forward to this single mex file.
 *
 *   h = tenscalc_b200_mex('create', uint8(image), batch, device [, moduleFile])
 *       moduleFile: the per-problem compiled module (tc_attach_module); omitted = tape interpreter
 *       tenscalc_b200_mex('set', h, setIndex, values)            values: [n x batch] or [n x 1]
 *   [status,iter,time] = tenscalc_b200_mex('solve', h, mu0, int32(maxIter), int32(saveIter), addEye2Hessian)
 *   y = tenscalc_b200_mex('getOutputs', h)                       cell array, y{k} is [n_k x batch]
 *       tenscalc_b200_mex('delete', h)
 *
 * MATLAB arrays are column-major, so an [n x batch] matrix IS the ABI's double[batch][n].
 * Build (on a machine with MATLAB):  mex -I../../include tenscalc_b200_mex.c -L../../tenscalc_b200 -ltenscalc_b200
 * NOT compiled in this repository's CI: no MATLAB/mex.h in the build image (syntax-checked
 * against a stub only).
 */
#include <string.h>

#include "mex.h"
#include "tenscalc_b200.h"

static void check(int rc) {
    if (rc != 0) mexErrMsgIdAndTxt("tenscalc_b200:abi", "%s (code %d)", tc_error_string(rc), rc);
}

static tc_solver *handle(const mxArray *a) { return (tc_solver *)(uintptr_t)(*(uint64_t *)mxGetData(a)); }

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof cmd)) mexErrMsgTxt("first argument must be a command string");
    if (!strcmp(cmd, "create")) {
        tc_solver *s = NULL;
        check(tc_create(&s, mxGetData(prhs[1]), mxGetNumberOfElements(prhs[1]), (int64_t)mxGetScalar(prhs[2]),
                        nrhs > 3 ? (int)mxGetScalar(prhs[3]) : 0));
        if (nrhs > 4) {
            char path[4096];
            if (mxGetString(prhs[4], path, sizeof path)) mexErrMsgTxt("create: bad module file name");
            check(tc_attach_module(s, path, NULL));
        }
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t *)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)s;
    } else if (!strcmp(cmd, "set")) {
        tc_solver *s = handle(prhs[1]);
        int idx;
        if (mxIsChar(prhs[2])) { /* by name, like the reference's <class>_set_<name> (writeCfunctionpergroup.m:524) */
            char name[256];
            if (mxGetString(prhs[2], name, sizeof name)) mexErrMsgTxt("set: bad name");
            idx = tc_find_set(s, name);
            if (idx < 0) mexErrMsgTxt("set: the class has no parameter or variable of that name");
        } else
            idx = (int)mxGetScalar(prhs[2]);
        int64_t n = tc_set_size(s, idx), total = (int64_t)mxGetNumberOfElements(prhs[3]);
        if (total != n && total != n * tc_batch(s)) mexErrMsgTxt("set: value must have n or n*batch entries");
        check(tc_set(s, idx, mxGetPr(prhs[3]), total == n ? 0 : n));
        check(tc_synchronize(s)); /* the MATLAB array may be freed as soon as we return */
    } else if (!strcmp(cmd, "solve")) {
        tc_solver *s = handle(prhs[1]);
        double mu0 = mxGetScalar(prhs[2]), ae[2] = {1e-9, 1e-9}, time = 0;
        int32_t maxIter = (int32_t)mxGetScalar(prhs[3]), saveIter = nrhs > 4 ? (int32_t)mxGetScalar(prhs[4]) : -1;
        if (nrhs > 5 && mxGetNumberOfElements(prhs[5]) >= 2) memcpy(ae, mxGetPr(prhs[5]), sizeof ae);
        plhs[0] = mxCreateNumericMatrix((mwSize)tc_batch(s), 1, mxINT32_CLASS, mxREAL);
        mxArray *it = mxCreateNumericMatrix((mwSize)tc_batch(s), 1, mxINT32_CLASS, mxREAL);
        check(tc_solve(s, &mu0, &maxIter, &saveIter, ae, (int32_t *)mxGetData(plhs[0]), (int32_t *)mxGetData(it), &time));
        if (nlhs > 1) plhs[1] = it; else mxDestroyArray(it);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar(time);
    } else if (!strcmp(cmd, "getOutputs")) {
        tc_solver *s = handle(prhs[1]);
        int k, n = tc_num_outputs(s);
        double *y[64];
        if (n > 64) mexErrMsgTxt("too many outputs");
        plhs[0] = mxCreateCellMatrix(1, (mwSize)n);
        for (k = 0; k < n; k++) {
            mxArray *a = mxCreateDoubleMatrix((mwSize)tc_output_size(s, k), (mwSize)tc_batch(s), mxREAL);
            mxSetCell(plhs[0], (mwIndex)k, a);
            y[k] = mxGetPr(a);
        }
        check(tc_get_outputs(s, y));
    } else if (!strcmp(cmd, "delete")) {
        check(tc_destroy(handle(prhs[1])));
    } else {
        mexErrMsgTxt("unknown command");
    }
}
