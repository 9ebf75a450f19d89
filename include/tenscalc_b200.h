/* tenscalc_b200 -- C ABI of the batched, Blackwell-native interior-point backend.
 *
 * This header is the drop-in boundary.  The reference's per-problem dynamic library
 * (built by CmexTools' createGateway from the generated C, lib/cmex2optimizeCS.m:403-433)
 * exports, with every argument passed by pointer:
 *
 *   void <class>_set_<name>(const double *input);              writeCfunctionpergroup.m:524
 *   void ipmPD_CSsolver(double *mu0, int32_t *maxIter, int32_t *saveIter,
 *                       double *addEye2Hessian, int32_t *status, int32_t *iter,
 *                       double *time);                          lib/ipmPD_CSsolver.c:132-141
 *   void <class>_getOutputs(double *y1, double *y2, ...);       writeCfunctionpergroup.m:607-635
 *
 * and keeps ONE problem instance in process-global storage.  The entry points below are
 * the same three operations on a handle that owns a BATCH of instances resident in the HBM
 * of one GPU.  Plain pointers and sizes only; every function returns 0 on success, a
 * negative TC_E* code for API misuse, or a positive cudaError_t.
 *
 * Array layout across this boundary: for a variable with n entries, a batch is
 * `double[batch][n]` (instance-major; each instance column-major like MATLAB), i.e. a
 * MATLAB array with the batch as trailing dimension.
 */
#ifndef TENSCALC_B200_H
#define TENSCALC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tc_solver tc_solver;

#define TC_EINVAL  (-1)  /* bad argument (index, batch size, null pointer) */
#define TC_EFORMAT (-2)  /* problem image not recognised                   */
#define TC_ENOGPU  (-3)  /* no CUDA device: there is no CPU fallback       */
#define TC_EMODULE (-4)  /* compiled problem module missing or mismatched  */

/* Library / device queries. */
int tc_version(void);
int tc_device_count(void);

/* Create a solver for `batch` instances of the problem described by `image` (the compiled
 * problem written by tenscalc_b200.tape.serialize; replaces loading <class>.so).  The
 * handle owns all device memory; one CUDA stream per handle; handles are independent. */
int tc_create(tc_solver **out, const void *image, size_t image_bytes, int64_t batch, int device);
int tc_destroy(tc_solver *s);

/* Attach the per-problem COMPILED module: the shared object that tenscalc_b200.build.build_module
 * produces from the sm_100a CUDA emitted for this problem (replaces the dynamic library that
 * cmex2optimizeCS compiles from generated C, lib/cmex2optimizeCS.m:403-433, and that the reference's
 * gateways dlopen per call, instructionsTableHeight4MEX.c:39-53).  After this call tc_solve* and
 * tc_get_outputs* run the module's kernels (straight-line code of every dependency group as native
 * SASS); without it they run the generic tape interpreter over the same image.  `fingerprint`
 * (may be NULL) must match the instruction-list hash the module was generated from. */
int tc_attach_module(tc_solver *s, const char *module_path, const char *fingerprint);
int tc_has_module(const tc_solver *s);

/* Introspection (what the generated class's help text lists). */
int64_t tc_batch(const tc_solver *s);
int tc_num_sets(const tc_solver *s);      /* parameters first, then optimisation variables */
int tc_num_outputs(const tc_solver *s);
int64_t tc_set_size(const tc_solver *s, int set_index);
int64_t tc_output_size(const tc_solver *s, int output_index);
int64_t tc_scratch_bytes(const tc_solver *s);
/* Name lookup.  The reference exports one symbol per parameter / optimisation variable, `<class>_set_<name>`
 * (lib/@csparse/writeCfunctionpergroup.m:524), and names its outputs (struct fields or y1, y2, ...:
 * lib/cmex2optimizeCS.m:337-355); here a set function or an output is an index, and these map <name> to it.
 * tc_find_*: index >= 0, or TC_EINVAL when the class has no such name; tc_*_name: NULL for a bad index. */
int tc_find_set(const tc_solver *s, const char *name);
int tc_find_output(const tc_solver *s, const char *name);
const char *tc_set_name(const tc_solver *s, int set_index);
const char *tc_output_name(const tc_solver *s, int output_index);

/* <class>_set_<name>: `values` is double[batch][n] on the host (tc_set) or on the device
 * (tc_set_device).  With batch_stride==0 the same n values are broadcast to all instances. */
int tc_set(tc_solver *s, int set_index, const double *values, int64_t batch_stride);
int tc_set_device(tc_solver *s, int set_index, const double *d_values, int64_t batch_stride);

/* ipmPD_CSsolver -- or ipmPDeq_CSsolver / ipmPDminmax_CSsolver, whichever driver the class was generated for
 * (lib/ipmPDeq_CSsolver.c:122-131, lib/ipmPDminmax_CSsolver.m:1) -- for every instance.  `addEye2Hessian` is
 * double[2] = {U, Eq} ({1, 2} for equilibrium classes) and double[3] = {U, D, Eq} for min-max classes.
 * Inputs by pointer as in the reference; `status` and
 * `iter` are int32[batch]; `time` receives the device time of the whole batch in seconds.
 * tc_solve copies status/iter to host memory; tc_solve_device leaves them on the device and
 * does not synchronise the host (`time` may be NULL).
 * `saveIter` (lib/ipmPD_CSsolver.c:136) only matters for a class generated with allowSave
 * (lib/cmex2optimizeCS.m:267,327): where the reference's driver calls saveWW__ -- iteration == saveIter, a
 * successful exit when saveIter == 0, a failed factorisation (:330-332,561-566,671-676,699-703,877-882) --
 * every instance keeps the values of its KKT matrix, and getOutputs returns them as one more, last, output
 * (structural non-zeros in column-major order of their subscripts).  Other classes ignore it, as the
 * reference does. */
int tc_solve(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
             const double *addEye2Hessian, int32_t *status, int32_t *iter, double *time);
int tc_solve_device(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                    const double *addEye2Hessian, int32_t *d_status, int32_t *d_iter);
/* tc_solve without the final host synchronisation: the solve and the copies of status/iter into the
 * (pinned) host arrays are only enqueued on the handle's stream; they are complete after the next
 * tc_get_outputs or tc_synchronize on this handle.  Lets a caller keep several handles in flight so
 * that the straggler tail of one batch overlaps the bulk of the next (bench.py --pipeline). */
int tc_solve_async(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                   const double *addEye2Hessian, int32_t *status, int32_t *iter);

/* Phased execution of a solve (default for the compiled back end's warp driver; TC_PHASED=0 in the
 * environment or tc_set_phased(s,0) selects the single-launch solve).  The Newton loop of
 * ipmPD_CSsolver.c:282-867 is launched in PHASES of a few iterations; between two phases the instances
 * that are still iterating are compacted into the leading scratchbook columns, so warps stay full although
 * instances need different numbers of iterations, and at the end every instance is moved back to its own
 * column.  Per instance the sequence of operations is unchanged (bit-identical results).
 * tc_set_phases: iteration budget of phase 0,1,... (the last entry repeats); n=0 restores the default. */
int tc_set_phased(tc_solver *s, int on);
int tc_set_phases(tc_solver *s, const int32_t *iterations_per_phase, int n);

/* <class>_getOutputs: y[k] is double[batch][tc_output_size(k)], caller-allocated. */
int tc_get_outputs(tc_solver *s, double *const *y);
int tc_get_outputs_device(tc_solver *s, double *const *d_y);
/* tc_get_outputs without the final host synchronisation: the gather kernels and the copies into the
 * (pinned) host arrays are only enqueued; the arrays are complete after tc_synchronize, or after an event
 * the caller records on tc_stream(s).  With tc_set + tc_solve_async + tc_get_outputs_async a caller can
 * queue the next batch on a handle before the previous one has finished (two sets of host buffers). */
int tc_get_outputs_async(tc_solver *s, double *const *y);

/* Device-resident warm start of a receding-horizon loop.  The reference's MPC examples feed outputs of
 * one solve back as the initial primal values of the next -- `out=getOutputs(obj); ...
 * setV_x(obj,out.xWarm); setV_u(obj,out.uWarm)` (examples/mpcmhe/mpc_dcmotor.m:108-126,192-225; lib/Tmpc.m) --
 * through host arrays.  tc_feedback does the same without leaving the device: it evaluates the output
 * expressions (the lazy evaluation of <class>_getOutputs) and writes output out_index[k] into set function
 * set_index[k] for every pair, invalidating what depends on it exactly as tc_set does.  All outputs are read
 * before anything is written (an output is typically a shifted view of the variable it feeds), the values are
 * exact copies, and tc_output_size(out_index[k]) must equal tc_set_size(set_index[k]). */
int tc_feedback(tc_solver *s, int n_pairs, const int32_t *out_index, const int32_t *set_index);

/* Stream control and measurement hooks used by bench.py. */
int tc_synchronize(tc_solver *s);
void *tc_stream(tc_solver *s);                 /* cudaStream_t of the handle */
int tc_last_solve_ms(tc_solver *s, float *ms); /* CUDA-event time of the last solve kernel */
int tc_kernel_launches(const tc_solver *s, int64_t *count);
int tc_total_newton_steps(tc_solver *s, int64_t *steps); /* sum of iter over the batch */
/* Profiling counters, the batched form of the reference's `profiling` option
 * (lib/@csparse/writeCfunctionpergroup.m:149-177): counts[g] = number of instance-executions
 * of dependency group g since creation / the last reset.  counts has tc_num_groups entries. */
int tc_num_groups(const tc_solver *s);
int tc_group_counts(tc_solver *s, int64_t *counts, int reset);
/* same, counting WARP-executions: counts[g]*32 / tc_group_counts[g] is the lane occupancy of group g */
int tc_group_warp_counts(tc_solver *s, int64_t *counts, int reset);
const char *tc_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Dense path: box-constrained dense least squares  min ||A x - b||^2/N  s.t. lo <= x <= hi  (BASELINE.json
 * config 5: N=4000, n=256; the problem class of examples/sls.m:136-166, whose generated class exposes
 * setP_A, setP_b, setV_x, solve, getOutputs -> {J, x}).  Same three operations as above on a handle that owns
 * a batch; the scalar instruction list is replaced by dense kernels (FP64 tensor-core Gram matrix and LDL'
 * trailing updates, csrc/tc_dense.cu) while the solver loop is ipmPD_CSsolver.c's.  n must be 8, 16, 32, 64,
 * 128 or 256.  Layouts: A double[batch][N*n] (every instance column-major, like the MATLAB array A(:,:,k)),
 * b double[batch][N], x double[batch][n].  on_device != 0: the pointer is device memory; A and b are then
 * BORROWED (no copy) and must stay valid until the next tc_dense_solve* has been enqueued. */
typedef struct tc_dense tc_dense;
int tc_dense_create(tc_dense **out, int64_t N, int64_t n, int64_t batch, double lo, double hi, int device);
int tc_dense_destroy(tc_dense *s);
int tc_dense_set_A(tc_dense *s, const double *A, int on_device);   /* <class>_set_A */
int tc_dense_set_b(tc_dense *s, const double *b, int on_device);   /* <class>_set_b */
int tc_dense_set_x(tc_dense *s, const double *x, int on_device);   /* <class>_set_x (primal initial value) */
/* ipmPD_CSsolver for every instance (arguments by pointer as in lib/ipmPD_CSsolver.c:132-141) */
int tc_dense_solve(tc_dense *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                   const double *addEye2Hessian, int32_t *status, int32_t *iter, double *time);
int tc_dense_solve_async(tc_dense *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                         const double *addEye2Hessian);
/* <class>_getOutputs: J double[batch], x double[batch][n]; status/iter int32[batch] or NULL */
int tc_dense_get_outputs(tc_dense *s, double *J, double *x, int32_t *status, int32_t *iter, int on_device);
int tc_dense_synchronize(tc_dense *s);
void *tc_dense_stream(tc_dense *s);
/* measurement hooks: CUDA-event times of the last solve (Gram kernel, Newton-loop kernel); kernel launches and
 * FP64 tensor-core flops issued by the LDL' trailing updates since creation */
int tc_dense_times_ms(tc_dense *s, float *gram_ms, float *solve_ms);
int tc_dense_counters(tc_dense *s, int64_t *launches, int64_t *dmma_flops);

#ifdef __cplusplus
}
#endif
#endif /* TENSCALC_B200_H */
