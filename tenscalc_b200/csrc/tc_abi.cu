// C ABI of the batched interior-point backend (see include/tenscalc_b200.h) and the kernels
// behind it.  Compiled for sm_100a only; there is no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <vector>

#include "../../include/tenscalc_b200.h"
#include "tc_kernels.cuh"

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// classname_set_<name>: values[batch][n] -> m[slot_k][instance]  (32x32 tiles through smem so
// that both the read and the write are coalesced), then groupStatus[children]=0
__global__ void tc_set_kernel(tc_device_image P, int fn, const double *__restrict__ values,
                              int64_t batch_stride) {
    __shared__ double tile[32][33];
    const tc_func f = P.funcs[fn];
    const uint32_t n = f.io_cnt;
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const uint32_t k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {       // rows = instances, cols = entries
        int64_t b = b0 + j;
        uint32_t k = k0 + threadIdx.x;
        if (b < P.batch && k < n) tile[j][threadIdx.x] = values[b * batch_stride + k];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {       // rows = entries, cols = instances
        uint32_t k = k0 + j;
        int64_t b = b0 + threadIdx.x;
        if (b < P.batch && k < n) TC_M(P, P.io[f.io_off + k], b) = tile[threadIdx.x][j];
    }
    if (blockIdx.y == 0 && threadIdx.y == 0) {
        int64_t b = b0 + threadIdx.x;
        if (b < P.batch) {
            P.group_status[b] &= ~f.children[0];
            P.group_status[P.stride + b] &= ~f.children[1];
        }
    }
}

// gather of one output: m[src_k][instance] -> y[batch][n]
__global__ void tc_gather_kernel(tc_device_image P, uint32_t io_off, uint32_t n, double *__restrict__ y) {
    __shared__ double tile[32][33];
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const uint32_t k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        uint32_t k = k0 + j;
        int64_t b = b0 + threadIdx.x;
        if (b < P.batch && k < n) {
            uint32_t w = P.io[io_off + k];
            double v = 0.0;
            if (w != TC_NO_SLOT)
                v = (w & TC_CONST_BIT) ? P.consts[w & TC_SLOT_MASK] : TC_M(P, w & TC_SLOT_MASK, b);
            tile[j][threadIdx.x] = v;
        }
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int64_t b = b0 + j;
        uint32_t k = k0 + threadIdx.x;
        if (b < P.batch && k < n) y[b * (int64_t)n + k] = tile[threadIdx.x][j];
    }
}

__global__ void tc_sum_iter_kernel(const int32_t *iter, int64_t n, unsigned long long *out) {
    unsigned long long s = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += (unsigned long long)iter[i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(TC_FULL, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct tc_solver {
    int device = 0;
    int64_t batch = 0, stride = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    tc_device_image img{};
    std::vector<void *> allocs;
    std::vector<tc_func> funcs;          // host copy
    std::vector<uint32_t> set_fids, out_sizes, out_offs;
    int32_t *d_status = nullptr, *d_iter = nullptr;
    unsigned long long *d_sum = nullptr;
    double *d_stage = nullptr;           // staging for host<->device I/O
    size_t stage_bytes = 0;
    int64_t n_slots = 0, launches = 0;
    float last_ms = 0.f;
    // compiled back end: per-problem module (tenscalc_b200/emit_cuda.py), NULL = tape interpreter
    void *module = nullptr;
    int (*mod_solve)(const tc_device_image *, double, int, double, double, int32_t *, int32_t *, void *) = nullptr;
    int (*mod_eval)(const tc_device_image *, int, void *) = nullptr;
};

#define TCM_ABI 1

#define TC_CUDA(x)                           \
    do {                                     \
        cudaError_t e_ = (x);                \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

extern "C" {

int tc_version(void) { return 100; }

int tc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *tc_error_string(int code) {
    switch (code) {
    case 0: return "ok";
    case TC_EINVAL: return "invalid argument";
    case TC_EFORMAT: return "unrecognised problem image";
    case TC_ENOGPU: return "no CUDA device (there is no CPU fallback)";
    case TC_EMODULE: return "compiled problem module missing, unloadable or built for another problem/ABI";
    }
    return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
}

static int tc_upload(tc_solver *s, const void *src, size_t bytes, size_t pad_bytes, const void **dst) {
    void *d = nullptr;
    TC_CUDA(cudaMalloc(&d, bytes + pad_bytes + 8));
    s->allocs.push_back(d);
    TC_CUDA(cudaMemsetAsync(d, 0, bytes + pad_bytes + 8, s->stream));
    TC_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, s->stream));
    *dst = d;
    return 0;
}

int tc_create(tc_solver **out, const void *image, size_t image_bytes, int64_t batch, int device) {
    if (!out || !image || batch <= 0 || image_bytes < 64 * 8 + 16 * 8) return TC_EINVAL;
    if (tc_device_count() <= device || device < 0) return TC_ENOGPU;
    const int64_t *ih = (const int64_t *)image;
    const double *dh = (const double *)(ih + 64);
    if (ih[0] != TC_MAGIC) return TC_EFORMAT;
    tc_solver *s = new (std::nothrow) tc_solver();
    if (!s) return TC_EINVAL;
    s->device = device;
    s->batch = batch;
    s->stride = (batch + 255) / 256 * 256;   // whole CTAs for every back end (TC_BLOCK, 32*TC_LOCKSTEP <= 256)
    const int64_t n_slots = ih[1], n_groups = ih[2], n_funcs = ih[3], n_sets = ih[4], n_out = ih[5],
                  tape_cells = ih[6], n_consts = ih[7], io_words = ih[8];
    s->n_slots = n_slots;
    auto pad16 = [](size_t b) { return (b + 15) / 16 * 16; };
    const char *p = (const char *)(dh + 16);
    const uint32_t *group_start = (const uint32_t *)p; p += pad16((n_groups + 1) * 4);
    const uint4 *tape = (const uint4 *)p;              p += pad16(tape_cells * sizeof(uint4));
    const double *consts = (const double *)p;          p += pad16(n_consts * 8);
    const tc_func *funcs = (const tc_func *)p;         p += pad16(n_funcs * sizeof(tc_func));
    const uint32_t *io = (const uint32_t *)p;          p += pad16(io_words * 4);
    const uint32_t *set_fids = (const uint32_t *)p;    p += pad16(n_sets * 4);
    const uint32_t *out_sizes = (const uint32_t *)p;   p += pad16(n_out * 4);
    if ((size_t)(p - (const char *)image) > image_bytes || n_groups > 128 || ih[9] != TC_RING ||
        ih[10] != TC_BUNDLE || tape_cells % TC_BUNDLE || n_consts > TC_MAXCONST) {
        delete s;
        return TC_EFORMAT;
    }

    int rc = 0;
#define TC_TRY(x) do { rc = (x); if (rc) { tc_destroy(s); return rc; } } while (0)
    TC_TRY((int)cudaSetDevice(device));
    TC_TRY((int)cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TC_TRY((int)cudaEventCreate(&s->ev0));
    TC_TRY((int)cudaEventCreate(&s->ev1));
    tc_device_image &I = s->img;
    TC_TRY(tc_upload(s, group_start, (n_groups + 1) * 4, 0, (const void **)&I.group_start));
    TC_TRY(tc_upload(s, tape, tape_cells * sizeof(uint4), 0, (const void **)&I.tape));
    TC_TRY(tc_upload(s, consts, n_consts * 8, 0, (const void **)&I.consts));
    TC_TRY(tc_upload(s, funcs, n_funcs * sizeof(tc_func), 0, (const void **)&I.funcs));
    TC_TRY(tc_upload(s, io, io_words * 4, 0, (const void **)&I.io));
    void *d = nullptr;
    const size_t mbytes = (size_t)n_slots * (size_t)s->stride * sizeof(double);
    TC_TRY((int)cudaMalloc(&d, mbytes));
    s->allocs.push_back(d);
    I.m = (double *)d;
    TC_TRY((int)cudaMemsetAsync(d, 0, mbytes, s->stream));
    TC_TRY((int)cudaMalloc(&d, 2 * s->stride * sizeof(uint64_t)));
    s->allocs.push_back(d);
    I.group_status = (uint64_t *)d;
    TC_TRY((int)cudaMemsetAsync(d, 0, 2 * s->stride * sizeof(uint64_t), s->stream));
    TC_TRY((int)cudaMalloc(&d, 256 * sizeof(unsigned long long)));   // [0,128): instance-, [128,256): warp-executions
    s->allocs.push_back(d);
    I.group_count = (unsigned long long *)d;
    TC_TRY((int)cudaMemsetAsync(d, 0, 256 * sizeof(unsigned long long), s->stream));
    TC_TRY((int)cudaMalloc(&d, s->stride * sizeof(int32_t) * 2 + 16));
    s->allocs.push_back(d);
    s->d_status = (int32_t *)d;
    s->d_iter = s->d_status + s->stride;
    s->d_sum = (unsigned long long *)(s->d_iter + s->stride);
    I.stride = s->stride;
    I.n_slots = n_slots;
    I.batch = batch;
    I.n_groups = (int32_t)n_groups;
    I.n_consts = (int32_t)n_consts;
    I.n_funcs = (int32_t)n_funcs;
    I.m_work = nullptr;
    I.queue = nullptr;
    I.work_tiles = 0;
    tc_defines &D = I.def;
    D.nU = (int32_t)ih[16]; D.nG = (int32_t)ih[17]; D.nF = (int32_t)ih[18];
    D.scaleCost = (int32_t)ih[19]; D.scaleInequalities = (int32_t)ih[20]; D.coupledAlphas = (int32_t)ih[21];
    D.setAddEye2Hessian = (int32_t)ih[22]; D.adjustAddEye2Hessian = (int32_t)ih[23];
    D.useInertia = (int32_t)ih[24]; D.skipAffine = (int32_t)ih[25]; D.delta = (int32_t)ih[26];
    D.gradTolerance = dh[0]; D.addEye2HessianUtolerance = dh[1]; D.equalTolerance = dh[2];
    D.desiredDualityGap = dh[3]; D.alphaMin = dh[4]; D.alphaMax = dh[5];
    D.muFactorAggressive = dh[6]; D.muFactorConservative = dh[7];
    if (!D.skipAffine) { tc_destroy(s); return TC_EFORMAT; }   // Mehrotra path: not emitted by cmex2optimizeCS
    s->funcs.assign(funcs, funcs + n_funcs);
    s->set_fids.assign(set_fids, set_fids + n_sets);
    s->out_sizes.assign(out_sizes, out_sizes + n_out);
    uint32_t off = s->funcs[FN_getOutputs].io_off;
    size_t max_io = 1;
    for (int64_t k = 0; k < n_out; k++) {
        s->out_offs.push_back(off);
        off += out_sizes[k];
        if (out_sizes[k] > max_io) max_io = out_sizes[k];
    }
    for (int64_t k = 0; k < n_sets; k++)
        if (s->funcs[set_fids[k]].io_cnt > max_io) max_io = s->funcs[set_fids[k]].io_cnt;
    s->stage_bytes = max_io * (size_t)batch * sizeof(double);
    TC_TRY((int)cudaMalloc(&d, s->stage_bytes));
    s->allocs.push_back(d);
    s->d_stage = (double *)d;
    TC_TRY((int)cudaStreamSynchronize(s->stream));
#undef TC_TRY
    *out = s;
    return 0;
}

int tc_destroy(tc_solver *s) {
    if (!s) return TC_EINVAL;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void *p : s->allocs) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->module) dlclose(s->module);
    delete s;
    return 0;
}

int tc_attach_module(tc_solver *s, const char *path, const char *fingerprint) {
    if (!s || !path || s->module) return TC_EINVAL;
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "tenscalc_b200: %s\n", dlerror());
        return TC_EMODULE;
    }
    auto abi = (int (*)(void))dlsym(h, "tcm_abi_version");
    auto img = (int (*)(void))dlsym(h, "tcm_image_bytes");
    auto blk = (int (*)(void))dlsym(h, "tcm_block");
    auto fp = (const char *(*)(void))dlsym(h, "tcm_fingerprint");
    auto solve = (decltype(s->mod_solve))dlsym(h, "tcm_solve");
    auto eval = (decltype(s->mod_eval))dlsym(h, "tcm_eval");
    auto load = (int (*)(void))dlsym(h, "tcm_load");
    auto work = (int (*)(void))dlsym(h, "tcm_work_tiles");   // > 0: work-queue driver, resident warps
    if (!load || !abi || !img || !blk || !fp || !solve || !eval || abi() != TCM_ABI || img() != (int)sizeof(tc_device_image) ||
        blk() <= 0 || s->stride % blk() || (fingerprint && strcmp(fingerprint, fp()))) {
        dlclose(h);
        return TC_EMODULE;
    }
    cudaSetDevice(s->device);
    int rc = load();
    if (rc) {
        dlclose(h);
        return rc;
    }
    if (work && work() > 0) {      // work scratchbook + queue counter of the work-queue driver
        int64_t tiles = work();
        if (tiles > s->stride / TC_TILE) tiles = s->stride / TC_TILE;
        void *d = nullptr;
        rc = (int)cudaMalloc(&d, (size_t)tiles * (size_t)s->n_slots * TC_TILE * sizeof(double));
        if (!rc) {
            s->allocs.push_back(d);
            s->img.m_work = (double *)d;
            rc = (int)cudaMalloc(&d, 64);
        }
        if (!rc) {
            s->allocs.push_back(d);
            s->img.queue = (unsigned long long *)d;
            s->img.work_tiles = tiles;
        }
        if (rc) {
            dlclose(h);
            return rc;
        }
    }
    s->module = h;
    s->mod_solve = solve;
    s->mod_eval = eval;
    // Warm-up launch over an EMPTY batch (one CTA, every lane invalid: nothing is read or written but
    // the lanes' own group-status words): the first launch of a module makes the driver size the
    // local-memory pool for the kernel's stack frames, 0.1-0.3 s that would otherwise be charged to the
    // first solve.
    tc_device_image w = s->img;
    w.batch = 0;
    w.stride = blk();
    if (w.queue) rc = (int)cudaMemsetAsync(w.queue, 0, sizeof(unsigned long long), s->stream);
    if (!rc) rc = solve(&w, 1.0, 1, 0.0, 0.0, s->d_status, s->d_iter, (void *)s->stream);
    if (!rc) rc = (int)cudaStreamSynchronize(s->stream);
    return rc;
}

int tc_has_module(const tc_solver *s) { return s ? (s->module != nullptr) : TC_EINVAL; }

int64_t tc_batch(const tc_solver *s) { return s ? s->batch : TC_EINVAL; }
int tc_num_sets(const tc_solver *s) { return s ? (int)s->set_fids.size() : TC_EINVAL; }
int tc_num_outputs(const tc_solver *s) { return s ? (int)s->out_sizes.size() : TC_EINVAL; }
int64_t tc_set_size(const tc_solver *s, int i) {
    if (!s || i < 0 || i >= (int)s->set_fids.size()) return TC_EINVAL;
    return s->funcs[s->set_fids[i]].io_cnt;
}
int64_t tc_output_size(const tc_solver *s, int i) {
    if (!s || i < 0 || i >= (int)s->out_sizes.size()) return TC_EINVAL;
    return s->out_sizes[i];
}
int64_t tc_scratch_bytes(const tc_solver *s) {
    return s ? (int64_t)((size_t)s->n_slots * (size_t)s->stride * sizeof(double)) : TC_EINVAL;
}
void *tc_stream(tc_solver *s) { return s ? (void *)s->stream : nullptr; }

int tc_set_device(tc_solver *s, int i, const double *d_values, int64_t batch_stride) {
    if (!s || !d_values || i < 0 || i >= (int)s->set_fids.size()) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    const uint32_t fn = s->set_fids[i];
    const uint32_t n = s->funcs[fn].io_cnt;
    dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
    tc_set_kernel<<<grid, block, 0, s->stream>>>(s->img, (int)fn, d_values, batch_stride);
    s->launches++;
    return (int)cudaGetLastError();
}

int tc_set(tc_solver *s, int i, const double *values, int64_t batch_stride) {
    if (!s || !values || i < 0 || i >= (int)s->set_fids.size()) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    const size_t n = s->funcs[s->set_fids[i]].io_cnt;
    const size_t bytes = (batch_stride == 0 ? n : n * (size_t)s->batch) * sizeof(double);
    if (batch_stride != 0 && batch_stride != (int64_t)n) return TC_EINVAL;   // host batches are dense
    TC_CUDA(cudaMemcpyAsync(s->d_stage, values, bytes, cudaMemcpyHostToDevice, s->stream));
    return tc_set_device(s, i, s->d_stage, batch_stride);
}

int tc_solve_device(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                    const double *addEye2Hessian, int32_t *d_status, int32_t *d_iter) {
    if (!s || !mu0 || !maxIter || !addEye2Hessian || !d_status || !d_iter) return TC_EINVAL;
    (void)saveIter;   // allowSave is a code-generation-time option of the reference (cmex2optimizeCS.m:326)
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA(cudaEventRecord(s->ev0, s->stream));
    if (s->module) {
        if (s->img.queue) TC_CUDA(cudaMemsetAsync(s->img.queue, 0, sizeof(unsigned long long), s->stream));
        TC_CUDA((cudaError_t)s->mod_solve(&s->img, *mu0, *maxIter, addEye2Hessian[0], addEye2Hessian[1], d_status,
                                          d_iter, (void *)s->stream));
    } else {
        tc_solve_kernel<<<(unsigned)(s->stride / TC_BLOCK), TC_BLOCK, 0, s->stream>>>(
            s->img, *mu0, *maxIter, addEye2Hessian[0], addEye2Hessian[1], d_status, d_iter);
        TC_CUDA(cudaGetLastError());
    }
    s->launches++;
    TC_CUDA(cudaEventRecord(s->ev1, s->stream));
    return 0;
}

int tc_solve_async(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                   const double *addEye2Hessian, int32_t *status, int32_t *iter) {
    if (!s || !status || !iter) return TC_EINVAL;
    int rc = tc_solve_device(s, mu0, maxIter, saveIter, addEye2Hessian, s->d_status, s->d_iter);
    if (rc) return rc;
    TC_CUDA(cudaMemcpyAsync(status, s->d_status, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaMemcpyAsync(iter, s->d_iter, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int tc_solve(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
             const double *addEye2Hessian, int32_t *status, int32_t *iter, double *time) {
    if (!s || !status || !iter) return TC_EINVAL;
    int rc = tc_solve_device(s, mu0, maxIter, saveIter, addEye2Hessian, s->d_status, s->d_iter);
    if (rc) return rc;
    TC_CUDA(cudaMemcpyAsync(status, s->d_status, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaMemcpyAsync(iter, s->d_iter, s->batch * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    TC_CUDA(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    if (time) *time = 1e-3 * (double)s->last_ms;
    return 0;
}

static int tc_launch_eval(tc_solver *s, int fn) {
    s->launches++;
    if (s->module) return s->mod_eval(&s->img, fn, (void *)s->stream);
    tc_eval_kernel<<<(unsigned)(s->stride / TC_BLOCK), TC_BLOCK, 0, s->stream>>>(s->img, fn);
    return (int)cudaGetLastError();
}

int tc_get_outputs_device(tc_solver *s, double *const *d_y) {
    if (!s || !d_y) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA((cudaError_t)tc_launch_eval(s, (int)FN_getOutputs));
    for (size_t k = 0; k < s->out_sizes.size(); k++) {
        const uint32_t n = s->out_sizes[k];
        dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
        tc_gather_kernel<<<grid, block, 0, s->stream>>>(s->img, s->out_offs[k], n, d_y[k]);
        s->launches++;
    }
    return (int)cudaGetLastError();
}

int tc_get_outputs(tc_solver *s, double *const *y) {
    if (!s || !y) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA((cudaError_t)tc_launch_eval(s, (int)FN_getOutputs));
    for (size_t k = 0; k < s->out_sizes.size(); k++) {
        const uint32_t n = s->out_sizes[k];
        dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
        tc_gather_kernel<<<grid, block, 0, s->stream>>>(s->img, s->out_offs[k], n, s->d_stage);
        s->launches++;
        TC_CUDA(cudaMemcpyAsync(y[k], s->d_stage, (size_t)n * s->batch * sizeof(double), cudaMemcpyDeviceToHost,
                                s->stream));
    }
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return (int)cudaGetLastError();
}

int tc_synchronize(tc_solver *s) {
    if (!s) return TC_EINVAL;
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

int tc_last_solve_ms(tc_solver *s, float *ms) {
    if (!s || !ms) return TC_EINVAL;
    TC_CUDA(cudaEventSynchronize(s->ev1));
    TC_CUDA(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    *ms = s->last_ms;
    return 0;
}

int tc_kernel_launches(const tc_solver *s, int64_t *count) {
    if (!s || !count) return TC_EINVAL;
    *count = s->launches;
    return 0;
}

int tc_num_groups(const tc_solver *s) { return s ? (int)s->img.n_groups : TC_EINVAL; }

static int tc_counts(tc_solver *s, int64_t *counts, int reset, int which) {
    if (!s || !counts) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    unsigned long long h[128];
    TC_CUDA(cudaMemcpyAsync(h, s->img.group_count + 128 * which, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    if (reset) TC_CUDA(cudaMemsetAsync(s->img.group_count + 128 * which, 0, sizeof(h), s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    for (int g = 0; g < s->img.n_groups; g++) counts[g] = (int64_t)h[g];
    return 0;
}

int tc_group_counts(tc_solver *s, int64_t *counts, int reset) { return tc_counts(s, counts, reset, 0); }
int tc_group_warp_counts(tc_solver *s, int64_t *counts, int reset) { return tc_counts(s, counts, reset, 1); }

int tc_total_newton_steps(tc_solver *s, int64_t *steps) {
    if (!s || !steps) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA(cudaMemsetAsync(s->d_sum, 0, sizeof(unsigned long long), s->stream));
    tc_sum_iter_kernel<<<148, 256, 0, s->stream>>>(s->d_iter, s->batch, s->d_sum);
    s->launches++;
    unsigned long long h = 0;
    TC_CUDA(cudaMemcpyAsync(&h, s->d_sum, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    *steps = (int64_t)h;
    return 0;
}

}  // extern "C"
