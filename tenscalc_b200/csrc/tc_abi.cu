// C ABI of the batched interior-point backend (see include/tenscalc_b200.h) and the kernels
// behind it.  Compiled for sm_100a only; there is no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <string>
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
#pragma unroll
            for (int w = 0; w < TC_GW; w++) P.group_status[w * P.stride + b] &= ~f.children[w];
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
// phased solve: compaction of the unfinished instances between two phase launches (tc_driver.cuh)
// ------------------------------------------------------------------------------------------
// What has to move with an instance: the PERSISTENT scratchbook slots -- every variable (I_set: written by
// set functions and copies: parameters, primal/dual variables, mu, step sizes, regularisation) and the outputs of
// the dependency groups that no solver-side set or copy invalidates (parameter-only subexpressions) -- plus
// the solver loop's locals (P.ps_d/ps_i).  Everything else is recomputed: the moved column's group status is
// reset to "only the static groups are valid".
struct tc_compact {
    const uint32_t *slots;       // persistent slots
    int32_t n_slots;
    uint64_t static_mask[TC_GW]; // groups that stay valid when an instance moves
    double *pool;                // [n_slots][batch]: final state of finished instances whose column was reused
    uint64_t *pool_status;       // [TC_GW][batch]: their group status (static groups only: evaluation is lazy, a
                                 // static group that nobody asked for yet is still stale and must stay so)
    uint8_t *saved;              // [batch]
    int32_t *holes, *movers;     // [stride] pairing of the current compaction
    int32_t *n_moves;
};

__global__ void tc_phase_reset_kernel(tc_device_image P, tc_compact C) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P.stride) {
        P.orig[i] = i < P.batch ? (int32_t)i : -1;
        P.ps_i[(size_t)1 * P.stride + i] = 0;
    }
    if (i < P.batch) C.saved[i] = 0;
    if (i == 0) { *P.n_live = (int32_t)P.batch; *C.n_moves = 0; }
}

// ONE CTA: ranks the columns [0, n_live) by "still running" and pairs the k-th finished column below the new
// n_live (a hole) with the k-th running column above it (a mover)
__global__ void __launch_bounds__(1024) tc_compact_plan_kernel(tc_device_image P, tc_compact C) {
    __shared__ int part[1024];
    __shared__ int s_new, s_lnn;
    const int n_old = *P.n_live, T = blockDim.x, tid = threadIdx.x;
    const int32_t *flags = P.ps_i + (size_t)1 * P.stride;            // PS_flags
    const int per = (n_old + T - 1) / T, a = min(tid * per, n_old), b = min(a + per, n_old);
    int cnt = 0;
    for (int p = a; p < b; p++) cnt += flags[p] & 1;
    part[tid] = cnt;
    __syncthreads();
    for (int o = 1; o < T; o <<= 1) {                                 // inclusive scan
        const int v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    if (tid == T - 1) s_new = part[tid];
    __syncthreads();
    const int n_new = s_new, L0 = part[tid] - cnt;                    // running columns before `a`
    int c2 = 0;                                                       // running columns in [a,b) below n_new
    for (int p = a; p < b && p < n_new; p++) c2 += flags[p] & 1;
    __syncthreads();
    part[tid] = c2;
    __syncthreads();
    for (int o = T >> 1; o; o >>= 1) {
        if (tid < o) part[tid] += part[tid + o];
        __syncthreads();
    }
    if (tid == 0) s_lnn = part[0];
    __syncthreads();
    const int lnn = s_lnn;
    int L = L0;
    for (int p = a; p < b; p++) {
        const int live = flags[p] & 1;
        if (p < n_new && !live) C.holes[p - L] = p;
        if (p >= n_new && live) C.movers[L - lnn] = p;
        L += live;
    }
    if (tid == 0) { *C.n_moves = n_new - lnn; *P.n_live = n_new; }
}

__device__ __forceinline__ void tc_column_to_pool(const tc_device_image &P, const tc_compact &C, int64_t col, int64_t id) {
    for (int k = threadIdx.x; k < C.n_slots; k += blockDim.x)
        C.pool[(size_t)k * P.batch + id] = TC_M(P, C.slots[k], col);
    for (int w = threadIdx.x; w < TC_GW; w += blockDim.x)
        C.pool_status[(size_t)w * P.batch + id] = P.group_status[(size_t)w * P.stride + col] & C.static_mask[w];
}

// one CTA per move (grid-stride): park the finished instance that sits in the hole, then move the mover in
__global__ void __launch_bounds__(256) tc_compact_move_kernel(tc_device_image P, tc_compact C) {
    const int n = *C.n_moves;
    for (int m = blockIdx.x; m < n; m += gridDim.x) {
        const int64_t h = C.holes[m], q = C.movers[m];
        const int32_t id_h = P.orig[h];
        if (id_h >= 0) tc_column_to_pool(P, C, h, id_h);
        __syncthreads();
        for (int k = threadIdx.x; k < C.n_slots; k += blockDim.x) {
            const uint32_t sl = C.slots[k];
            TC_M(P, sl, h) = TC_M(P, sl, q);
        }
        for (int k = threadIdx.x; k < TC_PS_ND; k += blockDim.x) P.ps_d[(size_t)k * P.stride + h] = P.ps_d[(size_t)k * P.stride + q];
        for (int k = threadIdx.x; k < TC_PS_NI; k += blockDim.x) P.ps_i[(size_t)k * P.stride + h] = P.ps_i[(size_t)k * P.stride + q];
        for (int w = threadIdx.x; w < TC_GW; w += blockDim.x)
            P.group_status[(size_t)w * P.stride + h] = P.group_status[(size_t)w * P.stride + q] & C.static_mask[w];
        __syncthreads();
        if (threadIdx.x == 0) {
            if (id_h >= 0) C.saved[id_h] = 1;
            P.orig[h] = P.orig[q];
            P.orig[q] = -1;
            P.ps_i[(size_t)1 * P.stride + q] = 0;                     // the vacated column holds nothing that runs
        }
        __syncthreads();
    }
}

// end of a phased solve: every instance goes back to its home column (getOutputs, and a later set +
// getOutputs without a new solve, evaluate the output groups there).  Two kernels: park every displaced
// instance, then restore everything that is parked.
__global__ void __launch_bounds__(256) tc_restore_park_kernel(tc_device_image P, tc_compact C) {
    for (int64_t col = blockIdx.x; col < P.stride; col += gridDim.x) {
        const int32_t id = P.orig[col];
        if (id >= 0 && id != col) {
            tc_column_to_pool(P, C, col, id);
            if (threadIdx.x == 0) C.saved[id] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) tc_restore_home_kernel(tc_device_image P, tc_compact C) {
    for (int64_t id = blockIdx.x; id < P.batch; id += gridDim.x) {
        if (!C.saved[id]) continue;
        for (int k = threadIdx.x; k < C.n_slots; k += blockDim.x)
            TC_M(P, C.slots[k], id) = C.pool[(size_t)k * P.batch + id];
        for (int w = threadIdx.x; w < TC_GW; w += blockDim.x)
            P.group_status[(size_t)w * P.stride + id] = C.pool_status[(size_t)w * P.batch + id];
    }
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
    std::vector<std::string> set_names, out_names;   // <name> of <class>_set_<name> / of the outputs (may be empty)
    int32_t *d_status = nullptr, *d_iter = nullptr;
    unsigned long long *d_sum = nullptr;
    double *d_stage = nullptr;           // staging for host<->device I/O
    double *d_feedback = nullptr;        // staging of tc_feedback
    size_t feedback_bytes = 0;
    size_t stage_bytes = 0;
    int64_t n_slots = 0, launches = 0;
    float last_ms = 0.f;
    // compiled back end: per-problem module (tenscalc_b200/emit_cuda.py), NULL = tape interpreter
    void *module = nullptr;
    int (*mod_solve)(const tc_device_image *, double, int, double, double, int32_t *, int32_t *, void *, int, int) = nullptr;
    // phased solve (tc_driver.cuh): compaction state and the iteration budget of every phase
    tc_compact cmp{};
    int phased = 0, mod_phased = 1;
    std::vector<int> phase_len;          // explicit schedule (tc_set_phases); empty: the default schedule
    int32_t *last_iter = nullptr;        // d_iter of the last solve (tc_total_newton_steps)
    int (*mod_eval)(const tc_device_image *, int, void *) = nullptr;
};

#define TCM_ABI 2

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
    s->stride = (batch + 1535) / 1536 * 1536;   // whole CTAs for every back end and launch shape: TC_BLOCK = 256,
                                                // 32 * lockstep (1,2,3,4,6,8,12) * instances per lane (1,2,4 | pair) divides 1536
    const int64_t n_slots = ih[1], n_groups = ih[2], n_funcs = ih[3], n_sets = ih[4], n_out = ih[5],
                  tape_cells = ih[6], n_consts = ih[7], io_words = ih[8], n_persist = ih[11];
    s->n_slots = n_slots;
    // a truncated or foreign image must fail with TC_EFORMAT, not with out-of-bounds reads
    if (n_slots <= 0 || n_groups < 0 || n_groups > TC_MAXGROUPS || n_funcs < FN_USER0 + n_sets || n_sets < 0 || n_out < 0 ||
        tape_cells < 0 || n_consts < 0 || io_words < 0 || n_persist < 0 || n_persist > n_slots || ih[9] != TC_RING ||
        ih[10] != TC_BUNDLE || tape_cells % TC_BUNDLE || ih[12] != TC_GW ||
        (uint64_t)n_funcs > image_bytes / sizeof(tc_func) || (uint64_t)tape_cells > image_bytes / sizeof(uint4) ||
        (uint64_t)io_words > image_bytes / 4 || (uint64_t)n_consts > image_bytes / 8) {
        delete s;
        return TC_EFORMAT;
    }
    auto pad16 = [](size_t b) { return (b + 15) / 16 * 16; };
    const char *p = (const char *)(dh + 16);
    const uint32_t *group_start = (const uint32_t *)p; p += pad16((n_groups + 1) * 4);
    const uint4 *tape = (const uint4 *)p;              p += pad16(tape_cells * sizeof(uint4));
    const double *consts = (const double *)p;          p += pad16(n_consts * 8);
    const tc_func *funcs = (const tc_func *)p;         p += pad16(n_funcs * sizeof(tc_func));
    const uint32_t *io = (const uint32_t *)p;          p += pad16(io_words * 4);
    const uint32_t *set_fids = (const uint32_t *)p;    p += pad16(n_sets * 4);
    const uint32_t *out_sizes = (const uint32_t *)p;   p += pad16(n_out * 4);
    const uint32_t *persist = (const uint32_t *)p;     p += pad16(n_persist * 4);
    const int64_t names_bytes = ih[13];                // 0: an image without names (tc_find_* then return TC_EINVAL)
    const char *names = p;
    if (names_bytes < 0 || (uint64_t)names_bytes > image_bytes) { delete s; return TC_EFORMAT; }
    p += pad16((size_t)names_bytes);
    bool ok = (size_t)(p - (const char *)image) <= image_bytes;
    if (ok && names_bytes > 0) {                       // n_sets + n_out NUL-terminated, non-empty names
        int64_t k = 0;
        for (const char *q = names; ok && q < names + names_bytes; k++) {
            const void *z = memchr(q, 0, (size_t)(names + names_bytes - q));
            ok = z != nullptr && z != q && k < n_sets + n_out;
            if (ok) {
                (k < n_sets ? s->set_names : s->out_names).emplace_back(q);
                q = (const char *)z + 1;
            }
        }
        ok = ok && k == n_sets + n_out;
    }
    for (int64_t g = 0; ok && g < n_groups; g++) ok = group_start[g] <= group_start[g + 1] && group_start[g + 1] <= tape_cells;
    for (int64_t f = 0; ok && f < n_funcs; f++)
        ok = (uint64_t)funcs[f].io_off + funcs[f].io_cnt <= (uint64_t)io_words;
    uint64_t out_total = 0;
    for (int64_t k = 0; ok && k < n_out; k++) out_total += out_sizes[k];
    for (int64_t k = 0; ok && k < n_sets; k++) ok = set_fids[k] >= FN_USER0 && set_fids[k] < n_funcs;
    for (int64_t k = 0; ok && k < n_persist; k++) ok = persist[k] < n_slots;
    if (ok) ok = out_total == funcs[FN_getOutputs].io_cnt;
    if (!ok) {
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
    TC_TRY((int)cudaMalloc(&d, TC_GW * s->stride * sizeof(uint64_t)));
    s->allocs.push_back(d);
    I.group_status = (uint64_t *)d;
    TC_TRY((int)cudaMemsetAsync(d, 0, TC_GW * s->stride * sizeof(uint64_t), s->stream));
    TC_TRY((int)cudaMalloc(&d, 2 * TC_MAXGROUPS * sizeof(unsigned long long)));   // instance-, then warp-executions
    s->allocs.push_back(d);
    I.group_count = (unsigned long long *)d;
    TC_TRY((int)cudaMemsetAsync(d, 0, 2 * TC_MAXGROUPS * sizeof(unsigned long long), s->stream));
    TC_TRY((int)cudaMalloc(&d, s->stride * sizeof(int32_t) * 2 + 16));
    s->allocs.push_back(d);
    TC_TRY((int)cudaMemsetAsync(d, 0, s->stride * sizeof(int32_t) * 2 + 16, s->stream));
    s->d_status = (int32_t *)d;
    s->d_iter = s->d_status + s->stride;
    s->d_sum = (unsigned long long *)(s->d_iter + s->stride);
    s->last_iter = s->d_iter;
    // phased solve: parked solver locals, column -> instance map, compaction work space, parking pool
    {
        const char *e = getenv("TC_PHASED");
        s->phased = (e && e[0] == '0') ? 0 : 1;
        const size_t st = (size_t)s->stride;
        TC_TRY((int)cudaMalloc(&d, TC_PS_ND * st * sizeof(double)));
        s->allocs.push_back(d);
        I.ps_d = (double *)d;
        TC_TRY((int)cudaMalloc(&d, (TC_PS_NI + 3) * st * sizeof(int32_t) + 64));
        s->allocs.push_back(d);
        TC_TRY((int)cudaMemsetAsync(d, 0, (TC_PS_NI + 3) * st * sizeof(int32_t) + 64, s->stream));
        I.ps_i = (int32_t *)d;
        I.orig = I.ps_i + TC_PS_NI * st;
        s->cmp.holes = I.orig + st;
        s->cmp.movers = s->cmp.holes + st;
        I.n_live = s->cmp.movers + st;
        s->cmp.n_moves = I.n_live + 1;
        TC_TRY(tc_upload(s, persist, n_persist * 4, 0, (const void **)&s->cmp.slots));
        s->cmp.n_slots = (int32_t)n_persist;
        for (int w = 0; w < TC_GW; w++) s->cmp.static_mask[w] = (uint64_t)ih[32 + w];
        TC_TRY((int)cudaMalloc(&d, (size_t)batch + 16));
        s->allocs.push_back(d);
        s->cmp.saved = (uint8_t *)d;
        s->cmp.pool = nullptr;           // allocated by the first phased solve (n_persist x batch doubles)
    }
    I.stride = s->stride;
    I.n_slots = n_slots;
    I.batch = batch;
    I.n_groups = (int32_t)n_groups;
    I.n_consts = (int32_t)n_consts;
    I.n_funcs = (int32_t)n_funcs;
    I.m_work = nullptr;
    I.queue = nullptr;
    I.work_tiles = 0;
    I.phased = 0;
    tc_defines &D = I.def;
    D.nU = (int32_t)ih[16]; D.nG = (int32_t)ih[17]; D.nF = (int32_t)ih[18];
    D.scaleCost = (int32_t)ih[19]; D.scaleInequalities = (int32_t)ih[20]; D.coupledAlphas = (int32_t)ih[21];
    D.setAddEye2Hessian = (int32_t)ih[22]; D.adjustAddEye2Hessian = (int32_t)ih[23];
    D.useInertia = (int32_t)ih[24]; D.skipAffine = (int32_t)ih[25]; D.delta = (int32_t)ih[26];
    D.driver = (int32_t)ih[27];
    if (D.driver < 0 || D.driver > 2) { tc_destroy(s); return TC_EFORMAT; }
    D.mpUdesired = (int32_t)ih[28]; D.mnDdesired = (int32_t)ih[29]; D.allowSave = (int32_t)ih[30];
    D.addEye2HessianDtolerance = dh[8];
    if (D.driver == 2 && (n_funcs <= FN_getHessUinertia || (D.setAddEye2Hessian && D.adjustAddEye2Hessian &&
                                                            !funcs[FN_getHessUinertia].io_cnt))) {
        tc_destroy(s);
        return TC_EFORMAT;
    }
    D.gradTolerance = dh[0]; D.addEye2HessianUtolerance = dh[1]; D.equalTolerance = dh[2];
    D.desiredDualityGap = dh[3]; D.alphaMin = dh[4]; D.alphaMax = dh[5];
    D.muFactorAggressive = dh[6]; D.muFactorConservative = dh[7];
    if (!D.skipAffine && D.nF > 0 && (n_funcs <= FN_getRho || !funcs[FN_getRho].io_cnt || !funcs[FN_getMinF_a].io_cnt)) {
        tc_destroy(s);                 // skipAffine=0 needs getMaxAlphas_a__/getMinF_a__/getRho__ (ipmPD_CS.m:313-315)
        return TC_EFORMAT;
    }
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
    // the per-problem module stays mapped until the process exits (its kernels are what ran: keep it visible to
    // tools that list the loaded native code after the handles are gone)
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
    auto phs = (int (*)(void))dlsym(h, "tcm_phased");        // 1: the module's driver can run a solve in phases
    auto ilv = (int (*)(void))dlsym(h, "tcm_interleave");    // 1: the two instances of a lane are interleaved (pair)
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
    s->img.ilv = ilv ? (ilv() == 1) : 0;   // scratchbook layout of the module (1: pair, 2: wide tiles); nothing has
    s->img.wide = ilv ? (ilv() == 2) : 0;  // been written yet (tc_create + attach)
    s->mod_solve = solve;
    s->mod_eval = eval;
    s->mod_phased = phs ? phs() : 0;
    auto occ = (int (*)())dlsym(h, "tcm_occupancy");
    fprintf(stderr, "tenscalc_b200: module %s attached (%s, %d CTAs of %d instances per SM)\n", path,
            s->mod_phased && s->phased ? "phased solve" : "single-launch solve", occ ? occ() : -1, blk());
    // Warm-up launch over an EMPTY batch (one CTA, every lane invalid: nothing is read or written but
    // the lanes' own group-status words): the first launch of a module makes the driver size the
    // local-memory pool for the kernel's stack frames, 0.1-0.3 s that would otherwise be charged to the
    // first solve.
    tc_device_image w = s->img;
    w.batch = 0;
    w.stride = blk();
    if (w.queue) rc = (int)cudaMemsetAsync(w.queue, 0, sizeof(unsigned long long), s->stream);
    if (!rc) rc = solve(&w, 1.0, 1, 0.0, 0.0, s->d_status, s->d_iter, (void *)s->stream, 1, 0);
    if (!rc) rc = (int)cudaStreamSynchronize(s->stream);
    return rc;
}

int tc_has_module(const tc_solver *s) { return s ? (s->module != nullptr) : TC_EINVAL; }

int64_t tc_batch(const tc_solver *s) { return s ? s->batch : TC_EINVAL; }
int tc_num_sets(const tc_solver *s) { return s ? (int)s->set_fids.size() : TC_EINVAL; }

static int tc_find_name(const std::vector<std::string> &v, const char *name) {
    for (size_t k = 0; k < v.size(); k++)
        if (v[k] == name) return (int)k;
    return TC_EINVAL;
}
// index of <class>_set_<name> / of output <name>; TC_EINVAL (negative) when there is no such name
int tc_find_set(const tc_solver *s, const char *name) { return (s && name) ? tc_find_name(s->set_names, name) : TC_EINVAL; }
int tc_find_output(const tc_solver *s, const char *name) { return (s && name) ? tc_find_name(s->out_names, name) : TC_EINVAL; }
const char *tc_set_name(const tc_solver *s, int i) {
    return (s && i >= 0 && (size_t)i < s->set_names.size()) ? s->set_names[i].c_str() : nullptr;
}
const char *tc_output_name(const tc_solver *s, int i) {
    return (s && i >= 0 && (size_t)i < s->out_names.size()) ? s->out_names[i].c_str() : nullptr;
}
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
    if (n == 0) return 0;
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

// iteration budget of phase k of a phased solve.  Default: one iteration (instances that start infeasible
// leave at once), then short phases while most instances finish, longer ones for the stragglers.
static int tc_phase_len(const tc_solver *s, int k, int covered) {
    if (!s->phase_len.empty()) return s->phase_len[k < (int)s->phase_len.size() ? k : (int)s->phase_len.size() - 1];
    if (covered == 0) return 1;
    if (covered < 64) return 2;
    if (covered < 128) return 4;
    if (covered < 256) return 8;
    if (covered < 512) return 16;
    return 32;
}

int tc_set_phases(tc_solver *s, const int32_t *lens, int n) {
    if (!s || n < 0 || (n > 0 && !lens)) return TC_EINVAL;
    for (int k = 0; k < n; k++)
        if (lens[k] <= 0) return TC_EINVAL;
    s->phase_len.assign(lens, lens + n);
    s->phased = n > 0 || s->phased;
    return 0;
}

int tc_set_phased(tc_solver *s, int on) {
    if (!s) return TC_EINVAL;
    s->phased = on != 0;
    return 0;
}

int tc_solve_device(tc_solver *s, const double *mu0, const int32_t *maxIter, const int32_t *saveIter,
                    const double *addEye2Hessian, int32_t *d_status, int32_t *d_iter) {
    if (!s || !mu0 || !maxIter || !addEye2Hessian || !d_status || !d_iter) return TC_EINVAL;
    // saveIter: when the solver was GENERATED with allowSave (cmex2optimizeCS.m:327) the reference dumps the KKT
    // matrix values of that iteration to a file (ipmPD_CSsolver.c:562-564,672-674); here every instance keeps them
    // in its save variables and getOutputs returns them as the last output (tc_driver.cuh).  Without allowSave the
    // argument is ignored, as in the reference.
    if (s->img.def.allowSave && !saveIter) return TC_EINVAL;
    s->img.save_iter = (s->img.def.allowSave && saveIter) ? *saveIter : -1;
    if (!s->module && s->img.n_consts > TC_MAXCONST) return TC_EFORMAT;   // interpreter: constant pool lives in shared memory
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA(cudaEventRecord(s->ev0, s->stream));
    s->last_iter = d_iter;
    const bool phased = s->phased && (!s->module || s->mod_phased) && !s->img.queue && s->cmp.n_slots > 0;
    // regularisers: {U, Eq} for ipmPD_CSsolver / {1, 2} for ipmPDeq_CSsolver; the min-max solver takes {U, D, Eq}
    // (ipmPDminmax_CSsolver.m:46-60)
    const bool mm = s->img.def.driver == 2;
    const double aeU = addEye2Hessian[0], aeEq = mm ? addEye2Hessian[2] : addEye2Hessian[1];
    s->img.aeD0 = mm ? addEye2Hessian[1] : 0.0;
    auto launch = [&](int first, int iters) -> int {
        if (s->module)
            return s->mod_solve(&s->img, *mu0, *maxIter, aeU, aeEq, d_status, d_iter,
                                (void *)s->stream, first, iters);
        tc_solve_kernel<<<(unsigned)(s->stride / TC_BLOCK), TC_BLOCK, 0, s->stream>>>(
            s->img, *mu0, *maxIter, aeU, aeEq, d_status, d_iter, first, iters);
        return (int)cudaGetLastError();
    };
    if (!phased) {
        s->img.phased = 0;
        if (s->img.queue) TC_CUDA(cudaMemsetAsync(s->img.queue, 0, sizeof(unsigned long long), s->stream));
        TC_CUDA((cudaError_t)launch(1, 0));
        s->launches++;
    } else {
        if (!s->cmp.pool) {
            void *d = nullptr;
            TC_CUDA(cudaMalloc(&d, ((size_t)s->cmp.n_slots + TC_GW) * (size_t)s->batch * sizeof(double)));
            s->allocs.push_back(d);
            s->cmp.pool = (double *)d;
            s->cmp.pool_status = (uint64_t *)(s->cmp.pool + (size_t)s->cmp.n_slots * (size_t)s->batch);
        }
        s->img.phased = 1;
        const unsigned nb = (unsigned)((s->stride + 255) / 256);
        tc_phase_reset_kernel<<<nb, 256, 0, s->stream>>>(s->img, s->cmp);
        s->launches++;
        // the loop body runs maxIter+1 times for an instance that hits the iteration limit (:284,308-313)
        const int64_t total = (int64_t)(*maxIter > 0 ? *maxIter : 0) + 1;
        int64_t covered = 0;
        for (int k = 0; covered < total; k++) {
            const int len = tc_phase_len(s, k, (int)covered);
            TC_CUDA((cudaError_t)launch(k == 0, len));
            covered += len;
            if (covered < total) {
                tc_compact_plan_kernel<<<1, 1024, 0, s->stream>>>(s->img, s->cmp);
                tc_compact_move_kernel<<<1184, 256, 0, s->stream>>>(s->img, s->cmp);
                s->launches += 2;
            }
            s->launches++;
        }
        tc_restore_park_kernel<<<1184, 256, 0, s->stream>>>(s->img, s->cmp);
        tc_restore_home_kernel<<<1184, 256, 0, s->stream>>>(s->img, s->cmp);
        s->launches += 2;
        TC_CUDA(cudaGetLastError());
    }
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
        if (n == 0) continue;
        dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
        tc_gather_kernel<<<grid, block, 0, s->stream>>>(s->img, s->out_offs[k], n, d_y[k]);
        s->launches++;
    }
    return (int)cudaGetLastError();
}

int tc_get_outputs_async(tc_solver *s, double *const *y) {
    if (!s || !y) return TC_EINVAL;
    TC_CUDA(cudaSetDevice(s->device));
    TC_CUDA((cudaError_t)tc_launch_eval(s, (int)FN_getOutputs));
    for (size_t k = 0; k < s->out_sizes.size(); k++) {
        const uint32_t n = s->out_sizes[k];
        if (n == 0) continue;
        dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
        tc_gather_kernel<<<grid, block, 0, s->stream>>>(s->img, s->out_offs[k], n, s->d_stage);
        s->launches++;
        TC_CUDA(cudaMemcpyAsync(y[k], s->d_stage, (size_t)n * s->batch * sizeof(double), cudaMemcpyDeviceToHost,
                                s->stream));
    }
    return (int)cudaGetLastError();
}

int tc_get_outputs(tc_solver *s, double *const *y) {
    int rc = tc_get_outputs_async(s, y);
    if (rc) return rc;
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return (int)cudaGetLastError();
}

int tc_feedback(tc_solver *s, int n_pairs, const int32_t *out_index, const int32_t *set_index) {
    if (!s || n_pairs < 0 || (n_pairs > 0 && (!out_index || !set_index))) return TC_EINVAL;
    size_t total = 0;
    for (int k = 0; k < n_pairs; k++) {
        if (out_index[k] < 0 || out_index[k] >= (int)s->out_sizes.size() || set_index[k] < 0 ||
            set_index[k] >= (int)s->set_fids.size())
            return TC_EINVAL;
        if (s->out_sizes[out_index[k]] != s->funcs[s->set_fids[set_index[k]]].io_cnt) return TC_EINVAL;
        total += s->out_sizes[out_index[k]];
    }
    if (total == 0) return 0;
    TC_CUDA(cudaSetDevice(s->device));
    const size_t bytes = total * (size_t)s->batch * sizeof(double);
    if (bytes > s->feedback_bytes) {                       // staging [pair][batch][n], grown on demand
        void *d = nullptr;
        TC_CUDA(cudaStreamSynchronize(s->stream));
        TC_CUDA(cudaMalloc(&d, bytes));
        s->allocs.push_back(d);                            // (the smaller buffer is released with the handle)
        s->d_feedback = (double *)d;
        s->feedback_bytes = bytes;
    }
    TC_CUDA((cudaError_t)tc_launch_eval(s, (int)FN_getOutputs));
    double *stage = s->d_feedback;
    for (int k = 0; k < n_pairs; k++) {                    // read every output first ...
        const uint32_t n = s->out_sizes[out_index[k]];
        if (n == 0) continue;
        dim3 grid((unsigned)((s->batch + 31) / 32), (n + 31) / 32), block(32, 8);
        tc_gather_kernel<<<grid, block, 0, s->stream>>>(s->img, s->out_offs[out_index[k]], n, stage);
        s->launches++;
        stage += (size_t)n * s->batch;
    }
    stage = s->d_feedback;
    for (int k = 0; k < n_pairs; k++) {                    // ... then write the set functions
        const uint32_t n = s->out_sizes[out_index[k]];
        if (n == 0) continue;
        int rc = tc_set_device(s, set_index[k], stage, n);
        if (rc) return rc;
        stage += (size_t)n * s->batch;
    }
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
    unsigned long long h[TC_MAXGROUPS];
    TC_CUDA(cudaMemcpyAsync(h, s->img.group_count + TC_MAXGROUPS * which, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    if (reset) TC_CUDA(cudaMemsetAsync(s->img.group_count + TC_MAXGROUPS * which, 0, sizeof(h), s->stream));
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
    tc_sum_iter_kernel<<<148, 256, 0, s->stream>>>(s->last_iter, s->batch, s->d_sum);   // iter[] of the last solve, wherever it went
    s->launches++;
    unsigned long long h = 0;
    TC_CUDA(cudaMemcpyAsync(&h, s->d_sum, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    *steps = (int64_t)h;
    return 0;
}

}  // extern "C"
