// libcu2b engine: device data layout, kernel dispatch, the device-resident training loop
// (replaces training.cu:21-217) and the C-ABI entry points that touch the GPU.
//
// HBM layout of a session
//   P  [rows x kp] fp32, Q [cols x kp] fp32, kp = n_factors rounded up to 4 (16-byte rows)
//   user_bias [rows], item_bias [cols] fp32
//   train / test matrices: indptr [rows+1] int32 + COO triplets (user,item,rating) 12 B each in
//     CSR order (the loss stream; the per-user sampler gathers (item,rating) from it)
//   update stream: [max_batch_segs x seg_pitch] triplets written by the sampler, consumed by
//     the SGD kernel through TMA bulk copies
//   DevState + metric log + loss partials: a few KB
// There is no CPU fallback anywhere in this file.

#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <climits>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "blocked_kernels.cuh"
#include "cu2b_internal.h"
#include "dsgd_kernels.cuh"
#include "loss_kernels.cuh"
#include "sgd_kernels.cuh"
#include "tiled_kernels.cuh"

using namespace cu2b;

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return cu2b_fail(CU2B_ERR_CUDA, "Cuda Error: %s (%s:%d)", cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                             \
    } while (0)

#define CU2B_TRY(expr)                      \
    do {                                    \
        cu2b_status s__ = (expr);           \
        if (s__ != CU2B_OK) return s__;     \
    } while (0)

namespace {

// CU2B_TRACE=1: wall-clock phase timings of the host-side session life cycle on stderr.
struct Trace {
    bool on;
    double t0, last;
    const char *what;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit Trace(const char *w) : on(getenv("CU2B_TRACE") != nullptr), t0(now()), last(t0), what(w) {}
    void mark(const char *phase) {
        if (!on) return;
        cudaDeviceSynchronize();
        const double t = now();
        fprintf(stderr, "[cu2b trace] %s: %-28s %8.2f ms (total %8.2f)\n", what, phase, t - last, t - t0);
        last = t;
    }
};

// ---------------------------------------------------------------------------------------
// small RAII pool: everything allocated through it is released when it goes out of scope
// ---------------------------------------------------------------------------------------
// The library's own stream-ordered memory pool of a device (one per device and process, created on first use).
// It is kept warm -- release threshold = never -- so that a train() call does not pay cudaMalloc / cudaFree of
// several GB every time; the host application's default pool and its cudaMalloc heap are left alone, and
// cu2b_release_cache() hands the cached memory back.
cudaMemPool_t *library_pools() {
    static cudaMemPool_t pools[64] = {nullptr};
    return pools;
}
cudaMemPool_t library_pool(int dev) {
    if (dev < 0 || dev >= 64) return nullptr;
    cudaMemPool_t *pools = library_pools();
    if (pools[dev]) return pools[dev];
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
    if (!supported) return nullptr;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t mp = nullptr;
    if (cudaMemPoolCreate(&mp, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    unsigned long long never = ~0ULL;
    cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &never);
    pools[dev] = mp;
    return mp;
}

struct DevPool {
    std::vector<void *> ptrs;
    // Stream-ordered allocation from the library's pool. DSGD contexts export their buffers through CUDA IPC
    // and therefore use plain cudaMalloc (async == false).
    bool async = false;
    cudaStream_t stream = nullptr;
    cudaMemPool_t mem_pool = nullptr;
    ~DevPool() { release(); }
    void use_async(cudaStream_t st) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return;
        mem_pool = library_pool(dev);
        if (!mem_pool) return;
        async = true;
        stream = st;
    }
    void release() {
        for (void *p : ptrs) {
            if (async) cudaFreeAsync(p, stream); else cudaFree(p);
        }
        ptrs.clear();
    }
    template <typename T>
    cu2b_status alloc(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = std::max<size_t>(16, count * sizeof(T));
        cudaError_t e = async ? cudaMallocFromPoolAsync(&p, bytes, mem_pool, stream) : cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            *out = nullptr;
            cudaGetLastError();
            return cu2b_fail(e == cudaErrorMemoryAllocation ? CU2B_ERR_NOMEM : CU2B_ERR_CUDA,
                             "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        ptrs.push_back(p);
        *out = (T *)p;
        return CU2B_OK;
    }
    void free_one(void *p) {
        auto it = std::find(ptrs.begin(), ptrs.end(), p);
        if (it != ptrs.end()) {
            if (async) cudaFreeAsync(p, stream); else cudaFree(p);
            ptrs.erase(it);
        }
    }
};

struct DevMatrix {
    int rows = 0, cols = 0;
    long long nnz = 0;
    int *indptr = nullptr;        // rows + 1
    cu2b_rating *coo = nullptr;   // nnz, padded to a multiple of 4 (+kChunkMax slack)
};

// thread per rating: user = last row whose indptr <= j
__global__ void __launch_bounds__(256)
expand_coo_kernel(const int *__restrict__ indptr, int rows, const int *__restrict__ indices,
                  const float *__restrict__ data, long long nnz, cu2b_rating *__restrict__ coo,
                  const int *__restrict__ item_pos) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz;
         j += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = rows;  // find the largest u in [0, rows) with indptr[u] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(indptr + mid) <= j) lo = mid; else hi = mid;
        }
        cu2b_rating r;
        r.user = lo;
        r.item = __ldg(indices + j);
        if (item_pos) r.item = __ldg(item_pos + r.item);  // the session's internal row order (ItemPlacement)
        r.rating = __ldg(data + j);
        coo[j] = r;
    }
}

// What the engine needs to know about a device, queried once per process and device
// (cudaGetDeviceProperties costs milliseconds; a train() call should not pay it twice).
struct DeviceFacts {
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    bool known = false;
};
cu2b_status device_facts(int dev, DeviceFacts *out) {
    static DeviceFacts cache[64];
    if (dev >= 0 && dev < 64 && cache[dev].known) { *out = cache[dev]; return CU2B_OK; }
    DeviceFacts f;
    CUDA_TRY(cudaDeviceGetAttribute(&f.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&f.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&f.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    f.known = true;
    if (dev >= 0 && dev < 64) cache[dev] = f;
    *out = f;
    return CU2B_OK;
}

cu2b_status check_device() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    DeviceFacts f;
    CU2B_TRY(device_facts(dev, &f));
    if (f.cc_major != 10)
        return cu2b_fail(CU2B_ERR_CUDA, "libcu2b is built for sm_100a only; device %d is sm_%d%d", dev, f.cc_major, f.cc_minor);
    return CU2B_OK;
}

cu2b_status validate_csr(const cu2b_csr *m, const char *what) {
    if (!m || !m->indptr || m->rows < 0 || m->cols < 0 || m->nonzeros < 0 ||
        (m->nonzeros > 0 && (!m->indices || !m->data)))
        return cu2b_fail(CU2B_ERR_INVALID, "%s: malformed cu2b_csr", what);
    return CU2B_OK;
}

// Orders `waiter` after everything enqueued on `signaller` so far (one reusable event).
cu2b_status stream_after(cudaStream_t waiter, cudaStream_t signaller, cudaEvent_t ev) {
    CUDA_TRY(cudaEventRecord(ev, signaller));
    CUDA_TRY(cudaStreamWaitEvent(waiter, ev, 0));
    return CU2B_OK;
}

// A CSR matrix on its way to the device: allocate, copy (H2D, or adopt device arrays), expand to
// COO triplets. The three phases are separate so that session creation can allocate everything
// first, stream all copies through one copy stream and run the expansions on the compute stream
// while the DMA engine is already moving the next buffer.
struct MatrixUpload {
    const cu2b_csr *m = nullptr;
    DevMatrix *out = nullptr;
    int *tmp_i = nullptr;
    float *tmp_d = nullptr;
};

// reuse: `out` already owns buffers of exactly this shape (session reload); only the staging
// buffers are allocated.
cu2b_status matrix_alloc(DevPool &pool, const cu2b_csr *m, DevMatrix *out, MatrixUpload *up, bool reuse = false) {
    CU2B_TRY(validate_csr(m, "upload_matrix"));
    const size_t nnz = (size_t)m->nonzeros;
    up->m = m;
    up->out = out;
    if (reuse) {
        if (out->rows != m->rows || out->cols != m->cols || out->nnz != m->nonzeros)
            return cu2b_fail(CU2B_ERR_INVALID, "reload: the matrix shape (%d x %d, %lld ratings) differs from the one the "
                             "session was created with (%d x %d, %lld)", m->rows, m->cols, (long long)m->nonzeros,
                             out->rows, out->cols, out->nnz);
    } else {
        out->rows = m->rows;
        out->cols = m->cols;
        out->nnz = m->nonzeros;
        CU2B_TRY(pool.alloc(&out->indptr, (size_t)m->rows + 1));
        CU2B_TRY(pool.alloc(&out->coo, nnz + kChunkMax + 4));
    }
    if (!m->on_device && nnz > 0) {
        CU2B_TRY(pool.alloc(&up->tmp_i, nnz));
        CU2B_TRY(pool.alloc(&up->tmp_d, nnz));
    }
    return CU2B_OK;
}

// part 0 = row pointers + item indices, part 1 = rating values, -1 = both
cu2b_status matrix_copy(cudaStream_t cst, const MatrixUpload &up, int part = -1) {
    const cu2b_csr *m = up.m;
    const size_t nnz = (size_t)m->nonzeros;
    const cudaMemcpyKind kind = m->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (part != 1) {
        CUDA_TRY(cudaMemcpyAsync(up.out->indptr, m->indptr, ((size_t)m->rows + 1) * sizeof(int), kind, cst));
        if (up.tmp_i) CUDA_TRY(cudaMemcpyAsync(up.tmp_i, m->indices, nnz * sizeof(int), cudaMemcpyHostToDevice, cst));
    }
    if (part != 0 && up.tmp_d) CUDA_TRY(cudaMemcpyAsync(up.tmp_d, m->data, nnz * sizeof(float), cudaMemcpyHostToDevice, cst));
    return CU2B_OK;
}

cu2b_status matrix_expand(DevPool &pool, cudaStream_t st, MatrixUpload &up, const int *item_pos = nullptr) {
    const cu2b_csr *m = up.m;
    const size_t nnz = (size_t)m->nonzeros;
    if (nnz == 0) return CU2B_OK;
    const int *indices = up.tmp_i ? up.tmp_i : m->indices;
    const float *data = up.tmp_d ? up.tmp_d : m->data;
    const int grid = (int)std::min<size_t>((nnz + 255) / 256, 148 * 16);
    expand_coo_kernel<<<grid, 256, 0, st>>>(up.out->indptr, m->rows, indices, data, (long long)nnz, up.out->coo, item_pos);
    CUDA_TRY(cudaGetLastError());
    if (up.tmp_i) {
        if (!pool.async) CUDA_TRY(cudaStreamSynchronize(st));  // stream-ordered frees need no sync
        pool.free_one(up.tmp_i);
        pool.free_one(up.tmp_d);
        up.tmp_i = nullptr;
        up.tmp_d = nullptr;
    }
    return CU2B_OK;
}

cu2b_status host_indptr(const cu2b_csr *m, cudaStream_t st, std::vector<int> *indptr_host) {
    indptr_host->resize((size_t)m->rows + 1);
    if (m->on_device) {
        CUDA_TRY(cudaMemcpyAsync(indptr_host->data(), m->indptr, ((size_t)m->rows + 1) * sizeof(int),
                                 cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        memcpy(indptr_host->data(), m->indptr, ((size_t)m->rows + 1) * sizeof(int));
    }
    return CU2B_OK;
}

// Single-stream form (kernel-level entry points).
cu2b_status upload_matrix(DevPool &pool, cudaStream_t st, const cu2b_csr *m, DevMatrix *out,
                          std::vector<int> *indptr_host) {
    MatrixUpload up;
    CU2B_TRY(matrix_alloc(pool, m, out, &up));
    CU2B_TRY(matrix_copy(st, up));
    if (indptr_host) CU2B_TRY(host_indptr(m, st, indptr_host));
    return matrix_expand(pool, st, up);
}

// dense [rows x k] host matrix <-> [rows x kp] device matrix
cu2b_status upload_dense(cudaStream_t st, float *dst, const float *src, int rows, int k, int kp) {
    if (rows == 0) return CU2B_OK;
    if (kp == k) {  // no padding: one contiguous copy
        CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)rows * k * sizeof(float), cudaMemcpyHostToDevice, st));
        return CU2B_OK;
    }
    CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)rows * kp * sizeof(float), st));
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)kp * sizeof(float), src, (size_t)k * sizeof(float),
                               (size_t)k * sizeof(float), (size_t)rows, cudaMemcpyHostToDevice, st));
    return CU2B_OK;
}
cu2b_status download_dense(cudaStream_t st, float *dst, const float *src, int rows, int k, int kp) {
    if (rows == 0) return CU2B_OK;
    if (kp == k) {
        CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)rows * k * sizeof(float), cudaMemcpyDeviceToHost, st));
        return CU2B_OK;
    }
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)k * sizeof(float), src, (size_t)kp * sizeof(float),
                               (size_t)k * sizeof(float), (size_t)rows, cudaMemcpyDeviceToHost, st));
    return CU2B_OK;
}

// ---------------------------------------------------------------------------------------
// kernel dispatch on the lane layout (L lanes per rating, V float4 per lane)
// ---------------------------------------------------------------------------------------
typedef void (*SgdKernel)(const SgdParams);
typedef void (*LossKernel)(const LossParams);
typedef void (*BlockedKernel)(const BlockedParams);
typedef void (*UserRunKernel)(const UserRunParams);
typedef void (*UserTileKernel)(const UserTileParams);
typedef void (*UserRoundKernel)(const UserRoundParams);

cu2b_status layout_for(int kp, int *L, int *V) {
    const int vecs = kp / 4;
    if (vecs < 1 || vecs > 128) return cu2b_fail(CU2B_ERR_UNSUPPORTED, "n_factors must be in [1, 512]");
    int l = 1;
    while (l < vecs && l < 32) l <<= 1;
    *L = l;
    *V = (vecs + l - 1) / l;
    return CU2B_OK;
}

// Hogwild kernel variants. WMODE 3 (item AND user rows updated with 128-bit L2 atomic adds) is
// the default: on B200 the red.v4.f32 path is 2-3x faster than read-modify-write stores to the
// same rows (profiles/r1_sweep1.jsonl) and no SGD step is lost under contention. One rating per
// group in flight (UNR 1) with >= 5 CTAs/SM beat deeper unrolling at lower occupancy for
// k = 128 (profiles/r1_sweep2.jsonl). CU2B_WMODE=0..3 switches the write path for A/B profiling.
template <int W>
SgdKernel pick_sgd_t(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_hogwild<1, 1, 2, W>;
        case 2: return mf_sgd_hogwild<2, 1, 2, W>;
        case 4: return mf_sgd_hogwild<4, 1, 2, W>;
        case 8: return mf_sgd_hogwild<8, 1, 2, W>;
        case 16: return mf_sgd_hogwild<16, 1, 2, W>;
        default:
            switch (V) {
                case 1: return mf_sgd_hogwild<32, 1, 1, W, 5>;
                case 2: return mf_sgd_hogwild<32, 2, 1, W>;
                case 3: return mf_sgd_hogwild<32, 3, 1, W>;
                default: return mf_sgd_hogwild<32, 4, 1, W>;
            }
    }
}
SgdKernel pick_sgd(int L, int V) {
    int w = 3;
    if (const char *e = getenv("CU2B_WMODE")) w = atoi(e) & 3;
    switch (w) {
        case 0: return pick_sgd_t<0>(L, V);
        case 1: return pick_sgd_t<1>(L, V);
        case 2: return pick_sgd_t<2>(L, V);
        default: return pick_sgd_t<3>(L, V);
    }
}
// The loss kernel is issue-bound, not memory-bound (profiles/r1_loss_summary.txt: SM 70 %, L2
// 23 %), so it uses a denser lane layout than the update kernel: up to four float4 per lane,
// i.e. 4 ratings per warp at k = 128 (L = 8, V = 4) -- fewer shuffles and fewer instructions per
// rating, and the 4 ratings of a pass usually share the user, so their P loads coalesce.
void loss_layout_for(int kp, int *L, int *V) {
    const int vecs = kp / 4, need = (vecs + 3) / 4;
    int l = 1;
    while (l < need) l <<= 1;
    *L = l;
    *V = (vecs + l - 1) / l;
}

template <int L>
LossKernel pick_loss_wide(int V, int U) {
    if (V == 3) return U >= 4 ? mf_loss_fused<L, 3, 4> : U == 2 ? mf_loss_fused<L, 3, 2> : mf_loss_fused<L, 3, 1>;
    return U >= 4 ? mf_loss_fused<L, 4, 4> : U == 2 ? mf_loss_fused<L, 4, 2> : mf_loss_fused<L, 4, 1>;
}

// Ratings in flight per lane group (loss_kernels.cuh): kLossUnroll by default, CU2B_LOSS_UNROLL = 1 | 2 | 4 for A/B runs.
constexpr int kLossUnroll = 1;

LossKernel pick_loss(int kp) {
    int L, V;
    loss_layout_for(kp, &L, &V);
    int U = kLossUnroll;
    if (const char *e = getenv("CU2B_LOSS_UNROLL")) U = atoi(e);
    switch (L) {
        case 1:
            switch (V) {
                case 1: return mf_loss_fused<1, 1>;
                case 2: return mf_loss_fused<1, 2>;
                case 3: return mf_loss_fused<1, 3>;
                default: return mf_loss_fused<1, 4>;
            }
        case 2: return pick_loss_wide<2>(V, U);
        case 4: return pick_loss_wide<4>(V, U);
        case 8: return pick_loss_wide<8>(V, U);
        case 16: return pick_loss_wide<16>(V, U);
        default: return pick_loss_wide<32>(V, U);
    }
}

BlockedKernel pick_blocked(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_blocked_round<1, 1>;
        case 2: return mf_sgd_blocked_round<2, 1>;
        case 4: return mf_sgd_blocked_round<4, 1>;
        case 8: return mf_sgd_blocked_round<8, 1>;
        case 16: return mf_sgd_blocked_round<16, 1>;
        default:
            switch (V) {
                case 1: return mf_sgd_blocked_round<32, 1>;
                case 2: return mf_sgd_blocked_round<32, 2>;
                case 3: return mf_sgd_blocked_round<32, 3>;
                default: return mf_sgd_blocked_round<32, 4>;
            }
    }
}

UserRunKernel pick_user_runs(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_user_runs<1, 1>;
        case 2: return mf_sgd_user_runs<2, 1>;
        case 4: return mf_sgd_user_runs<4, 1>;
        case 8: return mf_sgd_user_runs<8, 1>;
        case 16: return mf_sgd_user_runs<16, 1>;
        default:
            switch (V) {
                case 1: return mf_sgd_user_runs<32, 1>;
                case 2: return mf_sgd_user_runs<32, 2>;
                case 3: return mf_sgd_user_runs<32, 3>;
                default: return mf_sgd_user_runs<32, 4>;
            }
    }
}

// fused wait + sub-epoch + hand-off instantiations (multi-GPU DSGD, one rank per device)
template <bool THIN>
UserRunKernel pick_user_runs_linked(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_user_runs<1, 1, THIN, true>;
        case 2: return mf_sgd_user_runs<2, 1, THIN, true>;
        case 4: return mf_sgd_user_runs<4, 1, THIN, true>;
        case 8: return mf_sgd_user_runs<8, 1, THIN, true>;
        case 16: return mf_sgd_user_runs<16, 1, THIN, true>;
        default:
            switch (V) {
                case 1: return mf_sgd_user_runs<32, 1, THIN, true>;
                case 2: return mf_sgd_user_runs<32, 2, THIN, true>;
                case 3: return mf_sgd_user_runs<32, 3, THIN, true>;
                default: return mf_sgd_user_runs<32, 4, THIN, true>;
            }
    }
}

UserRunKernel pick_user_runs_thin(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_user_runs<1, 1, true>;
        case 2: return mf_sgd_user_runs<2, 1, true>;
        case 4: return mf_sgd_user_runs<4, 1, true>;
        case 8: return mf_sgd_user_runs<8, 1, true>;
        case 16: return mf_sgd_user_runs<16, 1, true>;
        default:
            switch (V) {
                case 1: return mf_sgd_user_runs<32, 1, true>;
                case 2: return mf_sgd_user_runs<32, 2, true>;
                case 3: return mf_sgd_user_runs<32, 3, true>;
                default: return mf_sgd_user_runs<32, 4, true>;
            }
    }
}

UserTileKernel pick_user_tiles(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_user_tiles<1, 1>;
        case 2: return mf_sgd_user_tiles<2, 1>;
        case 4: return mf_sgd_user_tiles<4, 1>;
        case 8: return mf_sgd_user_tiles<8, 1>;
        case 16: return mf_sgd_user_tiles<16, 1>;
        default:
            switch (V) {
                case 1: return mf_sgd_user_tiles<32, 1>;
                case 2: return mf_sgd_user_tiles<32, 2>;
                case 3: return mf_sgd_user_tiles<32, 3>;
                default: return mf_sgd_user_tiles<32, 4>;
            }
    }
}

// pf = 1: one item row of look-ahead per lane group (see mf_sgd_user_rounds).
template <int PF, int MINB>
UserRoundKernel pick_user_rounds_t(int L, int V) {
    switch (L) {
        case 1: return mf_sgd_user_rounds<1, 1, PF, MINB>;
        case 2: return mf_sgd_user_rounds<2, 1, PF, MINB>;
        case 4: return mf_sgd_user_rounds<4, 1, PF, MINB>;
        case 8: return mf_sgd_user_rounds<8, 1, PF, MINB>;
        case 16: return mf_sgd_user_rounds<16, 1, PF, MINB>;
        default:
            switch (V) {
                case 1: return mf_sgd_user_rounds<32, 1, PF, MINB>;
                case 2: return mf_sgd_user_rounds<32, 2, PF, MINB>;
                case 3: return mf_sgd_user_rounds<32, 3, PF, MINB>;
                default: return mf_sgd_user_rounds<32, 4, PF, MINB>;
            }
    }
}
UserRoundKernel pick_user_rounds(int L, int V, int pf, int minb) {
    if (pf) return minb >= 8 ? pick_user_rounds_t<1, 8>(L, V) : minb >= 6 ? pick_user_rounds_t<1, 6>(L, V) : minb == 5 ? pick_user_rounds_t<1, 5>(L, V) : pick_user_rounds_t<1, 4>(L, V);
    return minb >= 8 ? pick_user_rounds_t<0, 8>(L, V) : minb >= 6 ? pick_user_rounds_t<0, 6>(L, V) : minb == 5 ? pick_user_rounds_t<0, 5>(L, V) : pick_user_rounds_t<0, 4>(L, V);
}

// Host side of the deterministic mode: stable counting sort of the ratings by
// (round, user block) where round = (item block - user block) mod B.
struct BlockSchedule {
    int B = 0;
    std::vector<cu2b_rating> sched;
    std::vector<int> ptr;  // B*B + 1
};

int auto_blocks(int rows, int cols) { return std::max(1, std::min(std::min(rows, cols), 2048)); }

cu2b_status build_block_schedule(const cu2b_rating *coo, int64_t n, int rows, int cols, int B, BlockSchedule *out,
                                 int64_t *order = nullptr) {
    if (B < 1 || (int64_t)B * B > (1LL << 28)) return cu2b_fail(CU2B_ERR_INVALID, "n_blocks must be in [1, 16384]");
    const int ubs = std::max(1, (rows + B - 1) / B), ibs = std::max(1, (cols + B - 1) / B);
    out->B = B;
    out->ptr.assign((size_t)B * B + 1, 0);
    auto bucket = [&](const cu2b_rating &r) {
        const int ub = r.user / ubs, ib = r.item / ibs;
        const int s = ((ib - ub) % B + B) % B;
        return (size_t)s * B + ub;
    };
    for (int64_t t = 0; t < n; ++t) out->ptr[bucket(coo[t]) + 1]++;
    for (size_t b = 0; b < (size_t)B * B; ++b) out->ptr[b + 1] += out->ptr[b];
    std::vector<int> cursor(out->ptr.begin(), out->ptr.end() - 1);
    out->sched.resize((size_t)n);
    if (order) {
        for (int64_t t = 0; t < n; ++t) order[cursor[bucket(coo[t])]++] = t;
        for (int64_t t = 0; t < n; ++t) out->sched[(size_t)t] = coo[order[t]];
    } else {
        for (int64_t t = 0; t < n; ++t) out->sched[(size_t)cursor[bucket(coo[t])]++] = coo[t];
    }
    return CU2B_OK;
}

// Host version of item_draw_weight_kernel: w[i] = sum over every `user_stride`-th user u that rated i of
// floor(2^32 / deg(u)) -- the expected draws of item i per iteration under per-user sampling (one uniform draw
// per user per iteration, sgd.cu:27-37) in units of 2^-32. Exact integers: the device kernel, whatever its
// summation order, produces the same numbers.
std::vector<unsigned long long> item_draw_weights_host(const cu2b_csr *m, int user_stride) {
    std::vector<unsigned long long> w((size_t)m->cols, 0ULL);
    if (m->on_device) return w;
#pragma omp parallel
    {
        std::vector<unsigned long long> mine((size_t)m->cols, 0ULL);
#pragma omp for schedule(static) nowait
        for (int u = 0; u < m->rows; u += user_stride) {
            const int lo = m->indptr[u], hi = m->indptr[u + 1];
            if (hi > lo) {
                const unsigned long long share = 0x100000000ULL / (unsigned long long)(hi - lo);
                for (int j = lo; j < hi; ++j) mine[m->indices[j]] += share;
            }
        }
#pragma omp critical
        for (int i = 0; i < m->cols; ++i) w[i] += mine[i];
    }
    return w;
}

// Share of the draws of its item block that the most frequently drawn item receives.
double hot_item_share(const std::vector<unsigned long long> &w, const int *item_block_ptr, int n_blocks) {
    double hot = 0.0;
    const int cols = (int)w.size();
    for (int b = 0; b < n_blocks; ++b) {
        const int i0 = item_block_ptr ? item_block_ptr[b] : 0, i1 = item_block_ptr ? item_block_ptr[b + 1] : cols;
        double tot = 0.0, mx = 0.0;
        for (int i = i0; i < i1; ++i) { tot += (double)w[i]; mx = std::max(mx, (double)w[i]); }
        if (tot > 0) hot = std::max(hot, mx / tot);
    }
    return hot;
}

// The asynchronous-SGD stability load of an item: lr x (its share of the draws of its item block) x (user
// groups in flight) = lr x (expected number of this item's steps that a concurrent reader does not see yet).
// Measured on B200 (profiles/r2_dsgd_stability_map.md): runs with a load of 0.5 on the hottest item bias
// diverge or not depending on occupancy and chance; 0.25 and below never diverged. kStableLoad is that bound
// with its 2x margin; it is applied either by capping the groups in flight (single GPU: exact updates) or,
// on DSGD ranks, by applying only a fraction keep[i] = kStableLoad / load of the item's bias steps.
constexpr double kStableLoad = 0.25;

// keep[i] = fraction of item i's draws whose item-side step is applied so that load x keep <= budget.
std::vector<float> item_keep_fractions(const std::vector<unsigned long long> &w, const int *item_block_ptr, int n_blocks,
                                       float lr, int groups_in_flight, double budget) {
    const int cols = (int)w.size();
    std::vector<float> keep((size_t)cols, 1.0f);
    if (lr <= 0 || budget <= 0) return keep;
    for (int b = 0; b < n_blocks; ++b) {
        const int i0 = item_block_ptr ? item_block_ptr[b] : 0, i1 = item_block_ptr ? item_block_ptr[b + 1] : cols;
        double tot = 0.0;
        for (int i = i0; i < i1; ++i) tot += (double)w[i];
        if (tot <= 0) continue;
        for (int i = i0; i < i1; ++i) {
            const double load = (double)lr * ((double)w[i] / tot) * groups_in_flight;
            if (load > budget) keep[i] = (float)(budget / load);
        }
    }
    return keep;
}

// Bound on the concurrently processed ratings so that the hottest item's load stays <= budget.
int inflight_cap(double hot_share, float lr, double budget_default) {
    double budget = budget_default;
    if (const char *e = getenv("CU2B_INFLIGHT_LR")) budget = atof(e);
    if (hot_share <= 0 || budget <= 0 || lr <= 0) return INT32_MAX;
    const double cap = budget / ((double)lr * hot_share);
    return cap > 1e9 ? INT32_MAX : (int)std::max(32.0, cap);
}

StreamView flat_view(const cu2b_rating *base, long long n, int chunk) {
    StreamView sv;
    sv.base = base;
    sv.seg_pitch = (n + 3) & ~3LL;
    sv.seg_len = (int)n;
    sv.chunk = chunk;
    sv.chunks_per_seg = (int)((n + chunk - 1) / chunk);
    sv.num_chunks = sv.chunks_per_seg;
    return sv;
}

int pick_chunk(long long per_segment, int resident_ctas) {
    long long c = per_segment / std::max(1, resident_ctas);
    c &= ~3LL;
    return (int)std::min<long long>(kChunkMax, std::max<long long>(32, c));
}

struct Timing {  // CUDA-event stopwatch per kernel family, resolved after a stream sync
    enum Kind { SGD = 0, LOSS = 1, SAMPLER = 2, TOTAL = 3, WAIT = 4, SEND = 5, NKIND = 6 };
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    int begin(int kind, cudaStream_t st) {
        Span s{get(), get(), kind};
        cudaEventRecord(s.a, st);
        spans.push_back(s);
        return (int)spans.size() - 1;
    }
    void end(int id, cudaStream_t st) { cudaEventRecord(spans[id].b, st); }
    void collect(double ms[NKIND]) {  // call after the stream is idle
        for (Span &s : spans) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms[s.kind] += t;
            pool.push_back(s.a);
            pool.push_back(s.b);
        }
        spans.clear();
    }
    ~Timing() {
        for (Span &s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------
struct cu2b_session {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevPool pool;
    cu2b_config cfg;
    float mu = 0.f;
    int k = 0, kp = 0, L = 1, V = 1;
    int rows = 0, cols = 0;
    DevMatrix train, test;
    float *P = nullptr, *Q = nullptr, *ub = nullptr, *ib = nullptr;
    // ItemBiasLayout: item i's bias is ib[i * ibs]. ibs = 64 floats gives every item its own 256-byte
    // L2 hash granule: the 4-byte bias read + atomic add of every update then spread over as many L2
    // slices as the item rows do, instead of piling onto the few lines of a dense array (a DSGD rank's
    // item block of 2 221 items is 70 lines = 35 granules; measured in profiles/r2_l2_rows_micro.jsonl).
    // ib_dense: [cols] staging buffer for the dense host-side array (== ib when ibs == 1).
    int ibs = 1;
    float *ib_dense = nullptr;
    // ItemPlacement (Hogwild sessions): the item rows are stored in an internal order chosen for L2 slice balance
    // (cu2b_paired_slots over the items' popularity); item_pos[external id] = internal row. The rating
    // triplets carry internal ids; uploads / downloads of Q and item_bias translate. nullptr = identity.
    int *item_pos = nullptr;
    float *Q_stage = nullptr;  // [cols x k] dense staging for the permuting upload / download of Q
    bool want_placement = false;
    cudaStream_t out_stream = nullptr;  // cu2b_session_run_download: the model leaves while the last loss check runs
    cudaEvent_t ev_model_final = nullptr;
    // expected draws per iteration of every item (external ids), units of 2^-32 (item_draw_weight_kernel),
    // from every item_w_stride-th user; refreshed by every upload (create and reload)
    std::vector<unsigned long long> item_w;
    int item_w_stride = 1;
    unsigned long long *item_w_dev = nullptr;
    int *bad_ids = nullptr;  // device counter of item ids outside [0, cols) (check_item_ids_kernel)
    int *active = nullptr;
    int n_active = 0;
    int *user_ids = nullptr;  // DSGD: original user id of each local user (sampler key), else null
    // update stream
    cu2b_rating *stream_buf = nullptr;
    long long seg_pitch = 0;
    int chunk = 0, chunks_per_seg = 0, max_batch_segs = 0;
    int *gate = nullptr;
    int segs_done = 0;
    unsigned long long *counters = nullptr;
    int counter_slots = 0, counter_next = 0;
    // loss
    DevState *state = nullptr;
    cu2b_metrics *log_dev = nullptr;
    int log_cap = 0;
    double *part_train = nullptr, *part_test = nullptr;
    int loss_grid_max = 0, sgd_grid_max = 0;
    int loss_chunk = kChunkMax;
    SgdKernel sgd_kernel = nullptr;
    LossKernel loss_kernel = nullptr;
    BlockedKernel blocked_kernel = nullptr;
    // deterministic mode: block schedule on the device + "updates owed" accumulator
    cu2b_rating *sched = nullptr;
    int *bucket_ptr = nullptr;
    int B = 0;
    long long blocked_budget = 0;
    bool dsgd_child = false;  // created by cu2b_dsgd_create: no single-GPU update stream
    double hot_share = 0.0;   // hottest item's share of the draws (asynchronous-SGD stability bound)
    // iteration-tiled schedule (cfg.round_iters > 1)
    DsgdDraw *draws = nullptr;       // two buffers of n_active * draw_pitch draws
    size_t draws_stride = 0;
    int draw_pitch = 0, round_iters = 1, tiles_grid = 0;
    cudaStream_t sampler_stream = nullptr;  // the sampler of round r+1 overlaps the update kernel of round r
    cudaEvent_t ev_sampled[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool consumed_pending[2] = {false, false};
    long long rounds_enqueued = 0;
    UserTileKernel tiles_kernel = nullptr;
    // default: the sampler is fused into the update kernel (mf_sgd_user_rounds); CU2B_TILE_PIPE=tma
    // selects the separate sampler + TMA-fed tile kernel (mf_sgd_user_tiles) for A/B runs
    bool fused_sampler = false;
    UserRoundKernel rounds_kernel = nullptr;
    int rounds_grid = 0;
    int sgd_grid_occ = 0, rounds_grid_occ = 0, tiles_grid_occ = 0;  // occupancy-limited grids before the stability cap
    // experiment switches (environment): CU2B_TUNE_GATE=0 drops the per-user ordering gate,
    // CU2B_TUNE_CHUNK overrides the chunk size
    bool no_gate = false;
    int sm_count = 0;
    int iter_done = 0;  // the reference loop variable i (training.cu:107)
    Timing timing;
    cu2b_stats stats;
    ~cu2b_session() {
        if (sampler_stream) cudaStreamSynchronize(sampler_stream);
        pool.release();  // stream-ordered frees are enqueued before the stream goes away
        if (stream) {
            cudaStreamSynchronize(stream);
            cudaStreamDestroy(stream);
        }
        if (sampler_stream) cudaStreamDestroy(sampler_stream);
        if (out_stream) { cudaStreamSynchronize(out_stream); cudaStreamDestroy(out_stream); }
        if (ev_model_final) cudaEventDestroy(ev_model_final);
        for (int b = 0; b < 2; ++b) {
            if (ev_sampled[b]) cudaEventDestroy(ev_sampled[b]);
            if (ev_consumed[b]) cudaEventDestroy(ev_consumed[b]);
        }
    }
};

namespace {

// ItemBiasLayout helpers (see cu2b_session::ibs)
// pos (optional) = ItemPlacement: external item id -> internal row
__global__ void __launch_bounds__(256)
bias_scatter_kernel(const float *__restrict__ dense, float *__restrict__ padded, int n, int stride, const int *__restrict__ pos) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        padded[(size_t)(pos ? pos[i] : i) * stride] = dense[i];
}
__global__ void __launch_bounds__(256)
bias_gather_kernel(const float *__restrict__ padded, float *__restrict__ dense, int n, int stride, const int *__restrict__ pos) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dense[i] = __ldcg(padded + (size_t)(pos ? pos[i] : i) * stride);
}
// Item rows between the caller's dense [cols x k] order (staging buffer) and the device layout [cols x kp] in the
// internal row order. to_device: Q[pos[i]] = stage[i] (padding zeroed); else stage[i] = Q[pos[i]].
__global__ void __launch_bounds__(256)
permute_rows_kernel(float *__restrict__ Q, float *__restrict__ stage, const int *__restrict__ pos, int cols, int k, int kp,
                    int to_device) {
    const long long total = (long long)cols * kp;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / kp), f = (int)(t - (long long)i * kp);
        const size_t dev = (size_t)pos[i] * kp + f;
        if (to_device) Q[dev] = f < k ? stage[(size_t)i * k + f] : 0.f;
        else if (f < k) stage[(size_t)i * k + f] = __ldcg(Q + dev);
    }
}
// w[i] += floor(2^32 / deg(u)) for every rating (u, i) of every `user_stride`-th user: the expected number of draws
// of item i per iteration under per-user sampling (one uniform draw per user per iteration, sgd.cu:27-37) in units
// of 2^-32. Integer atomics => exact, order independent, identical to the host version (item_draw_weights_host).
__global__ void __launch_bounds__(256)
item_draw_weight_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int rows, int cols, int user_stride,
                        unsigned long long *__restrict__ w) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = warp * user_stride; u < rows; u += warps * user_stride) {
        const int lo = __ldg(indptr + u), hi = __ldg(indptr + u + 1);
        if (hi <= lo) continue;
        const unsigned long long share = 0x100000000ULL / (unsigned long long)(hi - lo);
        for (int j = lo + lane; j < hi; j += 32) {
            const int item = __ldg(indices + j);
            if ((unsigned)item < (unsigned)cols) atomicAdd(w + item, share);  // out-of-range ids are reported by check_item_ids_kernel
        }
    }
}
// Every item id of a rating matrix must lie in [0, cols): the kernels index Q and item_bias with them. One pass
// over the ids on the device (60 us for 90 M ratings); *bad counts the offenders.
__global__ void __launch_bounds__(256)
check_item_ids_kernel(const int *__restrict__ indices, long long nnz, int cols, int *bad) {
    int mine = 0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x)
        mine += (unsigned)__ldg(indices + j) >= (unsigned)cols;
    if (mine) atomicAdd(bad, mine);
}

int item_bias_stride() {
    int st = 64;
    if (const char *e = getenv("CU2B_IB_STRIDE")) st = atoi(e);
    return std::max(1, std::min(64, st));
}

cu2b_status alloc_item_bias(cu2b_session *s) {
    s->ibs = item_bias_stride();
    const size_t n = (size_t)std::max(1, s->cols);
    CU2B_TRY(s->pool.alloc(&s->ib, n * s->ibs));
    if (s->ibs == 1 && !s->want_placement) {
        s->ib_dense = s->ib;
    } else {
        CU2B_TRY(s->pool.alloc(&s->ib_dense, n));
        CUDA_TRY(cudaMemsetAsync(s->ib, 0, n * s->ibs * sizeof(float), s->stream));
    }
    return CU2B_OK;
}

// ib_dense (host order, dense) -> ib (one bias per line); call on a stream ordered after the H2D copy
cu2b_status scatter_item_bias(cu2b_session *s, cudaStream_t st) {
    if ((s->ibs == 1 && !s->item_pos) || s->cols == 0) return CU2B_OK;
    bias_scatter_kernel<<<std::max(1, std::min((s->cols + 255) / 256, 592)), 256, 0, st>>>(s->ib_dense, s->ib, s->cols, s->ibs, s->item_pos);
    CUDA_TRY(cudaGetLastError());
    return CU2B_OK;
}

cu2b_status download_item_bias(cu2b_session *s, float *host, cudaStream_t st) {
    if (s->cols == 0) return CU2B_OK;
    if (s->ibs != 1 || s->item_pos) {
        bias_gather_kernel<<<std::max(1, std::min((s->cols + 255) / 256, 592)), 256, 0, st>>>(s->ib, s->ib_dense, s->cols, s->ibs, s->item_pos);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(host, s->ib_dense, (size_t)s->cols * sizeof(float), cudaMemcpyDeviceToHost, st));
    return CU2B_OK;
}

cu2b_status check_device_error(cu2b_session *s) {
    DevState st;
    CUDA_TRY(cudaMemcpy(&st, s->state, sizeof(st), cudaMemcpyDeviceToHost));
    if (st.error < 0)
        return cu2b_fail(CU2B_ERR_DIVERGED, "the model became non-finite (NaN/inf RMSE at the loss check of iteration %d): "
                         "asynchronous SGD diverged; lower the learning rate or the number of concurrently applied updates "
                         "(CU2B_INFLIGHT_LR)", -st.error - 1);
    if (st.error)
        return cu2b_fail(CU2B_ERR_CUDA, "device-side wait timed out (code %d = site*1e6 + need*100 + seen): %s", st.error,
                         st.error == 2 ? "per-user ordering gate" : "DSGD peer did not deliver");
    return CU2B_OK;
}

cu2b_status launch_loss(cu2b_session *s, const DevMatrix &m, double *partials, int *nblk, float *err_out) {
    LossParams lp;
    lp.sv = flat_view(m.coo, m.nnz, s->loss_chunk);
    lp.P = s->P; lp.Q = s->Q; lp.user_bias = s->ub; lp.item_bias = s->ib;
    lp.ibs = s->ibs;
    lp.kp = s->kp;
    lp.mu = s->mu;
    lp.partials = partials;
    lp.err_out = err_out;
    const int grid = (int)std::max<long long>(1, std::min<long long>(lp.sv.num_chunks, s->loss_grid_max));
    s->loss_kernel<<<grid, kThreads, 0, s->stream>>>(lp);
    CUDA_TRY(cudaGetLastError());
    s->stats.kernel_launches++;
    *nblk = grid;
    return CU2B_OK;
}

cu2b_status enqueue_check(cu2b_session *s, int iteration_1based, int apply_schedule, bool log) {
    const int id = s->timing.begin(Timing::LOSS, s->stream);
    int nb_tr = 0, nb_te = 0;
    CU2B_TRY(launch_loss(s, s->train, s->part_train, &nb_tr, nullptr));
    CU2B_TRY(launch_loss(s, s->test, s->part_test, &nb_te, nullptr));
    loss_finalize_kernel<<<1, 256, 0, s->stream>>>(s->state, s->part_train, nb_tr, s->train.nnz,
                                                  s->part_test, nb_te, s->test.nnz, iteration_1based,
                                                  apply_schedule, log ? s->log_dev : nullptr);
    CUDA_TRY(cudaGetLastError());
    s->stats.kernel_launches++;
    s->timing.end(id, s->stream);
    return CU2B_OK;
}

cu2b_status launch_sgd(cu2b_session *s, const StreamView &sv, int *gate, int seg0, int serial,
                       const int *dyn_range = nullptr) {
    SgdParams sp;
    sp.sv = sv;
    sp.P = s->P; sp.Q = s->Q; sp.user_bias = s->ub; sp.item_bias = s->ib;
    sp.ibs = s->ibs;
    sp.kp = s->kp;
    sp.mu = s->mu;
    sp.lr = &s->state->lr;
    sp.P_reg = s->cfg.P_reg; sp.Q_reg = s->cfg.Q_reg;
    sp.ub_reg = s->cfg.user_bias_reg; sp.ib_reg = s->cfg.item_bias_reg;
    sp.is_train = s->cfg.is_train;
    if (s->counter_next == 0)
        CUDA_TRY(cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long) * s->counter_slots, s->stream));
    sp.chunk_counter = s->counters + s->counter_next;
    s->counter_next = (s->counter_next + 1) % s->counter_slots;
    sp.gate = gate;
    sp.seg0 = seg0;
    sp.serial = serial;
    sp.dyn_range = dyn_range;
    sp.error_flag = &s->state->error;
    if (s->no_gate) sp.gate = nullptr;
    const int grid = serial ? 1 : dyn_range ? s->sgd_grid_max
                                            : (int)std::max<long long>(1, std::min<long long>(sv.num_chunks, s->sgd_grid_max));
    s->sgd_kernel<<<grid, kThreads, 0, s->stream>>>(sp);
    CUDA_TRY(cudaGetLastError());
    s->stats.kernel_launches++;
    s->stats.sgd_launches++;
    return CU2B_OK;
}

// One deterministic pass over all training ratings: B rounds, one launch per round.
cu2b_status enqueue_blocked_pass(cu2b_session *s, const cu2b_rating *sched, const int *bucket_ptr, int B) {
    constexpr int kBlockedThreads = 256;
    const int G = 32 / s->L;
    const int warps = (B + G - 1) / G;
    const int grid = std::max(1, (warps + kBlockedThreads / 32 - 1) / (kBlockedThreads / 32));
    BlockedParams bp;
    bp.sched = sched;
    bp.B = B;
    SgdParams &sp = bp.model;
    memset(&sp, 0, sizeof(sp));
    sp.dyn_range = nullptr;
    sp.error_flag = nullptr;
    sp.P = s->P; sp.Q = s->Q; sp.user_bias = s->ub; sp.item_bias = s->ib;
    sp.ibs = s->ibs;
    sp.kp = s->kp;
    sp.mu = s->mu;
    sp.lr = &s->state->lr;
    sp.P_reg = s->cfg.P_reg; sp.Q_reg = s->cfg.Q_reg;
    sp.ub_reg = s->cfg.user_bias_reg; sp.ib_reg = s->cfg.item_bias_reg;
    sp.is_train = s->cfg.is_train;
    for (int r = 0; r < B; ++r) {
        bp.bucket_ptr = bucket_ptr + (size_t)r * B;
        s->blocked_kernel<<<grid, kBlockedThreads, 0, s->stream>>>(bp);
    }
    CUDA_TRY(cudaGetLastError());
    s->stats.kernel_launches += B;
    s->stats.sgd_launches += B;
    return CU2B_OK;
}

// Deterministic mode inside the training loop: iterations are converted to whole passes at
// equal update counts (n_seg * n_active updates are owed; a pass pays train.nnz of them).
cu2b_status enqueue_blocked_iterations(cu2b_session *s, int n_seg) {
    s->blocked_budget += (long long)n_seg * s->n_active;
    while (s->train.nnz > 0 && s->blocked_budget >= s->train.nnz) {
        const int id = s->timing.begin(Timing::SGD, s->stream);
        CU2B_TRY(enqueue_blocked_pass(s, s->sched, s->bucket_ptr, s->B));
        s->timing.end(id, s->stream);
        s->blocked_budget -= s->train.nnz;
        s->stats.updates += s->train.nnz;
    }
    return CU2B_OK;
}

// Iteration-tiled schedule: rounds of up to round_iters iterations, users processed in tiles.
// The draws of round r+1 are sampled on a second stream while the update kernel of round r runs
// (two draw buffers, two events per buffer); everything stays stream/event ordered on the device.
cu2b_status enqueue_fused_rounds(cu2b_session *s, int iter_abs, int n_seg) {
    UserRoundParams rp;
    rp.indptr = s->train.indptr;
    rp.coo = s->train.coo;
    rp.active_users = s->active;
    rp.user_ids = s->user_ids;
    rp.n_active = s->n_active;
    rp.seed = (uint32_t)s->cfg.seed;
    rp.P = s->P; rp.Q = s->Q; rp.user_bias = s->ub; rp.item_bias = s->ib;
    rp.ibs = s->ibs;
    rp.kp = s->kp;
    rp.mu = s->mu;
    rp.lr = &s->state->lr;
    rp.P_reg = s->cfg.P_reg; rp.Q_reg = s->cfg.Q_reg;
    rp.ub_reg = s->cfg.user_bias_reg; rp.ib_reg = s->cfg.item_bias_reg;
    rp.is_train = s->cfg.is_train;
    for (int done = 0; done < n_seg;) {
        const int nb = std::min(n_seg - done, s->round_iters);
        rp.iter0 = iter_abs + done;
        rp.nb = nb;
        if (s->counter_next == 0)
            CUDA_TRY(cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long) * s->counter_slots, s->stream));
        rp.tile_counter = s->counters + s->counter_next;
        s->counter_next = (s->counter_next + 1) % s->counter_slots;
        const int id = s->timing.begin(Timing::SGD, s->stream);
        s->rounds_kernel<<<s->rounds_grid, kRoundWarps * 32, 0, s->stream>>>(rp);
        CUDA_TRY(cudaGetLastError());
        s->timing.end(id, s->stream);
        s->stats.kernel_launches++;
        s->stats.sgd_launches++;
        s->stats.updates += (long long)nb * s->n_active;
        done += nb;
    }
    return CU2B_OK;
}

cu2b_status enqueue_tiled_iterations(cu2b_session *s, int iter_abs, int n_seg) {
    if (s->fused_sampler) return enqueue_fused_rounds(s, iter_abs, n_seg);
    const int TU = kConsumerWarps * (32 / s->L);
    // carve the segment into rounds
    std::vector<std::pair<int, int>> rounds;  // (first absolute iteration, count)
    for (int done = 0; done < n_seg;) {
        const int nb = std::min(n_seg - done, s->round_iters);
        rounds.push_back({iter_abs + done, nb});
        done += nb;
    }
    auto launch_sampler = [&](int abs_it, int nb, int buf) -> cu2b_status {
        if (s->consumed_pending[buf]) CUDA_TRY(cudaStreamWaitEvent(s->sampler_stream, s->ev_consumed[buf], 0));
        const long long draws = (long long)nb * s->n_active;
        const int grid = (int)std::max<long long>(1, std::min<long long>((draws + 255) / 256, (long long)s->sm_count * 16));
        const int tid = s->timing.begin(Timing::SAMPLER, s->sampler_stream);
        sample_user_major_kernel<<<grid, 256, 0, s->sampler_stream>>>(s->train.indptr, s->train.coo, s->active, s->user_ids,
                                                                     s->n_active, (uint32_t)s->cfg.seed, abs_it, nb,
                                                                     s->draw_pitch, s->draws + (size_t)buf * s->draws_stride);
        CUDA_TRY(cudaGetLastError());
        s->timing.end(tid, s->sampler_stream);
        CUDA_TRY(cudaEventRecord(s->ev_sampled[buf], s->sampler_stream));
        s->stats.kernel_launches++;
        return CU2B_OK;
    };
    // the sampler stream must not run ahead of earlier work on the main stream (model upload,
    // previous segment): fork it from the main stream once per segment
    cudaEvent_t fork = s->timing.get();
    CUDA_TRY(cudaEventRecord(fork, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->sampler_stream, fork, 0));
    s->timing.pool.push_back(fork);
    if (!rounds.empty()) CU2B_TRY(launch_sampler(rounds[0].first, rounds[0].second, (int)(s->rounds_enqueued & 1)));
    for (size_t r = 0; r < rounds.size(); ++r) {
        const int buf = (int)(s->rounds_enqueued & 1);
        const int nb = rounds[r].second;
        if (r + 1 < rounds.size()) CU2B_TRY(launch_sampler(rounds[r + 1].first, rounds[r + 1].second, buf ^ 1));
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_sampled[buf], 0));
        const int id = s->timing.begin(Timing::SGD, s->stream);
        UserTileParams tp;
        tp.draws = s->draws + (size_t)buf * s->draws_stride;
        tp.active_users = s->active;
        tp.n_active = s->n_active;
        tp.pitch = s->draw_pitch;
        tp.nb = nb;
        tp.n_tiles = (s->n_active + TU - 1) / TU;
        if (s->counter_next == 0)
            CUDA_TRY(cudaMemsetAsync(s->counters, 0, sizeof(unsigned long long) * s->counter_slots, s->stream));
        tp.tile_counter = s->counters + s->counter_next;
        s->counter_next = (s->counter_next + 1) % s->counter_slots;
        tp.P = s->P; tp.Q = s->Q; tp.user_bias = s->ub; tp.item_bias = s->ib;
        tp.ibs = s->ibs;
        tp.kp = s->kp;
        tp.mu = s->mu;
        tp.lr = &s->state->lr;
        tp.P_reg = s->cfg.P_reg; tp.Q_reg = s->cfg.Q_reg;
        tp.ub_reg = s->cfg.user_bias_reg; tp.ib_reg = s->cfg.item_bias_reg;
        tp.is_train = s->cfg.is_train;
        const int grid = std::max(1, std::min(tp.n_tiles, s->tiles_grid));
        s->tiles_kernel<<<grid, kThreads, 0, s->stream>>>(tp);
        CUDA_TRY(cudaGetLastError());
        s->stats.kernel_launches++;
        s->stats.sgd_launches++;
        s->timing.end(id, s->stream);
        CUDA_TRY(cudaEventRecord(s->ev_consumed[buf], s->stream));
        s->consumed_pending[buf] = true;
        s->rounds_enqueued++;
        s->stats.updates += (long long)nb * s->n_active;
    }
    return CU2B_OK;
}

// n_seg reference iterations starting at absolute iteration `iter_abs`
cu2b_status enqueue_sgd_iterations(cu2b_session *s, int iter_abs, int n_seg) {
    if (s->round_iters > 1) return enqueue_tiled_iterations(s, iter_abs, n_seg);
    while (n_seg > 0) {
        const int nb = std::min(n_seg, s->max_batch_segs);
        // 1. sampler: one draw per active user per iteration (sgd.cu:27-37), or the next n_active ratings of a
        //    shuffled pass over the rating list (per_rating)
        {
            const int id = s->timing.begin(Timing::SAMPLER, s->stream);
            const long long draws = (long long)nb * s->n_active;
            const int grid = (int)std::min<long long>((draws + 255) / 256, (long long)s->sm_count * 16);
            if (s->cfg.sampler == CU2B_SAMPLER_PER_RATING)
                sample_per_rating_kernel<<<grid, 256, 0, s->stream>>>(
                    s->train.coo, (unsigned long long)s->train.nnz, feistel_half_bits((unsigned long long)s->train.nnz),
                    (uint32_t)s->cfg.seed, (unsigned long long)iter_abs * (unsigned long long)s->n_active, draws, s->n_active,
                    s->stream_buf, s->seg_pitch);
            else
            sample_per_user_kernel<<<grid, 256, 0, s->stream>>>(
                s->train.indptr, s->train.coo, s->active, s->n_active, (uint32_t)s->cfg.seed, iter_abs,
                draws, s->stream_buf, s->seg_pitch, s->user_ids);
            CUDA_TRY(cudaGetLastError());
            s->stats.kernel_launches++;
            s->timing.end(id, s->stream);
        }
        // 2. Hogwild updates over the sampled stream
        {
            const int id = s->timing.begin(Timing::SGD, s->stream);
            StreamView sv;
            sv.base = s->stream_buf;
            sv.seg_pitch = s->seg_pitch;
            sv.seg_len = s->n_active;
            sv.chunk = s->chunk;
            sv.chunks_per_seg = s->chunks_per_seg;
            sv.num_chunks = (long long)nb * s->chunks_per_seg;
            // per_rating: a user may appear anywhere in any segment, so the per-user ordering gate (chunk j of
            // iteration t after chunk j of iteration t - 1) means nothing; both rows take their steps as L2 atomic adds
            CU2B_TRY(launch_sgd(s, sv, s->cfg.sampler == CU2B_SAMPLER_PER_RATING ? nullptr : s->gate, s->segs_done, 0));
            s->timing.end(id, s->stream);
        }
        s->segs_done += nb;
        s->stats.updates += (long long)nb * s->n_active;
        iter_abs += nb;
        n_seg -= nb;
    }
    return CU2B_OK;
}

}  // namespace

// Launch grids from the occupancy limits and the asynchronous-SGD stability bound (kStableLoad) for the item
// popularity of the matrix that was uploaded last (s->item_w): called at creation and after every reload.
static void apply_stability_cap(cu2b_session *s) {
    s->hot_share = hot_item_share(s->item_w, nullptr, 1);
    const int cap = inflight_cap(s->hot_share, s->cfg.learning_rate, kStableLoad);
    {   // ratings in flight per CTA: 8 consumer warps x (32/L) groups x unroll (2 for L < 32)
        const int per_cta = kConsumerWarps * (32 / s->L) * (s->L < 32 ? 2 : 1);
        s->sgd_grid_max = std::max(1, std::min(s->sgd_grid_occ, cap / per_cta));
    }
    if (s->rounds_kernel) {
        const int per_cta = kRoundWarps * (32 / s->L);
        s->rounds_grid = std::max(1, std::min(std::min(s->rounds_grid_occ, cap / per_cta), (s->n_active + per_cta - 1) / per_cta));
    }
    if (s->tiles_kernel) {
        const int per_cta = kConsumerWarps * (32 / s->L);
        s->tiles_grid = std::max(1, std::min(s->tiles_grid_occ, cap / per_cta));
    }
}

// Device-resident schedule state as train() starts it (training.cu:102-103) + host counters.
static cu2b_status session_reset_state(cu2b_session *s) {
    DevState st;
    memset(&st, 0, sizeof(st));
    st.lr = s->cfg.learning_rate;
    st.current_patience = (int)s->cfg.patience;  // training.cu:103
    st.patience0 = (int)s->cfg.patience;
    st.lr_decay = s->cfg.learning_rate_decay;
    st.validation_rmse = FLT_MAX;                // training.cu:102
    st.n_log = 0;
    st.log_cap = s->log_cap;
    CUDA_TRY(cudaMemcpyAsync(s->state, &st, sizeof(st), cudaMemcpyHostToDevice, s->stream));
    if (s->gate) CUDA_TRY(cudaMemsetAsync(s->gate, 0, (size_t)s->chunks_per_seg * sizeof(int), s->stream));
    s->iter_done = 0;
    s->segs_done = 0;
    s->blocked_budget = 0;
    memset(&s->stats, 0, sizeof(s->stats));
    return CU2B_OK;
}

// Host -> device traffic of a session: both rating matrices (copy + COO expansion), the active
// user list and the model. All copies go through one copy stream; the compute stream only waits
// for the buffer it is about to touch (the COO expansion of the training matrix overlaps the
// upload of the model and of the test matrix). reload = the session's buffers exist already.
static cu2b_status session_upload(cu2b_session *s, const cu2b_csr *train, const cu2b_csr *test, const float *P,
                                  const float *Q, const float *user_bias, const float *item_bias, bool reload) {
    struct CopyLane {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr, ev_vals = nullptr;
        ~CopyLane() {
            if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
            if (ev) cudaEventDestroy(ev);
            if (ev_vals) cudaEventDestroy(ev_vals);
        }
    } lane;
    CUDA_TRY(cudaStreamCreateWithFlags(&lane.st, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&lane.ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&lane.ev_vals, cudaEventDisableTiming));
    std::vector<int> indptr_host;
    CU2B_TRY(host_indptr(train, s->stream, &indptr_host));
    // users with at least one training rating (sgd.cu:35 skips the others)
    std::vector<int> active;
    active.reserve(s->rows);
    bool monotone = indptr_host[0] == 0 && indptr_host[s->rows] == train->nonzeros;
    for (int u = 0; u < s->rows; ++u) {
        monotone = monotone && indptr_host[u + 1] >= indptr_host[u];
        if (indptr_host[u + 1] > indptr_host[u]) active.push_back(u);
    }
    if (!monotone)
        return cu2b_fail(CU2B_ERR_INVALID, "train matrix: indptr must start at 0, never decrease and end at nonzeros (%d)", train->nonzeros);
    if (reload && (int)active.size() != s->n_active)
        return cu2b_fail(CU2B_ERR_INVALID, "reload: %d users have training ratings, the session was created with %d "
                         "(launch geometry depends on it)", (int)active.size(), s->n_active);
    s->n_active = (int)active.size();
    // 1. every allocation, in stream order on the compute stream
    MatrixUpload up_train, up_test;
    CU2B_TRY(matrix_alloc(s->pool, train, &s->train, &up_train, reload));
    CU2B_TRY(matrix_alloc(s->pool, test, &s->test, &up_test, reload));
    if (!reload) {
        CU2B_TRY(s->pool.alloc(&s->active, (size_t)std::max(1, s->rows)));
        CU2B_TRY(s->pool.alloc(&s->P, (size_t)s->rows * s->kp));
        CU2B_TRY(s->pool.alloc(&s->Q, (size_t)s->cols * s->kp));
        CU2B_TRY(s->pool.alloc(&s->ub, (size_t)s->rows));
        CU2B_TRY(alloc_item_bias(s));
        CU2B_TRY(s->pool.alloc(&s->item_w_dev, (size_t)std::max(1, s->cols)));
        CU2B_TRY(s->pool.alloc(&s->bad_ids, (size_t)1));
        if (s->want_placement) {
            CU2B_TRY(s->pool.alloc(&s->item_pos, (size_t)std::max(1, s->cols)));
            CU2B_TRY(s->pool.alloc(&s->Q_stage, (size_t)std::max(1, s->cols) * s->k));
        }
    }
    CU2B_TRY(stream_after(lane.st, s->stream, lane.ev));
    // 2. copies back to back on the copy stream; the compute stream waits only for what it is about to touch
    if (!active.empty())  // pageable source: staged synchronously, so it goes first
        CUDA_TRY(cudaMemcpyAsync(s->active, active.data(), active.size() * sizeof(int), cudaMemcpyHostToDevice, lane.st));
    CU2B_TRY(matrix_copy(lane.st, up_train, 0));  // row pointers + item ids
    CU2B_TRY(stream_after(s->stream, lane.st, lane.ev));
    // item draw weights from the ids alone, while the rating values and the model are still in flight
    {
        const int *ids = up_train.tmp_i ? up_train.tmp_i : train->indices;
        // single-GPU sessions only need the popularity order and the hottest share: visit <= ~16 M ratings;
        // DSGD strips derive per-item step fractions from the weights: every user (== cu2b_dsgd_item_keep)
        s->item_w_stride = s->dsgd_child ? 1 : (int)std::max<long long>(1, (long long)train->nonzeros >> 24);
        CUDA_TRY(cudaMemsetAsync(s->item_w_dev, 0, (size_t)std::max(1, s->cols) * sizeof(unsigned long long), s->stream));
        CUDA_TRY(cudaMemsetAsync(s->bad_ids, 0, sizeof(int), s->stream));
        if (train->nonzeros > 0) {
            const int warps = (s->rows + s->item_w_stride - 1) / s->item_w_stride;
            item_draw_weight_kernel<<<std::max(1, std::min((warps + 7) / 8, s->sm_count * 8)), 256, 0, s->stream>>>(
                s->train.indptr, ids, s->rows, s->cols, s->item_w_stride, s->item_w_dev);
            CUDA_TRY(cudaGetLastError());
            check_item_ids_kernel<<<std::max(1, std::min((train->nonzeros + 255) / 256, s->sm_count * 8)), 256, 0, s->stream>>>(
                ids, (long long)train->nonzeros, s->cols, s->bad_ids);
            CUDA_TRY(cudaGetLastError());
        }
    }
    CU2B_TRY(matrix_copy(lane.st, up_train, 1));  // rating values
    CUDA_TRY(cudaEventRecord(lane.ev_vals, lane.st));
    CU2B_TRY(upload_dense(lane.st, s->P, P, s->rows, s->k, s->kp));
    if (s->item_pos) CUDA_TRY(cudaMemcpyAsync(s->Q_stage, Q, (size_t)s->cols * s->k * sizeof(float), cudaMemcpyHostToDevice, lane.st));
    else CU2B_TRY(upload_dense(lane.st, s->Q, Q, s->cols, s->k, s->kp));
    CUDA_TRY(cudaMemcpyAsync(s->ub, user_bias, (size_t)s->rows * sizeof(float), cudaMemcpyHostToDevice, lane.st));
    CUDA_TRY(cudaMemcpyAsync(s->ib_dense, item_bias, (size_t)s->cols * sizeof(float), cudaMemcpyHostToDevice, lane.st));
    CU2B_TRY(matrix_copy(lane.st, up_test));
    // 3. weights back on the host (the copies above keep streaming meanwhile): hottest share, and at creation the
    //    internal row order
    s->item_w.assign((size_t)s->cols, 0ULL);
    if (s->cols > 0)
        CUDA_TRY(cudaMemcpyAsync(s->item_w.data(), s->item_w_dev, (size_t)s->cols * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    int bad_host = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad_host, s->bad_ids, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (bad_host)
        return cu2b_fail(CU2B_ERR_INVALID, "train matrix: %d item ids lie outside [0, %d) (a ratings file with itemId 0 "
                         "or ids beyond the declared column count?)", bad_host, s->cols);
    if (s->item_pos && !reload) {
        std::vector<int> order((size_t)s->cols), slot((size_t)s->cols), pos((size_t)s->cols);
        for (int i = 0; i < s->cols; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return s->item_w[a] > s->item_w[b]; });
        cu2b_paired_slots(s->cols, 0, cu2b_rows_per_l2_block(s->k), slot.data());
        for (int r = 0; r < s->cols; ++r) pos[order[r]] = slot[r];
        CUDA_TRY(cudaMemcpyAsync(s->item_pos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));  // pos is a local
    }
    // 4. expansions (ratings carry internal item ids from here on) and the item side in its internal order
    CUDA_TRY(cudaStreamWaitEvent(s->stream, lane.ev_vals, 0));  // the training matrix is complete
    CU2B_TRY(matrix_expand(s->pool, s->stream, up_train, s->item_pos));
    CU2B_TRY(stream_after(s->stream, lane.st, lane.ev));  // everything copied
    if (s->item_pos && s->cols > 0) {
        const long long total = (long long)s->cols * s->kp;
        permute_rows_kernel<<<(int)std::max<long long>(1, std::min<long long>((total + 255) / 256, s->sm_count * 8LL)), 256, 0, s->stream>>>(
            s->Q, s->Q_stage, s->item_pos, s->cols, s->k, s->kp, 1);
        CUDA_TRY(cudaGetLastError());
    }
    CU2B_TRY(scatter_item_bias(s, s->stream));
    if (test->nonzeros > 0) {
        check_item_ids_kernel<<<std::max(1, std::min((test->nonzeros + 255) / 256, s->sm_count * 8)), 256, 0, s->stream>>>(
            up_test.tmp_i ? up_test.tmp_i : test->indices, (long long)test->nonzeros, s->cols, s->bad_ids);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&bad_host, s->bad_ids, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (bad_host)
            return cu2b_fail(CU2B_ERR_INVALID, "test matrix: %d item ids lie outside [0, %d)", bad_host, s->cols);
    }
    CU2B_TRY(matrix_expand(s->pool, s->stream, up_test, s->item_pos));
    return CU2B_OK;
}

static cu2b_status session_create_impl(cu2b_session **out, int device, const cu2b_csr *train,
                                       const cu2b_csr *test, const cu2b_config *cfg, const float *P,
                                       const float *Q, const float *user_bias, const float *item_bias,
                                       float global_bias, bool alloc_stream) {
    if (!out || !cfg || !P || !Q || !user_bias || !item_bias)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_create: null argument");
    *out = nullptr;
    CU2B_TRY(validate_csr(train, "train"));
    CU2B_TRY(validate_csr(test, "test"));
    if (test->rows > train->rows || test->cols > train->cols)
        return cu2b_fail(CU2B_ERR_INVALID,
                         "test matrix (%d x %d) exceeds the model dimensions (%d x %d); size the train "
                         "matrix with max(train, test)", test->rows, test->cols, train->rows, train->cols);
    if (cfg->n_factors < 1) return cu2b_fail(CU2B_ERR_INVALID, "n_factors must be >= 1");
    if (cfg->check_error < 1) return cu2b_fail(CU2B_ERR_INVALID, "check_error must be >= 1");
    if (cfg->mode != CU2B_MODE_HOGWILD && cfg->mode != CU2B_MODE_DETERMINISTIC)
        return cu2b_fail(CU2B_ERR_INVALID, "unknown mode %d", cfg->mode);
    if (cfg->sampler != CU2B_SAMPLER_PER_USER && cfg->sampler != CU2B_SAMPLER_PER_RATING)
        return cu2b_fail(CU2B_ERR_INVALID, "unknown sampler %d", cfg->sampler);
    if (cfg->sampler == CU2B_SAMPLER_PER_RATING && !alloc_stream)
        return cu2b_fail(CU2B_ERR_UNSUPPORTED, "sampler=per_rating is a single-GPU schedule (DSGD draws per user)");
    CUDA_TRY(cudaSetDevice(device));
    CU2B_TRY(check_device());
    cu2b_session *s = new cu2b_session();
    std::unique_ptr<cu2b_session> guard(s);
    s->device = device;
    s->cfg = *cfg;
    s->mu = global_bias;
    s->k = cfg->n_factors;
    s->kp = cu2b_padded_factors(s->k);
    CU2B_TRY(layout_for(s->kp, &s->L, &s->V));
    s->rows = train->rows;
    s->cols = train->cols;
    memset(&s->stats, 0, sizeof(s->stats));
    DeviceFacts facts;
    CU2B_TRY(device_facts(device, &facts));
    s->sm_count = facts.sm_count;
    CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    if (alloc_stream && !getenv("CU2B_NO_MEMPOOL")) s->pool.use_async(s->stream);

    s->dsgd_child = !alloc_stream;
    // ItemPlacement: Hogwild sessions store the item rows in an L2-slice-balanced internal order. Not for the
    // deterministic mode (its block schedule, and the sequential replay it is compared with, cut the items in
    // contiguous ranges of the caller's ids) and not for DSGD strips (cu2b_dsgd_partition already places the rows
    // of every item block). CU2B_PLACEMENT=0 switches it off for A/B runs.
    s->want_placement = alloc_stream && cfg->mode == CU2B_MODE_HOGWILD && s->cols > 1;
    if (const char *e = getenv("CU2B_PLACEMENT")) s->want_placement = s->want_placement && atoi(e) != 0;
    Trace tr("session_create");
    tr.mark("setup");
    CU2B_TRY(session_upload(s, train, test, P, Q, user_bias, item_bias, false));
    tr.mark("uploads + COO expansion");
    // kernels and their persistent grid sizes
    s->sgd_kernel = pick_sgd(s->L, s->V);
    s->loss_kernel = pick_loss(s->kp);
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s->sgd_kernel, kThreads, 0));
    s->sgd_grid_occ = std::max(1, occ) * s->sm_count;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s->loss_kernel, kThreads, 0));
    s->loss_grid_max = std::max(1, occ) * s->sm_count;
    apply_stability_cap(s);  // sgd_grid_max (the iteration-synchronous kernel's grid sizes the update stream's chunks)

    tr.mark("occupancy + hot-item share");
    if (const char *e = getenv("CU2B_TUNE_GATE")) s->no_gate = e[0] == '0';
    // update stream geometry: one segment per reference iteration
    s->seg_pitch = ((long long)s->n_active + 3) & ~3LL;
    s->chunk = pick_chunk(s->n_active, s->sgd_grid_max);
    if (const char *e = getenv("CU2B_TUNE_CHUNK")) s->chunk = std::max(32, std::min(kChunkMax, atoi(e) & ~3));
    s->chunks_per_seg = std::max(1, (s->n_active + s->chunk - 1) / s->chunk);
    const long long cap_ratings = 48LL << 20;  // <= 576 MB of triplets per batch
    s->max_batch_segs = (int)std::max<long long>(1, std::min<long long>(cap_ratings / std::max<long long>(1, s->seg_pitch), 4096));
    s->round_iters = 1;
    if (alloc_stream && cfg->mode == CU2B_MODE_HOGWILD && cfg->round_iters > 1 && cfg->sampler == CU2B_SAMPLER_PER_USER) {
        const char *pipe = getenv("CU2B_TILE_PIPE");
        s->fused_sampler = !(pipe && strcmp(pipe, "tma") == 0);
        const int per_user_max = s->fused_sampler ? kRoundDrawsPerWarp / (32 / s->L)
                                                  : kTileDrawsMax / (kConsumerWarps * (32 / s->L));
        int r = std::min(cfg->round_iters, per_user_max);
        if (const char *e = getenv("CU2B_ROUND")) r = std::min(std::max(1, atoi(e)), per_user_max);
        s->round_iters = std::max(1, r);
    }
    if (s->round_iters > 1 && s->fused_sampler) {
        // Measured on B200 (profiles/r1_rounds_sweep.jsonl): 4, 5, 6 or 8 resident CTAs per SM and the
        // item-row look-ahead all land within 2 % of each other -- the kernel is bound by the L2
        // atomic units of the slices that hold the popular item rows, not by latency. Default: 5 CTAs
        // (46 registers, no spills), no look-ahead. CU2B_TUNE_PF / CU2B_TUNE_MINB switch for A/B runs.
        int pf = 0, minb = 5;
        if (const char *e = getenv("CU2B_TUNE_PF")) pf = atoi(e) != 0;
        if (const char *e = getenv("CU2B_TUNE_MINB")) minb = atoi(e);
        if (s->V > 1) minb = 4;  // k > 128: several float4 per lane, keep the registers
        s->rounds_kernel = pick_user_rounds(s->L, s->V, pf, minb);
        int occ_t = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, s->rounds_kernel, kRoundWarps * 32, 0));
        if (const char *e = getenv("CU2B_TUNE_OCC")) occ_t = std::max(1, std::min(occ_t, atoi(e)));
        s->rounds_grid_occ = std::max(1, occ_t) * s->sm_count;
    } else if (s->round_iters > 1) {
        const int TU = kConsumerWarps * (32 / s->L);
        s->draw_pitch = std::min((s->round_iters + 3) & ~3, kTileDrawsMax / TU);
        s->round_iters = std::min(s->round_iters, s->draw_pitch);
        s->draws_stride = ((size_t)std::max(1, s->n_active) * s->draw_pitch + kTileDrawsMax + 1) & ~(size_t)1;
        CU2B_TRY(s->pool.alloc(&s->draws, 2 * s->draws_stride));
        CUDA_TRY(cudaStreamCreateWithFlags(&s->sampler_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_sampled[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_consumed[b], cudaEventDisableTiming));
        }
        s->tiles_kernel = pick_user_tiles(s->L, s->V);
        int occ_t = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, s->tiles_kernel, kThreads, 0));
        s->tiles_grid_occ = std::max(1, occ_t) * s->sm_count;
    }
    apply_stability_cap(s);
    if (alloc_stream && s->round_iters == 1) {
        CU2B_TRY(s->pool.alloc(&s->stream_buf, (size_t)s->max_batch_segs * s->seg_pitch + kChunkMax + 4));
        CU2B_TRY(s->pool.alloc(&s->gate, (size_t)s->chunks_per_seg));
        CUDA_TRY(cudaMemsetAsync(s->gate, 0, (size_t)s->chunks_per_seg * sizeof(int), s->stream));
    }
    s->counter_slots = 256;
    CU2B_TRY(s->pool.alloc(&s->counters, (size_t)s->counter_slots));

    if (cfg->mode == CU2B_MODE_DETERMINISTIC && s->train.nnz > 0) {
        // block schedule: built on the host from the CSR order, uploaded once
        std::vector<cu2b_rating> coo((size_t)s->train.nnz);
        CUDA_TRY(cudaMemcpyAsync(coo.data(), s->train.coo, coo.size() * sizeof(cu2b_rating), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        BlockSchedule bs;
        CU2B_TRY(build_block_schedule(coo.data(), (int64_t)coo.size(), s->rows, s->cols,
                                      cfg->n_blocks > 0 ? cfg->n_blocks : auto_blocks(s->rows, s->cols), &bs));
        s->B = bs.B;
        CU2B_TRY(s->pool.alloc(&s->sched, bs.sched.size()));
        CU2B_TRY(s->pool.alloc(&s->bucket_ptr, bs.ptr.size()));
        CUDA_TRY(cudaMemcpyAsync(s->sched, bs.sched.data(), bs.sched.size() * sizeof(cu2b_rating), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->bucket_ptr, bs.ptr.data(), bs.ptr.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        s->blocked_kernel = pick_blocked(s->L, s->V);
    }

    // loss scratch + device-resident schedule state
    CU2B_TRY(s->pool.alloc(&s->part_train, (size_t)2 * s->loss_grid_max));
    CU2B_TRY(s->pool.alloc(&s->part_test, (size_t)2 * s->loss_grid_max));
    s->log_cap = cfg->total_iterations / cfg->check_error + 8;
    CU2B_TRY(s->pool.alloc(&s->log_dev, (size_t)s->log_cap));
    CU2B_TRY(s->pool.alloc(&s->state, 1));
    CU2B_TRY(session_reset_state(s));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    tr.mark("stream/loss buffers + state");
    *out = guard.release();
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_create(cu2b_session **out, int device, const cu2b_csr *train,
                                           const cu2b_csr *test, const cu2b_config *cfg,
                                           const float *P, const float *Q, const float *user_bias,
                                           const float *item_bias, float global_bias) {
    return session_create_impl(out, device, train, test, cfg, P, Q, user_bias, item_bias, global_bias, true);
}

namespace {
struct ModelOut {
    float *P, *Q, *user_bias, *item_bias;
};

// D2H of the model (each part optional) on stream st; Q and item_bias come back in the caller's item order.
cu2b_status enqueue_download(cu2b_session *s, cudaStream_t st, const ModelOut &o) {
    if (o.P) CU2B_TRY(download_dense(st, o.P, s->P, s->rows, s->k, s->kp));
    if (o.Q && s->item_pos && s->cols > 0) {
        const long long total = (long long)s->cols * s->kp;
        permute_rows_kernel<<<(int)std::max<long long>(1, std::min<long long>((total + 255) / 256, s->sm_count * 8LL)), 256, 0, st>>>(
            s->Q, s->Q_stage, s->item_pos, s->cols, s->k, s->kp, 0);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(o.Q, s->Q_stage, (size_t)s->cols * s->k * sizeof(float), cudaMemcpyDeviceToHost, st));
    } else if (o.Q) {
        CU2B_TRY(download_dense(st, o.Q, s->Q, s->cols, s->k, s->kp));
    }
    if (o.user_bias) CUDA_TRY(cudaMemcpyAsync(o.user_bias, s->ub, (size_t)s->rows * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (o.item_bias) CU2B_TRY(download_item_bias(s, o.item_bias, st));
    return CU2B_OK;
}

// The loop of training.cu:107-170, enqueued as a whole. out != nullptr: the model is also downloaded; when the call
// ends with a loss check (it does whenever it reaches total_iterations, training.cu:118) the D2H copies run on a
// second stream WHILE that check evaluates the final model -- the check only reads it.
cu2b_status session_run_impl(cu2b_session *s, int n_iterations, const ModelOut *out) {
    if (!s || n_iterations < 0) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_run: bad argument");
    if (s->dsgd_child) return cu2b_fail(CU2B_ERR_INVALID, "this session belongs to a DSGD context; use cu2b_dsgd_run");
    CUDA_TRY(cudaSetDevice(s->device));
    if (out && !s->out_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->out_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_model_final, cudaEventDisableTiming));
    }
    const int total = s->cfg.total_iterations, ce = s->cfg.check_error;
    auto is_check = [&](int i) {  // training.cu:118
        return (i + 1) % ce == 0 || i == 0 || (total > 0 && (i + 1) % total == 0);
    };
    const int tot_id = s->timing.begin(Timing::TOTAL, s->stream);
    int i = s->iter_done;
    const int end = i + n_iterations;
    bool downloading = false;
    while (i < end) {
        int j = i;
        while (j < end && !is_check(j)) ++j;  // next check iteration (or end)
        const int seg_end = std::min(j + 1, end);
        if (s->n_active > 0) {
            if (s->cfg.mode == CU2B_MODE_DETERMINISTIC)
                CU2B_TRY(enqueue_blocked_iterations(s, seg_end - i));
            else
                CU2B_TRY(enqueue_sgd_iterations(s, s->cfg.cur_iterations + i, seg_end - i));
        }
        if (out && seg_end == end && j < end) {  // the model is final: it leaves while the last check reads it
            CU2B_TRY(stream_after(s->out_stream, s->stream, s->ev_model_final));
            CU2B_TRY(enqueue_download(s, s->out_stream, *out));
            downloading = true;
        }
        if (j < end) CU2B_TRY(enqueue_check(s, j + 1, 1, true));
        i = seg_end;
    }
    s->timing.end(tot_id, s->stream);
    if (out && !downloading) CU2B_TRY(enqueue_download(s, s->stream, *out));
    CUDA_TRY(cudaStreamSynchronize(s->stream));  // the only host sync of the loop
    if (downloading) CUDA_TRY(cudaStreamSynchronize(s->out_stream));
    s->iter_done = end;
    double ms[Timing::NKIND] = {0, 0, 0, 0, 0, 0};
    s->timing.collect(ms);
    s->stats.sgd_ms += ms[Timing::SGD];
    s->stats.loss_ms += ms[Timing::LOSS];
    s->stats.sampler_ms += ms[Timing::SAMPLER];
    s->stats.total_ms += ms[Timing::TOTAL];
    return check_device_error(s);  // a timed-out device-side wait, or a non-finite loss check (CU2B_ERR_DIVERGED)
}
}  // namespace

extern "C" cu2b_status cu2b_session_run(cu2b_session *s, int n_iterations) { return session_run_impl(s, n_iterations, nullptr); }

extern "C" cu2b_status cu2b_session_run_download(cu2b_session *s, int n_iterations, float *P, float *Q, float *user_bias,
                                                 float *item_bias) {
    const ModelOut out{P, Q, user_bias, item_bias};
    return session_run_impl(s, n_iterations, &out);
}

static cu2b_status session_reload_impl(cu2b_session *s, const cu2b_csr *train, const cu2b_csr *test, const float *P,
                                       const float *Q, const float *user_bias, const float *item_bias,
                                       float global_bias) {
    if (!s || !P || !Q || !user_bias || !item_bias) return cu2b_fail(CU2B_ERR_INVALID, "reload: null argument");
    if (s->cfg.mode != CU2B_MODE_HOGWILD)
        return cu2b_fail(CU2B_ERR_UNSUPPORTED, "reload: the deterministic mode builds its block schedule at creation");
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->sampler_stream) CUDA_TRY(cudaStreamSynchronize(s->sampler_stream));
    double ms[Timing::NKIND] = {0, 0, 0, 0, 0, 0};
    s->timing.collect(ms);  // drop spans of the previous life
    Trace tr("session_reload");
    CU2B_TRY(session_upload(s, train, test, P, Q, user_bias, item_bias, true));
    apply_stability_cap(s);  // the reloaded matrix may concentrate its draws differently
    s->mu = global_bias;
    CU2B_TRY(session_reset_state(s));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    tr.mark("uploads + COO expansion + state");
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_reload(cu2b_session *s, const cu2b_csr *train, const cu2b_csr *test, const float *P,
                                           const float *Q, const float *user_bias, const float *item_bias,
                                           float global_bias) {
    if (s && s->dsgd_child) return cu2b_fail(CU2B_ERR_INVALID, "this session belongs to a DSGD context; use cu2b_dsgd_reload");
    return session_reload_impl(s, train, test, P, Q, user_bias, item_bias, global_bias);
}

extern "C" cu2b_status cu2b_session_eval(cu2b_session *s, float *train_mae, float *train_rmse,
                                         float *test_mae, float *test_rmse) {
    if (!s) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_eval: null session");
    CUDA_TRY(cudaSetDevice(s->device));
    CU2B_TRY(enqueue_check(s, 0, 0, false));
    DevState st;
    CUDA_TRY(cudaMemcpyAsync(&st, s->state, sizeof(st), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    double ms[Timing::NKIND] = {0, 0, 0, 0, 0, 0};
    s->timing.collect(ms);
    s->stats.loss_ms += ms[Timing::LOSS];
    if (train_rmse) *train_rmse = (float)sqrt(st.sums[0] / (double)s->train.nnz);
    if (train_mae) *train_mae = (float)(st.sums[1] / (double)s->train.nnz);
    if (test_rmse) *test_rmse = (float)sqrt(st.sums[2] / (double)s->test.nnz);
    if (test_mae) *test_mae = (float)(st.sums[3] / (double)s->test.nnz);
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_log(cu2b_session *s, cu2b_metrics *out, int cap, int *n) {
    if (!s || !n) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_log: bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    DevState st;
    CUDA_TRY(cudaMemcpy(&st, s->state, sizeof(st), cudaMemcpyDeviceToHost));
    const int have = std::min(st.n_log, s->log_cap);
    *n = have;
    const int take = std::min(have, cap);
    if (out && take > 0)
        CUDA_TRY(cudaMemcpy(out, s->log_dev, (size_t)take * sizeof(cu2b_metrics), cudaMemcpyDeviceToHost));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_download(cu2b_session *s, float *P, float *Q, float *user_bias,
                                             float *item_bias) {
    if (!s) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_download: null session");
    CUDA_TRY(cudaSetDevice(s->device));
    Trace tr("session_download");
    CU2B_TRY(enqueue_download(s, s->stream, ModelOut{P, Q, user_bias, item_bias}));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    tr.mark("D2H");
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_get_config(cu2b_session *s, cu2b_config *out) {
    if (!s || !out) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_get_config: bad argument");
    CUDA_TRY(cudaSetDevice(s->device));
    DevState st;
    CUDA_TRY(cudaMemcpy(&st, s->state, sizeof(st), cudaMemcpyDeviceToHost));
    *out = s->cfg;
    out->learning_rate = st.lr;                                   // training.cu:151
    out->cur_iterations = s->cfg.cur_iterations + s->iter_done;   // training.cu:170
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_session_stats(cu2b_session *s, cu2b_stats *out, int reset) {
    if (!s) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_session_stats: null session");
    if (out) *out = s->stats;
    if (reset) memset(&s->stats, 0, sizeof(s->stats));
    return CU2B_OK;
}

extern "C" void cu2b_session_destroy(cu2b_session *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    Trace tr("session_destroy");
    delete s;
    tr.mark("free");
}

namespace {
cu2b_status train_multi_gpu(const cu2b_csr *train, const cu2b_csr *test, cu2b_config *cfg, float *P, float *Q,
                            float *user_bias, float *item_bias, float global_bias, float *losses, cu2b_metrics *log,
                            int log_cap, int *n_log, cu2b_stats *stats);
}

extern "C" cu2b_status cu2b_train(const cu2b_csr *train, const cu2b_csr *test, cu2b_config *cfg,
                                  float *P, float *Q, float *user_bias, float *item_bias,
                                  float global_bias, int init_item_side, float *losses,
                                  cu2b_metrics *log, int log_cap, int *n_log, cu2b_stats *stats) {
    if (!train || !test || !cfg || !P || !Q || !user_bias || !item_bias)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_train: null argument");
    const int k = cfg->n_factors;
    if (k < 1) return cu2b_fail(CU2B_ERR_INVALID, "n_factors must be >= 1");
    // training.cu:212-213 (Q, item_bias) then training.cu:28,54 (P, user_bias): every array is
    // drawn from its own mt19937(42), so all four share a prefix of identical values.
    if (init_item_side) {
        cu2b_init_normal(Q, (int64_t)train->cols * k, k, 0.f, 1.f, 42);
        cu2b_init_normal(item_bias, train->cols, k, 0.f, 1.f, 42);
    }
    cu2b_init_normal(P, (int64_t)train->rows * k, k, 0.f, 1.f, 42);
    cu2b_init_normal(user_bias, train->rows, k, 0.f, 1.f, 42);
    if (cfg->n_gpus > 1)  // DSGD over n_gpus devices of this box (dsgd.inc)
        return train_multi_gpu(train, test, cfg, P, Q, user_bias, item_bias, global_bias, losses, log, log_cap, n_log, stats);
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    cu2b_session *s = nullptr;
    CU2B_TRY(cu2b_session_create(&s, dev, train, test, cfg, P, Q, user_bias, item_bias, global_bias));
    cu2b_status rc = cu2b_session_run_download(s, cfg->total_iterations, P, Q, user_bias, item_bias);
    std::vector<cu2b_metrics> rows_log;
    int have = 0;
    if (rc == CU2B_OK) {
        rows_log.resize((size_t)s->log_cap);
        rc = cu2b_session_log(s, rows_log.data(), s->log_cap, &have);
    }
    if (rc == CU2B_OK) {
        if (losses) {  // training.cu:158: only check iterations are written; we NaN the rest
            for (int i = 0; i < cfg->total_iterations; ++i) losses[i] = NAN;
            for (int r = 0; r < have; ++r) {
                const int it = rows_log[r].iteration;
                if (it >= 1 && it <= cfg->total_iterations) losses[it - 1] = rows_log[r].test_rmse;
            }
        }
        if (log) memcpy(log, rows_log.data(), (size_t)std::min(have, log_cap) * sizeof(cu2b_metrics));
        if (n_log) *n_log = have;
        if (stats) *stats = s->stats;
        cu2b_config after;
        rc = cu2b_session_get_config(s, &after);
        if (rc == CU2B_OK) {
            cfg->learning_rate = after.learning_rate;
            cfg->cur_iterations = after.cur_iterations;
        }
    }
    cu2b_session_destroy(s);
    return rc;
}

// ---------------------------------------------------------------------------------------
// kernel-level entry points (host buffers)
// ---------------------------------------------------------------------------------------
namespace {

// A throw-away model on the device for the host-buffer entry points below.
struct Scratch {
    cu2b_session s;  // reuses the session fields / launch helpers
    cu2b_status init(int rows, int cols, int k, const float *P, const float *Q, const float *ub,
                     const float *ib, float mu, const cu2b_config *cfg) {
        CU2B_TRY(check_device());
        CUDA_TRY(cudaGetDevice(&s.device));
        if (cfg) s.cfg = *cfg; else cu2b_config_default(&s.cfg);
        s.cfg.n_factors = k;
        s.mu = mu;
        s.k = k;
        s.kp = cu2b_padded_factors(k);
        CU2B_TRY(layout_for(s.kp, &s.L, &s.V));
        s.rows = rows;
        s.cols = cols;
        memset(&s.stats, 0, sizeof(s.stats));
        DeviceFacts facts;
        CU2B_TRY(device_facts(s.device, &facts));
        s.sm_count = facts.sm_count;
        CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU2B_TRY(s.pool.alloc(&s.P, (size_t)rows * s.kp));
        CU2B_TRY(s.pool.alloc(&s.Q, (size_t)cols * s.kp));
        CU2B_TRY(s.pool.alloc(&s.ub, (size_t)rows));
        CU2B_TRY(alloc_item_bias(&s));
        CU2B_TRY(upload_dense(s.stream, s.P, P, rows, k, s.kp));
        CU2B_TRY(upload_dense(s.stream, s.Q, Q, cols, k, s.kp));
        CUDA_TRY(cudaMemcpyAsync(s.ub, ub, (size_t)rows * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(s.ib_dense, ib, (size_t)cols * sizeof(float), cudaMemcpyHostToDevice, s.stream));
        CU2B_TRY(scatter_item_bias(&s, s.stream));
        s.sgd_kernel = pick_sgd(s.L, s.V);
        s.loss_kernel = pick_loss(s.kp);
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s.sgd_kernel, kThreads, 0));
        s.sgd_grid_max = std::max(1, occ) * s.sm_count;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s.loss_kernel, kThreads, 0));
        s.loss_grid_max = std::max(1, occ) * s.sm_count;
        s.counter_slots = 8;
        CU2B_TRY(s.pool.alloc(&s.counters, (size_t)s.counter_slots));
        CU2B_TRY(s.pool.alloc(&s.part_train, (size_t)2 * s.loss_grid_max));
        CU2B_TRY(s.pool.alloc(&s.state, 1));
        DevState st;
        memset(&st, 0, sizeof(st));
        st.lr = s.cfg.learning_rate;
        CUDA_TRY(cudaMemcpyAsync(s.state, &st, sizeof(st), cudaMemcpyHostToDevice, s.stream));
        return CU2B_OK;
    }
};

cu2b_status sum_partials(cudaStream_t st, const double *partials_dev, int nblk, double *sse, double *sae) {
    std::vector<double> h((size_t)2 * nblk);
    CUDA_TRY(cudaMemcpyAsync(h.data(), partials_dev, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    double a = 0, b = 0;
    for (int i = 0; i < nblk; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    *sse = a;
    *sae = b;
    return CU2B_OK;
}

cu2b_status loss_impl(const cu2b_csr *m, const float *P, const float *Q, const float *ub, const float *ib,
                      float mu, int k, float *mae, float *rmse, float *err) {
    CU2B_TRY(validate_csr(m, "cu2b_loss"));
    if (!P || !Q || !ub || !ib || k < 1) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_loss: bad argument");
    Scratch sc;
    CU2B_TRY(sc.init(m->rows, m->cols, k, P, Q, ub, ib, mu, nullptr));
    cu2b_session &s = sc.s;
    CU2B_TRY(upload_matrix(s.pool, s.stream, m, &s.train, nullptr));
    float *err_dev = nullptr;
    if (err) CU2B_TRY(s.pool.alloc(&err_dev, (size_t)m->nonzeros));
    int nblk = 0;
    CU2B_TRY(launch_loss(&s, s.train, s.part_train, &nblk, err_dev));
    double sse, sae;
    CU2B_TRY(sum_partials(s.stream, s.part_train, nblk, &sse, &sae));
    if (err && m->nonzeros > 0)
        CUDA_TRY(cudaMemcpy(err, err_dev, (size_t)m->nonzeros * sizeof(float), cudaMemcpyDeviceToHost));
    if (mae) *mae = (float)(sae / (double)m->nonzeros);
    if (rmse) *rmse = (float)sqrt(sse / (double)m->nonzeros);
    return CU2B_OK;
}

}  // namespace

extern "C" cu2b_status cu2b_loss(const cu2b_csr *m, const float *P, const float *Q, const float *ub,
                                 const float *ib, float mu, int k, float *mae, float *rmse) {
    return loss_impl(m, P, Q, ub, ib, mu, k, mae, rmse, nullptr);
}

extern "C" cu2b_status cu2b_residuals(const cu2b_csr *m, const float *P, const float *Q, const float *ub,
                                      const float *ib, float mu, int k, float *err) {
    if (!err) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_residuals: null output");
    return loss_impl(m, P, Q, ub, ib, mu, k, nullptr, nullptr, err);
}

extern "C" cu2b_status cu2b_error_metrics(const float *err, int64_t n, float *mae, float *rmse) {
    if (!err || n <= 0 || !mae || !rmse) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_error_metrics: bad argument");
    CU2B_TRY(check_device());
    DevPool pool;
    float *e_dev;
    double *part;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    CU2B_TRY(pool.alloc(&e_dev, (size_t)n));
    CU2B_TRY(pool.alloc(&part, (size_t)2 * grid));
    CUDA_TRY(cudaMemcpy(e_dev, err, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    error_metrics_kernel<<<grid, 256>>>(e_dev, (long long)n, part);
    CUDA_TRY(cudaGetLastError());
    double sse, sae;
    CU2B_TRY(sum_partials(0, part, grid, &sse, &sae));
    *mae = (float)(sae / (double)n);
    *rmse = (float)sqrt(sse / (double)n);
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_sample_per_user(const cu2b_csr *m, int seed, int iter0, int n_iter,
                                            cu2b_rating *out, int64_t *n_out) {
    CU2B_TRY(validate_csr(m, "cu2b_sample_per_user"));
    if (n_iter < 0 || !n_out) return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sample_per_user: bad argument");
    CU2B_TRY(check_device());
    DevPool pool;
    cudaStream_t st = 0;
    DevMatrix dm;
    std::vector<int> indptr;
    CU2B_TRY(upload_matrix(pool, st, m, &dm, &indptr));
    std::vector<int> active;
    for (int u = 0; u < m->rows; ++u)
        if (indptr[u + 1] > indptr[u]) active.push_back(u);
    const long long draws = (long long)n_iter * (long long)active.size();
    *n_out = draws;
    if (!out || draws == 0) return CU2B_OK;
    int *act_dev;
    cu2b_rating *out_dev;
    CU2B_TRY(pool.alloc(&act_dev, active.size()));
    CU2B_TRY(pool.alloc(&out_dev, (size_t)draws));
    CUDA_TRY(cudaMemcpy(act_dev, active.data(), active.size() * sizeof(int), cudaMemcpyHostToDevice));
    const int grid = (int)std::min<long long>((draws + 255) / 256, 148 * 16);
    sample_per_user_kernel<<<grid, 256, 0, st>>>(dm.indptr, dm.coo, act_dev, (int)active.size(), (uint32_t)seed,
                                                 iter0, draws, out_dev, (long long)active.size(), nullptr);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, out_dev, (size_t)draws * sizeof(cu2b_rating), cudaMemcpyDeviceToHost));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_sample_per_rating(const cu2b_csr *m, int seed, int64_t first_update, int64_t n_updates,
                                              cu2b_rating *out) {
    CU2B_TRY(validate_csr(m, "cu2b_sample_per_rating"));
    if (first_update < 0 || n_updates < 0 || (!out && n_updates > 0) || m->nonzeros <= 0)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sample_per_rating: bad argument");
    CU2B_TRY(check_device());
    if (n_updates == 0) return CU2B_OK;
    DevPool pool;
    cudaStream_t st = 0;
    DevMatrix dm;
    CU2B_TRY(upload_matrix(pool, st, m, &dm, nullptr));
    cu2b_rating *out_dev;
    CU2B_TRY(pool.alloc(&out_dev, (size_t)n_updates));
    const int grid = (int)std::min<long long>((n_updates + 255) / 256, 148 * 16);
    sample_per_rating_kernel<<<grid, 256, 0, st>>>(dm.coo, (unsigned long long)m->nonzeros, feistel_half_bits((unsigned long long)m->nonzeros),
                                                   (uint32_t)seed, (unsigned long long)first_update, (long long)n_updates,
                                                   (int)std::min<int64_t>(n_updates, INT32_MAX), out_dev, (long long)n_updates);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, out_dev, (size_t)n_updates * sizeof(cu2b_rating), cudaMemcpyDeviceToHost));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_sgd_apply(const cu2b_rating *stream, int64_t n, float *P, int rows, float *Q,
                                      int cols, float *ub, float *ib, float mu, const cu2b_config *cfg,
                                      int order) {
    if ((!stream && n > 0) || n < 0 || !P || !Q || !ub || !ib || !cfg || rows < 0 || cols < 0)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sgd_apply: bad argument");
    for (int64_t t = 0; t < n; ++t)
        if (stream[t].user < 0 || stream[t].user >= rows || stream[t].item < 0 || stream[t].item >= cols)
            return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sgd_apply: rating %ld out of range", (long)t);
    Scratch sc;
    CU2B_TRY(sc.init(rows, cols, cfg->n_factors, P, Q, ub, ib, mu, cfg));
    cu2b_session &s = sc.s;
    if (n > 0) {
        cu2b_rating *sdev;
        CU2B_TRY(s.pool.alloc(&sdev, (size_t)n + kChunkMax + 4));
        CUDA_TRY(cudaMemcpyAsync(sdev, stream, (size_t)n * sizeof(cu2b_rating), cudaMemcpyHostToDevice, s.stream));
        const int chunk = order == 1 ? kChunkMax : pick_chunk(n, s.sgd_grid_max);
        CU2B_TRY(launch_sgd(&s, flat_view(sdev, n, chunk), nullptr, 0, order == 1));
    }
    CU2B_TRY(download_dense(s.stream, P, s.P, rows, s.k, s.kp));
    CU2B_TRY(download_dense(s.stream, Q, s.Q, cols, s.k, s.kp));
    CUDA_TRY(cudaMemcpyAsync(ub, s.ub, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CU2B_TRY(download_item_bias(&s, ib, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_sgd_blocked(const cu2b_rating *coo, int64_t n, float *P, int rows, float *Q, int cols,
                                        float *ub, float *ib, float mu, const cu2b_config *cfg, int n_blocks,
                                        int n_passes) {
    if ((!coo && n > 0) || n < 0 || !P || !Q || !ub || !ib || !cfg || rows < 0 || cols < 0 || n_passes < 0)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sgd_blocked: bad argument");
    for (int64_t t = 0; t < n; ++t)
        if (coo[t].user < 0 || coo[t].user >= rows || coo[t].item < 0 || coo[t].item >= cols)
            return cu2b_fail(CU2B_ERR_INVALID, "cu2b_sgd_blocked: rating %ld out of range", (long)t);
    Scratch sc;
    CU2B_TRY(sc.init(rows, cols, cfg->n_factors, P, Q, ub, ib, mu, cfg));
    cu2b_session &s = sc.s;
    if (n > 0 && n_passes > 0) {
        BlockSchedule bs;
        CU2B_TRY(build_block_schedule(coo, n, rows, cols, n_blocks > 0 ? n_blocks : auto_blocks(rows, cols), &bs));
        cu2b_rating *sched;
        int *ptr;
        CU2B_TRY(s.pool.alloc(&sched, bs.sched.size()));
        CU2B_TRY(s.pool.alloc(&ptr, bs.ptr.size()));
        CUDA_TRY(cudaMemcpyAsync(sched, bs.sched.data(), bs.sched.size() * sizeof(cu2b_rating), cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(cudaMemcpyAsync(ptr, bs.ptr.data(), bs.ptr.size() * sizeof(int), cudaMemcpyHostToDevice, s.stream));
        s.blocked_kernel = pick_blocked(s.L, s.V);
        for (int pass = 0; pass < n_passes; ++pass) CU2B_TRY(enqueue_blocked_pass(&s, sched, ptr, bs.B));
    }
    CU2B_TRY(download_dense(s.stream, P, s.P, rows, s.k, s.kp));
    CU2B_TRY(download_dense(s.stream, Q, s.Q, cols, s.k, s.kp));
    CUDA_TRY(cudaMemcpyAsync(ub, s.ub, (size_t)rows * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
    CU2B_TRY(download_item_bias(&s, ib, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_block_schedule_order(const cu2b_rating *coo, int64_t n, int rows, int cols, int n_blocks,
                                                 int64_t *order, int *n_blocks_used) {
    if ((!coo && n > 0) || n < 0 || !order || rows < 0 || cols < 0)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_block_schedule_order: bad argument");
    for (int64_t t = 0; t < n; ++t)
        if (coo[t].user < 0 || coo[t].user >= rows || coo[t].item < 0 || coo[t].item >= cols)
            return cu2b_fail(CU2B_ERR_INVALID, "cu2b_block_schedule_order: rating %ld out of range", (long)t);
    BlockSchedule bs;
    const int B = n_blocks > 0 ? n_blocks : auto_blocks(rows, cols);
    CU2B_TRY(build_block_schedule(coo, n, rows, cols, B, &bs, order));
    if (n_blocks_used) *n_blocks_used = B;
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_release_cache(void) {
    cudaMemPool_t *pools = library_pools();
    for (int dev = 0; dev < 64; ++dev)
        if (pools[dev]) CUDA_TRY(cudaMemPoolTrimTo(pools[dev], 0));
    return CU2B_OK;
}

extern "C" cu2b_status cu2b_device_info(int device, char *name, int name_cap, int *sm_count, int *cc_major,
                                        int *cc_minor, int64_t *free_bytes, int64_t *total_bytes) {
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (name && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (free_bytes || total_bytes) {
        CUDA_TRY(cudaSetDevice(device));
        size_t f = 0, t = 0;
        CUDA_TRY(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = (int64_t)f;
        if (total_bytes) *total_bytes = (int64_t)t;
    }
    return CU2B_OK;
}

#include "dsgd.inc"
