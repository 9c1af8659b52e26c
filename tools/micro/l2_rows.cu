// L2 row-traffic micro-benchmarks behind the SGD kernels' roofline:
// what can one B200 sustain of the access pattern "warp reads a 512-byte item row from L2
// (ld.global.cg.v4), then adds a 512-byte step to it with red.global.add.v4.f32, plus one 4-byte
// item-bias read and one 4-byte bias red", for uniform and Zipf-Mandelbrot row popularity, for a
// dense and a line-padded bias array, and for a handful of rows hammered by every warp.
//
// One warp = one row access at a time (like a lane group of the k = 128 update kernels); `unroll`
// independent accesses per warp give the memory-level parallelism of deeper look-ahead.
// Output: one JSON line per configuration on stdout.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o l2_rows l2_rows.cu
// run:   ./l2_rows [quick]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e__ = (x);                                                                   \
        if (e__ != cudaSuccess) {                                                                \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                             \
        }                                                                                        \
    } while (0)

__device__ __forceinline__ float4 ld_cg4(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cg1(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red4(float4 *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red1(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

enum Mode { GATHER = 0, RED = 1, GATHER_RED = 2, SGD_LIKE = 3, BIAS_ONLY = 4 };

// idx: [n] row ids; warp w processes idx[w * per_warp ... ), UN accesses in flight.
// PLANAR: line j of every row lives in plane j ([4][n_rows][128 bytes]) instead of row-major.
template <int MODE, int UN, int PLANAR = 0>
__global__ void __launch_bounds__(256)
rows_kernel(float4 *rows, float *bias, int bias_stride, const int *__restrict__ idx, long long n, float *sink, int n_rows,
            float4 *hot, int n_hot) {
    const int lane = threadIdx.x & 31;
    // PLANAR == 2: "hot table": rows with id < n_hot (the popular ones; ids are popularity ranks here) are stored
    // sector-scattered -- sector j (32 bytes, lanes 2j and 2j+1) of hot row h starts 1 KB block h * 16 + j of a
    // separate table, so one hot row spreads over 16 L2 slice pairs; the other rows stay row-major.
    auto at = [&](int r) -> float4 * {
        if (PLANAR == 2) {
            if (r < n_hot) return hot + (((size_t)r * 16 + (lane >> 1)) * 64 + (lane & 1));
            return rows + (size_t)r * 32 + lane;
        }
        return PLANAR ? rows + ((size_t)(lane >> 3) * n_rows + r) * 8 + (lane & 7) : rows + (size_t)r * 32 + lane;
    };
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (long long i = warp * UN; i + UN <= n; i += warps * UN) {
        int r[UN];
        float4 v[UN];
        float b[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) r[u] = __ldg(idx + i + u);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            if (MODE == SGD_LIKE || MODE == BIAS_ONLY) b[u] = ld_cg1(bias + (size_t)r[u] * bias_stride);
            if (MODE == GATHER || MODE == GATHER_RED || MODE == SGD_LIKE) v[u] = ld_cg4(at(r[u]));
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            float s = 0.f;
            if (MODE == GATHER || MODE == GATHER_RED || MODE == SGD_LIKE) s = v[u].x + v[u].y + v[u].z + v[u].w;
            if (MODE == SGD_LIKE || MODE == BIAS_ONLY) s += b[u];
            if (MODE == GATHER_RED || MODE == SGD_LIKE) {
                // the butterfly of the dot product: the step depends on the whole row
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            }
            acc += s;
            const float e = s * 1e-30f;  // data dependent, numerically nil
            if (MODE == RED || MODE == GATHER_RED || MODE == SGD_LIKE)
                red4(at(r[u]), make_float4(e, e, e, e));
            if ((MODE == SGD_LIKE || MODE == BIAS_ONLY) && lane == 0) red1(bias + (size_t)r[u] * bias_stride, e);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

static uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// dist: 0 uniform, 1 Zipf-Mandelbrot (weight 1/(rank + c), the bench generator's law) with
// scattered ids, 2 the same with id == popularity rank (popular items adjacent), 3 "paired": rank r of the
// popular half in row 2r, the other half fills the odd rows (one popular row per 1 KB), 4 "oct": rank r of the
// popular eighth in row 8r, the rest fills the other seven (one popular row per 8 rows; for the planar layout).
// law_c: Mandelbrot offset (17770 / 300 + 1 for the whole catalogue; a DSGD block of 1/G of the items keeps every
// G-th rank, i.e. the law 1 / (G * rank + c) = (1/G) / (rank + c / G)).
static std::vector<int> make_indices(long long n, int N, int dist, uint64_t seed, double law_c) {
    std::vector<int> out((size_t)n);
    std::vector<double> cdf(N);
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[i] = i;
    if (dist == 1)
        for (int i = N - 1; i > 0; --i) std::swap(perm[i], perm[(int)(mix64(seed + 77 + i) % (uint64_t)(i + 1))]);
    if (dist == 3) {
        const int h = (N + 1) / 2;
        for (int r = 0; r < N; ++r) perm[r] = r < h ? 2 * r : 2 * (r - h) + 1;
    }
    if (dist == 4) {
        const int h = (N + 7) / 8;
        std::vector<char> used(N, 0);
        for (int r = 0; r < h; ++r) { perm[r] = std::min(8 * r, N - 1); }
        for (int r = 0; r < h; ++r) used[perm[r]] = 1;
        int slot = 0;
        for (int r = h; r < N; ++r) { while (used[slot]) ++slot; perm[r] = slot; used[slot] = 1; }
    }
    if (dist != 0) {
        const double c = law_c;
        double acc = 0;
        for (int i = 0; i < N; ++i) { acc += 1.0 / (i + c); cdf[i] = acc; }
        for (int i = 0; i < N; ++i) cdf[i] /= acc;
        cdf[N - 1] = 1.0;
    }
#pragma omp parallel for schedule(static)
    for (long long j = 0; j < n; ++j) {
        const uint64_t h = mix64(mix64(seed) + (uint64_t)j);
        if (dist == 0) {
            out[j] = (int)(h % (uint64_t)N);
        } else {
            const double x = ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
            int rk = (int)(std::lower_bound(cdf.begin(), cdf.end(), x) - cdf.begin());
            out[j] = perm[std::min(rk, N - 1)];
        }
    }
    return out;
}

typedef void (*Kern)(float4 *, float *, int, const int *, long long, float *, int, float4 *, int);
static Kern pick(int mode, int un, int planar) {
    if (planar == 2) return mode == GATHER_RED ? (Kern)rows_kernel<GATHER_RED, 1, 2> : (Kern)rows_kernel<SGD_LIKE, 1, 2>;
    if (planar) return mode == GATHER_RED ? (Kern)rows_kernel<GATHER_RED, 1, 1> : (Kern)rows_kernel<SGD_LIKE, 1, 1>;
#define ROW(M) (un == 1 ? (Kern)rows_kernel<M, 1> : un == 2 ? (Kern)rows_kernel<M, 2> : (Kern)rows_kernel<M, 4>)
    switch (mode) {
        case GATHER: return ROW(GATHER);
        case RED: return ROW(RED);
        case GATHER_RED: return ROW(GATHER_RED);
        case SGD_LIKE: return ROW(SGD_LIKE);
        default: return ROW(BIAS_ONLY);
    }
#undef ROW
}

int main(int argc, char **argv) {
    const bool quick = argc > 1 && !strcmp(argv[1], "quick");
    const bool placement = argc > 1 && !strcmp(argv[1], "placement");
    const long long n = quick ? (1LL << 22) : (1LL << 24);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int maxN = 17770, max_stride = 64;
    float4 *rows;
    float *bias, *sink;
    int *idx_dev;
    CK(cudaMalloc(&rows, (size_t)maxN * 512));
    CK(cudaMemset(rows, 0, (size_t)maxN * 512));
    CK(cudaMalloc(&bias, (size_t)maxN * max_stride * 4));
    CK(cudaMemset(bias, 0, (size_t)maxN * max_stride * 4));
    CK(cudaMalloc(&sink, 4));
    float4 *hot;
    const int max_hot = 1024;
    CK(cudaMalloc(&hot, (size_t)max_hot * 16 * 1024));
    CK(cudaMemset(hot, 0, (size_t)max_hot * 16 * 1024));
    int n_hot = 0;
    CK(cudaMalloc(&idx_dev, (size_t)n * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char *mode_name[] = {"gather", "red", "gather_red", "sgd_like", "bias_only"};
    const char *dist_name[] = {"uniform", "zipf_scattered", "zipf_sorted_ids", "zipf_paired", "zipf_oct"};
    auto run = [&](int mode, int dist, int N, int stride, int ctas_per_sm, int un, int planar = 0) {
        Kern k = pick(mode, un, planar);
        const int grid = sms * ctas_per_sm;
        for (int w = 0; w < 2; ++w) k<<<grid, 256>>>(rows, bias, stride, idx_dev, n, sink, N, hot, n_hot);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            k<<<grid, 256>>>(rows, bias, stride, idx_dev, n, sink, N, hot, n_hot);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
        }
        CK(cudaGetLastError());
        const double rows_s = (double)(n / un * un) / (best * 1e-3);
        double bytes = 0;  // L2 bytes moved per access (row read 512, row red 512, bias 4 + 4)
        if (mode == GATHER) bytes = 512;
        if (mode == RED) bytes = 512;
        if (mode == GATHER_RED) bytes = 1024;
        if (mode == SGD_LIKE) bytes = 1032;
        if (mode == BIAS_ONLY) bytes = 8;
        printf("{\"mode\": \"%s\", \"dist\": \"%s\", \"layout\": \"%s\", \"n_hot\": %d, \"rows\": %d, \"bias_stride_floats\": %d, \"ctas_per_sm\": %d, "
               "\"unroll\": %d, \"ms\": %.4f, \"G_rows_per_s\": %.3f, \"l2_TB_per_s\": %.3f}\n",
               mode_name[mode], dist_name[dist], planar == 2 ? "hot_table" : planar ? "planar" : "row_major", planar == 2 ? n_hot : 0, N, stride,
               ctas_per_sm, un, best, rows_s / 1e9,
               rows_s * bytes / 1e12);
        fflush(stdout);
    };
    auto load = [&](int N, int dist) {
        std::vector<int> h = make_indices(n, N, dist, 20240607, N / 300.0 + 1.0);
        CK(cudaMemcpy(idx_dev, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    };
    if (argc > 1 && !strcmp(argv[1], "peak")) {
        // the roofline denominator bench.py uses: row read + atomic row add, uniformly random rows of the whole catalogue
        load(17770, 0);
        for (int un : {1, 4})
            for (int c : {4, 8}) run(GATHER_RED, 0, 17770, 1, c, un);
        for (int c : {4, 8}) run(SGD_LIKE, 0, 17770, 64, c, 1);
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "hot")) {
        // hot table: ids are popularity ranks (dist 2); the cold rows keep the plain popularity order
        for (int N : {17770, 4442, 2221}) {
            load(N, 2);
            for (int h : {0, 8, 32, 128, 512}) {
                n_hot = std::min(h, N);
                for (int c : {4, 8}) {
                    run(GATHER_RED, 2, N, 64, c, 1, h ? 2 : 0);
                    run(SGD_LIKE, 2, N, 64, c, 1, h ? 2 : 0);
                }
            }
        }
        return 0;
    }
    if (placement) {
        // Which internal row order balances the L2 slices? (bias padded to one per 256 bytes throughout)
        for (int N : {17770, 8885, 4442, 2221}) {
            for (int dist : {0, 1, 2, 3, 4}) {
                load(N, dist);
                for (int planar : {0, 1})
                    for (int c : {4, 8}) {
                        run(GATHER_RED, dist, N, 64, c, 1, planar);
                        run(SGD_LIKE, dist, N, 64, c, 1, planar);
                    }
            }
        }
        return 0;
    }
    // 1. peaks on the full catalogue: uniform vs Zipf, each traffic class alone and combined
    for (int dist = 0; dist < 3; ++dist) {
        load(17770, dist);
        for (int mode : {GATHER, RED, GATHER_RED}) {
            if (dist == 2 && mode != GATHER_RED) continue;
            for (int un : {1, 4})
                for (int c : {4, 8}) run(mode, dist, 17770, 1, c, un);
        }
        // the SGD access pattern with a dense and with padded bias arrays
        for (int stride : {1, 8, 32, 64})
            for (int c : {4, 8}) run(SGD_LIKE, dist, 17770, stride, c, 1);
        for (int stride : {1, 8, 32, 64}) run(BIAS_ONLY, dist, 17770, stride, 8, 4);
    }
    // 2. one DSGD item block (1/4 and 1/8 of the catalogue)
    for (int N : {4442, 2221}) {
        for (int dist : {0, 1}) {
            load(N, dist);
            for (int c : {2, 4, 8}) run(GATHER_RED, dist, N, 1, c, 1);
            for (int stride : {1, 32, 64})
                for (int c : {2, 4, 8}) run(SGD_LIKE, dist, N, stride, c, 1);
        }
    }
    // 3. a handful of rows hammered by every warp: per-row / per-line serialisation of red.v4
    for (int N : {1, 2, 4, 16, 64, 256}) {
        load(N, 0);
        run(RED, 0, N, 1, 8, 4);
        run(GATHER_RED, 0, N, 1, 8, 1);
        run(BIAS_ONLY, 0, N, 1, 8, 4);
        run(BIAS_ONLY, 0, N, 64, 8, 4);
    }
    return 0;
}
