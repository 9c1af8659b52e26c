// Dependent-load latency of a 512-byte row (one warp, float4 per lane) through different load
// flavours, with and without the 128-bit atomic add the SGD kernels issue after each row read.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldlat ldlat.cu ; run: ./ldlat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ld_cg(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_ca(const float4 *p) { return *p; }
__device__ __forceinline__ float4 ld_relaxed(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_weak_na(const float4 *p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_v4(float4 *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int MODE, int RED>
__global__ void chase(float4 *rows, int n_rows, int iters, long long *out) {
    const int lane = threadIdx.x & 31;
    unsigned idx = 12345u + blockIdx.x * 977u + (threadIdx.x >> 5) * 131u;
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        idx = idx * 1664525u + 1013904223u;
        const unsigned r = (idx >> 8) % (unsigned)n_rows;
        float4 *p = rows + (size_t)r * 32 + lane;
        float4 v = MODE == 0 ? ld_cg(p) : MODE == 1 ? ld_ca(p) : MODE == 2 ? ld_relaxed(p) : ld_weak_na(p);
        float s = v.x + v.y + v.z + v.w;
        // warp butterfly like the kernels (keeps the dependency chain comparable)
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        acc += s;
        idx += (unsigned)(__float_as_int(s) & 1);  // next address depends on the loaded data
        if (RED == 1) red_v4(p, make_float4(0.f, 0.f, 0.f, 0.f));
        if (RED == 2) red_v4(rows + (size_t)((r + 7777u) % (unsigned)n_rows) * 32 + lane, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}
template <int MODE, int RED>
void run(const char *name, float4 *rows, int n_rows, long long *out, int grid, int block) {
    const int iters = 20000;
    chase<MODE, RED><<<grid, block>>>(rows, n_rows, 2000, out);
    chase<MODE, RED><<<grid, block>>>(rows, n_rows, iters, out);
    long long h[2];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-34s grid %4d x %3d : %7.1f cycles / dependent row\n", name, grid, block, (double)h[0] / iters);
}
int main() {
    const int n_rows = 17770;
    float4 *rows; long long *out;
    cudaMalloc(&rows, (size_t)n_rows * 512); cudaMemset(rows, 0, (size_t)n_rows * 512); cudaMalloc(&out, 16);
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int grid = cfg == 0 ? 1 : cfg == 1 ? 148 : 740, block = cfg == 0 ? 32 : 256;
        run<0, 0>("ld.cg", rows, n_rows, out, grid, block);
        run<1, 0>("ld (default .ca)", rows, n_rows, out, grid, block);
        run<2, 0>("ld.relaxed.gpu", rows, n_rows, out, grid, block);
        run<3, 0>("ld.L1::no_allocate", rows, n_rows, out, grid, block);
        run<0, 1>("ld.cg + red.v4 same row", rows, n_rows, out, grid, block);
        run<0, 2>("ld.cg + red.v4 other row", rows, n_rows, out, grid, block);
        run<3, 1>("ld.no_allocate + red.v4 same row", rows, n_rows, out, grid, block);
        run<2, 1>("ld.relaxed.gpu + red.v4 same row", rows, n_rows, out, grid, block);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
