// Single-pass RMSE / MAE (replaces loss_kernel + two total_loss_kernel launches + host sums,
// loss.cu:19-35,58-128,150-200). One pass over the rating triplets: the residual is formed
// with the same lane layout / shuffle reduction as the SGD kernel and accumulated straight
// into double precision |err| and err^2 sums; nothing is written per rating unless the caller
// asks for the residual vector (the calculate_loss_gpu contract).
//
// Determinism: chunks are assigned statically (chunk = blockIdx.x + i * gridDim.x), lanes and
// warps are reduced in a fixed order, per-CTA partials are summed by finalize kernels in a
// fixed order => bitwise reproducible sums for a given model.
#ifndef CU2B_LOSS_KERNELS_CUH_
#define CU2B_LOSS_KERNELS_CUH_

#include "sgd_kernels.cuh"

namespace cu2b {

struct LossParams {
    StreamView sv;  // flat stream over the matrix' COO triplets
    const float *P, *Q, *user_bias, *item_bias;
    int kp;
    int ibs;           // item_bias stride in floats
    float mu;
    double *partials;  // [gridDim.x][2] = {sum err^2, sum |err|}
    float *err_out;    // optional residual vector (stream order), may be nullptr
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// U = ratings a lane group keeps in flight (their Q-row slices are requested back to back before the first one is
// used). Measured on the Netflix shape at k = 128 (profiles/r2_loss_kernel.md): U = 1 at 4 CTAs/SM 3.4 ms per check,
// U = 2 at 3 CTAs/SM 5.5 ms, U = 4 at 2 CTAs/SM 8.6 ms -- the kernel is bound by the L2 -> SM row traffic (14.7 of the
// 20.5 TB/s gather peak), not by latency, and every spilled register adds to exactly that traffic. U = 1 is the product.
constexpr int loss_min_ctas(int U) { return U >= 4 ? 2 : U == 2 ? 3 : 4; }

template <int L, int V, int U = 1>
__global__ void __launch_bounds__(kThreads, loss_min_ctas(U))
mf_loss_fused(const LossParams p) {
    __shared__ StreamSmem sm;
    __shared__ double wsum[kConsumerWarps][2];
    pipe_init(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {
        if (lane == 0) {
            long long next = blockIdx.x;
            pipe_produce(
                sm, p.sv,
                [&]() -> long long {
                    const long long c = next;
                    next += gridDim.x;
                    return c < p.sv.num_chunks ? c : -1LL;
                },
                [](long long, int, int) {});
        }
        return;
    }
    constexpr int G = 32 / L;
    constexpr int kGroups = kConsumerWarps * G;
    const int g = lane / L, l = lane % L;
    const int vecs = p.kp >> 2;
    double sse = 0.0, sae = 0.0;
    // The ratings arrive in CSR order and a lane group takes a CONTIGUOUS run of every chunk, so consecutive ratings of
    // a group mostly belong to the same user: its slice of that user's P row and the user bias stay in registers
    // until the user changes (same operands, same operation order per rating, same residual bits).
    float4 pu[V];
    float ub_cur = 0.f;
    int cur_user = -1;
    for (int it = 0;; ++it) {
        const int s = it % kStages;
        mbar_wait(&sm.full[s], (it / kStages) & 1);
        const int cnt = sm.count[s];
        if (cnt < 0) break;
        const long long c = sm.chunk_id[s];
        const int per = (cnt + kGroups - 1) / kGroups;  // warp-uniform trip count: the group sums are shuffles
        const int j0 = (warp * G + g) * per, j1 = min(cnt, j0 + per);
        for (int t = 0; t < per; t += U) {
            cu2b_rating rt[U];
            bool ok[U];
            float4 q[U][V];
            float ib[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + t + u;
                ok[u] = j < j1;
                rt[u] = sm.stage[s][ok[u] ? j : 0];
                const float4 *qrow = reinterpret_cast<const float4 *>(p.Q + (size_t)rt[u].item * p.kp);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int idx = v * L + l;
                    q[u][v] = (ok[u] && idx < vecs) ? __ldg(qrow + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                ib[u] = ok[u] ? __ldg(p.item_bias + (size_t)rt[u].item * p.ibs) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u] && rt[u].user != cur_user) {
                    const float4 *prow = reinterpret_cast<const float4 *>(p.P + (size_t)rt[u].user * p.kp);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int idx = v * L + l;
                        pu[v] = idx < vecs ? __ldg(prow + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    ub_cur = __ldg(p.user_bias + rt[u].user);
                    cur_user = rt[u].user;
                }
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int idx = v * L + l;
                    if (ok[u] && idx < vecs) {
                        const float4 a = pu[v], b = q[u][v];
                        acc = __fmaf_rn(a.x, b.x, acc);
                        acc = __fmaf_rn(a.y, b.y, acc);
                        acc = __fmaf_rn(a.z, b.z, acc);
                        acc = __fmaf_rn(a.w, b.w, acc);
                    }
                }
                const float ub = ok[u] ? ub_cur : 0.f;
                const float dot = group_sum<L>(acc);
                const float pred = __fadd_rn(__fadd_rn(__fadd_rn(p.mu, ub), ib[u]), dot);
                const float err = __fsub_rn(rt[u].rating, pred);
                if (ok[u] && l == 0) {
                    sse += (double)err * (double)err;
                    sae += (double)fabsf(err);
                    if (p.err_out) p.err_out[c * p.sv.chunk + j0 + t + u] = err;  // flat stream: seg_pitch unused
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    sse = warp_sum_d(sse);
    sae = warp_sum_d(sae);
    if (lane == 0) { wsum[warp][0] = sse; wsum[warp][1] = sae; }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");  // consumers only
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kConsumerWarps; ++w) { a += wsum[w][0]; b += wsum[w][1]; }
        p.partials[2 * blockIdx.x] = a;
        p.partials[2 * blockIdx.x + 1] = b;
    }
}

// get_error_metrics_gpu contract (loss.cu:196-200) on an explicit error vector.
__global__ void __launch_bounds__(256)
error_metrics_kernel(const float *__restrict__ err, long long n, double *partials) {
    __shared__ double wsum[8][2];
    double sse = 0.0, sae = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float e = __ldg(err + i);
        sse += (double)e * (double)e;
        sae += (double)fabsf(e);
    }
    sse = warp_sum_d(sse);
    sae = warp_sum_d(sae);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { wsum[warp][0] = sse; wsum[warp][1] = sae; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += wsum[w][0]; b += wsum[w][1]; }
        partials[2 * blockIdx.x] = a;
        partials[2 * blockIdx.x + 1] = b;
    }
}

// Device-resident training state (replaces the host variables of training.cu:95-104 and the
// nine cudaMemcpyToSymbol calls per learning-rate change, config.cu:24-35).
struct DevState {
    float lr;
    int current_patience;   // training.cu:103 (int copy of the float cfg->patience)
    int patience0;
    float lr_decay;
    float validation_rmse;  // training.cu:102 (starts at FLT_MAX)
    int n_log;
    int log_cap;
    int error;              // set by a device-side wait that timed out (never hang the GPU)
    double sums[4];         // last evaluated {train sse, train sae, test sse, test sae}
};

// Divergence guard: asynchronous SGD can blow up (too large a learning rate for the number of
// concurrently applied updates on a popular item). A non-finite metric at a loss check
// is recorded as error = -(iteration + 1); the host turns it into CU2B_ERR_DIVERGED, so a diverged
// run can never report throughput with rc 0. (A NaN also blinds the reference's patience rule,
// training.cu:146: `last < NaN` is false.)
__device__ __forceinline__ void flag_non_finite(DevState *st, float train_rmse, float test_rmse, long long n_train,
                                                long long n_test, int iteration) {
    const bool bad = (n_train > 0 && !isfinite(train_rmse)) || (n_test > 0 && !isfinite(test_rmse));
    if (bad) atomicCAS(&st->error, 0, -(max(iteration, 0) + 1));
}

// Sums the per-CTA partials of one matrix in a fixed order. One block of 256 threads.
__device__ __forceinline__ void reduce_partials(const double *partials, int nblk, double *out2,
                                                double (*sh)[2]) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += partials[2 * i]; b += partials[2 * i + 1]; }
    sh[threadIdx.x][0] = a;
    sh[threadIdx.x][1] = b;
    __syncthreads();
    for (int w = blockDim.x / 2; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) { sh[threadIdx.x][0] += sh[threadIdx.x + w][0]; sh[threadIdx.x][1] += sh[threadIdx.x + w][1]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sh[0][0]; out2[1] = sh[0][1]; }
    __syncthreads();
}

// The check step of training.cu:118-158 on the device: metrics from the partial sums, patience
// bookkeeping, learning-rate decay, one log row. apply_schedule == 0 => evaluation only.
__global__ void __launch_bounds__(256)
loss_finalize_kernel(DevState *st, const double *part_train, int nblk_train, long long n_train,
                     const double *part_test, int nblk_test, long long n_test, int iteration,
                     int apply_schedule, cu2b_metrics *log) {
    __shared__ double sh[256][2];
    __shared__ double tot[4];
    reduce_partials(part_train, nblk_train, &tot[0], sh);
    reduce_partials(part_test, nblk_test, &tot[2], sh);
    if (threadIdx.x == 0) {
        st->sums[0] = tot[0]; st->sums[1] = tot[1]; st->sums[2] = tot[2]; st->sums[3] = tot[3];
        const float train_rmse = (float)sqrt(tot[0] / (double)n_train);
        const float train_mae = (float)(tot[1] / (double)n_train);
        const float test_rmse = (float)sqrt(tot[2] / (double)n_test);
        const float test_mae = (float)(tot[3] / (double)n_test);
        flag_non_finite(st, train_rmse, test_rmse, n_train, n_test, iteration);
        if (apply_schedule) {
            const float last = st->validation_rmse;
            st->validation_rmse = test_rmse;
            if (last < test_rmse) st->current_patience--;
            if (st->current_patience <= 0) {
                st->current_patience = st->patience0;
                st->lr = st->lr * st->lr_decay;
            }
        }
        if (log && st->n_log < st->log_cap) {
            cu2b_metrics m;
            m.iteration = iteration;
            m.train_mae = train_mae; m.train_rmse = train_rmse;
            m.test_mae = test_mae; m.test_rmse = test_rmse;
            m.learning_rate = st->lr;
            log[st->n_log] = m;
        }
        if (log) st->n_log++;
    }
}

}  // namespace cu2b
#endif  // CU2B_LOSS_KERNELS_CUH_
