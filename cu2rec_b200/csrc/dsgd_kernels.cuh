// Device side of the multi-GPU DSGD path: bucketing of a sampled round by item block, the
// peer-memory hand-off of an item block (Q rows + item_bias) to the next rank, and the
// all-rank combine of the loss partial sums. All inter-rank traffic is P2P stores into the
// peer's HBM over NVLink followed by a system-scope release flag; receivers poll their own
// memory. No host round trip anywhere in a round.
#ifndef CU2B_DSGD_KERNELS_CUH_
#define CU2B_DSGD_KERNELS_CUH_

#include "loss_kernels.cuh"

namespace cu2b {

constexpr int kMaxWorld = 16;
constexpr long long kSpinTimeoutCycles = 20LL * 2000000000LL;  // ~20 s at 2 GHz: never hang the box

__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Spins (one thread) until *flag >= need. On timeout records an error and returns so that the
// stream drains instead of hanging.
__global__ void dsgd_wait_kernel(const int *flag, int need, int *error_flag) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < need) {
        __nanosleep(200);
        if (clock64() - t0 > kSpinTimeoutCycles) {
            atomicExch(error_flag, 1);
            return;
        }
    }
}

__device__ __forceinline__ int item_block_of(int item, const int *sh_ptr, int world) {
    int b = 0;
#pragma unroll 1
    while (b + 1 < world && item >= sh_ptr[b + 1]) ++b;
    return b;
}

// Pass 1: how many draws of this round fall into each item block.
__global__ void __launch_bounds__(256)
dsgd_bucket_count_kernel(const cu2b_rating *__restrict__ stream, long long n_draws, int seg_len,
                         long long seg_pitch, const int *__restrict__ item_block_ptr, int world,
                         int *counts) {
    __shared__ int sh_ptr[kMaxWorld + 1];
    __shared__ int sh_cnt[kMaxWorld];
    if (threadIdx.x <= world) sh_ptr[threadIdx.x] = item_block_ptr[threadIdx.x];
    if (threadIdx.x < world) sh_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_draws;
         i += (long long)gridDim.x * blockDim.x) {
        const long long seg = i / seg_len;
        const int item = __ldg(&stream[seg * seg_pitch + (i - seg * seg_len)].item);
        atomicAdd(&sh_cnt[item_block_of(item, sh_ptr, world)], 1);
    }
    __syncthreads();
    if (threadIdx.x < world && sh_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sh_cnt[threadIdx.x]);
}

// Bucket b occupies [ranges[2b], ranges[2b] + ranges[2b+1]) of the bucket buffer; starts are
// multiples of 4 ratings (TMA alignment). Also resets the scatter cursors and the counts.
__global__ void dsgd_bucket_scan_kernel(int *counts, int world, int *ranges, int *cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int off = 0;
        for (int b = 0; b < world; ++b) {
            ranges[2 * b] = off;
            ranges[2 * b + 1] = counts[b];
            cursor[b] = off;
            off += (counts[b] + 3) & ~3;
            counts[b] = 0;
        }
    }
}

// Pass 2: tile-wise scatter; a CTA reserves one contiguous range per bucket for its tile.
constexpr int kScatterTile = 2048;
__global__ void __launch_bounds__(256)
dsgd_bucket_scatter_kernel(const cu2b_rating *__restrict__ stream, long long n_draws, int seg_len,
                           long long seg_pitch, const int *__restrict__ item_block_ptr, int world,
                           int *cursor, cu2b_rating *__restrict__ out) {
    __shared__ int sh_ptr[kMaxWorld + 1];
    __shared__ int sh_cnt[kMaxWorld];
    __shared__ int sh_base[kMaxWorld];
    if (threadIdx.x <= world) sh_ptr[threadIdx.x] = item_block_ptr[threadIdx.x];
    constexpr int PER = kScatterTile / 256;
    for (long long tile = blockIdx.x; tile * kScatterTile < n_draws; tile += gridDim.x) {
        if (threadIdx.x < world) sh_cnt[threadIdx.x] = 0;
        __syncthreads();
        cu2b_rating r[PER];
        int bk[PER], loc[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const long long i = tile * kScatterTile + e * 256 + threadIdx.x;
            bk[e] = -1;
            if (i < n_draws) {
                const long long seg = i / seg_len;
                r[e] = stream[seg * seg_pitch + (i - seg * seg_len)];
                bk[e] = item_block_of(r[e].item, sh_ptr, world);
                loc[e] = atomicAdd(&sh_cnt[bk[e]], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x < world) sh_base[threadIdx.x] = sh_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], sh_cnt[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < PER; ++e)
            if (bk[e] >= 0) out[sh_base[bk[e]] + loc[e]] = r[e];
        __syncthreads();
    }
}

// Hands item block rows [row0, row1) (Q rows + item_bias) to a peer: plain 128-bit stores into
// the peer's arrays (NVLink P2P), then the last CTA publishes flag = value with system scope.
// dst_flag == nullptr => plain local copy without signalling.
__global__ void __launch_bounds__(256)
dsgd_send_block_kernel(const float4 *__restrict__ src_q, float4 *__restrict__ dst_q, long long n_vec,
                       const float *__restrict__ src_ib, float *__restrict__ dst_ib, int n_items,
                       int *ticket, int *dst_flag, int value) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        dst_q[i] = __ldcg(src_q + i);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride)
        dst_ib[i] = __ldcg(src_ib + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        if (t == (int)gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
            if (dst_flag) st_release_sys(dst_flag, value);
        }
    }
}

// Loss check across ranks. Every rank reduces its per-CTA partials (train strip, test strip) in
// a fixed order, writes its four sums into slot [parity][rank] of every rank (peers through
// P2P), publishes flag = check_no, waits for all ranks, then adds the slots in rank order --
// identical bits on every rank -- and applies the schedule of training.cu:146-155.
struct DsgdLossPeers {
    double *slots[kMaxWorld];  // each: [2][kMaxWorld][4] doubles in that rank's memory
    int *flags[kMaxWorld];     // each: [kMaxWorld] ints in that rank's memory
};

__global__ void __launch_bounds__(256)
dsgd_loss_combine_kernel(DevState *st, const double *part_train, int nblk_train, const double *part_test,
                         int nblk_test, long long n_train_global, long long n_test_global, int iteration,
                         int apply_schedule, cu2b_metrics *log, DsgdLossPeers peers, int rank, int world,
                         int check_no, int *error_flag) {
    __shared__ double sh[256][2];
    __shared__ double tot[4];
    reduce_partials(part_train, nblk_train, &tot[0], sh);
    reduce_partials(part_test, nblk_test, &tot[2], sh);
    const int parity = check_no & 1;
    if ((int)threadIdx.x < world) {
        double *dst = peers.slots[threadIdx.x] + ((size_t)parity * kMaxWorld + rank) * 4;
        dst[0] = tot[0]; dst[1] = tot[1]; dst[2] = tot[2]; dst[3] = tot[3];
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, check_no);
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(peers.flags[rank] + threadIdx.x) < check_no) {
            __nanosleep(200);
            if (clock64() - t0 > kSpinTimeoutCycles) { atomicExch(error_flag, 1); break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->sums[0] = tot[0]; st->sums[1] = tot[1]; st->sums[2] = tot[2]; st->sums[3] = tot[3];  // local
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        const double *mine = peers.slots[rank] + (size_t)parity * kMaxWorld * 4;
        for (int r = 0; r < world; ++r)
            for (int c = 0; c < 4; ++c) g[c] += ((const volatile double *)mine)[r * 4 + c];
        const float train_rmse = (float)sqrt(g[0] / (double)n_train_global);
        const float train_mae = (float)(g[1] / (double)n_train_global);
        const float test_rmse = (float)sqrt(g[2] / (double)n_test_global);
        const float test_mae = (float)(g[3] / (double)n_test_global);
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
#endif  // CU2B_DSGD_KERNELS_CUH_
