// Device side of the multi-GPU DSGD path: per-user runs of a sampled round grouped by item block, the
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
__global__ void dsgd_wait_kernel(const int *flag, int need, int *error_flag, int site, long long timeout_cycles) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < need) {
        __nanosleep(200);
        if (clock64() - t0 > timeout_cycles) {
            atomicCAS(error_flag, 0, site * 1000000 + need * 100 + (ld_acquire_sys(flag) % 100));
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

// A sampled draw inside a user's run: the user is implied by the row. With item-step thinning
// (see dsgd_sample_runs_kernel) bit 31 of `item` marks a draw whose item ROW step is
// skipped and bit 30 one whose item BIAS step is skipped; item ids are < 2^30 then.
struct __align__(8) DsgdDraw {
    int item;
    float rating;
};
constexpr int kClaimStride = 64;   // claim counters 512 bytes apart: atomics on one L2 line serialise
constexpr int kClaimRanges = 64;
constexpr int kDrawsPerLane = 8;  // dsgd_sample_runs_kernel keeps a round's draws in registers: rounds of <= 256 iterations
constexpr int kDrawRowFrozen = (int)0x80000000u;
constexpr int kDrawBiasFrozen = 0x40000000;
constexpr int kDrawItemMask = 0x3fffffff;

// Issue-order-pinned load of a draw (volatile asm, like ldcg_pinned): the prefetch of the next
// draw stays ahead of the current update's item-row loads.
__device__ __forceinline__ DsgdDraw ld_draw_pinned(const DsgdDraw *p) {
    DsgdDraw d;
    asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(d.item), "=f"(d.rating) : "l"(p) : "memory");
    return d;
}

// Sampler of one DSGD round (nb reference iterations): one WARP per active user draws that
// user's nb ratings (same Philox stream as the single-GPU sampler, keyed by the original user
// id) and writes them to the user's row of `draws` grouped by item block, iteration order kept
// inside a block (ballot ranking => deterministic). row_off[a*(world+1)+b] = start of block b
// inside row a. In sub-epoch b the update kernel walks exactly run [row_off[b], row_off[b+1]).
//
// keep_bias (default for DSGD ranks, CU2B_DSGD_THIN_BIAS) / keep_row (opt-in, CU2B_DSGD_THIN): keep[i] in (0, 1] is the
// fraction of item i's draws whose item ROW step / item BIAS step is applied; a draw that skips one is
// flagged with kDrawRowFrozen / kDrawBiasFrozen (the user side always moves). The decision is the second
// Philox word of the draw's own counter compared with keep[item], so it is a pure function of (seed, user,
// iteration, keep[item]) and the two events are nested. This keeps the number of not-yet-visible steps on
// a popular item under the asynchronous-SGD stability bound without limiting how many user groups run
// (model: tools/async_sim: the bound is set by the item BIAS, whose curvature is 1).
__global__ void __launch_bounds__(256)
dsgd_sample_runs_kernel(const int *__restrict__ indptr, const cu2b_rating *__restrict__ coo,
                        const int *__restrict__ active_users, const int *__restrict__ user_ids,
                        int n_active, uint32_t seed, int iter0, int nb, int pitch,
                        const int *__restrict__ item_block_ptr, int world, DsgdDraw *__restrict__ draws,
                        int *__restrict__ row_off, const float *__restrict__ keep_row,
                        const float *__restrict__ keep_bias) {
    __shared__ int sh_ptr[kMaxWorld + 1];
    if (threadIdx.x <= world) sh_ptr[threadIdx.x] = item_block_ptr[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; a < n_active; a += warps) {
        const int u = __ldg(&active_users[a]);
        const uint32_t uid = user_ids ? (uint32_t)__ldg(&user_ids[u]) : (uint32_t)u;
        const int lo = __ldg(&indptr[u]), n = __ldg(&indptr[u + 1]) - lo;
        // The lane's draws (iterations lane, lane + 32, ...) are evaluated once and kept in registers: Philox,
        // the gather of (item, rating), the item block and the thinning flags. kDrawsPerLane * 32 bounds the round.
        DsgdDraw mine[kDrawsPerLane];
        int blk_of[kDrawsPerLane];
        int cnt[kMaxWorld];
#pragma unroll
        for (int b = 0; b < kMaxWorld; ++b) cnt[b] = 0;
#pragma unroll
        for (int q = 0; q < kDrawsPerLane; ++q) {
            const int t = q * 32 + lane;
            mine[q].item = 0; mine[q].rating = 0.f;
            blk_of[q] = -1;
            if (t < nb) {
                const uint2 r = philox4x32_10_xy(uid, (uint32_t)(iter0 + t), 0u, PHILOX_TAG, seed, PHILOX_KEY1);
                const int j = lo + (int)__umulhi(r.x, (uint32_t)n);
                DsgdDraw d;
                d.item = __ldg(&coo[j].item);
                d.rating = __ldg(&coo[j].rating);
                const int blk = item_block_of(d.item, sh_ptr, world);
                if (keep_row || keep_bias) {
                    const float x = (float)(r.y >> 8) * (1.0f / 16777216.0f);
                    int flags = 0;
                    if (keep_row && x >= __ldg(&keep_row[d.item])) flags |= kDrawRowFrozen;
                    if (keep_bias && x >= __ldg(&keep_bias[d.item])) flags |= kDrawBiasFrozen;
                    d.item |= flags;
                }
                mine[q] = d;
                blk_of[q] = blk;
#pragma unroll
                for (int b = 0; b < kMaxWorld; ++b) cnt[b] += (blk == b);
            }
        }
        int off[kMaxWorld + 1];
        off[0] = 0;
#pragma unroll
        for (int b = 0; b < kMaxWorld; ++b) {
            int c = cnt[b];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            off[b + 1] = off[b] + c;
        }
        if (lane <= world) {
            int v = 0;
#pragma unroll
            for (int b = 0; b <= kMaxWorld; ++b) v = (lane == b) ? off[b] : v;
            row_off[(size_t)a * (world + 1) + lane] = v;
        }
        // place the draws, iteration order kept inside a block (ballot ranking)
        DsgdDraw *row = draws + (size_t)a * pitch;
#pragma unroll
        for (int q = 0; q < kDrawsPerLane; ++q) {
            if (q * 32 >= nb) break;  // warp-uniform
#pragma unroll
            for (int b = 0; b < kMaxWorld; ++b) {
                if (b < world) {
                    const unsigned m = __ballot_sync(0xffffffffu, blk_of[q] == b);
                    if (blk_of[q] == b) row[off[b] + __popc(m & ((1u << lane) - 1u))] = mine[q];
                    off[b] += __popc(m);
                }
            }
        }
    }
}

// Update kernel of one DSGD sub-epoch: a lane group owns one USER, keeps that user's P row and
// bias in registers, applies the user's run of draws for the resident item block strictly in
// order (exact sequential semantics on the user side, P traffic once per run instead of once per
// update), and adds the item-side steps with 128-bit L2 atomics (other users update the same
// item rows concurrently). Same arithmetic as sgd_update_slots.
// What ties a sub-epoch kernel to its neighbours on the ring when wait, update and hand-off run as ONE launch
// (mf_sgd_user_runs<..., LINKED = true>): the prologue polls this rank's own arrive word for the item block
// (written by the upstream rank through NVLink), the epilogue -- executed by whichever CTA leaves last -- copies
// the block's rows and biases into the downstream rank's arrays with peer stores and publishes its arrive word.
struct SubEpochLink {
    const int *wait_flag;       // nullptr: the block is already here (first sub-epoch of the first round)
    int wait_need;
    const float4 *src_q;        // nullptr: nothing to hand off
    float4 *dst_q;
    long long n_vec;
    const float *src_ib;
    float *dst_ib;
    int n_items;
    int *dst_flag;
    int value;
    unsigned int *done;         // CTAs that have left; the last one resets it
    int *error_flag;
    long long timeout_cycles;
    int site;
};

struct UserRunParams {
    const DsgdDraw *draws;
    const int *row_off;   // [n_active][world + 1]
    const int *active_users;
    int n_active, pitch, world, block;
    int claim, n_ranges;  // tiles per claim; claim counters (one per contiguous range of users, kClaimStride apart)
    unsigned long long *tile_counter;  // dynamic tile claims (a tile = the G users of one warp); zeroed before launch
    float *P, *Q, *user_bias, *item_bias;
    int kp, ibs;
    float mu;
    const float *lr;
    float P_reg, Q_reg, ub_reg, ib_reg;
    int is_train;
    SubEpochLink link;          // used by the LINKED instantiations only
};

// THIN: the draws may carry kDrawRowFrozen / kDrawBiasFrozen (item-step thinning); the default
// instantiation does not look at the bits. LINKED: fused wait + sub-epoch + hand-off (SubEpochLink).
template <int L, int V, bool THIN = false, bool LINKED = false>
__global__ void __launch_bounds__(256, (V == 1 ? 5 : 4))
mf_sgd_user_runs(const UserRunParams p) {
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31, g = lane / L, l = lane % L;
    if (LINKED) {
        if (p.link.wait_flag && threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_acquire_sys(p.link.wait_flag) < p.link.wait_need) {
                __nanosleep(100);
                if (clock64() - t0 > p.link.timeout_cycles) {
                    atomicCAS(p.link.error_flag, 0, p.link.site * 1000000 + p.link.wait_need * 100 + (ld_acquire_sys(p.link.wait_flag) % 100));
                    break;
                }
            }
        }
        __syncthreads();
    }
    const int vecs = p.kp >> 2;
    const float lr = __ldg(p.lr);
    const StepCoef sc = step_coef(lr, p.P_reg, p.Q_reg, p.ub_reg, p.ib_reg);
    float4 *const Pv = reinterpret_cast<float4 *>(p.P);
    float4 *const Qv = reinterpret_cast<float4 *>(p.Q);
    // Dynamic tile claims (the SMs do not run at one speed; a static split leaves the fast ones idle at the end).
    // The users are cut into n_ranges contiguous ranges with one claim counter each, 512 bytes apart; a warp claims
    // `claim` tiles at a time from its home range (the next claim is issued one chunk ahead) and, when a range is
    // used up, moves on to the next one. Why not one counter: a sub-epoch of a multi-GPU rank is short (a few tiles
    // per warp), so the claims should be small to balance the warps, but small claims on ONE address cost: atomics
    // on one L2 line retire at ~0.5 G/s (profiles/r2_l2_roofline.md), and at 8 GPUs single-tile claims on one
    // counter ran at 16.4 G updates/s against 17.9 with four tiles per claim (profiles/r2c_bench_n8_*.json).
    // Several counters take the pressure off: 16 ranges x 2 tiles per claim is the measured best (18.3; r2d_*).
    // n_ranges = 1 is the single-counter scheme.
    const int n_tiles = (p.n_active + G - 1) / G;
    const int n_ranges = max(1, p.n_ranges), claim_sz = max(1, p.claim);
    const int range_len = (n_tiles + n_ranges - 1) / n_ranges;
    int range = (int)((((unsigned)blockIdx.x * blockDim.x + threadIdx.x) >> 5) % (unsigned)n_ranges), ranges_tried = 0;
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(p.tile_counter + (size_t)range * kClaimStride, (unsigned long long)claim_sz);
    for (;;) {
        const long long t = (long long)__shfl_sync(0xffffffffu, claim, 0);
        const int base = range * range_len;
        const int lim = min(range_len, n_tiles - base);  // tiles in this range (<= 0 for trailing empty ranges)
        if (t >= lim) {                                  // range used up: try the next one, give up after a full circle
            if (++ranges_tried >= n_ranges) break;
            range = range + 1 == n_ranges ? 0 : range + 1;
            if (lane == 0) claim = atomicAdd(p.tile_counter + (size_t)range * kClaimStride, (unsigned long long)claim_sz);
            continue;
        }
        if (lane == 0) claim = atomicAdd(p.tile_counter + (size_t)range * kClaimStride, (unsigned long long)claim_sz);
        const int t0 = base + (int)t, t1 = base + (int)min((long long)lim, t + claim_sz);
    for (int tile = (int)t0; tile < t1; ++tile) {
        const int a0 = tile * G;
        const int a = a0 + g;
        int j = 0, end = 0, u = 0;
        if (a < p.n_active) {
            j = __ldg(p.row_off + (size_t)a * (p.world + 1) + p.block);
            end = __ldg(p.row_off + (size_t)a * (p.world + 1) + p.block + 1);
            u = __ldg(p.active_users + a);
        }
        if (!__any_sync(0xffffffffu, j < end)) continue;
        const bool mine = j < end;
        float4 pv[V];
        const size_t po = (size_t)u * vecs + l;
#pragma unroll
        for (int v = 0; v < V; ++v)
            pv[v] = (mine && v * L + l < vecs) ? __ldcg(Pv + po + v * L) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ub = mine ? __ldcg(p.user_bias + u) : 0.f;
        const DsgdDraw *row = p.draws + (size_t)a * p.pitch;
        // The draws are fetched two updates ahead: a draw is not part of the item row's
        // read -> atomic-add window, and with two updates of slack its L2 round trip never
        // delays the row request of the update that consumes it.
        DsgdDraw nxt, nxt2;
        nxt.item = 0; nxt.rating = 0.f;
        nxt2 = nxt;
        if (mine) nxt = ld_draw_pinned(row + j);
        if (j + 1 < end) nxt2 = ld_draw_pinned(row + j + 1);
        while (__any_sync(0xffffffffu, j < end)) {
            const bool ok = j < end;
            const DsgdDraw d = nxt;
            nxt = nxt2;
            if (j + 2 < end) nxt2 = ld_draw_pinned(row + j + 2);
            if (THIN)
                user_side_update<L, V, true>(pv, ub, d.item & kDrawItemMask, d.rating, ok, l, vecs, Qv, p.item_bias, p.ibs, p.mu, lr, sc,
                                             p.is_train ? ((d.item < 0 ? 0 : 1) | ((d.item & kDrawBiasFrozen) ? 0 : 2)) : 0);
            else
                user_side_update<L, V>(pv, ub, d.item, d.rating, ok, l, vecs, Qv, p.item_bias, p.ibs, p.mu, lr, sc, p.is_train);
            ++j;
        }
        if (mine) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v * L + l < vecs) __stcg(Pv + po + v * L, pv[v]);
            if (l == 0) __stcg(p.user_bias + u, ub);
        }
    }
    }
    if (LINKED) {
        if (p.link.src_q == nullptr) return;
        // Hand-off by the CTAs that leave last. Every thread's atomic adds are performed device-wide before its
        // CTA counts itself out (fence, barrier, counter). A CTA whose ticket is not among the last kCopiers simply
        // exits (its SM slots go to the next round's sampler, which runs on its own stream); the last kCopiers
        // wait until the counter reaches the grid size -- the block is final then -- and copy one share each into
        // the downstream rank's arrays with peer stores; the last copier publishes the arrive flag. The handful
        // of waiting CTAs finish their own work within the same tail of the kernel, so the wait is short; it is
        // bounded by the same timeout as the other device-side waits anyway.
        constexpr unsigned kCopiers = 32;
        __shared__ unsigned ticket_sh;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) ticket_sh = atomicAdd(p.link.done, 1u);
        __syncthreads();
        const unsigned copiers = min(kCopiers, gridDim.x), first = gridDim.x - copiers;
        if (ticket_sh < first) return;
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu((const int *)p.link.done) < (int)gridDim.x) {
                __nanosleep(32);
                if (clock64() - t0 > p.link.timeout_cycles) {
                    atomicCAS(p.link.error_flag, 0, (20 + p.link.site) * 1000000);
                    break;
                }
            }
        }
        __syncthreads();
        const long long share = ticket_sh - first, stride = (long long)copiers * blockDim.x;
        for (long long i = share * blockDim.x + threadIdx.x; i < p.link.n_vec; i += 2 * stride) {
            const bool two = i + stride < p.link.n_vec;
            const float4 v0 = __ldcg(p.link.src_q + i);
            float4 v1 = v0;
            if (two) v1 = __ldcg(p.link.src_q + i + stride);
            p.link.dst_q[i] = v0;
            if (two) p.link.dst_q[i + stride] = v1;
        }
        for (long long i = share * blockDim.x + threadIdx.x; i < p.link.n_items; i += stride)
            p.link.dst_ib[i * p.ibs] = __ldcg(p.link.src_ib + i * p.ibs);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (atomicAdd(p.link.done + 1, 1u) == copiers - 1) {  // last copier: publish, reset both counters
                p.link.done[0] = 0u;
                p.link.done[1] = 0u;
                __threadfence_system();
                st_release_sys(p.link.dst_flag, p.link.value);
            }
        }
    }
}

// Hands item block rows [row0, row1) (Q rows + item_bias) to a peer: plain 128-bit stores into
// the peer's arrays (NVLink P2P), then the last CTA publishes flag = value with system scope.
// dst_flag == nullptr => plain local copy without signalling.
__global__ void __launch_bounds__(256)
dsgd_send_block_kernel(const float4 *__restrict__ src_q, float4 *__restrict__ dst_q, long long n_vec,
                       const float *__restrict__ src_ib, float *__restrict__ dst_ib, int n_items, int ibs,
                       int *ticket, int *dst_flag, int value) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride)
        dst_q[i] = __ldcg(src_q + i);
    // the biases live one per line (ItemBiasLayout): only the 4-byte values travel
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride)
        dst_ib[i * ibs] = __ldcg(src_ib + i * ibs);
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
                         int check_no, int *error_flag, long long timeout_cycles) {
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
            if (clock64() - t0 > timeout_cycles) { atomicCAS(error_flag, 0, 5000000 + check_no * 100 + threadIdx.x); break; }
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
        flag_non_finite(st, train_rmse, test_rmse, n_train_global, n_test_global, iteration);
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
