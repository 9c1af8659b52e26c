// Iteration-tiled Hogwild schedule for one GPU: `round` consecutive reference iterations of a
// user are applied back to back by one lane group that keeps the user's P row and bias in
// registers (P traffic once per round instead of once per update; item rows via 128-bit L2
// atomic adds as in mf_sgd_hogwild). Same draws as the iteration-synchronous kernel (same
// Philox stream), same arithmetic; only the interleaving between users changes. round = 1 is
// the reference's iteration-synchronous order and uses mf_sgd_hogwild instead.
//
// The draws of a round are stored user-major ([active user][pitch] of (item, rating)); a CTA
// consumes them in tiles of kConsumerWarps * (32/L) users -- one user per lane group -- that the
// producer warp streams into a shared-memory ring with 1-D TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), exactly like the rating stream
// of mf_sgd_hogwild.
#ifndef CU2B_TILED_KERNELS_CUH_
#define CU2B_TILED_KERNELS_CUH_

#include "dsgd_kernels.cuh"

namespace cu2b {

constexpr int kTileStages = 4;
constexpr int kTileDrawsMax = 1024;  // draws per stage (8 KB): tile users * pitch must fit

struct __align__(16) TileSmem {
    DsgdDraw stage[kTileStages][kTileDrawsMax];
    unsigned long long full[kTileStages];
    unsigned long long empty[kTileStages];
    int tile_id[kTileStages];  // -1 => no more tiles
};

struct UserTileParams {
    const DsgdDraw *draws;      // [n_active][pitch]
    const int *active_users;
    int n_active, pitch, nb;    // nb = iterations in this round (<= pitch)
    int n_tiles;
    unsigned long long *tile_counter;
    float *P, *Q, *user_bias, *item_bias;
    int kp, ibs;
    float mu;
    const float *lr;
    float P_reg, Q_reg, ub_reg, ib_reg;
    int is_train;
};

// Sampler of one round for the tiled schedule: thread i draws rating (user a = i / nb, iteration
// t = i % nb) -- same Philox counter as sample_per_user_kernel -- and stores it user-major.
__global__ void __launch_bounds__(256)
sample_user_major_kernel(const int *__restrict__ indptr, const cu2b_rating *__restrict__ coo,
                         const int *__restrict__ active_users, const int *__restrict__ user_ids, int n_active,
                         uint32_t seed, int iter0, int nb, int pitch, DsgdDraw *__restrict__ draws) {
    const long long n = (long long)n_active * nb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(i / nb), t = (int)(i - (long long)a * nb);
        const int u = __ldg(&active_users[a]);
        const uint32_t uid = user_ids ? (uint32_t)__ldg(&user_ids[u]) : (uint32_t)u;
        const int lo = __ldg(&indptr[u]), hi = __ldg(&indptr[u + 1]);
        const uint32_t r = philox4x32_10_x(uid, (uint32_t)(iter0 + t), 0u, PHILOX_TAG, seed, PHILOX_KEY1);
        const int j = lo + (int)__umulhi(r, (uint32_t)(hi - lo));
        DsgdDraw d;
        d.item = __ldg(&coo[j].item);
        d.rating = __ldg(&coo[j].rating);
        draws[(size_t)a * pitch + t] = d;
    }
}

template <int L, int V>
__global__ void __launch_bounds__(kThreads)
mf_sgd_user_tiles(const UserTileParams p) {
    __shared__ TileSmem sm;
    constexpr int G = 32 / L;
    constexpr int TU = kConsumerWarps * G;  // users per tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTileStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == kConsumerWarps) {  // producer warp: one lane streams the tiles
        if (lane == 0) {
            for (int it = 0;; ++it) {
                const int s = it % kTileStages;
                if (it >= kTileStages) mbar_wait(&sm.empty[s], ((it / kTileStages) - 1) & 1);
                const unsigned long long t = atomicAdd(p.tile_counter, 1ULL);
                if (t >= (unsigned long long)p.n_tiles) {
                    sm.tile_id[s] = -1;
                    mbar_arrive(&sm.full[s]);
                    break;
                }
                sm.tile_id[s] = (int)t;
                const int users = min(TU, p.n_active - (int)t * TU);
                const uint32_t bytes = (uint32_t)(users * p.pitch * (int)sizeof(DsgdDraw));  // pitch % 2 == 0 => 16 B multiple
                mbar_arrive_expect_tx(&sm.full[s], bytes);
                tma_load_1d(&sm.stage[s][0], p.draws + (size_t)t * TU * p.pitch, bytes, &sm.full[s]);
            }
        }
        return;
    }
    const int g = lane / L, l = lane % L;
    const int vecs = p.kp >> 2;
    const float lr = __ldg(p.lr);
    const StepCoef sc = step_coef(lr, p.P_reg, p.Q_reg, p.ub_reg, p.ib_reg);
    float4 *const Pv = reinterpret_cast<float4 *>(p.P);
    float4 *const Qv = reinterpret_cast<float4 *>(p.Q);
    for (int it = 0;; ++it) {
        const int s = it % kTileStages;
        mbar_wait(&sm.full[s], (it / kTileStages) & 1);
        const int tile = sm.tile_id[s];
        if (tile < 0) break;
        const int slot = warp * G + g;              // this group's user inside the tile
        const int a = tile * TU + slot;
        const bool mine = a < p.n_active;
        const int u = mine ? __ldg(p.active_users + a) : 0;
        const size_t po = (size_t)u * vecs + l;
        float4 pv[V];
#pragma unroll
        for (int v = 0; v < V; ++v)
            pv[v] = (mine && v * L + l < vecs) ? __ldcg(Pv + po + v * L) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ub = mine ? __ldcg(p.user_bias + u) : 0.f;
        const DsgdDraw *row = &sm.stage[s][slot * p.pitch];
        for (int j = 0; j < p.nb; ++j) {  // every lane runs nb steps (the shuffles are warp-wide)
            DsgdDraw d;
            d.item = 0; d.rating = 0.f;
            if (mine) d = row[j];
            user_side_update<L, V>(pv, ub, d.item, d.rating, mine, l, vecs, Qv, p.item_bias, p.ibs, p.mu, lr, sc, p.is_train);
        }
        if (mine) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v * L + l < vecs) __stcg(Pv + po + v * L, pv[v]);
            if (l == 0) __stcg(p.user_bias + u, ub);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// mf_sgd_user_rounds: the same schedule with the sampler fused into the update kernel.
//
// The draws of a round do not depend on the model (counter-based Philox keyed by (seed), counter
// (user, iteration)), so the warp that owns a user can draw that user's `nb` ratings itself: the
// 32 lanes evaluate the G * nb Philox counters of the warp's G users in parallel, gather
// (item, rating) from the user's CSR row, and park the draws in a warp-private strip of shared
// memory; the update loop then reads them back one per step (LDS broadcast). No sampler kernel,
// no draw buffer in HBM, no producer warp: per round the kernel reads 2 indptr words + nb gathered
// ratings per user and moves the P row once. Bit-identical draws to sample_user_major_kernel /
// sample_per_user_kernel, hence the same updates as mf_sgd_user_tiles.
// Work distribution: a warp claims kRoundClaim consecutive tiles (a tile = the warp's G users) at
// a time from a global counter. The work per tile is uniform, the SMs are not (ncu: with a static
// split some SMs idle for 20 % of the launch while others still run), so the claim is dynamic;
// the next claim is issued one chunk ahead so that its round trip never stalls the warp.
// ---------------------------------------------------------------------------------------------
constexpr int kRoundWarps = 8;
constexpr int kRoundDrawsPerWarp = 128;   // G * nb <= 128  (nb <= 128 at k = 128, 64 at k = 64, ...)
constexpr int kRoundStrip = kRoundDrawsPerWarp + 32;  // + one pad draw per user row (bank spread)
constexpr int kRoundClaim = 4;            // tiles per claim

struct UserRoundParams {
    const int *indptr;          // train CSR row pointers
    const cu2b_rating *coo;     // train triplets in CSR order
    const int *active_users;
    const int *user_ids;        // DSGD strips: original user id (sampler key); else null
    int n_active;
    unsigned long long *tile_counter;  // zeroed before the launch
    uint32_t seed;
    int iter0, nb;              // absolute first iteration of the round, iterations in it
    float *P, *Q, *user_bias, *item_bias;
    int kp, ibs;
    float mu;
    const float *lr;
    float P_reg, Q_reg, ub_reg, ib_reg;
    int is_train;
};

// PF = 1: the item row of update j+1 is requested before update j is computed (one row of
// look-ahead per lane group; a repeated item is re-read after the atomic add so that a user's
// updates keep their exact sequential semantics). MINB = resident CTAs per SM the register
// allocation is bounded for.
template <int L, int V, int PF, int MINB>
__global__ void __launch_bounds__(kRoundWarps * 32, MINB)
mf_sgd_user_rounds(const UserRoundParams p) {
    __shared__ DsgdDraw sh[kRoundWarps][kRoundStrip];
    constexpr int G = 32 / L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / L, l = lane % L;
    const int vecs = p.kp >> 2;
    const float lr = __ldg(p.lr);
    const StepCoef sc = step_coef(lr, p.P_reg, p.Q_reg, p.ub_reg, p.ib_reg);
    float4 *const Pv = reinterpret_cast<float4 *>(p.P);
    float4 *const Qv = reinterpret_cast<float4 *>(p.Q);
    DsgdDraw *const strip = sh[warp];
    const int pitch = p.nb + 1;
    const int total = G * p.nb;
    const int n_tiles = (p.n_active + G - 1) / G;
    unsigned long long claim = 0;
    if (lane == 0) claim = atomicAdd(p.tile_counter, (unsigned long long)kRoundClaim);
    for (;;) {
        const long long t0 = (long long)__shfl_sync(0xffffffffu, claim, 0);
        if (t0 >= n_tiles) break;
        if (lane == 0) claim = atomicAdd(p.tile_counter, (unsigned long long)kRoundClaim);  // consumed next time round
        const int t1 = (int)min((long long)n_tiles, t0 + kRoundClaim);
    for (int tile = (int)t0; tile < t1; ++tile) {
        const int a0 = tile * G;
        const int a = a0 + g;
        const bool mine = a < p.n_active;
        const int u = mine ? __ldg(p.active_users + a) : 0;
        // the P row does not depend on the draws: its loads fly while the lanes sample
        const size_t po = (size_t)u * vecs + l;
        float4 pv[V];
#pragma unroll
        for (int v = 0; v < V; ++v)
            pv[v] = (mine && v * L + l < vecs) ? __ldcg(Pv + po + v * L) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ub = mine ? __ldcg(p.user_bias + u) : 0.f;
        __syncwarp();  // the previous users' draws have been consumed
        for (int i = lane; i < total; i += 32) {
            const int slot = i / p.nb, t = i - slot * p.nb;
            DsgdDraw d;
            d.item = 0; d.rating = 0.f;
            if (a0 + slot < p.n_active) {
                const int uu = __ldg(p.active_users + a0 + slot);
                const uint32_t uid = p.user_ids ? (uint32_t)__ldg(p.user_ids + uu) : (uint32_t)uu;
                const int lo = __ldg(p.indptr + uu), hi = __ldg(p.indptr + uu + 1);
                const uint32_t r = philox4x32_10_x(uid, (uint32_t)(p.iter0 + t), 0u, PHILOX_TAG, p.seed, PHILOX_KEY1);
                const int j = lo + (int)__umulhi(r, (uint32_t)(hi - lo));
                d.item = __ldg(&p.coo[j].item);
                d.rating = __ldg(&p.coo[j].rating);
            }
            strip[slot * pitch + t] = d;
        }
        __syncwarp();
        const DsgdDraw *row = strip + g * pitch;
        if (PF) {
            DsgdDraw d = row[0];
            ItemSide<V> cur;
            item_side_load<L, V>(cur, d.item, mine, l, vecs, Qv, p.item_bias, p.ibs);
            for (int j = 0; j < p.nb; ++j) {  // every lane runs nb steps (the shuffles are warp-wide)
                const bool more = j + 1 < p.nb;
                const DsgdDraw dn = row[more ? j + 1 : j];
                const bool repeat = dn.item == d.item;
                ItemSide<V> nxt;
                item_side_load<L, V>(nxt, dn.item, mine && more && !repeat, l, vecs, Qv, p.item_bias, p.ibs);
                user_side_apply<L, V>(pv, ub, cur, d.item, d.rating, mine, l, vecs, Qv, p.item_bias, p.ibs, p.mu, lr, sc, p.is_train);
                if (repeat) item_side_load<L, V>(nxt, dn.item, mine && more, l, vecs, Qv, p.item_bias, p.ibs);
                cur = nxt;
                d = dn;
            }
        } else {
            for (int j = 0; j < p.nb; ++j) {
                const DsgdDraw d = row[j];
                user_side_update<L, V>(pv, ub, d.item, d.rating, mine, l, vecs, Qv, p.item_bias, p.ibs, p.mu, lr, sc, p.is_train);
            }
        }
        if (mine) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v * L + l < vecs) __stcg(Pv + po + v * L, pv[v]);
            if (l == 0) __stcg(p.user_bias + u, ub);
        }
    }
    }
}

}  // namespace cu2b
#endif  // CU2B_TILED_KERNELS_CUH_
