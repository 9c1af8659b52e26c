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
    int kp;
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
            const size_t qo = (size_t)d.item * vecs + l;
            float4 qv[V];
#pragma unroll
            for (int v = 0; v < V; ++v)
                qv[v] = (mine && v * L + l < vecs) ? __ldcg(Qv + qo + v * L) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ib = mine ? __ldcg(p.item_bias + d.item) : 0.f;
            float acc = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                acc = __fmaf_rn(pv[v].x, qv[v].x, acc);
                acc = __fmaf_rn(pv[v].y, qv[v].y, acc);
                acc = __fmaf_rn(pv[v].z, qv[v].z, acc);
                acc = __fmaf_rn(pv[v].w, qv[v].w, acc);
            }
            const float dot = group_sum<L>(acc);
            const float pred = __fadd_rn(__fadd_rn(__fadd_rn(p.mu, ub), ib), dot);
            const float err = __fsub_rn(d.rating, pred);
            if (mine) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4 x = pv[v], y = qv[v];
                    float4 nq;
                    nq.x = __fmul_rn(lr, __fsub_rn(__fmul_rn(err, x.x), __fmul_rn(p.Q_reg, y.x)));
                    nq.y = __fmul_rn(lr, __fsub_rn(__fmul_rn(err, x.y), __fmul_rn(p.Q_reg, y.y)));
                    nq.z = __fmul_rn(lr, __fsub_rn(__fmul_rn(err, x.z), __fmul_rn(p.Q_reg, y.z)));
                    nq.w = __fmul_rn(lr, __fsub_rn(__fmul_rn(err, x.w), __fmul_rn(p.Q_reg, y.w)));
                    pv[v].x = __fadd_rn(x.x, __fmul_rn(lr, __fsub_rn(__fmul_rn(err, y.x), __fmul_rn(p.P_reg, x.x))));
                    pv[v].y = __fadd_rn(x.y, __fmul_rn(lr, __fsub_rn(__fmul_rn(err, y.y), __fmul_rn(p.P_reg, x.y))));
                    pv[v].z = __fadd_rn(x.z, __fmul_rn(lr, __fsub_rn(__fmul_rn(err, y.z), __fmul_rn(p.P_reg, x.z))));
                    pv[v].w = __fadd_rn(x.w, __fmul_rn(lr, __fsub_rn(__fmul_rn(err, y.w), __fmul_rn(p.P_reg, x.w))));
                    if (p.is_train && v * L + l < vecs) red_add_v4(Qv + qo + v * L, nq);
                }
                if (p.is_train && l == 0)
                    red_add_f32(p.item_bias + d.item, __fmul_rn(lr, __fsub_rn(err, __fmul_rn(p.ib_reg, ib))));
                ub = __fadd_rn(ub, __fmul_rn(lr, __fsub_rn(err, __fmul_rn(p.ub_reg, ub))));
            }
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

}  // namespace cu2b
#endif  // CU2B_TILED_KERNELS_CUH_
