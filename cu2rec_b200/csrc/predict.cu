// Batched predict + top-k (BASELINE config 5): score every user against the
// whole catalogue and keep each user's best unrated items. Replaces the per-user CPU loops of
// the reference (predict_ratings predict.cu:17-29, get_recommendations predict.cu:49-63).
//
// Scoring all 480 189 x 17 770 pairs at k = 128 is a genuine dense GEMM (2.2 TFLOP), so this is
// the one place the 5th-generation tensor cores are used:
//   pass 1  predict_candidates_kernel (tcgen05 / TMEM / TMA, one CTA per SM, persistent):
//           D[128 users x 128 items] = P_tile . Q_tile^T in TF32 (kind::tf32 reads the fp32 rows
//           as they are), fp32 accumulators in TMEM, double buffered. The epilogue warps read the
//           accumulators with tcgen05.ld, add the item bias, drop items the user has already
//           rated (one 128-bit word of a per-user bitmap per tile) and keep a sorted list of the
//           KC best candidates per user in registers. Nothing of the 8.5 G score matrix is
//           ever written.
//   pass 2  predict_rescore_kernel: exact fp32 scores of the KC candidates in the reference's
//           op order (serial dot, predict.cu:22-26), final ordering (score desc, item asc), top-k.
// Each of the two epilogue groups keeps KC >= k candidates from its half of the item tiles, so
// TF32 rounding can only matter if a true top-k item is not among the KC best TF32 scores of its
// own half (the lists hold 24-32 candidates for a top-10).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 and
// 6-9 two epilogue groups, one per accumulator buffer (TMEM lane quadrant = warp % 4).
//
// Any n_factors: the factor rows are zero-padded on the device to whole 128-byte swizzle rows (k = 50, the
// reference's default, config.h:27, runs as 64). Up to 128 factors the user tile stays resident in shared memory
// for the whole sweep over the catalogue; beyond that (k = 300 of experiments/cu2rec.sh:10 runs as 320) user and
// item tiles stream through a three-stage ring in chunks of 64 factors and the MMAs of all chunks accumulate into
// the same TMEM tile. Any top-k: a pass yields the 16 best items of a user exactly; further passes with those
// items added to the exclusion bitmap yield the next 16, and a final kernel orders the union.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cu2b_internal.h"

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return cu2b_fail(CU2B_ERR_CUDA, "Cuda Error: %s (%s:%d)", cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                             \
    } while (0)

namespace {

constexpr int BM = 128;        // users per tile (MMA M, = TMEM lanes)
constexpr int BN = 128;        // items per tile (MMA N, = TMEM columns per accumulator)
constexpr int SLAB_K = 32;     // floats per 128-byte swizzle row
constexpr int SLAB_BYTES = BM * SLAB_K * 4;  // 16 KB: 128 rows x 128 B
constexpr int kThreadsPredict = 320;  // TMA warp, MMA warp, 2 x 4 epilogue warps

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void bar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(s32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(s32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(s32(bar))
        : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 UMMA): start address and
// stride between 8-row groups (1024 B) in 16-byte units, version 1, layout type SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte off.  bits [16,30) (unused for SW128 K-major)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;       // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// v[c] for a per-lane column index: a five-level select tree (31 selects, depth 5) instead of a 32-step chain.
__device__ __forceinline__ float pick32(const float (&v)[32], int c) {
    float a[16], b[8], d[4];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (c & 1) ? v[2 * j + 1] : v[2 * j];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (c & 2) ? a[2 * j + 1] : a[2 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = (c & 4) ? b[2 * j + 1] : b[2 * j];
    const float e0 = (c & 8) ? d[1] : d[0], e1 = (c & 8) ? d[3] : d[2];
    return (c & 16) ? e1 : e0;
}

// Pending candidates per epilogue thread (shared memory behind PredictSmemCtl, [slot][thread]): columns that beat a
// user's threshold are parked here and merged into the sorted register list in batches, when some lane's ring is full.
// A warp executes the merge for all its lanes at once, so batching raises its lane efficiency from ~17 % (one hit per
// lane and pass) to ~45 %; the merge sees the candidates in item order and re-checks the threshold, so the list is the
// one immediate insertion builds.
constexpr int kRing = 8;
constexpr int kEpilogueThreads = 256;
constexpr size_t kRingBytes = (size_t)2 * kRing * kEpilogueThreads * sizeof(float);

// Accumulator tiles in TMEM: 4 x 128 columns = all 512. Each epilogue group owns two of them alternately, so the MMAs
// of a group's next tile run while the group is still scanning the current one (with one tile per group the chain
// epilogue -> acc_empty -> MMA -> acc_full -> epilogue left the epilogue warps waiting 47 % of the time).
constexpr int kAccBufs = 4;

struct PredictSmemCtl {
    uint64_t a_full, a_empty;
    uint64_t b_full[2], b_empty[2];
    uint64_t acc_full[kAccBufs], acc_empty[kAccBufs];
    uint64_t st_full[3], st_empty[3];  // STREAM: ring of {user chunk, item chunk} stages
    uint32_t tmem_base;
    uint32_t pad;
    float ib[2][2][BN];  // [epilogue group][tile parity][column]
};

constexpr int CHUNK_SLABS = 2;   // STREAM: factors per stage = 64 (2 x 16 KB of users + 2 x 16 KB of items)
constexpr int STREAM_STAGES = 3;

struct PredictParams {
    int users, items, kslabs;
    int n_user_tiles, n_item_tiles;
    const float *item_bias;
    const uint32_t *bitmap;  // [users][mask_pitch] words, bit i set => item i already rated; nullptr => none
    int mask_pitch;          // words per user (multiple of 4)
    int32_t *cand_items;     // [users][2][KC]: one sorted list per epilogue group
    float *cand_scores;      // [users][2][KC] (TF32 scores incl. item bias; diagnostics)
};

// -DCU2B_PREDICT_PROFILE: cycles one epilogue warp of each group spends per phase (printed by CTA 0; tools/r2 README).
#ifdef CU2B_PREDICT_PROFILE
#define PROF_DECL long long prof[6] = {0, 0, 0, 0, 0, 0}; long long prof_t = clock64();
#define PROF(i) { const long long now_ = clock64(); prof[i] += now_ - prof_t; prof_t = now_; }
#else
#define PROF_DECL
#define PROF(i)
#endif

template <int KC, bool STREAM>
__global__ void __launch_bounds__(kThreadsPredict, 1)
predict_candidates_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_q,
                          const PredictParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SW128 needs 1024-byte alignment
    uint8_t *smem_a = base;                                        // kslabs slabs
    uint8_t *smem_b = base + (size_t)p.kslabs * SLAB_BYTES;        // 2 stages x kslabs slabs
    // STREAM: STREAM_STAGES stages of {CHUNK_SLABS user slabs, CHUNK_SLABS item slabs} from `base`
    constexpr size_t kStageBytes = (size_t)2 * CHUNK_SLABS * SLAB_BYTES;
    PredictSmemCtl *ctl = STREAM ? (PredictSmemCtl *)(base + STREAM_STAGES * kStageBytes)
                                 : (PredictSmemCtl *)(smem_b + (size_t)2 * p.kslabs * SLAB_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t slab_tx = (uint32_t)p.kslabs * SLAB_BYTES;

    if (threadIdx.x == 0) {
        bar_init(&ctl->a_full, 1);
        bar_init(&ctl->a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            bar_init(&ctl->b_full[s], 1);
            bar_init(&ctl->b_empty[s], 1);
        }
        for (int s = 0; s < kAccBufs; ++s) {
            bar_init(&ctl->acc_full[s], 1);
            bar_init(&ctl->acc_empty[s], 4);
        }
        for (int s = 0; s < STREAM_STAGES; ++s) {
            bar_init(&ctl->st_full[s], 1);
            bar_init(&ctl->st_empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: kAccBufs accumulators x 128 columns
        const uint32_t ncols = kAccBufs * BN;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&ctl->tmem_base)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = ctl->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && STREAM) {
            const int n_chunks = p.kslabs / CHUNK_SLABS;
            int st = 0;  // global stage counter
            for (int ut = blockIdx.x; ut < p.n_user_tiles; ut += gridDim.x)
                for (int j = 0; j < p.n_item_tiles; ++j)
                    for (int c = 0; c < n_chunks; ++c, ++st) {
                        const int s = st % STREAM_STAGES;
                        if (st >= STREAM_STAGES) bar_wait(&ctl->st_empty[s], ((st / STREAM_STAGES) - 1) & 1);
                        uint8_t *stage = base + (size_t)s * kStageBytes;
                        bar_expect(&ctl->st_full[s], (uint32_t)kStageBytes);
                        for (int sl = 0; sl < CHUNK_SLABS; ++sl) {
                            tma_load_2d(stage + (size_t)sl * SLAB_BYTES, &map_p, (c * CHUNK_SLABS + sl) * SLAB_K, ut * BM, &ctl->st_full[s]);
                            tma_load_2d(stage + (size_t)(CHUNK_SLABS + sl) * SLAB_BYTES, &map_q, (c * CHUNK_SLABS + sl) * SLAB_K, j * BN,
                                        &ctl->st_full[s]);
                        }
                    }
        } else if (lane == 0) {
            int it = 0, n = 0;
            for (int ut = blockIdx.x; ut < p.n_user_tiles; ut += gridDim.x, ++n) {
                if (n > 0) bar_wait(&ctl->a_empty, (n - 1) & 1);
                bar_expect(&ctl->a_full, slab_tx);
                for (int sl = 0; sl < p.kslabs; ++sl)
                    tma_load_2d(smem_a + (size_t)sl * SLAB_BYTES, &map_p, sl * SLAB_K, ut * BM, &ctl->a_full);
                for (int j = 0; j < p.n_item_tiles; ++j, ++it) {
                    // (one barrier pair per 16 KB slab instead of per tile was measured and is slower: 5.9 against 5.2 ms)
                    const int s = it & 1;
                    if (it >= 2) bar_wait(&ctl->b_empty[s], ((it >> 1) - 1) & 1);
                    bar_expect(&ctl->b_full[s], slab_tx);
                    for (int sl = 0; sl < p.kslabs; ++sl)
                        tma_load_2d(smem_b + ((size_t)s * p.kslabs + sl) * SLAB_BYTES, &map_q, sl * SLAB_K, j * BN,
                                    &ctl->b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0 && STREAM) {
            const uint32_t idesc = umma_idesc_tf32(BM, BN);
            const int n_chunks = p.kslabs / CHUNK_SLABS;
            int st = 0, it = 0;
            for (int ut = blockIdx.x; ut < p.n_user_tiles; ut += gridDim.x)
                for (int j = 0; j < p.n_item_tiles; ++j, ++it) {
                    const int b = it % kAccBufs;
                    if (it >= kAccBufs) bar_wait(&ctl->acc_empty[b], ((it / kAccBufs) - 1) & 1);
                    const uint32_t d_tmem = tmem_base + (uint32_t)b * BN;
                    for (int c = 0; c < n_chunks; ++c, ++st) {
                        const int s = st % STREAM_STAGES;
                        bar_wait(&ctl->st_full[s], (st / STREAM_STAGES) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint8_t *stage = base + (size_t)s * kStageBytes;
                        for (int sl = 0; sl < CHUNK_SLABS; ++sl) {
                            const uint32_t a0 = s32(stage + (size_t)sl * SLAB_BYTES);
                            const uint32_t b0 = s32(stage + (size_t)(CHUNK_SLABS + sl) * SLAB_BYTES);
#pragma unroll
                            for (int ks = 0; ks < SLAB_K / 8; ++ks)
                                umma_tf32(d_tmem, umma_desc_sw128(a0 + ks * 32), umma_desc_sw128(b0 + ks * 32), idesc,
                                          (uint32_t)((c | sl | ks) != 0));
                        }
                        umma_commit(&ctl->st_empty[s]);  // the stage may be refilled once these MMAs retire
                    }
                    umma_commit(&ctl->acc_full[b]);
                }
        } else if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BM, BN);
            int it = 0, n = 0;
            for (int ut = blockIdx.x; ut < p.n_user_tiles; ut += gridDim.x, ++n) {
                bar_wait(&ctl->a_full, n & 1);
                for (int j = 0; j < p.n_item_tiles; ++j, ++it) {
                    const int s = it & 1;
                    const int b = it % kAccBufs;
                    bar_wait(&ctl->b_full[s], (it >> 1) & 1);
                    if (it >= kAccBufs) bar_wait(&ctl->acc_empty[b], ((it / kAccBufs) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)b * BN;
                    for (int sl = 0; sl < p.kslabs; ++sl) {
                        const uint32_t a0 = s32(smem_a + (size_t)sl * SLAB_BYTES);
                        const uint32_t b0 = s32(smem_b + ((size_t)s * p.kslabs + sl) * SLAB_BYTES);
#pragma unroll
                        for (int ks = 0; ks < SLAB_K / 8; ++ks)  // K = 8 tf32 (32 bytes) per instruction
                            umma_tf32(d_tmem, umma_desc_sw128(a0 + ks * 32), umma_desc_sw128(b0 + ks * 32), idesc,
                                      (uint32_t)((sl | ks) != 0));
                    }
                    umma_commit(&ctl->b_empty[s]);   // smem stage may be refilled once these MMAs retire
                    umma_commit(&ctl->acc_full[b]);  // accumulator ready for the epilogue
                }
                umma_commit(&ctl->a_empty);
            }
        }
    } else {
        // ===== epilogue: two groups of 4 warps; group g owns accumulator buffer g, i.e. every
        // second item tile. One TMEM lane (= one user) per thread; each group keeps its own sorted
        // candidate list per user (disjoint item tiles, so the two lists never overlap). =====
        const int grp = (warp - 2) >> 2;        // 0 or 1
        const int quad = warp & 3;              // TMEM lane quadrant this warp may read
        const int et = ((warp - 2) & 3) * 32 + lane;  // 0..127 inside the group (stages the item-bias tile)
        float *ring_sc = reinterpret_cast<float *>(ctl + 1);
        int *ring_it = reinterpret_cast<int *>(ring_sc + kRing * kEpilogueThreads);
        const int rt = grp * 128 + et;
        int n = 0;
        PROF_DECL
        for (int ut = blockIdx.x; ut < p.n_user_tiles; ut += gridDim.x, ++n) {
            const int u = ut * BM + quad * 32 + lane;
            float cs[KC];
            int ci[KC];
#pragma unroll
            for (int i = 0; i < KC; ++i) { cs[i] = -INFINITY; ci[i] = -1; }
            int cnt = 0;  // entries in this thread's ring
            // Merge the parked candidates, oldest first. Position = number of kept scores >= sc (ties keep the earlier
            // item first). Scores: new[i] = max(old[i], min(old[i-1], sc)) is exactly "keep / insert here / shift down"
            // for a descending list; items follow with the two comparisons of the neighbours.
            auto merge_pending = [&]() {
                const int most = __reduce_max_sync(0xffffffffu, cnt);
#pragma unroll 1
                for (int r = 0; r < most; ++r) {
                    if (r < cnt) {
                        const float sc = ring_sc[r * kEpilogueThreads + rt];
                        const int item = ring_it[r * kEpilogueThreads + rt];
                        if (sc > cs[KC - 1]) {
                            bool ge_hi = cs[KC - 1] >= sc;
#pragma unroll
                            for (int i = KC - 1; i > 0; --i) {
                                const bool ge_lo = cs[i - 1] >= sc;
                                ci[i] = ge_hi ? ci[i] : (ge_lo ? item : ci[i - 1]);
                                cs[i] = fmaxf(cs[i], fminf(cs[i - 1], sc));
                                ge_hi = ge_lo;
                            }
                            ci[0] = ge_hi ? ci[0] : item;
                            cs[0] = fmaxf(cs[0], sc);
                        }
                    }
                }
                cnt = 0;
            };
            const uint4 *mrow = (p.bitmap && u < p.users) ? reinterpret_cast<const uint4 *>(p.bitmap + (size_t)u * p.mask_pitch) : nullptr;
            const int it0 = n * p.n_item_tiles;  // global tile counter at the start of this user tile
            // first item tile of this user tile that belongs to my buffer
            int j = ((it0 & 1) == grp) ? 0 : 1;
            uint4 mask = (mrow && j < p.n_item_tiles) ? __ldg(mrow + j) : make_uint4(0u, 0u, 0u, 0u);
            for (int t = 0; j < p.n_item_tiles; j += 2, ++t) {
                const int it = it0 + j;
                const int n0 = j * BN;
                float *ibs = ctl->ib[grp][t & 1];  // double buffered: a slow warp may still read the previous tile's
                ibs[et] = (n0 + et < p.items) ? __ldg(p.item_bias + n0 + et) : 0.f;
                const uint4 mask_next = (mrow && j + 2 < p.n_item_tiles) ? __ldg(mrow + j + 2) : make_uint4(0u, 0u, 0u, 0u);
                if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
                else          asm volatile("bar.sync 2, 128;" ::: "memory");
                const int ab = it % kAccBufs;  // this group's tiles alternate between accumulators grp and grp + 2
                bar_wait(&ctl->acc_full[ab], (it / kAccBufs) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                PROF(0)
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * BN + ch * 32), v);
                    PROF(1)
                    const uint32_t mword = ch == 0 ? mask.x : ch == 1 ? mask.y : ch == 2 ? mask.z : mask.w;
                    const int valid = p.items - (n0 + ch * 32);  // columns [0, valid) of this chunk exist
                    const uint32_t live = (valid >= 32 ? 0xffffffffu : valid <= 0 ? 0u : ((1u << valid) - 1u)) & ~mword;
                    // cheap unrolled pass: add the item bias, mark the columns that beat the current
                    // KC-th best score. After the first few tiles almost no column does.
                    // Three instructions per column: the bias add, thr - score (negative <=> the column beats the
                    // threshold) and a funnel shift that collects the sign bits (column 31 first, so column c ends up in bit c).
                    const float thr = cs[KC - 1];
                    uint32_t hits = 0;
#pragma unroll
                    for (int c = 31; c >= 0; --c) {
                        v[c] += ibs[ch * 32 + c];
                        hits = __funnelshift_l(__float_as_uint(thr - v[c]), hits, 1);
                    }
                    hits &= live;
                    PROF(2)
                    // park the hits (the loop body exists once, not 128 times); merge when a ring is full
                    while (__any_sync(0xffffffffu, hits != 0)) {
                        if (hits) {
                            const int c = __ffs(hits) - 1;
                            hits &= hits - 1;
                            const float sc = pick32(v, c);
                            if (sc > cs[KC - 1]) {
                                ring_sc[cnt * kEpilogueThreads + rt] = sc;
                                ring_it[cnt * kEpilogueThreads + rt] = n0 + ch * 32 + c;
                                ++cnt;
                            }
                        }
                        if (__any_sync(0xffffffffu, cnt == kRing)) { PROF(3) merge_pending(); PROF(4) }
                    }
                    PROF(3)
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(&ctl->acc_empty[ab]);
                mask = mask_next;
            }
            merge_pending();
            PROF(4)
            if (u < p.users) {
#pragma unroll
                for (int i = 0; i < KC; ++i) {
                    p.cand_items[((size_t)u * 2 + grp) * KC + i] = ci[i];
                    p.cand_scores[((size_t)u * 2 + grp) * KC + i] = cs[i];
                }
            }
            PROF(5)
        }
#ifdef CU2B_PREDICT_PROFILE
        if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6))
            printf("PROF warp %d: wait %lld ld %lld scan %lld loop %lld merge %lld store %lld cycles\n", warp, prof[0], prof[1], prof[2],
                   prof[3], prof[4], prof[5]);
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ncols = kAccBufs * BN;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// One bit per (user, item) already rated.
__global__ void __launch_bounds__(256)
rated_bitmap_kernel(const int *__restrict__ indptr, const int *__restrict__ indices, int users, long long nnz,
                    uint32_t *bitmap, int mask_pitch) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = users;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(indptr + mid) <= j) lo = mid; else hi = mid;
        }
        const int item = __ldg(indices + j);
        atomicOr(bitmap + (size_t)lo * mask_pitch + (item >> 5), 1u << (item & 31));
    }
}

// Exact fp32 score of every candidate in the reference's op order (predict.cu:22-26), then the
// order inside this pass: score descending, item ascending on ties. One warp per user, lane = candidate.
// The pass's best min(2 KC, 16) candidates go to slots [slot0, slot0 + 16) of the user's collected list; when
// `bitmap` is given their bits are set so that the next pass looks past them. kp = row pitch (k padded).
template <int KC>
__global__ void __launch_bounds__(256)
predict_rescore_kernel(const float *__restrict__ P, const float *__restrict__ Q, const float *__restrict__ user_bias,
                       const float *__restrict__ item_bias, float mu, int k, int kp, int users,
                       const int32_t *__restrict__ cand_items, int slot0, int slots, int32_t *out_items, float *out_scores,
                       uint32_t *bitmap, int mask_pitch) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= users) return;
    static_assert(2 * KC <= 32, "one lane per candidate");
    int item = lane < 2 * KC ? cand_items[(size_t)u * 2 * KC + lane] : -1;
    float score = -INFINITY;
    if (item >= 0) {
        // rows are padded with zeros to kp (a multiple of 32): 128-bit loads, products still added one by one in
        // ascending f; the zero products past k leave the sum unchanged
        const float4 *pu = reinterpret_cast<const float4 *>(P + (size_t)u * kp);
        const float4 *qi = reinterpret_cast<const float4 *>(Q + (size_t)item * kp);
        float pred = __fadd_rn(__fadd_rn(mu, __ldg(user_bias + u)), __ldg(item_bias + item));
        const int vec = (k + 3) >> 2;
#pragma unroll 4
        for (int f = 0; f < vec; ++f) {
            const float4 a = __ldg(qi + f), b = __ldg(pu + f);
            pred = __fadd_rn(pred, __fmul_rn(a.x, b.x));
            if (4 * f + 1 < k) pred = __fadd_rn(pred, __fmul_rn(a.y, b.y));
            if (4 * f + 2 < k) pred = __fadd_rn(pred, __fmul_rn(a.z, b.z));
            if (4 * f + 3 < k) pred = __fadd_rn(pred, __fmul_rn(a.w, b.w));
        }
        score = pred;
    }
    int rank = 0;
#pragma unroll
    for (int o = 0; o < 2 * KC; ++o) {
        const float so = __shfl_sync(0xffffffffu, score, o);
        const int io = __shfl_sync(0xffffffffu, item, o);
        if (io >= 0 && (so > score || (so == score && io < item))) ++rank;
    }
    if (item >= 0 && rank < 16) {
        out_items[(size_t)u * slots + slot0 + rank] = item;
        out_scores[(size_t)u * slots + slot0 + rank] = score;
        if (bitmap) atomicOr(bitmap + (size_t)u * mask_pitch + (item >> 5), 1u << (item & 31));
    }
}

// Final order of a user's collected candidates (<= 128 slots from up to 8 passes): score descending, item
// ascending on ties, the first topk go out. One warp per user, four slots per lane, rank by counting.
__global__ void __launch_bounds__(256)
predict_order_kernel(const int32_t *__restrict__ col_items, const float *__restrict__ col_scores, int slots, int users, int topk,
                     int32_t *out_items, float *out_scores) {
    const int lane = threadIdx.x & 31;
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= users) return;
    int it[4];
    float sc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sidx = q * 32 + lane;
        it[q] = sidx < slots ? col_items[(size_t)u * slots + sidx] : -1;
        sc[q] = sidx < slots ? col_scores[(size_t)u * slots + sidx] : -INFINITY;
    }
    int rank[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q * 32 >= slots) break;
        for (int o = 0; o < 32; ++o) {
            const float so = __shfl_sync(0xffffffffu, sc[q], o);
            const int io = __shfl_sync(0xffffffffu, it[q], o);
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (io >= 0 && it[m] >= 0 && (so > sc[m] || (so == sc[m] && io < it[m]))) ++rank[m];
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
        if (it[m] >= 0 && rank[m] < topk) {
            out_items[(size_t)u * topk + rank[m]] = it[m];
            out_scores[(size_t)u * topk + rank[m]] = sc[m];
        }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cu2b_status make_row_major_map(EncodeTiledFn encode, CUtensorMap *map, const float *base, int rows, int k) {
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)SLAB_K, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cu2b_fail(CU2B_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return CU2B_OK;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cu2b_status alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return cu2b_fail(CU2B_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return CU2B_OK;
    }
    template <typename T> T *as() { return (T *)p; }
};

// Row-major [rows x k] host matrix -> [rows x kp] device matrix, zero padded.
cu2b_status upload_padded(DevBuf &buf, const float *src, int rows, int k, int kp) {
    cu2b_status rc = buf.alloc((size_t)rows * kp * 4);
    if (rc != CU2B_OK) return rc;
    if (kp == k) {
        CUDA_TRY(cudaMemcpy(buf.p, src, (size_t)rows * k * 4, cudaMemcpyHostToDevice));
    } else {
        CUDA_TRY(cudaMemset(buf.p, 0, (size_t)rows * kp * 4));
        CUDA_TRY(cudaMemcpy2D(buf.p, (size_t)kp * 4, src, (size_t)k * 4, (size_t)k * 4, (size_t)rows, cudaMemcpyHostToDevice));
    }
    return CU2B_OK;
}

template <bool STREAM>
cu2b_status launch_candidates(const CUtensorMap &mp, const CUtensorMap &mq, const PredictParams &pp, int sm_count) {
    const size_t smem = (STREAM ? (size_t)STREAM_STAGES * 2 * CHUNK_SLABS * SLAB_BYTES : (size_t)3 * pp.kslabs * SLAB_BYTES) +
                        sizeof(PredictSmemCtl) + kRingBytes + 1024;
    CUDA_TRY(cudaFuncSetAttribute(predict_candidates_kernel<16, STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(pp.n_user_tiles, sm_count));
    predict_candidates_kernel<16, STREAM><<<grid, kThreadsPredict, smem>>>(mp, mq, pp);
    CUDA_TRY(cudaGetLastError());
    return CU2B_OK;
}

}  // namespace

extern "C" cu2b_status cu2b_predict_topk(const float *P, int rows, const float *Q, int cols, const float *user_bias,
                                         const float *item_bias, float global_bias, int n_factors,
                                         const cu2b_csr *exclude, int topk, int32_t *out_items, float *out_scores,
                                         double *ms_out) {
    if (!P || !Q || !user_bias || !item_bias || !out_items || !out_scores || rows < 1 || cols < 1 || topk < 1)
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_predict_topk: bad argument");
    if (n_factors < 1 || n_factors > 512)
        return cu2b_fail(CU2B_ERR_UNSUPPORTED, "cu2b_predict_topk: n_factors must be in [1, 512] (got %d)", n_factors);
    if (topk > 128) return cu2b_fail(CU2B_ERR_UNSUPPORTED, "cu2b_predict_topk: topk <= 128 (got %d)", topk);
    if (exclude && (exclude->on_device || exclude->rows > rows || exclude->cols > cols))
        return cu2b_fail(CU2B_ERR_INVALID, "cu2b_predict_topk: exclude matrix must be a host CSR within the model dimensions");
    int dev = 0, cc_major = 0, sm_count = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (cc_major != 10) return cu2b_fail(CU2B_ERR_CUDA, "cu2b_predict_topk needs an sm_100 device (tcgen05)");
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) return cu2b_fail(CU2B_ERR_CUDA, "cuTensorMapEncodeTiled is not available");

    const int k = n_factors;
    // factor rows padded to whole 128-byte swizzle rows; beyond 128 factors to whole 64-factor chunks (STREAM)
    const bool stream = k > 128;
    const int kp = stream ? ((k + 63) / 64) * 64 : ((k + 31) / 32) * 32;
    // 16 candidates per epilogue group and pass; both groups together hand 32 candidates to the exact rescoring
    // pass, whose 16 best are final. A group sees every second item tile, so each list can hold a pass's whole yield.
    const int kc = 16;
    const int passes = (topk + 15) / 16, slots = passes * 16;
    DevBuf dP, dQ, dub, dib, dcand_i, dcand_s, dcol_i, dcol_s, dout_i, dout_s, dmask, dptr, dind;
    cu2b_status rc;
    if ((rc = upload_padded(dP, P, rows, k, kp)) || (rc = upload_padded(dQ, Q, cols, k, kp)) || (rc = dub.alloc((size_t)rows * 4)) ||
        (rc = dib.alloc((size_t)cols * 4)) || (rc = dcand_i.alloc((size_t)rows * 2 * kc * 4)) ||
        (rc = dcand_s.alloc((size_t)rows * 2 * kc * 4)) || (rc = dcol_i.alloc((size_t)rows * slots * 4)) ||
        (rc = dcol_s.alloc((size_t)rows * slots * 4)) || (rc = dout_i.alloc((size_t)rows * topk * 4)) ||
        (rc = dout_s.alloc((size_t)rows * topk * 4)))
        return rc;
    CUDA_TRY(cudaMemcpy(dub.p, user_bias, (size_t)rows * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dib.p, item_bias, (size_t)cols * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(dcol_i.p, 0xFF, (size_t)rows * slots * 4));  // item -1 = "no such candidate"
    CUDA_TRY(cudaMemset(dout_i.p, 0xFF, (size_t)rows * topk * 4));
    {
        std::vector<float> nanv((size_t)rows * topk, NAN);
        CUDA_TRY(cudaMemcpy(dout_s.p, nanv.data(), nanv.size() * 4, cudaMemcpyHostToDevice));
    }
    PredictParams pp;
    pp.users = rows;
    pp.items = cols;
    pp.kslabs = kp / SLAB_K;
    pp.n_user_tiles = (rows + BM - 1) / BM;
    pp.n_item_tiles = (cols + BN - 1) / BN;
    pp.item_bias = dib.as<float>();
    pp.bitmap = nullptr;
    pp.mask_pitch = pp.n_item_tiles * 4;
    pp.cand_items = dcand_i.as<int32_t>();
    pp.cand_scores = dcand_s.as<float>();
    const bool have_exclude = exclude && exclude->nonzeros > 0;
    if (have_exclude || passes > 1) {
        if ((rc = dmask.alloc((size_t)rows * pp.mask_pitch * 4))) return rc;
        CUDA_TRY(cudaMemset(dmask.p, 0, (size_t)rows * pp.mask_pitch * 4));
        pp.bitmap = dmask.as<uint32_t>();
    }
    if (have_exclude) {
        if ((rc = dptr.alloc(((size_t)exclude->rows + 1) * 4)) || (rc = dind.alloc((size_t)exclude->nonzeros * 4))) return rc;
        CUDA_TRY(cudaMemcpy(dptr.p, exclude->indptr, ((size_t)exclude->rows + 1) * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dind.p, exclude->indices, (size_t)exclude->nonzeros * 4, cudaMemcpyHostToDevice));
        const int grid = (int)std::min<long long>(((long long)exclude->nonzeros + 255) / 256, (long long)sm_count * 16);
        rated_bitmap_kernel<<<grid, 256>>>(dptr.as<int>(), dind.as<int>(), exclude->rows, exclude->nonzeros,
                                          dmask.as<uint32_t>(), pp.mask_pitch);
        CUDA_TRY(cudaGetLastError());
    }
    CUtensorMap mp, mq;
    if ((rc = make_row_major_map(encode, &mp, dP.as<float>(), rows, kp)) || (rc = make_row_major_map(encode, &mq, dQ.as<float>(), cols, kp)))
        return rc;
    cudaEvent_t ev[3];
    for (cudaEvent_t &e : ev) cudaEventCreate(&e);
    float ms_c = 0.f, ms_r = 0.f;
    const int warps_per_cta = 8, rescore_grid = (rows + warps_per_cta - 1) / warps_per_cta;
    for (int pass = 0; pass < passes && rc == CU2B_OK; ++pass) {
        cudaEventRecord(ev[0]);
        rc = stream ? launch_candidates<true>(mp, mq, pp, sm_count) : launch_candidates<false>(mp, mq, pp, sm_count);
        if (rc != CU2B_OK) break;
        cudaEventRecord(ev[1]);
        predict_rescore_kernel<16><<<rescore_grid, 256>>>(dP.as<float>(), dQ.as<float>(), dub.as<float>(), dib.as<float>(),
                                                         global_bias, k, kp, rows, pp.cand_items, pass * 16, slots,
                                                         dcol_i.as<int32_t>(), dcol_s.as<float>(),
                                                         passes > 1 ? dmask.as<uint32_t>() : nullptr, pp.mask_pitch);
        cudaEventRecord(ev[2]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = cu2b_fail(CU2B_ERR_CUDA, "predict kernels: %s", cudaGetErrorString(e)); break; }
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        ms_c += a;
        ms_r += b;
    }
    if (rc == CU2B_OK) {
        cudaEventRecord(ev[0]);
        predict_order_kernel<<<rescore_grid, 256>>>(dcol_i.as<int32_t>(), dcol_s.as<float>(), slots, rows, topk,
                                                   dout_i.as<int32_t>(), dout_s.as<float>());
        cudaEventRecord(ev[1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = cu2b_fail(CU2B_ERR_CUDA, "predict order kernel: %s", cudaGetErrorString(e));
        float a = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        ms_r += a;
    }
    for (cudaEvent_t &e : ev) cudaEventDestroy(e);
    if (rc != CU2B_OK) return rc;
    CUDA_TRY(cudaMemcpy(out_items, dout_i.p, (size_t)rows * topk * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(out_scores, dout_s.p, (size_t)rows * topk * 4, cudaMemcpyDeviceToHost));
    if (ms_out) { ms_out[0] = ms_c; ms_out[1] = ms_r; }
    return CU2B_OK;
}
