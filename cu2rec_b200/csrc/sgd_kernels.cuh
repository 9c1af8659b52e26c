// Hogwild SGD update kernel (replaces the reference's sgd_update, sgd.cu:22-75, one thread per
// user walking a whole factor row) and the per-user sampler (replaces initCurand + the
// curand_uniform draw, sgd.cu:11-16,36-37).
//
// Mapping: a group of L lanes owns one rating; each lane holds V float4 slices of the P row
// and of the Q row (L*V*4 >= kp). k=128 -> L=32,V=1: one 128-bit load per lane per row, the
// whole warp touches one contiguous 512-byte row. Smaller k packs 32/L ratings into a warp.
// The dot product is reduced with an xor butterfly over the L lanes (every lane ends with the
// same value), the rating error is formed once, and both rows are written back in place.
// Update arithmetic (sgd_step below): mf_sequential.cu:129-141 with the learning rate folded into
// the error and the regularisers, one multiply + one fma per step; a sequential replay on the
// CPU reproduces it bit for bit (oracle flavour KERNEL), the reference's own op order agrees to
// fp32 rounding (stated tolerance in the tests).
#ifndef CU2B_SGD_KERNELS_CUH_
#define CU2B_SGD_KERNELS_CUH_

#include "stream_pipe.cuh"

namespace cu2b {

struct SgdParams {
    StreamView sv;              // the update stream (device memory)
    float *P, *Q, *user_bias, *item_bias;
    int kp;                     // row pitch in floats (multiple of 4)
    int ibs;                    // item_bias stride in floats: item i's bias is item_bias[i * ibs] (ItemBiasLayout)
    float mu;
    const float *lr;            // device scalar: current learning rate (decayed on device)
    float P_reg, Q_reg, ub_reg, ib_reg;
    int is_train;               // 0 => Q / item_bias frozen
    unsigned long long *chunk_counter;  // dynamic chunk scheduler, zeroed before launch
    // Per-user mode ordering gate: gate[j] = number of segments (reference iterations) whose
    // chunk j is complete. Chunk j of segment a starts only when gate[j] == a, so a user's
    // updates from consecutive iterations never overlap (the reference separates iterations
    // by kernel launches, training.cu:107-112). nullptr => ungated flat stream.
    int *gate;
    int seg0;                   // absolute index of segment 0 of this launch
    int serial;                 // 1 => a single group processes the stream strictly in order
    // DSGD: the stream range {first rating, count} is only known on the device (bucket sizes of
    // a sampled round). When set, sv.base / seg_len / chunk counts are derived from it in-kernel.
    const int *dyn_range;
    int *error_flag;            // DevState::error
};

#define PHILOX_TAG 0x53474431u
#define PHILOX_KEY1 0x43553242u

__device__ __forceinline__ uint32_t philox4x32_10_x(uint32_t c0, uint32_t c1, uint32_t c2,
                                                    uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

// Same generator, first two output words (the second one decides the item-step thinning of the
// DSGD sampler; the first is the rating draw and is bit-identical to philox4x32_10_x).
__device__ __forceinline__ uint2 philox4x32_10_xy(uint32_t c0, uint32_t c1, uint32_t c2,
                                                  uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint2(c0, c1);
}

// One draw per (iteration, active user): thread i handles draw i of n_iter * n_active and
// writes it to segment t = i / n_active of the update stream (segment pitch seg_pitch).
__global__ void __launch_bounds__(256)
sample_per_user_kernel(const int *__restrict__ indptr, const cu2b_rating *__restrict__ coo,
                       const int *__restrict__ active_users, int n_active, uint32_t seed,
                       int iter0, long long n_draws, cu2b_rating *__restrict__ out,
                       long long seg_pitch, const int *__restrict__ user_ids) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n_draws; i += stride) {
        const int t = (int)(i / n_active);
        const int a = (int)(i - (long long)t * n_active);
        const int u = __ldg(&active_users[a]);
        const int lo = __ldg(&indptr[u]), hi = __ldg(&indptr[u + 1]);
        // the counter is keyed by the ORIGINAL user id (DSGD strips renumber users locally)
        const uint32_t uid = user_ids ? (uint32_t)__ldg(&user_ids[u]) : (uint32_t)u;
        const uint32_t r = philox4x32_10_x(uid, (uint32_t)(iter0 + t), 0u, PHILOX_TAG, seed, PHILOX_KEY1);
        const int j = lo + (int)__umulhi(r, (uint32_t)(hi - lo));
        cu2b_rating v;
        v.user = u;
        v.item = __ldg(&coo[j].item);
        v.rating = __ldg(&coo[j].rating);
        out[(long long)t * seg_pitch + a] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// per_rating sampler (CU2B_SAMPLER_PER_RATING; no reference counterpart -- the reference only samples per
// user, sgd.cu:27-37): the update stream is a shuffled pass over the rating list. Update number q of a
// run (q = iteration x active users + position) applies rating perm_e(q mod nnz) of pass e = q / nnz,
// perm_e a keyed bijection of [0, nnz): a 4-round Feistel network on 2h bits (2^(2h) >= nnz) with cycle
// walking, round function = one Philox-style multiply-xor keyed by (seed, pass, round). Integer only,
// so the CPU oracle (orc_rating_permutation) reproduces it bit for bit.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t feistel_round(uint32_t x, uint32_t key) {
    uint32_t v = (x ^ key) * 0x9E3779B1u;
    v ^= v >> 15;
    v *= 0x85EBCA77u;
    v ^= v >> 13;
    return v;
}
__host__ __device__ __forceinline__ unsigned long long rating_permutation(unsigned long long j, unsigned long long n, int half_bits,
                                                                          uint32_t seed, uint32_t pass) {
    const uint32_t mask = half_bits >= 32 ? 0xffffffffu : ((1u << half_bits) - 1u);
    unsigned long long x = j;
    do {
        uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
        for (uint32_t round = 0; round < 4; ++round) {
            const uint32_t t = l ^ (feistel_round(r, seed ^ (pass * 0x632BE5ABu) ^ (round * 0xB5297A4Du + 0x68E31DA4u)) & mask);
            l = r;
            r = t;
        }
        x = ((unsigned long long)l << half_bits) | r;
    } while (x >= n);  // cycle walking: a bijection of [0, 2^(2h)) restricted to [0, n) stays a bijection
    return x;
}
__host__ __device__ __forceinline__ int feistel_half_bits(unsigned long long n) {
    int bits = 1;
    while (bits < 64 && (1ULL << bits) < n) ++bits;
    return (bits + 1) / 2;
}

// Thread i writes update q0 + i of the run to segment (i / seg_len) of the update stream.
__global__ void __launch_bounds__(256)
sample_per_rating_kernel(const cu2b_rating *__restrict__ coo, unsigned long long nnz, int half_bits, uint32_t seed,
                         unsigned long long q0, long long n_draws, int seg_len, cu2b_rating *__restrict__ out,
                         long long seg_pitch) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n_draws; i += stride) {
        const unsigned long long q = q0 + (unsigned long long)i;
        const unsigned long long r = rating_permutation(q % nnz, nnz, half_bits, seed, (uint32_t)(q / nnz));
        cu2b_rating v;
        v.user = __ldg(&coo[r].user);
        v.item = __ldg(&coo[r].item);
        v.rating = __ldg(&coo[r].rating);
        const long long t = i / seg_len;
        out[t * seg_pitch + (i - t * seg_len)] = v;
    }
}

// One SGD step of a row element, learning rate folded in:
//   lr * (err * other - reg * self)  ==  fma(a, other, -(c * self)),  a = lr * err, c = lr * reg
// (mf_sequential.cu:134-137 / sgd.cu:55-61). The row then takes self + step with ONE rounding,
// either as an FADD in registers or as the L2 atomic add, so both write paths agree bit for bit.
struct StepCoef {
    float cP, cQ, cU, cI;
};
__device__ __forceinline__ StepCoef step_coef(float lr, float P_reg, float Q_reg, float ub_reg, float ib_reg) {
    StepCoef c;
    c.cP = __fmul_rn(lr, P_reg);
    c.cQ = __fmul_rn(lr, Q_reg);
    c.cU = __fmul_rn(lr, ub_reg);
    c.cI = __fmul_rn(lr, ib_reg);
    return c;
}
__device__ __forceinline__ float sgd_step(float a, float other, float c, float self) {
    return __fmaf_rn(a, other, -__fmul_rn(c, self));
}
// bias: lr * (err - reg * b) == fma(-c, b, a)   (mf_sequential.cu:140-141)
__device__ __forceinline__ float bias_step(float a, float c, float self) { return __fmaf_rn(-c, self, a); }

template <int L>
__device__ __forceinline__ float group_sum(float a) {
#pragma unroll
    for (int off = L / 2; off >= 1; off >>= 1) a = a + __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

// 128-bit fire-and-forget float add at L2 (SASS REDG.E.ADD.F32x4): concurrent Hogwild updates of
// one row accumulate instead of overwriting each other.
__device__ __forceinline__ void red_add_v4(float4 *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add_f32(float *addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// One update with the user side held in registers (the user-major kernels: mf_sgd_user_rounds,
// mf_sgd_user_tiles, mf_sgd_user_runs): the lane group's P slice `pv` and the user bias `ub` are
// updated in place, the item row and item bias take their steps as L2 atomic adds. `ok` false =>
// the group idles through the warp-wide shuffles.
// Item-side operands of one update: the lane's slice of the Q row and the item bias.
template <int V>
struct ItemSide {
    float4 q[V];
    float ib;
};
// Issue-order-pinned L2 loads (ld.global.cg as volatile asm): the item bias and the item row are
// requested back to back at the top of an update. Left to the scheduler, the bias load of the
// DSGD kernel was issued only after the dot-product butterfly, which put a second L2 round trip
// into every update's read -> atomic-add window (measured: 0.77 us -> see profiles/r1_dsgd_window.jsonl).
__device__ __forceinline__ float ldcg_pinned(const float *p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ldcg_pinned(const float4 *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <int L, int V>
__device__ __forceinline__ void item_side_load(ItemSide<V> &it, int item, bool ok, int l, int vecs, const float4 *Qv,
                                               const float *item_bias, int ibs) {
    const size_t qo = (size_t)item * vecs + l;
    it.ib = 0.f;
    if (ok) it.ib = ldcg_pinned(item_bias + (size_t)item * ibs);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        it.q[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && v * L + l < vecs) it.q[v] = ldcg_pinned(Qv + qo + v * L);
    }
}
// The arithmetic of one update given its operands; P slice and user bias are updated in place,
// the item row and item bias take their steps as L2 atomic adds.
// MASKED (the thinning instantiations of the DSGD kernel only): is_train is a bit mask, bit 0 = the item
// row takes its step, bit 1 = the item bias takes its step; otherwise it is the plain 0 / 1 flag.
template <int L, int V, bool MASKED = false>
__device__ __forceinline__ void user_side_apply(float4 (&pv)[V], float &ub, const ItemSide<V> &it, int item, float rating,
                                                bool ok, int l, int vecs, float4 *Qv, float *item_bias, int ibs, float mu,
                                                float lr, const StepCoef &sc, int is_train) {
    const size_t qo = (size_t)item * vecs + l;
    float acc = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        acc = __fmaf_rn(pv[v].x, it.q[v].x, acc);
        acc = __fmaf_rn(pv[v].y, it.q[v].y, acc);
        acc = __fmaf_rn(pv[v].z, it.q[v].z, acc);
        acc = __fmaf_rn(pv[v].w, it.q[v].w, acc);
    }
    const float dot = group_sum<L>(acc);
    const float pred = __fadd_rn(__fadd_rn(__fadd_rn(mu, ub), it.ib), dot);
    const float err = __fsub_rn(rating, pred);
    const float ea = __fmul_rn(lr, err);
    if (ok) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 x = pv[v], y = it.q[v];
            float4 nq;
            nq.x = sgd_step(ea, x.x, sc.cQ, y.x);
            nq.y = sgd_step(ea, x.y, sc.cQ, y.y);
            nq.z = sgd_step(ea, x.z, sc.cQ, y.z);
            nq.w = sgd_step(ea, x.w, sc.cQ, y.w);
            pv[v].x = __fadd_rn(x.x, sgd_step(ea, y.x, sc.cP, x.x));
            pv[v].y = __fadd_rn(x.y, sgd_step(ea, y.y, sc.cP, x.y));
            pv[v].z = __fadd_rn(x.z, sgd_step(ea, y.z, sc.cP, x.z));
            pv[v].w = __fadd_rn(x.w, sgd_step(ea, y.w, sc.cP, x.w));
            if ((MASKED ? (is_train & 1) : is_train) && v * L + l < vecs) red_add_v4(Qv + qo + v * L, nq);
        }
        if ((MASKED ? (is_train & 2) : is_train) && l == 0) red_add_f32(item_bias + (size_t)item * ibs, bias_step(ea, sc.cI, it.ib));
        ub = __fadd_rn(ub, bias_step(ea, sc.cU, ub));
    }
}
// One update with the user side held in registers (the user-major kernels: mf_sgd_user_rounds,
// mf_sgd_user_tiles, mf_sgd_user_runs). `ok` false => the group idles through the warp-wide shuffles.
template <int L, int V, bool MASKED = false>
__device__ __forceinline__ void user_side_update(float4 (&pv)[V], float &ub, int item, float rating, bool ok, int l,
                                                 int vecs, float4 *Qv, float *item_bias, int ibs, float mu, float lr,
                                                 const StepCoef &sc, int is_train) {
    ItemSide<V> it;
    item_side_load<L, V>(it, item, ok, l, vecs, Qv, item_bias, ibs);
    user_side_apply<L, V, MASKED>(pv, ub, it, item, rating, ok, l, vecs, Qv, item_bias, ibs, mu, lr, sc, is_train);
}

// Model rows are read-write data shared by every SM: they are read with ld.global.cg and, on
// the non-atomic path, written with st.global.cg (L2 only; SASS LDG/STG .STRONG.GPU) so that no
// stale copy can sit in the non-coherent L1. Measured on B200 (profiles/r1_sweep2.jsonl): weak
// L1::no_allocate loads are ~18 % slower, weak stores change nothing.
// Processes UNR ratings (one per slot) for this lane's group. `ok[x]` false => slot idle.
// WMODE bit 0: item-side updates (Q row, item_bias) are applied as atomic adds of the SGD step
// instead of read-modify-write stores; bit 1: the same for the user side (P row, user_bias).
// Same arithmetic (x + step, one rounding) when a row is touched by one update at a time;
// under contention no step is lost.
template <int L, int V, int UNR, int WMODE>
__device__ __forceinline__ void sgd_update_slots(const SgdParams &p, const cu2b_rating *rt,
                                                 const bool *ok, int l, int vecs, float lr) {
    constexpr bool ATOMQ = (WMODE & 1) != 0, ATOMP = (WMODE & 2) != 0;
    float4 pv[UNR][V], qv[UNR][V];
    float ub[UNR], ib[UNR];
    float4 *const Pv = reinterpret_cast<float4 *>(p.P);
    float4 *const Qv = reinterpret_cast<float4 *>(p.Q);
    const StepCoef sc = step_coef(lr, p.P_reg, p.Q_reg, p.ub_reg, p.ib_reg);
#pragma unroll
    for (int x = 0; x < UNR; ++x) {
        const size_t po = (size_t)rt[x].user * vecs + l, qo = (size_t)rt[x].item * vecs + l;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (ok[x] && v * L + l < vecs) {
                pv[x][v] = __ldcg(Pv + po + v * L);
                qv[x][v] = __ldcg(Qv + qo + v * L);
            } else {
                pv[x][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                qv[x][v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        ub[x] = ok[x] ? __ldcg(p.user_bias + rt[x].user) : 0.f;
        ib[x] = ok[x] ? __ldcg(p.item_bias + (size_t)rt[x].item * p.ibs) : 0.f;
    }
#pragma unroll
    for (int x = 0; x < UNR; ++x) {
        float acc = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            acc = __fmaf_rn(pv[x][v].x, qv[x][v].x, acc);
            acc = __fmaf_rn(pv[x][v].y, qv[x][v].y, acc);
            acc = __fmaf_rn(pv[x][v].z, qv[x][v].z, acc);
            acc = __fmaf_rn(pv[x][v].w, qv[x][v].w, acc);
        }
        const float dot = group_sum<L>(acc);
        const float pred = __fadd_rn(__fadd_rn(__fadd_rn(p.mu, ub[x]), ib[x]), dot);
        const float err = __fsub_rn(rt[x].rating, pred);
        const float ea = __fmul_rn(lr, err);
        const size_t po = (size_t)rt[x].user * vecs + l, qo = (size_t)rt[x].item * vecs + l;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float4 a = pv[x][v], b = qv[x][v];
            float4 na, nb;
            na.x = sgd_step(ea, b.x, sc.cP, a.x);
            na.y = sgd_step(ea, b.y, sc.cP, a.y);
            na.z = sgd_step(ea, b.z, sc.cP, a.z);
            na.w = sgd_step(ea, b.w, sc.cP, a.w);
            nb.x = sgd_step(ea, a.x, sc.cQ, b.x);
            nb.y = sgd_step(ea, a.y, sc.cQ, b.y);
            nb.z = sgd_step(ea, a.z, sc.cQ, b.z);
            nb.w = sgd_step(ea, a.w, sc.cQ, b.w);
            if (ok[x] && v * L + l < vecs) {
                if (ATOMP) {
                    red_add_v4(Pv + po + v * L, na);
                } else {
                    na.x = __fadd_rn(a.x, na.x); na.y = __fadd_rn(a.y, na.y);
                    na.z = __fadd_rn(a.z, na.z); na.w = __fadd_rn(a.w, na.w);
                    __stcg(Pv + po + v * L, na);
                }
                if (p.is_train) {
                    if (ATOMQ) {
                        red_add_v4(Qv + qo + v * L, nb);
                    } else {
                        nb.x = __fadd_rn(b.x, nb.x); nb.y = __fadd_rn(b.y, nb.y);
                        nb.z = __fadd_rn(b.z, nb.z); nb.w = __fadd_rn(b.w, nb.w);
                        __stcg(Qv + qo + v * L, nb);
                    }
                }
            }
        }
        if (ok[x] && l == 0) {
            const float ustep = bias_step(ea, sc.cU, ub[x]);
            if (ATOMP) red_add_f32(p.user_bias + rt[x].user, ustep);
            else __stcg(p.user_bias + rt[x].user, __fadd_rn(ub[x], ustep));
            if (p.is_train) {
                const float step = bias_step(ea, sc.cI, ib[x]);
                if (ATOMQ) red_add_f32(p.item_bias + (size_t)rt[x].item * p.ibs, step);
                else __stcg(p.item_bias + (size_t)rt[x].item * p.ibs, __fadd_rn(ib[x], step));
            }
        }
    }
}

template <int L, int V, int UNR, int WMODE, int OCC = 1>
__global__ void __launch_bounds__(kThreads, OCC)
mf_sgd_hogwild(const SgdParams p_in) {
    __shared__ StreamSmem sm;
    SgdParams p = p_in;
    if (p.dyn_range) {
        const int first = __ldg(p.dyn_range), cnt = __ldg(p.dyn_range + 1);
        p.sv.base += first;
        p.sv.seg_len = cnt;
        p.sv.chunks_per_seg = (cnt + p.sv.chunk - 1) / p.sv.chunk;
        p.sv.num_chunks = p.sv.chunks_per_seg;
    }
    pipe_init(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kConsumerWarps) {  // producer warp
        if (lane == 0) {
            pipe_produce(
                sm, p.sv,
                [&]() -> long long {
                    const unsigned long long c = atomicAdd(p.chunk_counter, 1ULL);
                    return c < (unsigned long long)p.sv.num_chunks ? (long long)c : -1LL;
                },
                [&](long long seg, int j, int) {
                    if (p.gate) {
                        const int want = p.seg0 + (int)seg;
                        const long long t0 = clock64();
                        while (ld_acquire_gpu(p.gate + j) < want) {
                            __nanosleep(32);
                            if (clock64() - t0 > (8LL << 30)) {  // ~4 s: report, never hang
                                if (p.error_flag) atomicExch(p.error_flag, 2);
                                break;
                            }
                        }
                    }
                });
        }
        return;
    }
    constexpr int G = 32 / L;  // ratings per warp pass
    const int g = lane / L, l = lane % L;
    const int vecs = p.kp >> 2;
    const float lr = __ldg(p.lr);
    const int groups = kConsumerWarps * G;
    for (int it = 0;; ++it) {
        const int s = it % kStages;
        mbar_wait(&sm.full[s], (it / kStages) & 1);
        const int cnt = sm.count[s];
        if (cnt < 0) break;
        if (p.serial) {
            // strictly sequential replay (grid of one CTA): warp 0 / group 0 applies one rating
            // at a time; __syncwarp orders lane 0's bias store before the next rating's loads.
            if (warp == 0) {
                const bool ok = (g == 0);
                for (int j = 0; j < cnt; ++j) {
                    const cu2b_rating rt = sm.stage[s][j];
                    sgd_update_slots<L, V, 1, 0>(p, &rt, &ok, l, vecs, lr);
                    __syncwarp();
                }
            }
        } else {
            // all lanes of a warp run the same trip count (the shuffles are warp-wide)
            for (int base = warp * G; base < cnt; base += groups * UNR) {
                cu2b_rating rt[UNR];
                bool ok[UNR];
#pragma unroll
                for (int x = 0; x < UNR; ++x) {
                    const int j = base + x * groups + g;
                    ok[x] = j < cnt;
                    rt[x] = sm.stage[s][ok[x] ? j : 0];
                }
                sgd_update_slots<L, V, UNR, WMODE>(p, rt, ok, l, vecs, lr);
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (p.gate) {
                // Last consumer warp out publishes "chunk j of this segment is complete".
                // Release chain: each warp's row updates -> (acq_rel at CTA scope on the shared
                // counter) -> the last warp -> st.release.gpu on the gate word (cumulative).
                if (atom_add_acq_rel_cta_shared(&sm.done[s], 1) == kConsumerWarps - 1) {
                    sm.done[s] = 0;
                    const long long c = sm.chunk_id[s];
                    const long long seg = c / p.sv.chunks_per_seg;
                    const int j = (int)(c - seg * p.sv.chunks_per_seg);
                    st_release_gpu(p.gate + j, p.seg0 + (int)seg + 1);
                }
            }
            mbar_arrive(&sm.empty[s]);
        }
    }
}

}  // namespace cu2b
#endif  // CU2B_SGD_KERNELS_CUH_
