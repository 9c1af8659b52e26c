// oracle/mf_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain single-threaded CPU restatement of cu2rec's matrix-factorisation hot path
// (reference: nickgreenquist/cu2rec, paths below are relative to its repository root).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library. The product (libcu2b.so, bin/mf) never links or calls it.
//
// Parity pinning: this restatement is checked (tests/test_oracle_pins.py) against
//   * every golden value the reference's own tests hold for this path
//     (tests/test_loss.cu:90 "74.0", tests/test_loss.cu:137-138 all-ones mae==rmse==1,
//      tests/test_util.cu:28-31,43,123-125,170-172, tests/test_config.cu:14-15),
//   * outputs of the reference itself, compiled unmodified from /root/reference into
//     oracle/_ref/ by oracle/Makefile (ref_harness: initialize_normal_array, readCSV,
//     read_config; mf_cpu: final RMSE distribution; on the GPU box: loss_kernel,
//     total_loss_kernel, sgd_update, train()).
//
// Two arithmetic "flavours" of the single update / prediction are provided:
//   ORC_FLAVOUR_REF    : the op order of matrix_factorization/mf_sequential.cu:114-141
//                        (serial ascending-f dot product, unfused mul/add);
//   ORC_FLAVOUR_KERNEL : the op order of our CUDA update kernels (per-lane fmaf partial sums,
//                        xor-butterfly reduction over L lanes; update steps with the learning
//                        rate folded in, one multiply + one fma each -- see orc_sgd_update_one).
//                        Used for the bit-exact check of the deterministic mode.
//   ORC_FLAVOUR_LOSS_KERNEL : same, with the lane layout of the loss kernel (up to four float4
//                        per lane so that several ratings share a warp).
// Build with -ffp-contract=off so the compiler introduces no FMAs of its own.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

extern "C" {

enum { ORC_FLAVOUR_REF = 0, ORC_FLAVOUR_KERNEL = 1, ORC_FLAVOUR_LOSS_KERNEL = 2 };

typedef struct {
    int32_t user;
    int32_t item;
    float rating;
} orc_triplet;

typedef struct {
    int n_factors;
    float learning_rate;
    float P_reg, Q_reg, user_bias_reg, item_bias_reg;
    int is_train;  // 0 => Q and item_bias frozen (predict.cu:105 intent); 1 => both trained;
                   // 2 => only item_bias, 3 => only the Q row (the two halves of our DSGD item-step thinning)
} orc_hyper;

// ---------------------------------------------------------------------------------------
// Initialisation. Follows util.cu:124-144: mt19937(seed) -> normal_distribution<float>(mean,
// stddev / n_factors), sequential fill. Uses the same libstdc++ as the reference build.
// ---------------------------------------------------------------------------------------
void orc_init_normal(float *out, int size, int n_factors, float mean, float stddev, int seed) {
    std::mt19937 gen(seed);
    std::normal_distribution<float> dist(mean, stddev / n_factors);
    for (int i = 0; i < size; ++i) out[i] = dist(gen);
}

// ---------------------------------------------------------------------------------------
// Config file. Follows config.cu:7-13: nine whitespace separated tokens
//   cur_iterations total_iterations n_factors learning_rate seed P_reg Q_reg ub_reg ib_reg
// A short file leaves the remaining fields untouched (operator>> stops at first failure).
// Returns the number of fields parsed.
// ---------------------------------------------------------------------------------------
int orc_read_config(const char *path, int *cur_it, int *total_it, int *n_factors, float *lr,
                    int *seed, float *P_reg, float *Q_reg, float *ub_reg, float *ib_reg) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int n = 0;
    do {
        if (fscanf(f, "%d", cur_it) != 1) break; ++n;
        if (fscanf(f, "%d", total_it) != 1) break; ++n;
        if (fscanf(f, "%d", n_factors) != 1) break; ++n;
        if (fscanf(f, "%f", lr) != 1) break; ++n;
        if (fscanf(f, "%d", seed) != 1) break; ++n;
        if (fscanf(f, "%f", P_reg) != 1) break; ++n;
        if (fscanf(f, "%f", Q_reg) != 1) break; ++n;
        if (fscanf(f, "%f", ub_reg) != 1) break; ++n;
        if (fscanf(f, "%f", ib_reg) != 1) break; ++n;
    } while (0);
    fclose(f);
    return n;
}

// ---------------------------------------------------------------------------------------
// Ratings CSV. Follows util.cu:17-45: skip one header line, then rows of
// "int <char> int <char> float"; ids are 1-based in the file and 0-based in memory;
// rows = max userId, cols = max itemId, global_bias = (float)(sum / count) with a double sum.
// Two-call protocol: pass out == NULL to count.
// ---------------------------------------------------------------------------------------
long orc_read_csv(const char *path, orc_triplet *out, long cap, int *rows, int *cols,
                  float *global_bias) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int c;
    int skipped = 0;  // util.cu:29 ignore(1000, '\n')
    while (skipped < 1000 && (c = fgetc(f)) != EOF) {
        ++skipped;
        if (c == '\n') break;
    }
    long n = 0;
    int max_row = 0, max_col = 0;
    double sum = 0.0;
    int u, i;
    char d1, d2;
    float r;
    while (fscanf(f, "%d %c %d %c %f", &u, &d1, &i, &d2, &r) == 5) {
        if (out && n < cap) {
            out[n].user = u - 1;
            out[n].item = i - 1;
            out[n].rating = r;
        }
        if (u > max_row) max_row = u;
        if (i > max_col) max_col = i;
        sum += r;
        ++n;
    }
    fclose(f);
    *rows = max_row;
    *cols = max_col;
    *global_bias = (float)(sum / (1.0 * (double)n));
    return n;
}

// ---------------------------------------------------------------------------------------
// CSR build. Follows util.cu:152-179 (createSparseMatrix) / mf_sequential.cu:42-58: input grouped
// by ascending user; a missing user repeats the previous indptr entry. indptr must hold
// rows+1 ints. Returns the number of indptr entries written.
// ---------------------------------------------------------------------------------------
int orc_build_csr(const orc_triplet *ratings, long n, int rows, int *indptr, int *indices,
                  float *data) {
    int written = 0;
    int last_user = -1;
    for (long k = 0; k < n; ++k) {
        while (last_user != ratings[k].user) {
            if (written <= rows) indptr[written] = (int)k;
            ++written;
            ++last_user;
        }
        indices[k] = ratings[k].item;
        data[k] = ratings[k].rating;
    }
    // trailing users with no ratings + the closing entry
    while (written <= rows) indptr[written++] = (int)n;
    return written;
}

// ---------------------------------------------------------------------------------------
// Prediction. REF: util.cu:199-204 / mf_sequential.cu:125-127 (serial). KERNEL: lane layout.
// ---------------------------------------------------------------------------------------
static void kernel_layout(int k, int *L, int *V) {
    int kp = (k + 3) & ~3;  // rows are padded to a multiple of 4 floats
    int vecs = kp / 4;      // float4 per row
    int l = 1;
    while (l < vecs && l < 32) l <<= 1;
    *L = l;
    *V = (vecs + l - 1) / l;
}

// The loss kernel packs more ratings into a warp: up to 4 float4 per lane.
static void loss_layout(int k, int *L, int *V) {
    int vecs = ((k + 3) & ~3) / 4;
    int need = (vecs + 3) / 4;  // lanes needed at 4 float4 per lane
    int l = 1;
    while (l < need) l <<= 1;
    *L = l;
    *V = (vecs + l - 1) / l;
}

static float dot_layout(const float *p, const float *q, int k, int L, int V) {
    float acc[32];
    for (int l = 0; l < L; ++l) {
        float a = 0.0f;
        for (int v = 0; v < V; ++v) {
            int base = v * 4 * L + 4 * l;
            for (int e = 0; e < 4; ++e) {
                int f = base + e;
                float pv = f < k ? p[f] : 0.0f;
                float qv = f < k ? q[f] : 0.0f;
                a = fmaf(pv, qv, a);
            }
        }
        acc[l] = a;
    }
    for (int off = L / 2; off >= 1; off >>= 1) {
        float nxt[32];
        for (int l = 0; l < L; ++l) nxt[l] = acc[l] + acc[l ^ off];
        memcpy(acc, nxt, sizeof(float) * L);
    }
    return acc[0];
}

static float dot_kernel_flavour(const float *p, const float *q, int k) {
    int L, V;
    kernel_layout(k, &L, &V);
    return dot_layout(p, q, k, L, V);
}

float orc_predict(const float *p, const float *q, int k, float ub, float ib, float mu,
                  int flavour) {
    if (flavour == ORC_FLAVOUR_REF) {
        float pred = mu + ub + ib;
        for (int f = 0; f < k; ++f) pred += q[f] * p[f];
        return pred;
    }
    float base = (mu + ub) + ib;
    if (flavour == ORC_FLAVOUR_LOSS_KERNEL) {
        int L, V;
        loss_layout(k, &L, &V);
        return base + dot_layout(p, q, k, L, V);
    }
    return base + dot_kernel_flavour(p, q, k);
}

// ---------------------------------------------------------------------------------------
// One SGD update. Follows mf_sequential.cu:114-141 (== sgd.cu:40-72 without the early-bird
// gate and with Q / item_bias updated in place): right-hand sides use the pre-update p, q,
// ub, ib. Returns the rating error.
// ---------------------------------------------------------------------------------------
float orc_sgd_update_one(float *p, float *q, float *ub, float *ib, float rating, float mu,
                         const orc_hyper *h, int flavour) {
    const int k = h->n_factors;
    const float lr = h->learning_rate;
    float ub0 = *ub, ib0 = *ib;
    const bool train_row = h->is_train == 1 || h->is_train == 3, train_bias = h->is_train == 1 || h->is_train == 2;
    float err = rating - orc_predict(p, q, k, ub0, ib0, mu, flavour);
    if (flavour != ORC_FLAVOUR_REF) {
        // Op order of the CUDA kernels (sgd_kernels.cuh, sgd_step): the learning rate is folded
        // into the error and the regularisers once (a = lr*err, c = lr*reg), every step is one
        // multiply + one fused multiply-add, and the row takes `old + step` with one rounding
        // (an FADD in registers or the L2 atomic add). Algebraically mf_sequential.cu:129-141.
        const float a = lr * err;
        const float cP = lr * h->P_reg, cQ = lr * h->Q_reg;
        const float cU = lr * h->user_bias_reg, cI = lr * h->item_bias_reg;
        for (int f = 0; f < k; ++f) {
            float p_old = p[f];
            float q_old = q[f];
            p[f] = p_old + fmaf(a, q_old, -(cP * p_old));
            if (train_row) q[f] = q_old + fmaf(a, p_old, -(cQ * q_old));
        }
        *ub = ub0 + fmaf(-cU, ub0, a);
        if (train_bias) *ib = ib0 + fmaf(-cI, ib0, a);
        return err;
    }
    for (int f = 0; f < k; ++f) {
        float p_old = p[f];
        float q_old = q[f];
        float gp = err * q_old;
        float rp = h->P_reg * p_old;
        p[f] = p_old + lr * (gp - rp);
        if (train_row) {
            float gq = err * p_old;
            float rq = h->Q_reg * q_old;
            q[f] = q_old + lr * (gq - rq);
        }
    }
    *ub = ub0 + lr * (err - h->user_bias_reg * ub0);
    if (train_bias) *ib = ib0 + lr * (err - h->item_bias_reg * ib0);
    return err;
}

// Sequential replay of an explicit update stream (the comparator for the deterministic mode).
void orc_sgd_apply_stream(const orc_triplet *stream, long n, float *P, float *Q, float *user_bias,
                          float *item_bias, float mu, const orc_hyper *h, int flavour) {
    const int k = h->n_factors;
    for (long t = 0; t < n; ++t) {
        int u = stream[t].user, i = stream[t].item;
        orc_sgd_update_one(P + (size_t)u * k, Q + (size_t)i * k, user_bias + u, item_bias + i,
                           stream[t].rating, mu, h, flavour);
    }
}

// ---------------------------------------------------------------------------------------
// Residuals and metrics. loss.cu:19-35 (error[i] = data[i] - prediction) and
// loss.cu:58-69,185-200 (double accumulation of err^2 / |err|, result cast to float).
// ---------------------------------------------------------------------------------------
void orc_residuals(int rows, const int *indptr, const int *indices, const float *data,
                   const float *P, const float *Q, const float *user_bias, const float *item_bias,
                   float mu, int k, float *err, int flavour) {
    for (int u = 0; u < rows; ++u) {
        const float *p = P + (size_t)u * k;
        float ub = user_bias[u];
        for (int j = indptr[u]; j < indptr[u + 1]; ++j) {
            int it = indices[j];
            err[j] = data[j] - orc_predict(p, Q + (size_t)it * k, k, ub, item_bias[it], mu, flavour);
        }
    }
}

void orc_error_metrics(const float *err, long n, float *mae, float *rmse) {
    double sae = 0.0, sse = 0.0;
    for (long j = 0; j < n; ++j) {
        sae += (double)fabsf(err[j]);
        sse += (double)err[j] * (double)err[j];
    }
    *mae = (float)(sae / (double)n);
    *rmse = (float)sqrt(sse / (double)n);
}

void orc_loss(int rows, const int *indptr, const int *indices, const float *data, const float *P,
              const float *Q, const float *user_bias, const float *item_bias, float mu, int k,
              float *mae, float *rmse, double *sse_out, double *sae_out, int flavour) {
    double sae = 0.0, sse = 0.0;
    long n = indptr[rows];
    for (int u = 0; u < rows; ++u) {
        const float *p = P + (size_t)u * k;
        float ub = user_bias[u];
        for (int j = indptr[u]; j < indptr[u + 1]; ++j) {
            int it = indices[j];
            float e = data[j] - orc_predict(p, Q + (size_t)it * k, k, ub, item_bias[it], mu, flavour);
            sae += (double)fabsf(e);
            sse += (double)e * (double)e;
        }
    }
    *mae = (float)(sae / (double)n);
    *rmse = (float)sqrt(sse / (double)n);
    if (sse_out) *sse_out = sse;
    if (sae_out) *sae_out = sae;
}

// The same sums for the loss checks of orc_train on large problems (at-size parity tests): users are cut into
// contiguous chunks, one thread per chunk, chunk sums added in chunk order (deterministic; differs from the
// sequential double sum by rounding of order 1e-16). Small problems keep the sequential loop above.
static void loss_chunked(int rows, const int *indptr, const int *indices, const float *data, const float *P,
                         const float *Q, const float *user_bias, const float *item_bias, float mu, int k,
                         float *mae, float *rmse, int flavour) {
    const long n = indptr[rows];
    unsigned hw = std::thread::hardware_concurrency();
    const int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
    if (n < (1L << 21) || nt == 1) {
        orc_loss(rows, indptr, indices, data, P, Q, user_bias, item_bias, mu, k, mae, rmse, nullptr, nullptr, flavour);
        return;
    }
    std::vector<double> sse(nt, 0.0), sae(nt, 0.0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            // chunk bounds by rating count so that the threads finish together
            const long lo_n = n * t / nt, hi_n = n * (t + 1) / nt;
            const int u0 = (int)(std::lower_bound(indptr, indptr + rows + 1, (int)lo_n) - indptr);
            const int u1 = t + 1 == nt ? rows : (int)(std::lower_bound(indptr, indptr + rows + 1, (int)hi_n) - indptr);
            double a = 0.0, b = 0.0;
            for (int u = u0; u < u1; ++u) {
                const float *p = P + (size_t)u * k;
                const float ub = user_bias[u];
                for (int j = indptr[u]; j < indptr[u + 1]; ++j) {
                    const int it = indices[j];
                    const float e = data[j] - orc_predict(p, Q + (size_t)it * k, k, ub, item_bias[it], mu, flavour);
                    a += (double)fabsf(e);
                    b += (double)e * (double)e;
                }
            }
            sae[t] = a;
            sse[t] = b;
        });
    for (auto &x : th) x.join();
    double a = 0.0, b = 0.0;
    for (int t = 0; t < nt; ++t) { a += sae[t]; b += sse[t]; }
    *mae = (float)(a / (double)n);
    *rmse = (float)sqrt(b / (double)n);
}

// mf_sequential.cu:146-201 accumulates the same sums in float; kept for comparing against the
// compiled mf_cpu's printed lines.
void orc_loss_float_acc(int rows, const int *indptr, const int *indices, const float *data,
                        const float *P, const float *Q, const float *user_bias,
                        const float *item_bias, float mu, int k, float *mae, float *rmse) {
    float sae = 0.0f, sse = 0.0f;
    long n = indptr[rows];
    for (int u = 0; u < rows; ++u) {
        const float *p = P + (size_t)u * k;
        float ub = user_bias[u];
        for (int j = indptr[u]; j < indptr[u + 1]; ++j) {
            int it = indices[j];
            float e = data[j] - orc_predict(p, Q + (size_t)it * k, k, ub, item_bias[it], mu,
                                            ORC_FLAVOUR_REF);
            sae += fabsf(e);
            sse += e * e;
        }
    }
    *mae = sae / n;
    *rmse = sqrtf(sse / n);
}

// ---------------------------------------------------------------------------------------
// Per-user sampler. Distribution of sgd.cu:27-37: every user with >=1 rating draws one of its
// ratings uniformly per iteration. The reference draws from a per-user XORWOW curandState
// (sgd.cu:11-16,36); we replace the 48-byte stateful generator with the stateless counter
// based Philox4x32-10 (Salmon et al., SC'11), counter = (user, iteration, 0, TAG),
// key = (seed, KEY1), and map the first output word to [lo, hi) by multiply-shift.
// Integer arithmetic only => the CUDA sampler must match this bit for bit.
// ---------------------------------------------------------------------------------------
#define ORC_PHILOX_TAG 0x53474431u  /* "SGD1" */
#define ORC_PHILOX_KEY1 0x43553242u /* "CU2B" */

void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0;
        k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline int sample_index(uint32_t seed, uint32_t user, uint32_t iteration, int lo, int hi) {
    uint32_t ctr[4] = {user, iteration, 0u, ORC_PHILOX_TAG};
    uint32_t key[2] = {seed, ORC_PHILOX_KEY1};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    uint32_t n = (uint32_t)(hi - lo);
    return lo + (int)(((uint64_t)r[0] * (uint64_t)n) >> 32);
}

// Emits, for iterations [iter0, iter0+n_iter), one triplet per active user in ascending user
// order. Returns the number of triplets written (n_iter * n_active).
long orc_sample_per_user(int rows, const int *indptr, const int *indices, const float *data,
                         int seed, int iter0, int n_iter, orc_triplet *out) {
    long w = 0;
    for (int t = 0; t < n_iter; ++t) {
        for (int u = 0; u < rows; ++u) {
            int lo = indptr[u], hi = indptr[u + 1];
            if (lo == hi) continue;  // sgd.cu:35
            int j = sample_index((uint32_t)seed, (uint32_t)u, (uint32_t)(iter0 + t), lo, hi);
            out[w].user = u;
            out[w].item = indices[j];
            out[w].rating = data[j];
            ++w;
        }
    }
    return w;
}

// ---------------------------------------------------------------------------------------
// per_rating sampler (an extension of ours; the reference only samples per user, sgd.cu:27-37): update q of a run
// applies rating perm_e(q mod nnz) of pass e = q / nnz, perm_e a 4-round Feistel network on 2h bits with cycle
// walking. Restated here from the definition in include/cu2b.h / sgd_kernels.cuh so that the tests can compare
// the CUDA stream with an independent evaluation.
// ---------------------------------------------------------------------------------------
static inline uint32_t orc_feistel_round(uint32_t x, uint32_t key) {
    uint32_t v = (x ^ key) * 0x9E3779B1u;
    v ^= v >> 15;
    v *= 0x85EBCA77u;
    v ^= v >> 13;
    return v;
}
unsigned long long orc_rating_permutation(unsigned long long j, unsigned long long n, uint32_t seed, uint32_t pass) {
    int bits = 1;
    while (bits < 64 && (1ULL << bits) < n) ++bits;
    const int h = (bits + 1) / 2;
    const uint32_t mask = h >= 32 ? 0xffffffffu : ((1u << h) - 1u);
    unsigned long long x = j;
    do {
        uint32_t l = (uint32_t)(x >> h) & mask, r = (uint32_t)x & mask;
        for (uint32_t round = 0; round < 4; ++round) {
            const uint32_t t = l ^ (orc_feistel_round(r, seed ^ (pass * 0x632BE5ABu) ^ (round * 0xB5297A4Du + 0x68E31DA4u)) & mask);
            l = r;
            r = t;
        }
        x = ((unsigned long long)l << h) | r;
    } while (x >= n);
    return x;
}
// triplets (CSR order) of updates [first, first + count)
void orc_sample_per_rating(int rows, const int *indptr, const int *indices, const float *data, int seed, long long first,
                           long long count, orc_triplet *out) {
    const unsigned long long n = (unsigned long long)indptr[rows];
    std::vector<int> user_of((size_t)n);
    for (int u = 0; u < rows; ++u)
        for (int j = indptr[u]; j < indptr[u + 1]; ++j) user_of[j] = u;
    for (long long i = 0; i < count; ++i) {
        const unsigned long long q = (unsigned long long)(first + i);
        const unsigned long long r = orc_rating_permutation(q % n, n, (uint32_t)seed, (uint32_t)(q / n));
        out[i].user = user_of[r];
        out[i].item = indices[r];
        out[i].rating = data[r];
    }
}

// ---------------------------------------------------------------------------------------
// Whole training loop. Update loop: mf_sequential.cu:102-143 (one sampled rating per user per
// iteration, Q / item_bias in place). Check cadence: training.cu:118 == mf_sequential.cu:146.
// Patience / learning-rate decay: training.cu:103,129,146-155 (only when use_decay != 0;
// mf_sequential ignores it). Sampling: orc_sample_per_user above instead of
// mf_sequential.cu:109-112 (random_device per update, inclusive upper bound -- SURVEY A8).
// log rows: {iteration(1-based), train_mae, train_rmse, test_mae, test_rmse, lr_after}.
// Returns the number of log rows written.
// ---------------------------------------------------------------------------------------
typedef struct {
    int iteration;
    float train_mae, train_rmse, test_mae, test_rmse, learning_rate;
} orc_log_row;

int orc_train(int rows, const int *tr_indptr, const int *tr_indices, const float *tr_data,
              int te_rows, const int *te_indptr, const int *te_indices, const float *te_data,
              float *P, float *Q, float *user_bias, float *item_bias, float mu, orc_hyper *h,
              int seed, int iter0, int total_iterations, int check_error, float patience,
              float lr_decay, int use_decay, int flavour, orc_log_row *log, int log_cap) {
    const int k = h->n_factors;
    int n_log = 0;
    float validation_rmse = std::numeric_limits<float>::max(), last_validation_rmse;
    int current_patience = (int)patience;  // training.cu:103
    for (int i = 0; i < total_iterations; ++i) {
        for (int u = 0; u < rows; ++u) {
            int lo = tr_indptr[u], hi = tr_indptr[u + 1];
            if (lo == hi) continue;
            int j = sample_index((uint32_t)seed, (uint32_t)u, (uint32_t)(iter0 + i), lo, hi);
            int it = tr_indices[j];
            orc_sgd_update_one(P + (size_t)u * k, Q + (size_t)it * k, user_bias + u,
                               item_bias + it, tr_data[j], mu, h, flavour);
        }
        if ((i + 1) % check_error == 0 || i == 0 || (i + 1) % total_iterations == 0) {
            float tr_mae, tr_rmse, te_mae, te_rmse;
            loss_chunked(rows, tr_indptr, tr_indices, tr_data, P, Q, user_bias, item_bias, mu, k,
                         &tr_mae, &tr_rmse, flavour);
            loss_chunked(te_rows, te_indptr, te_indices, te_data, P, Q, user_bias, item_bias, mu, k,
                         &te_mae, &te_rmse, flavour);
            last_validation_rmse = validation_rmse;
            validation_rmse = te_rmse;
            if (use_decay) {
                if (last_validation_rmse < validation_rmse) current_patience--;
                if (current_patience <= 0) {
                    current_patience = (int)patience;
                    h->learning_rate *= lr_decay;
                }
            }
            if (n_log < log_cap) {
                log[n_log].iteration = i + 1;
                log[n_log].train_mae = tr_mae;
                log[n_log].train_rmse = tr_rmse;
                log[n_log].test_mae = te_mae;
                log[n_log].test_rmse = te_rmse;
                log[n_log].learning_rate = h->learning_rate;
            }
            ++n_log;
        }
    }
    return n_log;
}

// ---------------------------------------------------------------------------------------
// Deterministic block schedule (no reference counterpart; it restates the ordering contract
// of our conflict-free mode so that a sequential replay can be compared with it):
// users are cut into B contiguous blocks of ceil(rows/B), items into B blocks of ceil(cols/B);
// round s in [0,B) holds the blocks (b, (b+s) mod B); the canonical sequential order is
// round-major, then user block, then original position. Writes the permutation of [0,n).
// ---------------------------------------------------------------------------------------
void orc_block_schedule_order(const orc_triplet *coo, long n, int rows, int cols, int B,
                              long *order) {
    int ubs = (rows + B - 1) / B, ibs = (cols + B - 1) / B;
    std::vector<std::vector<long>> buckets((size_t)B * B);
    for (long t = 0; t < n; ++t) {
        int ub = coo[t].user / ubs, ib = coo[t].item / ibs;
        int s = ((ib - ub) % B + B) % B;
        buckets[(size_t)s * B + ub].push_back(t);
    }
    long w = 0;
    for (size_t b = 0; b < buckets.size(); ++b)
        for (long t : buckets[b]) order[w++] = t;
}

// ---------------------------------------------------------------------------------------
// Batched predict. predict.cu:17-29 (score = mu + b_u + b_i + sum_f Q_i[f]*P_u[f], serial) and
// predict.cu:49-63 (drop rated items, sort high to low) for every user; ties by ascending item.
// ---------------------------------------------------------------------------------------
void orc_predict_topk(int rows, int cols, int k, const float *P, const float *Q, const float *user_bias,
                      const float *item_bias, float mu, const int *ex_indptr, const int *ex_indices, int ex_rows,
                      int topk, int32_t *out_items, float *out_scores) {
    std::vector<char> rated((size_t)cols);
    std::vector<std::pair<float, int>> sc;
    for (int u = 0; u < rows; ++u) {
        std::fill(rated.begin(), rated.end(), 0);
        if (ex_indptr && u < ex_rows)
            for (int j = ex_indptr[u]; j < ex_indptr[u + 1]; ++j) rated[ex_indices[j]] = 1;
        sc.clear();
        for (int i = 0; i < cols; ++i) {
            if (rated[i]) continue;
            sc.push_back({orc_predict(P + (size_t)u * k, Q + (size_t)i * k, k, user_bias[u], item_bias[i], mu, ORC_FLAVOUR_REF), i});
        }
        const int take = (int)std::min<size_t>((size_t)topk, sc.size());
        std::partial_sort(sc.begin(), sc.begin() + take, sc.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &b) {
            return a.first > b.first || (a.first == b.first && a.second < b.second);
        });
        for (int r = 0; r < topk; ++r) {
            out_items[(size_t)u * topk + r] = r < take ? sc[r].second : -1;
            out_scores[(size_t)u * topk + r] = r < take ? sc[r].first : NAN;
        }
    }
}

}  // extern "C"
