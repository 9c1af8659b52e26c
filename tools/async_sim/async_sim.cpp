// tools/async_sim/async_sim.cpp -- CPU ANALYSIS TOOL. Not product code, not the parity oracle.
//
// A model of asynchronous (Hogwild) SGD with a bounded number of updates in flight, used to study
// the stability limit DESIGN.md 6.1 measured on 8 x B200 and to evaluate remedies before GPU time
// is spent on them. One "rank" has `inflight` lane groups; a group owns one user at a time and
// applies that user's draws of the round in order with the user's row held privately (exact
// sequential semantics on the user side, as mf_sgd_user_rounds / mf_sgd_user_runs do). The groups
// start their updates round-robin. An update reads its item row and item bias when it starts,
// computes its steps from what it read, and the item-side steps (the L2 atomic adds) land only
// after `stale_factor x busy groups` further updates have started: every update that reads the
// same item in between works from a row that does not contain them yet. stale_factor = 1 is a
// read -> add window of one update time per group; 2 models a kernel that requests the next
// update's item row before the current update's adds have landed (the look-ahead loads of
// mf_sgd_user_runs). inflight = 1, stale_factor = 1 is plain sequential SGD in user-major order.
//
// DSGD (G > 1): users and items are cut into G blocks; in sub-epoch s rank g applies those draws
// of its users that fall into item block (g + s) mod G. Ranks never share a row, so they are
// simulated one after the other.
//
// item_scale (optional): per-item factor m <= 1 for the items whose load exceeds the budget; `mode` says how
// it is used (0 scale both item-side steps, 1 thin both, 2 scale the bias step only, 3 thin the bias step
// only, 4 thin the row step only).
//
// The update itself follows mf_sequential.cu:114-141 (right-hand sides use the pre-update values).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {
void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3], k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// same counter layout as the product sampler: one uniform draw per (user, iteration)
inline int sample_index(uint32_t seed, uint32_t user, uint32_t iteration, int lo, int hi) {
    uint32_t ctr[4] = {user, iteration, 0u, 0x53474431u}, key[2] = {seed, 0x43553242u}, r[4];
    philox4x32_10(ctr, key, r);
    return lo + (int)(((uint64_t)r[0] * (uint64_t)(uint32_t)(hi - lo)) >> 32);
}
double rmse(int rows, const int *indptr, const int *indices, const float *data, const float *P, const float *Q,
            const float *ub, const float *ib, float mu, int k) {
    double sse = 0;
    for (int u = 0; u < rows; ++u)
        for (int j = indptr[u]; j < indptr[u + 1]; ++j) {
            const float *p = P + (size_t)u * k, *q = Q + (size_t)indices[j] * k;
            float pred = mu + ub[u] + ib[indices[j]];
            for (int f = 0; f < k; ++f) pred += q[f] * p[f];
            const double e = (double)data[j] - pred;
            sse += e * e;
        }
    const long n = indptr[rows];
    return n ? std::sqrt(sse / (double)n) : 0.0;
}
struct Group {
    int user = -1;  // -1 = idle
    int pos = 0;    // next iteration offset inside the round
};
}  // namespace

extern "C" int async_sim_train(int rows, int cols, const int *indptr, const int *indices, const float *data,
                               const int *te_indptr, const int *te_indices, const float *te_data, float *P, float *Q,
                               float *ub, float *ib, float mu, int k, float lr, float reg, int seed,
                               int total_iterations, int check_error, int G, const int *user_block,
                               const int *item_block, int round_iters, int inflight, float stale_factor, const float *item_scale, int mode,
                               double *log /* rows of {iteration, test_rmse, max unseen steps at a read} */, int log_cap) {
    (void)cols;
    int n_log = 0;
    std::vector<std::vector<int>> users_of(G);
    for (int u = 0; u < rows; ++u)
        if (indptr[u + 1] > indptr[u]) users_of[G > 1 ? user_block[u] : 0].push_back(u);
    std::vector<int> draw((size_t)rows * round_iters);  // rating index per (user, offset in round)
    std::vector<Group> groups((size_t)inflight);
    // FIFO of item-side steps that have been computed but have not landed yet
    const size_t ring_cap = (size_t)(stale_factor * inflight) + 2;
    std::vector<int> pend_item(ring_cap);
    std::vector<long> pend_start(ring_cap);
    std::vector<float> pend_step(ring_cap * (k + 1));
    std::vector<int> unseen((size_t)cols, 0);  // steps in the FIFO per item
    long max_conc = 0;
    for (int it0 = 0; it0 < total_iterations; it0 += round_iters) {
        const int T = std::min(round_iters, total_iterations - it0);
        for (int u = 0; u < rows; ++u) {
            const int lo = indptr[u], hi = indptr[u + 1];
            if (hi > lo)
                for (int t = 0; t < T; ++t) draw[(size_t)u * round_iters + t] = sample_index((uint32_t)seed, (uint32_t)u, (uint32_t)(it0 + t), lo, hi);
        }
        for (int s = 0; s < G; ++s)
            for (int g = 0; g < G; ++g) {
                const int b = (g + s) % G;
                const std::vector<int> &mine = users_of[g];
                size_t next_user = 0;
                // advance a group to its next draw inside block b (claiming new users as needed)
                auto advance = [&](Group &gr) {
                    for (;;) {
                        if (gr.user >= 0) {
                            while (gr.pos < T && G > 1 && item_block[indices[draw[(size_t)gr.user * round_iters + gr.pos]]] != b) ++gr.pos;
                            if (gr.pos < T) return true;
                        }
                        if (next_user >= mine.size()) { gr.user = -1; return false; }
                        gr.user = mine[next_user++];
                        gr.pos = 0;
                    }
                };
                int busy = 0;
                for (Group &gr : groups) { gr.user = -1; gr.pos = 0; busy += advance(gr); }
                size_t head = 0, tail = 0;  // ring indices (monotonic)
                long started = 0;
                auto land = [&](long threshold) {
                    while (head < tail && started - pend_start[head % ring_cap] >= threshold) {
                        const size_t e = head % ring_cap;
                        const int i = pend_item[e];
                        const float *step = pend_step.data() + e * (k + 1);
                        float *q = Q + (size_t)i * k;
                        for (int f = 0; f < k; ++f) q[f] += step[f];
                        ib[i] += step[k];
                        --unseen[i];
                        ++head;
                    }
                };
                while (busy > 0) {
                    for (Group &gr : groups) {
                        if (gr.user < 0) continue;
                        land(std::max(1L, (long)(stale_factor * busy)));
                        const int u = gr.user, j = draw[(size_t)u * round_iters + gr.pos], i = indices[j];
                        float *p = P + (size_t)u * k;
                        const float *q = Q + (size_t)i * k;
                        float pred = mu + ub[u] + ib[i];
                        for (int f = 0; f < k; ++f) pred += q[f] * p[f];
                        const float err = data[j] - pred;
                        const float m = item_scale ? item_scale[i] : 1.0f;
                        if (unseen[i] > max_conc) max_conc = unseen[i];
                        // what the item side of this draw takes: mode 0 scales both steps by m; 1 applies both for
                        // a fraction m of the draws (thinning); 2 / 3 scale / thin the bias step only; 4 thins the
                        // row step only
                        float m_row = 1.0f, m_bias = 1.0f;
                        if (m < 1.0f) {
                            bool keep = true;
                            if (mode == 1 || mode == 3 || mode == 4) {
                                uint32_t ctr[4] = {(uint32_t)u, (uint32_t)(it0 + gr.pos), 1u, 0x53474431u}, key[2] = {(uint32_t)seed, 0x43553242u}, r[4];
                                philox4x32_10(ctr, key, r);
                                keep = (float)(r[0] >> 8) * (1.0f / 16777216.0f) < m;
                            }
                            if (mode == 0) m_row = m_bias = m;
                            else if (mode == 1) m_row = m_bias = keep ? 1.0f : 0.0f;
                            else if (mode == 2) m_bias = m;
                            else if (mode == 3) m_bias = keep ? 1.0f : 0.0f;
                            else if (mode == 4) m_row = keep ? 1.0f : 0.0f;
                        }
                        if (m_row == 0.0f && m_bias == 0.0f) {  // nothing to add: only the user side moves
                            for (int f = 0; f < k; ++f) p[f] = p[f] + lr * (err * q[f] - reg * p[f]);
                            ub[u] = ub[u] + lr * (err - reg * ub[u]);
                            ++started;
                            ++gr.pos;
                            if (!advance(gr)) --busy;
                            continue;
                        }
                        const size_t e = tail % ring_cap;
                        float *step = pend_step.data() + e * (k + 1);
                        for (int f = 0; f < k; ++f) {
                            const float p_old = p[f], q_old = q[f];
                            p[f] = p_old + lr * (err * q_old - reg * p_old);
                            step[f] = m_row * (lr * (err * p_old - reg * q_old));
                        }
                        step[k] = m_bias * (lr * (err - reg * ib[i]));
                        ub[u] = ub[u] + lr * (err - reg * ub[u]);
                        pend_item[e] = i;
                        pend_start[e] = started;
                        ++unseen[i];
                        ++tail;
                        ++started;
                        ++gr.pos;
                        if (!advance(gr)) --busy;
                    }
                }
                started += (long)ring_cap * 4;
                land(1);  // end of the sub-epoch: everything lands before the block moves on
            }
        const int done = it0 + T;
        if (done % check_error == 0 || done == total_iterations) {
            const double r = rmse(rows, te_indptr, te_indices, te_data, P, Q, ub, ib, mu, k);
            if (n_log < log_cap) { log[3 * n_log] = done; log[3 * n_log + 1] = r; log[3 * n_log + 2] = (double)max_conc; }
            ++n_log;
            max_conc = 0;
            if (!std::isfinite(r) || r > 100.0) break;  // diverged
        }
    }
    return n_log;
}
