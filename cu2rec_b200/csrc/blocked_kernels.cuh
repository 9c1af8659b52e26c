// Deterministic conflict-free mode (no counterpart in the reference; ordering idea as in
// LIBMF / cuMF_SGD block scheduling, made static so that the result is reproducible).
//
// Users are cut into B contiguous blocks, items into B contiguous blocks. Round s in [0, B)
// holds the B rating buckets (b, (b + s) mod B): no two of them share a user block or an item
// block, so they can run concurrently without touching a common row. Inside a bucket one lane
// group applies the ratings strictly in order. The result therefore equals a sequential replay
// in round-major / user-block / original order, bit for bit, whatever the launch geometry.
// One kernel launch per round; the kernel boundary is the inter-round barrier.
#ifndef CU2B_BLOCKED_KERNELS_CUH_
#define CU2B_BLOCKED_KERNELS_CUH_

#include "sgd_kernels.cuh"

namespace cu2b {

struct BlockedParams {
    const cu2b_rating *sched;  // ratings sorted by (round, user block, original position)
    const int *bucket_ptr;     // this round's B + 1 offsets into sched
    int B;
    SgdParams model;           // P, Q, biases, hyper-parameters (stream fields unused)
};

template <int L, int V>
__global__ void __launch_bounds__(256)
mf_sgd_blocked_round(const BlockedParams bp) {
    constexpr int G = 32 / L;
    const int lane = threadIdx.x & 31, g = lane / L, l = lane % L;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_groups = ((gridDim.x * blockDim.x) >> 5) * G;
    const int vecs = bp.model.kp >> 2;
    const float lr = __ldg(bp.model.lr);
    // all groups of a warp advance together (the dot-product shuffles are warp-wide)
    for (int b0 = warp_global * G; b0 < bp.B; b0 += n_groups) {
        const int b = b0 + g;
        int j = 0, end = 0;
        if (b < bp.B) { j = __ldg(bp.bucket_ptr + b); end = __ldg(bp.bucket_ptr + b + 1); }
        while (__any_sync(0xffffffffu, j < end)) {
            const bool ok = j < end;
            cu2b_rating rt;
            rt.user = 0; rt.item = 0; rt.rating = 0.f;
            if (ok) {
                rt.user = __ldg(&bp.sched[j].user);
                rt.item = __ldg(&bp.sched[j].item);
                rt.rating = __ldg(&bp.sched[j].rating);
            }
            sgd_update_slots<L, V, 1, 0>(bp.model, &rt, &ok, l, vecs, lr);
            __syncwarp();  // lane 0's bias stores are visible to the group's next rating
            ++j;
        }
    }
}

}  // namespace cu2b
#endif  // CU2B_BLOCKED_KERNELS_CUH_
