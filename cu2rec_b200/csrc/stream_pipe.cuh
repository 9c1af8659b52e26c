// TMA-fed rating-stream pipeline shared by the SGD and loss kernels (sm_100a).
//
// A CTA owns a ring of kStages shared-memory buffers. A dedicated producer warp (one elected
// lane) claims chunks of the rating stream and pulls each one into the ring with a single
// 1-D bulk async copy (cp.async.bulk.shared::cluster.global -> SASS UBLKCP) that signals an
// mbarrier with its byte count; consumer warps wait on that "full" barrier, read the
// (user, item, rating) triplets from shared memory, and release the buffer through an
// "empty" barrier (one arrival per consumer warp). No __syncthreads in the steady state.
//
// A stream is a sequence of `n_seg` segments of `seg_len` ratings stored with a pitch of
// `seg_pitch` ratings (a multiple of 4, so that every chunk starts 16-byte aligned); each
// segment is cut into `chunks_per_seg` chunks of `chunk` ratings (the last one shorter).
// In the per-user sampling mode a segment is one reference iteration (sgd.cu: one update per
// user), otherwise the stream is a single flat segment.
#ifndef CU2B_STREAM_PIPE_CUH_
#define CU2B_STREAM_PIPE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "cu2b.h"

namespace cu2b {

constexpr int kStages = 4;
constexpr int kChunkMax = 512;                         // ratings per stage (6 KB)
constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;    // + 1 producer warp

struct StreamView {
    const cu2b_rating *base;
    long long seg_pitch;   // ratings between segment starts (multiple of 4)
    int seg_len;           // valid ratings per segment
    int chunk;             // ratings per chunk (multiple of 4, <= kChunkMax)
    int chunks_per_seg;
    long long num_chunks;  // n_seg * chunks_per_seg
};

struct __align__(16) StreamSmem {
    cu2b_rating stage[kStages][kChunkMax];
    unsigned long long full[kStages];
    unsigned long long empty[kStages];
    long long chunk_id[kStages];  // -1 => end of stream
    int count[kStages];
    int done[kStages];            // consumer warps finished with the stage (gated mode)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int atom_add_acq_rel_cta_shared(int *p, int v) {
    int old;
    asm volatile("atom.acq_rel.cta.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
// Call with all threads of the CTA before the role split.
__device__ __forceinline__ void pipe_init(StreamSmem &sm) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 2);  // TMA issue (+ tx bytes) and the ordering gate
            mbar_init(&sm.empty[s], kConsumerWarps);
            sm.done[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// Producer loop (one lane). next_chunk() yields the next chunk index for this CTA or -1;
// gate(seg, j, cnt) blocks until chunk j of segment `seg` may be processed (no-op when
// ungated). The triplet copy is issued BEFORE the gate is polled so its latency overlaps the
// wait; consumers are released by the second arrival on the stage's full barrier.
template <typename NextChunk, typename Gate>
__device__ __forceinline__ void pipe_produce(StreamSmem &sm, const StreamView &sv,
                                             NextChunk next_chunk, Gate gate) {
    for (int it = 0;; ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&sm.empty[s], ((it / kStages) - 1) & 1);
        const long long c = next_chunk();
        sm.chunk_id[s] = c;
        if (c < 0) {
            sm.count[s] = -1;
            mbar_arrive(&sm.full[s]);
            mbar_arrive(&sm.full[s]);
            break;
        }
        const long long seg = c / sv.chunks_per_seg;
        const int j = (int)(c - seg * sv.chunks_per_seg);
        const int left = sv.seg_len - j * sv.chunk;
        const int cnt = left < sv.chunk ? left : sv.chunk;
        sm.count[s] = cnt;
        const uint32_t bytes = (uint32_t)(((cnt + 3) & ~3) * (int)sizeof(cu2b_rating));
        mbar_arrive_expect_tx(&sm.full[s], bytes);
        tma_load_1d(&sm.stage[s][0], sv.base + seg * sv.seg_pitch + (long long)j * sv.chunk, bytes,
                    &sm.full[s]);
        gate(seg, j, cnt);
        mbar_arrive(&sm.full[s]);
    }
}

}  // namespace cu2b
#endif  // CU2B_STREAM_PIPE_CUH_
