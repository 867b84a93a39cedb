// timeline.cuh -- device-side launch timeline of the fused step kernels (rcn_cuda_timeline_enable / _read).
// When a model's timeline is enabled its step kernels receive a pointer to this block and record, per launch, the
// earliest CTA start and the latest CTA end in %globaltimer nanoseconds: the durations of, and the gaps between, the
// kernels of a REPLAYED CUDA graph -- which neither CUDA events around eager launches nor ncu's serialised replay can
// see -- become measurable (bench.py takes the roofline durations from here).  Disabled = null pointer = one predicated-off
// branch on a kernel parameter.
#pragma once
#include <cstdint>

namespace rcn {

constexpr int kTlKernels = 4;    // 0 kernel A, 1 kernel B, 2 exchange / update kernel, 3 spare
constexpr int kTlRing = 64;      // launches remembered per kernel

struct Timeline {
    unsigned long long t0[kTlKernels][kTlRing];   // min over CTAs of the start stamp (reset to ~0)
    unsigned long long t1[kTlKernels][kTlRing];   // max over CTAs of the end stamp (reset to 0)
    unsigned int seq[kTlKernels];                 // launches completed
    unsigned int done[kTlKernels];                // CTA ticket of the running launch
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// call from ONE thread per CTA
__device__ __forceinline__ unsigned tl_begin(Timeline* tl, int k) {
    const unsigned seq = *(volatile unsigned*)&tl->seq[k];
    atomicMin(&tl->t0[k][seq % kTlRing], tl_now());
    return seq;
}
__device__ __forceinline__ void tl_end(Timeline* tl, int k, unsigned seq, unsigned nblocks) {
    atomicMax(&tl->t1[k][seq % kTlRing], tl_now());
    __threadfence();
    if (atomicAdd(&tl->done[k], 1u) == nblocks - 1) {
        tl->done[k] = 0;
        // the slot this ring position will hold kTlRing launches from now must start clean
        tl->t0[k][(seq + 1) % kTlRing] = ~0ull;
        tl->t1[k][(seq + 1) % kTlRing] = 0ull;
        __threadfence();
        *(volatile unsigned*)&tl->seq[k] = seq + 1;
    }
}
#define RCN_TL_BEGIN(tl, k) unsigned _tl_seq = 0; if ((tl) && threadIdx.x == 0 && threadIdx.y == 0) _tl_seq = ::rcn::tl_begin((tl), (k))
#define RCN_TL_END(tl, k) do { if ((tl) && threadIdx.x == 0 && threadIdx.y == 0) ::rcn::tl_end((tl), (k), _tl_seq, gridDim.x * gridDim.y * gridDim.z); } while (0)
#endif

}  // namespace rcn
