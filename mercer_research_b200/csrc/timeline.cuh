// timeline.cuh -- optional device-side launch timeline (profiling aid, compiled only with -DRCN_TIMELINE by
// profiles/build_phases.sh; the product library contains none of it).  Every instrumented kernel records, per launch,
// the earliest CTA start and the latest CTA end in %globaltimer nanoseconds, so the gaps between the kernels of a
// replayed CUDA graph -- which neither CUDA events nor ncu's serialised replay can see -- become visible.
#pragma once
#ifdef RCN_TIMELINE
namespace rcn_tl {
constexpr int kKernels = 4, kRing = 64;
static __device__ unsigned long long g_t0[kKernels][kRing], g_t1[kKernels][kRing];
static __device__ unsigned int g_seq[kKernels], g_done[kKernels];
__device__ __forceinline__ unsigned long long now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned begin(int k) {
    const unsigned seq = *(volatile unsigned*)&g_seq[k];
    atomicMin(&g_t0[k][seq % kRing], now());
    return seq;
}
__device__ __forceinline__ void end(int k, unsigned seq, unsigned nblocks) {
    atomicMax(&g_t1[k][seq % kRing], now());
    __threadfence();
    if (atomicAdd(&g_done[k], 1u) == nblocks - 1) {
        g_done[k] = 0;
        __threadfence();
        *(volatile unsigned*)&g_seq[k] = seq + 1;
    }
}
static int reset_host() {
    static unsigned long long ones[kKernels][kRing];
    memset(ones, 0xff, sizeof(ones));
    cudaMemcpyToSymbol(g_t0, ones, sizeof(ones));
    memset(ones, 0, sizeof(ones));
    cudaMemcpyToSymbol(g_t1, ones, sizeof(ones));
    unsigned z[kKernels] = {};
    cudaMemcpyToSymbol(g_seq, z, sizeof(z));
    cudaMemcpyToSymbol(g_done, z, sizeof(z));
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 4;
}
static int read_host(unsigned long long* out /* [2][kKernels][kRing] */, unsigned* seq /* [kKernels] */) {
    cudaMemcpyFromSymbol(out, g_t0, sizeof(g_t0));
    cudaMemcpyFromSymbol(out + kKernels * kRing, g_t1, sizeof(g_t1));
    return cudaMemcpyFromSymbol(seq, g_seq, sizeof(g_seq)) == cudaSuccess ? 0 : 4;
}
}  // namespace rcn_tl
#define RCN_TL_BEGIN(k) unsigned _tl_seq = 0; if (threadIdx.x == 0 && threadIdx.y == 0) _tl_seq = rcn_tl::begin(k)
#define RCN_TL_END(k) do { if (threadIdx.x == 0 && threadIdx.y == 0) rcn_tl::end(k, _tl_seq, gridDim.x * gridDim.y * gridDim.z); } while (0)
#else
#define RCN_TL_BEGIN(k) do { } while (0)
#define RCN_TL_END(k) do { } while (0)
#endif
