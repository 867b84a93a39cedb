// dp.cu -- the exchange step of rcn's minibatch gradient sum, scaled from host threads to the GPUs of one box.
//
// Reference: rcn/src/rcn.rs:190-205 sums the per-sample gradients of a minibatch over rayon worker threads behind two
// mutexes, then rcn.rs:210-222 applies W -= (eta/B) * sum.  Data-parallel on N GPUs the same sum runs over ranks:
// every rank holds gradient SUMS of its shard in the flat buffer [dW0|db0|dW1|db1|...], the ranks' buffers are added
// and every replica applies the identical update.
//
// For the latency-bound regime (rcn's canonical 784-30-10 network has 23 860 parameters = 190 KB) a library
// all-reduce costs more than the whole training step, so exchange and update are ONE kernel over NVLink / NVSwitch
// peer memory:
//   * every rank owns a communication block (cudaMalloc, exported to the peers with cudaIpcGetMemHandle or, inside
//     one process, mapped with cudaDeviceEnablePeerAccess) holding 2 x world receive slots of n doubles, pre-filled
//     with a SENTINEL bit pattern (all ones -- a NaN payload no arithmetic produces);
//   * each thread PUSHES its elements of the local gradient into slot [step parity][my rank] of every peer (plain
//     8-byte stores to peer addresses travel over NVLink; an 8-byte store is single-copy atomic, so the value itself
//     is the arrival flag -- the low-latency protocol NCCL calls LL, without a separate flag, fence or barrier);
//   * the same thread then polls its elements of the peers' slots in ITS OWN block until they differ from the sentinel,
//     puts the sentinel back, adds the ranks in RANK ORDER (so all replicas compute bit-identical sums and stay
//     bit-identical to each other) and applies the update.  Latency = one NVLink traversal + a few L2 polls.
// Threads are independent: no grid-wide, CTA-wide or end-of-kernel barrier.  The receive slots are double buffered by
// step parity: a peer can only push step s+2 into the parity that held step s after it has received every rank's
// step s+1 push, which a rank issues only after its step-s kernel (reads and sentinel resets included) has finished.
// The step counter lives in device memory, so a captured CUDA graph replays the kernel unchanged.  Large gradient
// buffers (e.g. the 268 MB of the 4096-wide MLP) are bandwidth-bound and stay on NCCL (trainer.py picks by size).
#include "dp.cuh"

#include <cstdlib>

namespace rcn {

constexpr size_t kDpCtrlBytes = 256;

size_t dp_block_bytes(int world, size_t n) { return kDpCtrlBytes + 2 * (size_t)world * n * sizeof(double); }

int dp_alloc(DpState& st, int world, int rank, size_t n, cudaStream_t stream) {
    if (world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world) return fail(RCN_ERR_INVALID, "bad world / rank (%d / %d)", world, rank);
    dp_release(st);
    st.world = world; st.rank = rank; st.n = n;
    st.n_chunks = (n + kDpChunk - 1) / kDpChunk;
    st.bytes = dp_block_bytes(world, n);
    RCN_CUDA_TRY(cudaMalloc(&st.block, st.bytes));
    alloc_generation().fetch_add(1, std::memory_order_relaxed);   // peer pointers are baked into captured step graphs too
    RCN_CUDA_TRY(cudaMemsetAsync(st.block, 0xFF, st.bytes, stream));      // every receive slot = sentinel
    RCN_CUDA_TRY(cudaMemsetAsync(st.block, 0, kDpCtrlBytes, stream));
    RCN_CUDA_TRY(cudaStreamSynchronize(stream));
    st.peers[rank] = st.block;
    return RCN_OK;
}

void dp_release(DpState& st) {
    for (int q = 0; q < kDpMaxWorld; ++q) {
        if (st.imported[q] && st.peers[q]) cudaIpcCloseMemHandle(st.peers[q]);
        st.peers[q] = nullptr; st.imported[q] = false;
    }
    if (st.block) { cudaFree(st.block); alloc_generation().fetch_add(1, std::memory_order_relaxed); }
    st.block = nullptr; st.connected = false; st.world = 1; st.rank = 0;
    cudaGetLastError();
}

static_assert(kDpCtrlBytes == kDpCtrlBytesPub, "slot offset mismatch");

DpPush dp_push_desc(const DpState& st) {
    DpPush d{};
    d.world = st.connected ? st.world : 1;
    d.rank = st.rank;
    d.n = st.n;
    for (int q = 0; q < st.world && q < kDpMaxWorld; ++q) d.peers[q] = (char*)st.peers[q];
    return d;
}

struct DpPeers { char* p[kDpMaxWorld]; };

constexpr unsigned long long kDpSentinel = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long kDpQuietNaN = 0x7FF8000000000000ull;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int WORLD>   // 0 = runtime world size
__global__ void __launch_bounds__(256) dp_allreduce_sgd_kernel(const DpPeers peers, int world_rt, int rank, size_t n,
                                                               double* __restrict__ params, double* __restrict__ grads,
                                                               double scale, long long* __restrict__ cursor, long long batch,
                                                               long long n_samples, const double* __restrict__ stats,
                                                               double* __restrict__ stats_ring, int already_pushed, int pipe) {
    // programmatic dependent launch (no-ops unless the host asked for it): the next kernel's launch may proceed under this
    // one; this one waits for the gradient kernel to complete before it touches global memory
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int world = WORLD ? WORLD : world_rt;
    char* self = peers.p[rank];
    DpCtrl* ctrl = reinterpret_cast<DpCtrl*>(self);
    const long long step = *reinterpret_cast<volatile long long*>(&ctrl->step) + 1;
    const size_t par_off = (size_t)(step & 1) * world * n;
    const int tid = threadIdx.x;
    const size_t lo = (size_t)blockIdx.x * kDpChunk;
    const size_t hi = lo + kDpChunk < n ? lo + kDpChunk : n;
    constexpr int E = kDpChunk / 256;            // elements per thread
    double g[E];
    // ---- push my elements of the local gradient sums into my slot on every peer (stores travel over NVLink) ---------
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const size_t i = lo + tid + (size_t)e * 256;
        g[e] = 0.0;
        if (i < hi) {
            g[e] = grads[i];
            if (!already_pushed) {
                unsigned long long bits = (unsigned long long)__double_as_longlong(g[e]);
                if (bits == kDpSentinel) bits = kDpQuietNaN;          // a NaN either way; never send the sentinel itself
                for (int q = 0; q < world; ++q)
                    if (q != rank)
                        reinterpret_cast<unsigned long long*>(peers.p[q] + kDpCtrlBytes)[par_off + (size_t)rank * n + i] = bits;
            }
        }
    }
    // ---- receive: poll my elements of every peer's slot in my own block, restore the sentinel, add in rank order ------
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(self + kDpCtrlBytes) + par_off;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const size_t i = lo + tid + (size_t)e * 256;
        if (i < hi) {
            double s = 0.0;
            for (int q = 0; q < world; ++q) {
                double v;
                if (q == rank) v = g[e];
                else {
                    unsigned long long* p = slots + (size_t)q * n + i;
                    unsigned long long bits;
                    while ((bits = ld_volatile_u64(p)) == kDpSentinel) { }
                    *p = kDpSentinel;
                    v = __longlong_as_double((long long)bits);
                }
                s = (q == 0) ? v : s + v;
            }
            grads[i] = s;                        // the gradient buffer ends up holding the global sum, like an all-reduce
            params[i] = sgd_apply(params[i], scale, s);   // W -= (eta / B) * sum   (rcn.rs:214,221)
        }
    }
    // ---- the last CTA to finish advances the device-side step counter (and the epoch cursor) ----------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&ctrl->done_ctas, 1u);
        if (t == gridDim.x - 1) {
            ctrl->done_ctas = 0;
            *reinterpret_cast<volatile long long*>(&ctrl->step) = step;
            if (cursor) {
                if (stats_ring) {
                    double* dst = stats_ring + 2 * (*cursor / batch);
                    dst[0] = stats[0];
                    dst[1] = stats[1];
                }
                long long cc = *cursor + batch;
                if (cc + batch > n_samples) cc = 0;
                *cursor = cc;
                if (pipe) cursor[1] += 1;   // pipelined epoch mode: training steps done
            }
            __threadfence();
        }
    }
}

int launch_dp_allreduce_sgd(const DpState& st, double* params, double* grads, double scale, cudaStream_t stream,
                            long long* cursor, long long batch, long long n_samples, const double* stats, double* stats_ring,
                            bool already_pushed, bool pipe) {
    if (!st.connected) return fail(RCN_ERR_STATE, "data-parallel group is not connected");
    if (st.n == 0) return RCN_OK;
    DpPeers pp{};
    for (int q = 0; q < st.world; ++q) pp.p[q] = (char*)st.peers[q];
    const unsigned grid = (unsigned)st.n_chunks;
    static const bool pdl = []() { const char* e = getenv("RCN_CUDA_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    const int world_i = st.world, rank_i = st.rank, pushed_i = already_pushed ? 1 : 0, pipe_i = pipe ? 1 : 0;
    const size_t n_i = st.n;
#define RCN_DP_LAUNCH(W)                                                                                                       \
    RCN_LAUNCH("dp_allreduce_sgd_kernel", stream,                                                                              \
               cudaLaunchKernelEx(&cfg, dp_allreduce_sgd_kernel<W>, pp, world_i, rank_i, n_i, params, grads, scale, cursor, batch, \
                                  n_samples, stats, stats_ring, pushed_i, pipe_i))
    switch (st.world) {
        case 2: RCN_DP_LAUNCH(2); break;
        case 4: RCN_DP_LAUNCH(4); break;
        case 8: RCN_DP_LAUNCH(8); break;
        default: RCN_DP_LAUNCH(0); break;
    }
#undef RCN_DP_LAUNCH
    return RCN_OK;
}

}  // namespace rcn
