// dp.cuh -- data-parallel gradient exchange fused with the SGD update, over NVLink peer memory (see dp.cu).
#pragma once
#include "common.cuh"

namespace rcn {

constexpr int kDpMaxWorld = 16;
constexpr int kDpChunk = 256;      // parameters per CTA of the fused kernel (one per thread: every push is in flight at once)

// One rank's communication block (device memory, exported to the peers):
//   [ctrl: step | done_ctas | pad to 256 B][receive slots: 2 (step parity) x world x n doubles, sentinel-filled]
struct DpState {
    bool connected = false;
    int world = 1, rank = 0;
    size_t n = 0, n_chunks = 0, bytes = 0;
    void* block = nullptr;                 // this rank's block
    void* peers[kDpMaxWorld] = {};         // every rank's block as mapped into THIS process (peers[rank] == block)
    bool imported[kDpMaxWorld] = {};       // mapped with cudaIpcOpenMemHandle (must be closed)
};

// Handed to a gradient-producing kernel so that it can PUSH its final gradient values into the peers' receive slots
// itself (the exchange then overlaps that kernel's tail and the next launch); world == 1 means "do not push".
struct DpPush {
    char* peers[kDpMaxWorld];
    int world, rank;
    unsigned long long n;
};
DpPush dp_push_desc(const DpState& st);
constexpr size_t kDpCtrlBytesPub = 256;   // offset of the receive slots inside a communication block

// Device side of the push: slot [step parity][my rank][i] on every peer = v (the sentinel itself is never sent).
__device__ __forceinline__ void dp_push_value(const DpPush& dp, size_t par_off, size_t i, double v) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    if (bits == 0xFFFFFFFFFFFFFFFFull) bits = 0x7FF8000000000000ull;
    for (int q = 0; q < dp.world; ++q)
        if (q != dp.rank)
            reinterpret_cast<unsigned long long*>(dp.peers[q] + kDpCtrlBytesPub)[par_off + (size_t)dp.rank * dp.n + i] = bits;
}
// Control words at the head of a communication block.
struct DpCtrl {
    long long step;              // exchanges completed by this rank (device-side: CUDA-graph replayable)
    unsigned int done_ctas;      // ticket of the kernel that is finishing the current exchange
};
constexpr unsigned long long kDpSentinelBits = 0xFFFFFFFFFFFFFFFFull;

// Device side of the receive for ONE element: wait for every peer's value of element i in my own block, put the sentinel
// back and add the ranks in RANK ORDER (mine = `own`), so every replica computes the bit-identical global sum.
__device__ __forceinline__ double dp_receive_sum(const DpPush& dp, size_t par_off, size_t i, double own) {
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(dp.peers[dp.rank] + kDpCtrlBytesPub) + par_off;
    double s = 0.0;
    for (int q = 0; q < dp.world; ++q) {
        double v = own;
        if (q != dp.rank) {
            unsigned long long* p = slots + (size_t)q * dp.n + i;
            unsigned long long bits;
            do { asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(bits) : "l"(p) : "memory"); } while (bits == kDpSentinelBits);
            *p = kDpSentinelBits;
            v = __longlong_as_double((long long)bits);
        }
        s = (q == 0) ? v : s + v;
    }
    return s;
}
// The last CTA (of `total`) to finish an exchange publishes the step counter; call from ONE thread per CTA after the CTA's
// receives are done.  `step` is the value dp_push_parity_offset() derived the parity from.
__device__ __forceinline__ void dp_finish_step(const DpPush& dp, unsigned total_ctas) {
    DpCtrl* ctrl = reinterpret_cast<DpCtrl*>(dp.peers[dp.rank]);
    __threadfence();
    if (atomicAdd(&ctrl->done_ctas, 1u) == total_ctas - 1) {
        ctrl->done_ctas = 0;
        const long long step = *reinterpret_cast<volatile long long*>(&ctrl->step) + 1;
        *reinterpret_cast<volatile long long*>(&ctrl->step) = step;
        __threadfence();
    }
}

// parity offset of the step this rank is about to exchange (reads the device-side step counter of its own block)
__device__ __forceinline__ size_t dp_push_parity_offset(const DpPush& dp) {
    const long long step = *reinterpret_cast<volatile long long*>(dp.peers[dp.rank]) + 1;
    return (size_t)(step & 1) * dp.world * dp.n;
}

size_t dp_block_bytes(int world, size_t n);
int dp_alloc(DpState& st, int world, int rank, size_t n, cudaStream_t stream);
void dp_release(DpState& st);

// params[i] -= scale * sum_r grads_r[i], r ascending, identical on every rank; grads[i] <- the global sum.
// cursor / batch / n_samples: optional epoch cursor advanced like sgd_update_kernel does.
// already_pushed: the gradient kernel pushed this step's values itself (DpPush); the kernel then only receives.
int launch_dp_allreduce_sgd(const DpState& st, double* params, double* grads, double scale, cudaStream_t stream,
                            long long* cursor, long long batch, long long n_samples, const double* stats = nullptr,
                            double* stats_ring = nullptr, bool already_pushed = false, bool pipe = false);

}  // namespace rcn
