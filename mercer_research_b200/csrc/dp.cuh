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

size_t dp_block_bytes(int world, size_t n);
int dp_alloc(DpState& st, int world, int rank, size_t n, cudaStream_t stream);
void dp_release(DpState& st);

// params[i] -= scale * sum_r grads_r[i], r ascending, identical on every rank; grads[i] <- the global sum.
// cursor / batch / n_samples: optional epoch cursor advanced like sgd_update_kernel does.
int launch_dp_allreduce_sgd(const DpState& st, double* params, double* grads, double scale, cudaStream_t stream,
                            long long* cursor, long long batch, long long n_samples, const double* stats = nullptr,
                            double* stats_ring = nullptr);

}  // namespace rcn
