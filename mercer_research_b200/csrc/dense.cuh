// dense.cuh -- fully-connected layer kernels (f64): forward, backward-data, backward-weight, SGD step.
// Reference: rcn/src/rcn.rs:105-116 (classify_test), :260-314 (backprop), :176-223 (train_batch).
#pragma once
#include "common.cuh"

namespace rcn {

struct OzakiWorkspace;   // ozaki.cuh: scratch of the tcgen05 integer-slice GEMM (NULL = never take that path)

// env RCN_CUDA_GEMM: (unset) auto = tcgen05 integer slices for large dense layers, DMMA otherwise; "dmma"; "tc" forces the
// tcgen05 path for every GEMM; "simt" selects the plain SIMT cross-check kernels.
enum GemmImpl { GEMM_AUTO = 0, GEMM_DMMA = 1, GEMM_SIMT = 2, GEMM_TC = 3 };
GemmImpl gemm_impl();

// A_out (M x N) = sigmoid(W (M x K) * A_in (K x N) + b 1^T)                         rcn.rs:113 / :287-289
// If delta_out != NULL (last layer): delta_out = (A_out - Y) .* A_out .* (1 - A_out)  rcn.rs:299
//   with Y given either as one-hot matrix (M x N) or as labels (N).
int launch_dense_forward(const double* W, const double* b, const double* A_in, size_t M, size_t K, size_t N,
                         double* A_out, double* delta_out, const double* onehot, const int64_t* labels,
                         cudaStream_t stream, OzakiWorkspace* oz = nullptr);

// delta_out (M x N) = (W_up^T (M x K) * delta_up (K x N)) .* A (M x N) .* (1 - A)    rcn.rs:305-309
// W_up is stored K x M (rows = upper layer width).
int launch_dense_backward_data(const double* W_up, const double* delta_up, const double* A, size_t M, size_t K,
                               size_t N, double* delta_out, cudaStream_t stream, OzakiWorkspace* oz = nullptr);

// dW (M x N) = delta (M x Kb) * A_prev (N x Kb)^T, db (M) = delta * 1                 rcn.rs:302-303,309-310 summed
// over the batch (rcn.rs:190-205).  workspace: split-K partials.
int launch_dense_backward_weight(const double* delta, const double* A_prev, size_t M, size_t N, size_t Kb, double* dW,
                                 double* db, DevBuf& workspace, ReduceScratch& rs, cudaStream_t stream, OzakiWorkspace* oz = nullptr);

// out[i] = sum_p partials[p*n + i], p ascending (deterministic split-K combine)
int launch_reduce_splits(const double* partials, int splits, size_t n, double* out, cudaStream_t stream);
// db[m] = sum_n delta[m + n*M]  (delta M x N column-major), fixed summation order
int launch_bias_grad(const double* delta, size_t M, size_t N, double* db, ReduceScratch& rs, cudaStream_t stream);

// params -= scale * grads   (rcn.rs:210-222, scale = eta / batch formed first)
// cursor (optional): device-side position of the epoch walk, advanced by `batch` with chunks_exact wrap-around.
// stats / stats_ring (optional, with cursor): the step's {cost, hits} pair is copied to stats_ring + 2 * (cursor / batch)
// (pinned host memory in rcn_cuda_train_epoch_host) before the cursor advances.
int launch_sgd_update(double* params, const double* grads, size_t n, double scale, cudaStream_t stream,
                      long long* cursor = nullptr, long long batch = 0, long long n_samples = 0, const double* stats = nullptr,
                      double* stats_ring = nullptr);

// labels[b] = argmax_i acts[i, b], last maximal element wins (rcn.rs:92-97)
int launch_argmax_last(const double* acts, size_t n, size_t B, int64_t* labels, cudaStream_t stream);

// stats[0] = sum_b 0.5*|a_b - y_b|^2 (as double), stats[1] = bit pattern of uint64 hit count (rcn.rs:153-157)
int launch_batch_stats(const double* acts, size_t n, size_t B, const double* onehot, const int64_t* labels,
                       double* stats_dev, ReduceScratch& rs, cudaStream_t stream);

}  // namespace rcn
