// smallnet.cuh -- fused training step for narrow sigmoid networks (every layer width <= 32), see smallnet.cu.
#pragma once
#include "common.cuh"

namespace rcn {

constexpr int kSmallNetMaxLayers = 4;

struct SmallNetDesc {
    int n_layers;
    int n_in;                          // input width of layer 0 (feature length)
    int rows[kSmallNetMaxLayers];      // layer widths
    int w_off[kSmallNetMaxLayers];     // offsets into the flat [W0|b0|W1|b1|...] buffer
    int b_off[kSmallNetMaxLayers];
    int n_params;
};

bool smallnet_eligible(const SmallNetDesc& d);

// acts: sum(rows) x B (layer after layer, rows_l x B column-major each), same layout for deltas.
int launch_smallnet_forward(const SmallNetDesc& d, const double* params, const double* feats, size_t B, double* acts,
                            cudaStream_t stream);
// Full backprop of a minibatch: fills acts, deltas, the flat gradient buffer (batch sums) and stats[0..1]
// (quadratic cost, hit count as uint64 bits).
int launch_smallnet_backprop(const SmallNetDesc& d, const double* params, const double* feats, size_t B,
                             const double* onehot, const int64_t* labels, double* acts, double* deltas, double* grads,
                             double* stats, DevBuf& workspace, cudaStream_t stream);

}  // namespace rcn
