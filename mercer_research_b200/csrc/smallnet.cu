// smallnet.cu -- fused training step for rcn's canonical network family: one wide input layer (n_in = feature
// length, e.g. 784) into narrow sigmoid layers (every layer width <= 32, e.g. 784-30-10, main.rs:51-62).
//
// Reference: rcn/src/rcn.rs:260-314 (backprop), :176-223 (batch sum), :105-116 (forward), :152-157 (accuracy).
// At these sizes the generic tiled GEMMs of dense.cu are launch/latency bound (M = 30 or 10), so the whole
// backprop of a minibatch runs as three kernels with no shared-memory operand staging at all:
//
//   A  smallnet_fwd_bwd_kernel    one CTA per 8 samples.  z1 = W1 a0 on the f64 tensor path (DMMA.8x8x4): each of the
//                                 8 warps owns a K-range of the 784-deep contraction and streams its W1 / a0
//                                 fragments straight from L2 into registers; partial sums meet in 16 KB of shared
//                                 memory; the narrow layers, the output delta, the backward-data chain and the
//                                 batch statistics finish in registers/shared memory.  Writes a_l, delta_l.
//   B  smallnet_wgrad_kernel      dW1 = Delta1 A0^T split over (64-column group, K-split of the batch): DMMA again,
//                                 operands straight from L2; an extra CTA per K-split does db1 and the narrow
//                                 layers' dW/db.  Writes per-split partial gradient vectors.
//   C  smallnet_reduce_kernel     sums the K-split partials in a fixed order into the flat gradient buffer (the
//                                 all-reduce target) and finalises the batch statistics.
#include "smallnet.cuh"

namespace rcn {

__device__ __forceinline__ double sn_sigmoid(double z) { return 1.0 / (1.0 + exp(-z)); }  // rcn.rs:478-483

__device__ __forceinline__ void sn_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int SN_TB = 8;        // samples per CTA in kernel A (one DMMA n-fragment)
constexpr int SN_THREADS = 256;
constexpr int SN_WARPS = 8;

// ------------------------------------------------------------------------------------------------
// Kernel A
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SN_THREADS) smallnet_fwd_bwd_kernel(const __grid_constant__ SmallNetDesc d,
                                                                      const double* __restrict__ params,
                                                                      const double* __restrict__ feats, int B,
                                                                      const double* __restrict__ onehot,
                                                                      const int64_t* __restrict__ labels,
                                                                      double* __restrict__ acts, double* __restrict__ deltas,
                                                                      double* __restrict__ stats_partial, int backward) {
    __shared__ double zpart[SN_WARPS][32][SN_TB];           // per-warp partial z1 (16 KB)
    __shared__ double s_act[kSmallNetMaxLayers][SN_TB][33]; // a_l for this CTA's samples
    __shared__ double s_del[kSmallNetMaxLayers][SN_TB][33];
    __shared__ double s_cost[SN_TB];
    __shared__ unsigned long long s_hit[SN_TB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int s0 = blockIdx.x * SN_TB;
    const int L = d.n_in, R0 = d.rows[0];
    const int mf = (R0 + 7) >> 3;

    // ---- layer 0: z = W0 a0 on DMMA, K split across warps ------------------------------------------------------
    {
        const double* __restrict__ W0 = params + d.w_off[0];
        const int ksteps = (L + 3) >> 2;
        const int per = (ksteps + SN_WARPS - 1) / SN_WARPS;
        const int ks_begin = warp * per;
        const int ks_end = min(ksteps, ks_begin + per);
        double acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
        const int sample = s0 + g;
        const bool sample_ok = sample < B;
        const double* __restrict__ frow = feats + (size_t)(sample_ok ? sample : 0) * L;
        constexpr int U = 4;
        for (int ks = ks_begin; ks < ks_end; ks += U) {
            double af[U][4], bf[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = (ks + u) * 4 + t;
                const bool kok = (ks + u) < ks_end && k < L;
                bf[u] = (kok && sample_ok) ? __ldg(frow + k) : 0.0;           // B frag: row t (k), col g (sample)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = i * 8 + g;
                    af[u][i] = (kok && i < mf && m < R0) ? __ldg(W0 + (size_t)k * R0 + m) : 0.0;  // A frag: row g (m), col t (k)
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < mf) sn_dmma(acc[i][0], acc[i][1], af[u][i], bf[u]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // C frag: row g (m), cols 2t, 2t+1 (sample)
            zpart[warp][i * 8 + g][2 * t] = acc[i][0];
            zpart[warp][i * 8 + g][2 * t + 1] = acc[i][1];
        }
    }
    __syncthreads();

    const int n = tid >> 5;       // sample within the tile
    const int m = tid & 31;       // neuron
    const int sample = s0 + n;
    const bool live = sample < B;
    {
        double z = 0.0;
#pragma unroll
        for (int w = 0; w < SN_WARPS; ++w) z += zpart[w][m][n];
        if (m < R0) {
            z = z + params[d.b_off[0] + m];             // w * a + b      (rcn.rs:287)
            s_act[0][n][m] = sn_sigmoid(z);             // sigmoid(&z)    (rcn.rs:289)
        }
    }
    __syncthreads();
    // ---- narrow layers ---------------------------------------------------------------------------------------------
    for (int l = 1; l < d.n_layers; ++l) {
        const int R = d.rows[l], C = d.rows[l - 1];
        if (m < R) {
            const double* __restrict__ W = params + d.w_off[l];
            double z = 0.0;
            for (int k = 0; k < C; ++k) z = fma(W[(size_t)k * R + m], s_act[l - 1][n][k], z);
            z = z + params[d.b_off[l] + m];
            s_act[l][n][m] = sn_sigmoid(z);
        }
        __syncthreads();
    }
    const int last = d.n_layers - 1;
    const int RL = d.rows[last];
    // ---- activations out ---------------------------------------------------------------------------------------------
    {
        size_t off = 0;
        for (int l = 0; l < d.n_layers; ++l) {
            if (live && m < d.rows[l]) acts[off * B + (size_t)sample * d.rows[l] + m] = s_act[l][n][m];
            off += d.rows[l];
        }
    }
    if (!backward) return;
    // ---- output delta (rcn.rs:299) and batch statistics (rcn.rs:152-157) --------------------------------------------
    double y = 0.0;
    if (m < RL && live) y = onehot ? onehot[(size_t)sample * RL + m] : ((labels[sample] == (int64_t)m) ? 1.0 : 0.0);
    if (m < RL) {
        const double a = s_act[last][n][m];
        s_del[last][n][m] = (a - y) * (a * (1.0 - a));
    }
    if (m == 0) {
        double cost = 0.0;
        unsigned long long hit = 0;
        if (live) {
            double mx = s_act[last][n][0];
            for (int i = 1; i < RL; ++i) mx = fmax(mx, s_act[last][n][i]);
            bool ok = true;
            for (int i = 0; i < RL; ++i) {
                const double yi = onehot ? onehot[(size_t)sample * RL + i] : ((labels[sample] == (int64_t)i) ? 1.0 : 0.0);
                const double a = s_act[last][n][i];
                const double df = a - yi;
                cost += df * df;
                ok = ok && (((a == mx) ? 1.0 : 0.0) == yi);
            }
            cost *= 0.5;
            hit = ok ? 1ull : 0ull;
        }
        s_cost[n] = cost;
        s_hit[n] = hit;
    }
    __syncthreads();
    // ---- backward-data chain (rcn.rs:305-309) -------------------------------------------------------------------------
    for (int l = last - 1; l >= 0; --l) {
        const int R = d.rows[l], Ru = d.rows[l + 1];
        if (m < R) {
            const double* __restrict__ Wu = params + d.w_off[l + 1];  // Ru x R column-major: (k, m) at m*Ru + k
            double v = 0.0;
            for (int k = 0; k < Ru; ++k) v = fma(Wu[(size_t)m * Ru + k], s_del[l + 1][n][k], v);
            const double a = s_act[l][n][m];
            s_del[l][n][m] = v * (a * (1.0 - a));
        }
        __syncthreads();
    }
    {
        size_t off = 0;
        for (int l = 0; l < d.n_layers; ++l) {
            if (live && m < d.rows[l]) deltas[off * B + (size_t)sample * d.rows[l] + m] = s_del[l][n][m];
            off += d.rows[l];
        }
    }
    if (tid == 0) {
        double c = 0.0;
        unsigned long long h = 0;
        for (int i = 0; i < SN_TB; ++i) { c += s_cost[i]; h += s_hit[i]; }
        stats_partial[2 * blockIdx.x] = c;
        reinterpret_cast<unsigned long long*>(stats_partial)[2 * blockIdx.x + 1] = h;
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel B: per-K-split partial gradients.  blockIdx.x < col_groups: 64 columns of dW0 (one n-fragment per warp);
// blockIdx.x == col_groups: db0 and the narrow layers.  blockIdx.y = K-split (a range of `ksplit` samples).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SN_THREADS) smallnet_wgrad_kernel(const __grid_constant__ SmallNetDesc d,
                                                                    const double* __restrict__ feats,
                                                                    const double* __restrict__ acts,
                                                                    const double* __restrict__ deltas, int B, int ksplit,
                                                                    int col_groups, double* __restrict__ partial) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int L = d.n_in, R0 = d.rows[0];
    const int b_begin = blockIdx.y * ksplit;
    const int b_end = min(B, b_begin + ksplit);
    double* __restrict__ out = partial + (size_t)blockIdx.y * d.n_params;

    if ((int)blockIdx.x < col_groups) {
        const int mf = (R0 + 7) >> 3;
        const int col = blockIdx.x * 64 + warp * 8 + g;     // B frag column (feature index)
        const bool col_ok = col < L;
        double acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
        constexpr int U = 4;
        for (int kb = b_begin; kb < b_end; kb += 4 * U) {
            double af[U][4], bf[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int b = kb + u * 4 + t;
                const bool bok = b < b_end;
                bf[u] = (bok && col_ok) ? feats[(size_t)b * L + col] : 0.0;           // B frag: row t (sample), col g (feature)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = i * 8 + g;
                    af[u][i] = (bok && i < mf && m < R0) ? deltas[(size_t)b * R0 + m] : 0.0;  // A frag: row g (m), col t (sample)
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < mf) sn_dmma(acc[i][0], acc[i][1], af[u][i], bf[u]);
        }
        const int c0 = blockIdx.x * 64 + warp * 8 + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = i * 8 + g;
            if (m < R0) {
                if (c0 < L) out[d.w_off[0] + (size_t)c0 * R0 + m] = acc[i][0];
                if (c0 + 1 < L) out[d.w_off[0] + (size_t)(c0 + 1) * R0 + m] = acc[i][1];
            }
        }
        return;
    }
    // ---- db0 and the narrow layers: one thread per output, samples in index order ----------------------------------
    size_t act_off = 0;   // rows before layer l (activations/deltas are stored layer after layer, rows_l x B each)
    for (int l = 0; l < d.n_layers; ++l) {
        const int R = d.rows[l];
        const double* __restrict__ dl = deltas + act_off * B;
        for (int mrow = tid; mrow < R; mrow += SN_THREADS) {       // db_l = sum_b delta_l   (rcn.rs:302,309)
            double s = 0.0;
            for (int b = b_begin; b < b_end; ++b) s += dl[(size_t)b * R + mrow];
            out[d.b_off[l] + mrow] = s;
        }
        if (l >= 1) {                                              // dW_l = sum_b delta_l a_{l-1}^T   (rcn.rs:303,310)
            const int C = d.rows[l - 1];
            const double* __restrict__ ap = acts + (act_off - C) * B;
            for (int o = tid; o < R * C; o += SN_THREADS) {
                const int mrow = o % R, k = o / R;
                double s = 0.0;
                for (int b = b_begin; b < b_end; ++b) s = fma(dl[(size_t)b * R + mrow], ap[(size_t)b * C + k], s);
                out[d.w_off[l] + (size_t)k * R + mrow] = s;
            }
        }
        act_off += R;
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel C: fixed-order sum of the K-split partials -> flat gradient buffer; batch statistics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) smallnet_reduce_kernel(const double* __restrict__ partial, int splits, int n_params,
                                                             double* __restrict__ grads,
                                                             const double* __restrict__ stats_partial, int n_stat,
                                                             double* __restrict__ stats) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_params; i += gridDim.x * blockDim.x) {
        double s = partial[i];
        for (int p = 1; p < splits; ++p) s += partial[(size_t)p * n_params + i];
        grads[i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32 && stats) {
        // 32 lanes, each a contiguous chunk in order, then a fixed-order shuffle tree => deterministic
        const int per = (n_stat + 31) / 32;
        double c = 0.0;
        unsigned long long h = 0;
        for (int i = threadIdx.x * per; i < min(n_stat, (int)(threadIdx.x + 1) * per); ++i) {
            c += stats_partial[2 * i];
            h += reinterpret_cast<const unsigned long long*>(stats_partial)[2 * i + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_down_sync(0xffffffffu, c, o);
            h += __shfl_down_sync(0xffffffffu, h, o);
        }
        if (threadIdx.x == 0) {
            stats[0] = c;
            reinterpret_cast<unsigned long long*>(stats)[1] = h;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
bool smallnet_eligible(const SmallNetDesc& d) {
    if (d.n_layers < 1 || d.n_layers > kSmallNetMaxLayers) return false;
    for (int l = 0; l < d.n_layers; ++l)
        if (d.rows[l] < 1 || d.rows[l] > 32) return false;
    return d.n_in >= 1;
}

int smallnet_ksplit(size_t B) {
    size_t ks = (B + 31) / 32;
    ks = (ks + 3) / 4 * 4;
    if (ks < 64) ks = 64;
    return (int)ks;
}

int launch_smallnet_forward(const SmallNetDesc& d, const double* params, const double* feats, size_t B, double* acts,
                            cudaStream_t stream) {
    if (B == 0) return RCN_OK;
    RCN_LAUNCH("smallnet_fwd_bwd_kernel", stream,
               smallnet_fwd_bwd_kernel<<<cdiv(B, SN_TB), SN_THREADS, 0, stream>>>(d, params, feats, (int)B, nullptr, nullptr,
                                                                                  acts, nullptr, nullptr, 0));
    return RCN_OK;
}

int launch_smallnet_backprop(const SmallNetDesc& d, const double* params, const double* feats, size_t B,
                             const double* onehot, const int64_t* labels, double* acts, double* deltas, double* grads,
                             double* stats, DevBuf& workspace, cudaStream_t stream) {
    if (B == 0) return RCN_OK;
    if (B > 0x7fffffff / 64) return fail(RCN_ERR_INVALID, "batch too large for the fused small-network path");
    const int ksplit = smallnet_ksplit(B);
    const int splits = (int)cdiv(B, ksplit);
    const int n_tiles = (int)cdiv(B, SN_TB);
    // workspace: [splits][n_params] partial gradients | [n_tiles][2] statistics partials
    const size_t part_elems = (size_t)splits * d.n_params;
    RCN_TRY(workspace.reserve((part_elems + 2 * (size_t)n_tiles) * sizeof(double)));
    double* partial = workspace.as<double>();
    double* stats_partial = partial + part_elems;
    RCN_LAUNCH("smallnet_fwd_bwd_kernel", stream,
               smallnet_fwd_bwd_kernel<<<n_tiles, SN_THREADS, 0, stream>>>(d, params, feats, (int)B, onehot, labels, acts,
                                                                           deltas, stats_partial, 1));
    const int col_groups = (int)cdiv(d.n_in, 64);
    dim3 grid(col_groups + 1, splits);
    RCN_LAUNCH("smallnet_wgrad_kernel", stream,
               smallnet_wgrad_kernel<<<grid, SN_THREADS, 0, stream>>>(d, feats, acts, deltas, (int)B, ksplit, col_groups, partial));
    unsigned rgrid = cdiv(d.n_params, 256);
    if (rgrid > (unsigned)kNumSMs) rgrid = kNumSMs;
    RCN_LAUNCH("smallnet_reduce_kernel", stream,
               smallnet_reduce_kernel<<<rgrid, 256, 0, stream>>>(partial, splits, d.n_params, grads, stats_partial, n_tiles, stats));
    return RCN_OK;
}

}  // namespace rcn
