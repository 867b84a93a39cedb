// ext_oracle.cpp -- CPU definitions of the EXTENSION ops (TEST INFRASTRUCTURE ONLY, same rules as rcn_oracle.cpp).
//
// BASELINE.json's north_star names operations that the reference (jtstrader/mercer-research, crate `rcn`) does
// NOT implement: learned multi-channel convolution forward / backward-data /
// backward-weight, average pooling, pooling backward, softmax + cross-entropy.  There is no reference code, test
// or golden vector for any of them: **PARITY UNPINNED**.  This file states their textbook definitions in plain f64
// loops so the CUDA kernels have something independent to be checked against; conventions are chosen to extend
// the reference's own:
//   * convolution is a cross-correlation (no kernel flip), like Convolve2D::convolve_2d (kernel.rs:110-194);
//     Padding::Same zero-pads kh/2, kw/2 on every side (the un-quirked 3x3 behaviour pinned by kernel.rs:434-441),
//     Padding::None is the valid convolution (kernel.rs:171-192).
//   * pooling is the 2x2 / stride-2 window of Pool2D::pool_2d (kernel.rs:245-349): Same pads one zero row / column
//     at the bottom / right for odd sizes, None drops the trailing row / column; window order
//     [(0,0),(0,1),(1,0),(1,1)], max ties -> LAST maximal element (kernel.rs:273-281), argmax index = 2*dy+dx.
//     Average divides by 4 always (the zero padding counts), the natural completion of kernel.rs:253-261.
//   * activations / targets of the dense head stay column-major (classes x B), sample b contiguous.
// Tensors with channels are NHWC f64: x[((b*H + y)*W + x)*C + c].
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace {
enum : int { PAD_NONE = 0, PAD_SAME = 1 };
enum : int { POOL_AVERAGE = 0, POOL_MAX = 1 };
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

inline double act_fwd(double z, int act) {
    if (act == ACT_RELU) return z >= 0.0 ? z : 0.0;              // kernel.rs:209-216 rule (x >= 0 ? x : 0)
    if (act == ACT_SIGMOID) return 1.0 / (1.0 + std::exp(-z));   // rcn.rs:478-483
    return z;
}
// derivative expressed through the stored OUTPUT y = act(z)
inline double act_bwd_from_output(double y, int act) {
    if (act == ACT_RELU) return y > 0.0 ? 1.0 : 0.0;
    if (act == ACT_SIGMOID) return y * (1.0 - y);                 // rcn.rs:490-492
    return 1.0;
}
}  // namespace

extern "C" {

// y[b,oy,ox,co] = act( bias[co] + sum_{ky,kx,ci} x[b, oy+ky-ph, ox+kx-pw, ci] * w[co,ky,kx,ci] )
int ext_conv2d_forward(const double* x, size_t B, size_t H, size_t W, size_t Ci, const double* w, const double* bias,
                       size_t Co, size_t kh, size_t kw, int padding, int act, double* y) {
    if (kh == 0 || kw == 0 || kh > H || kw > W) return 1;
    if (padding == PAD_SAME && (kh % 2 == 0 || kw % 2 == 0)) return 1;
    const long ph = padding == PAD_SAME ? (long)kh / 2 : 0, pw = padding == PAD_SAME ? (long)kw / 2 : 0;
    const size_t Ho = padding == PAD_SAME ? H : H - kh + 1, Wo = padding == PAD_SAME ? W : W - kw + 1;
    for (size_t b = 0; b < B; ++b)
        for (size_t oy = 0; oy < Ho; ++oy)
            for (size_t ox = 0; ox < Wo; ++ox)
                for (size_t co = 0; co < Co; ++co) {
                    double acc = 0.0;
                    for (size_t ky = 0; ky < kh; ++ky)
                        for (size_t kx = 0; kx < kw; ++kx) {
                            const long iy = (long)oy + (long)ky - ph, ix = (long)ox + (long)kx - pw;
                            if (iy < 0 || ix < 0 || iy >= (long)H || ix >= (long)W) continue;
                            const double* xp = x + ((b * H + iy) * W + ix) * Ci;
                            const double* wp = w + ((co * kh + ky) * kw + kx) * Ci;
                            for (size_t ci = 0; ci < Ci; ++ci) acc += xp[ci] * wp[ci];
                        }
                    if (bias) acc += bias[co];
                    y[((b * Ho + oy) * Wo + ox) * Co + co] = act_fwd(acc, act);
                }
    return 0;
}

// dz = dy .* act'(y)   (elementwise, n values)
void ext_activation_backward(const double* y, const double* dy, size_t n, int act, double* dz) {
    for (size_t i = 0; i < n; ++i) dz[i] = dy[i] * act_bwd_from_output(y[i], act);
}

// dx[b,iy,ix,ci] = sum_{ky,kx,co} dz[b, iy-ky+ph, ix-kx+pw, co] * w[co,ky,kx,ci]
int ext_conv2d_backward_data(const double* dz, size_t B, size_t H, size_t W, size_t Ci, const double* w, size_t Co,
                             size_t kh, size_t kw, int padding, double* dx) {
    const long ph = padding == PAD_SAME ? (long)kh / 2 : 0, pw = padding == PAD_SAME ? (long)kw / 2 : 0;
    const size_t Ho = padding == PAD_SAME ? H : H - kh + 1, Wo = padding == PAD_SAME ? W : W - kw + 1;
    for (size_t b = 0; b < B; ++b)
        for (size_t iy = 0; iy < H; ++iy)
            for (size_t ix = 0; ix < W; ++ix)
                for (size_t ci = 0; ci < Ci; ++ci) {
                    double acc = 0.0;
                    for (size_t ky = 0; ky < kh; ++ky)
                        for (size_t kx = 0; kx < kw; ++kx) {
                            const long oy = (long)iy - (long)ky + ph, ox = (long)ix - (long)kx + pw;
                            if (oy < 0 || ox < 0 || oy >= (long)Ho || ox >= (long)Wo) continue;
                            const double* dp = dz + ((b * Ho + oy) * Wo + ox) * Co;
                            for (size_t co = 0; co < Co; ++co) acc += dp[co] * w[((co * kh + ky) * kw + kx) * Ci + ci];
                        }
                    dx[((b * H + iy) * W + ix) * Ci + ci] = acc;
                }
    return 0;
}

// dw[co,ky,kx,ci] = sum_{b,oy,ox} dz[b,oy,ox,co] * x[b, oy+ky-ph, ox+kx-pw, ci] ;  db[co] = sum dz[.,co]
int ext_conv2d_backward_weight(const double* x, const double* dz, size_t B, size_t H, size_t W, size_t Ci, size_t Co,
                               size_t kh, size_t kw, int padding, double* dw, double* db) {
    const long ph = padding == PAD_SAME ? (long)kh / 2 : 0, pw = padding == PAD_SAME ? (long)kw / 2 : 0;
    const size_t Ho = padding == PAD_SAME ? H : H - kh + 1, Wo = padding == PAD_SAME ? W : W - kw + 1;
    for (size_t i = 0; i < Co * kh * kw * Ci; ++i) dw[i] = 0.0;
    if (db) for (size_t i = 0; i < Co; ++i) db[i] = 0.0;
    for (size_t b = 0; b < B; ++b)
        for (size_t oy = 0; oy < Ho; ++oy)
            for (size_t ox = 0; ox < Wo; ++ox) {
                const double* dp = dz + ((b * Ho + oy) * Wo + ox) * Co;
                for (size_t co = 0; co < Co; ++co) {
                    const double d = dp[co];
                    if (db) db[co] += d;
                    for (size_t ky = 0; ky < kh; ++ky)
                        for (size_t kx = 0; kx < kw; ++kx) {
                            const long iy = (long)oy + (long)ky - ph, ix = (long)ox + (long)kx - pw;
                            if (iy < 0 || ix < 0 || iy >= (long)H || ix >= (long)W) continue;
                            const double* xp = x + ((b * H + iy) * W + ix) * Ci;
                            double* wp = dw + ((co * kh + ky) * kw + kx) * Ci;
                            for (size_t ci = 0; ci < Ci; ++ci) wp[ci] += d * xp[ci];
                        }
                }
            }
    return 0;
}

// 2x2 / stride 2 pooling, NHWC.  argmax (max only, optional): 2*dy+dx of the chosen element, last max wins.
int ext_pool2d_forward(const double* x, size_t B, size_t H, size_t W, size_t C, int padding, int pooling, double* y,
                       uint8_t* argmax) {
    if (H < 2 || W < 2) return 1;
    const size_t Ho = padding == PAD_SAME ? (H + 1) / 2 : H / 2, Wo = padding == PAD_SAME ? (W + 1) / 2 : W / 2;
    for (size_t b = 0; b < B; ++b)
        for (size_t oy = 0; oy < Ho; ++oy)
            for (size_t ox = 0; ox < Wo; ++ox)
                for (size_t c = 0; c < C; ++c) {
                    double p[4];
                    for (int i = 0; i < 4; ++i) {
                        const size_t iy = 2 * oy + (i >> 1), ix = 2 * ox + (i & 1);
                        p[i] = (iy < H && ix < W) ? x[((b * H + iy) * W + ix) * C + c] : 0.0;
                    }
                    const size_t o = ((b * Ho + oy) * Wo + ox) * C + c;
                    if (pooling == POOL_MAX) {
                        int best = 0;
                        for (int i = 1; i < 4; ++i) {
                            if (p[i] != p[i] || p[best] != p[best]) return 4;  // NaN: partial_cmp().unwrap() panics
                            if (!(p[i] < p[best])) best = i;                    // last maximal element wins
                        }
                        y[o] = p[best];
                        if (argmax) argmax[o] = (uint8_t)best;
                    } else {
                        y[o] = (((p[0] + p[1]) + p[2]) + p[3]) * 0.25;
                    }
                }
    return 0;
}

int ext_pool2d_backward(const double* dy, const uint8_t* argmax, size_t B, size_t H, size_t W, size_t C, int padding,
                        int pooling, double* dx) {
    const size_t Ho = padding == PAD_SAME ? (H + 1) / 2 : H / 2, Wo = padding == PAD_SAME ? (W + 1) / 2 : W / 2;
    for (size_t i = 0; i < B * H * W * C; ++i) dx[i] = 0.0;
    for (size_t b = 0; b < B; ++b)
        for (size_t oy = 0; oy < Ho; ++oy)
            for (size_t ox = 0; ox < Wo; ++ox)
                for (size_t c = 0; c < C; ++c) {
                    const size_t o = ((b * Ho + oy) * Wo + ox) * C + c;
                    for (int i = 0; i < 4; ++i) {
                        const size_t iy = 2 * oy + (i >> 1), ix = 2 * ox + (i & 1);
                        if (iy >= H || ix >= W) continue;             // gradient routed into the zero padding is dropped
                        if (pooling == POOL_MAX) { if (argmax[o] == i) dx[((b * H + iy) * W + ix) * C + c] = dy[o]; }
                        else dx[((b * H + iy) * W + ix) * C + c] = dy[o] * 0.25;
                    }
                }
    return 0;
}

// softmax + cross-entropy on column-major logits z (n x B): p = softmax(z_b), loss_b = -sum_i y_i log p_i,
// delta = p - y (the gradient of loss_b with respect to z_b).  Targets: onehot (n x B) or labels (B).
int ext_softmax_xent(const double* z, size_t n, size_t B, const double* onehot, const int64_t* labels, double* probs,
                     double* loss, double* delta) {
    for (size_t b = 0; b < B; ++b) {
        const double* zb = z + b * n;
        double m = zb[0];
        for (size_t i = 1; i < n; ++i) m = zb[i] > m ? zb[i] : m;
        double s = 0.0;
        for (size_t i = 0; i < n; ++i) s += std::exp(zb[i] - m);
        const double ls = std::log(s);
        double l = 0.0;
        for (size_t i = 0; i < n; ++i) {
            const double y = onehot ? onehot[b * n + i] : (labels[b] == (int64_t)i ? 1.0 : 0.0);
            const double p = std::exp(zb[i] - m) / s;
            if (probs) probs[b * n + i] = p;
            if (delta) delta[b * n + i] = p - y;
            if (y != 0.0) l += y * ((ls + m) - zb[i]);
        }
        if (loss) loss[b] = l;
    }
    return 0;
}

}  // extern "C"
