// rcn_oracle.cpp -- CPU restatement of rcn's training hot path (TEST INFRASTRUCTURE ONLY).
//
// This file is the *checker*, never the product: only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load it. Nothing under
// mercer_research_b200/ links, imports or calls it.
//
// It restates, operation for operation and in f64, the reference's CPU algorithm
// (jtstrader/mercer-research, crate `rcn`). Every function cites the reference file:line it follows.
// The reference itself cannot be built here (no rustc/cargo in the image),
// and its dense arithmetic lives in the un-vendored crate nalgebra 0.31.x (Cargo.toml:11), whose
// published gemv algorithm (column-sweep axpy, separate multiply and add) is restated in
// `matvec_colsweep` below.
//
// PARITY STATUS: pinned only where the reference's own tests pin it --
//   * kernel.rs:400-417  verify_separated_sobels    -> orc_sobel_separated / orc_sobel_full
//   * kernel.rs:434-441  convolve_2d_padding_same   -> orc_convolve_2d_i32 (30x30 i32 ramp, identity)
//   * rcn.rs:530-538     weight_init (shape only)   -> orc_layer_shapes
// Everything else (separated-SAME conv, relu, pooling, flatten order, standardisation, sigmoid,
// backprop, batch reduction, SGD step, argmax) is "PARITY UNPINNED": the reference holds no golden
// vector, KAT or fixture for it, so this restatement (cross-checked against an independent numpy
// restatement, oracle/rcn_oracle_np.py, and the derived KATs of SURVEY.md Appendix B) is the pin.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, so the separate
// multiply / add order of the reference is kept).
//
// All matrices are COLUMN-MAJOR f64 (nalgebra storage; kernel.rs:514-516 acknowledges it).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#ifndef M_E
#define M_E 2.71828182845904523536
#endif

namespace {

enum : int {
    ORC_OK = 0,
    ORC_PANIC_SHAPE = 1,       // reference would panic!() on a shape contract (kernel.rs:127,133,200,247)
    ORC_PANIC_OOB = 2,         // reference would panic on an out-of-bounds index (kernel.rs:156, >=5-wide SAME)
    ORC_PANIC_NOT_IMPL = 3,    // panic!("Not implemented") (kernel.rs:284,342)
    ORC_PANIC_NAN = 4,         // partial_cmp(..).unwrap() on NaN (kernel.rs:280,338)
    ORC_PANIC_DIM = 5,         // nalgebra dimension mismatch in `w * a`
    ORC_BAD_ARG = 6,
};

// Padding / Pooling / SeparableOperator discriminants, declaration order of kernel.rs:16-35.
enum : int { PAD_NONE = 0, PAD_SAME = 1 };
enum : int { POOL_AVERAGE = 0, POOL_MAX = 1 };
enum : int { OP_TOP = 0, OP_BOTTOM = 1, OP_LEFT = 2, OP_RIGHT = 3 };
// RCNLayer codes used by this repo's C ABI (rcn.rs:35-38 flattened with the inner enum).
enum : int { LAYER_CONV_NONE = 0, LAYER_CONV_SAME = 1, LAYER_POOL_AVERAGE = 2, LAYER_POOL_MAX = 3 };

struct Mat {
    size_t r = 0, c = 0;
    std::vector<double> d;  // column-major
    Mat() = default;
    Mat(size_t r_, size_t c_) : r(r_), c(c_), d(r_ * c_, 0.0) {}
    double& at(size_t i, size_t j) { return d[j * r + i]; }
    double at(size_t i, size_t j) const { return d[j * r + i]; }
};

// kernel.rs:110-194 -- generic over the scalar type (the reference tests use i32 and f64).
template <typename N>
int convolve_2d_t(const N* x, size_t H, size_t W, const N* k, size_t kh, size_t kw, int padding,
                  std::vector<N>& out, size_t& oh, size_t& ow) {
    // kernel.rs:123-128
    if ((kh == 0 && kw == 0) || kh > H || kw > W || kh == 0 || kw == 0) return ORC_PANIC_SHAPE;
    // kernel.rs:131-135
    if ((kh % 2 == 0 || kw % 2 == 0) && padding == PAD_SAME) return ORC_PANIC_SHAPE;
    auto X = [&](size_t i, size_t j) { return x[j * H + i]; };
    auto K = [&](size_t i, size_t j) { return k[j * kh + i]; };
    if (padding == PAD_SAME) {
        // kernel.rs:138-170
        oh = H; ow = W;
        const size_t ph = kh / 2, pw = kw / 2;
        const size_t PH = H + 2 * ph, PW = W + 2 * pw;
        std::vector<N> P(PH * PW, N(0));
        // kernel.rs:154-158: the copy hard-codes offset 1 in both dimensions.
        for (size_t cy = 1; cy < H + ph; ++cy)
            for (size_t cx = 1; cx < W + pw; ++cx) {
                if (cy - 1 >= H || cx - 1 >= W) return ORC_PANIC_OOB;  // self[(cy-1,cx-1)] out of range
                P[cx * PH + cy] = X(cy - 1, cx - 1);
            }
        out.assign(oh * ow, N(0));
        // kernel.rs:160-168: conv += p*k, ky outer, kx inner, starting from zero.
        for (size_t cy = 0; cy < oh; ++cy)
            for (size_t cx = 0; cx < ow; ++cx) {
                N acc = N(0);
                for (size_t ky = 0; ky < kh; ++ky)
                    for (size_t kx = 0; kx < kw; ++kx) acc += P[(cx + kx) * PH + (cy + ky)] * K(ky, kx);
                out[cx * oh + cy] = acc;
            }
        return ORC_OK;
    }
    // kernel.rs:171-192
    oh = H - kh + 1; ow = W - kw + 1;
    out.assign(oh * ow, N(0));
    for (size_t cy = 0; cy < oh; ++cy)
        for (size_t cx = 0; cx < ow; ++cx) {
            N acc = N(0);
            for (size_t ky = 0; ky < kh; ++ky)
                for (size_t kx = 0; kx < kw; ++kx) acc += X(cy + ky, cx + kx) * K(ky, kx);
            out[cx * oh + cy] = acc;
        }
    return ORC_OK;
}

// kernel.rs:38-53 -- (3x1 column vector, 1x3 row vector) per operator.
void sobel_separated(int op, double v[3], double h[3]) {
    switch (op) {
        case OP_TOP:    v[0] = 1;  v[1] = 0; v[2] = -1; h[0] = 1;  h[1] = 2; h[2] = 1;  break;
        case OP_BOTTOM: v[0] = -1; v[1] = 0; v[2] = 1;  h[0] = 1;  h[1] = 2; h[2] = 1;  break;
        case OP_LEFT:   v[0] = 1;  v[1] = 2; v[2] = 1;  h[0] = 1;  h[1] = 0; h[2] = -1; break;
        default:        v[0] = 1;  v[1] = 2; v[2] = 1;  h[0] = -1; h[1] = 0; h[2] = 1;  break;  // Right
    }
}

// kernel.rs:209-216
void relu_inplace(std::vector<double>& m) {
    for (double& f : m) f = (f >= 0.0) ? f : 0.0;
}

// kernel.rs:196-207
int convolve_2d_separated(const Mat& m, int op, int padding, Mat& out) {
    if (m.r < 3 || m.c < 3) return ORC_PANIC_SHAPE;  // kernel.rs:199-201
    double v[3], h[3];
    sobel_separated(op, v, h);
    std::vector<double> t1, t2;
    size_t h1, w1, h2, w2;
    int rc = convolve_2d_t<double>(m.d.data(), m.r, m.c, v, 3, 1, padding, t1, h1, w1);  // 3x1 first
    if (rc) return rc;
    rc = convolve_2d_t<double>(t1.data(), h1, w1, h, 1, 3, padding, t2, h2, w2);          // then 1x3
    if (rc) return rc;
    relu_inplace(t2);
    out.r = h2; out.c = w2; out.d.swap(t2);
    return ORC_OK;
}

// Rust's Iterator::max_by keeps the LAST maximal element (kernel.rs:278-281).
// Window order pooler[py + 2*px] = m[(2ry+px, 2rx+py)] (kernel.rs:273-277): (0,0),(0,1),(1,0),(1,1).
int max4_last(const double p[4], double& val, int& idx) {
    int best = 0;
    for (int i = 1; i < 4; ++i) {
        if (std::isnan(p[i]) || std::isnan(p[best])) return ORC_PANIC_NAN;
        if (!(p[i] < p[best])) best = i;  // Ordering::Less keeps old; Equal/Greater takes the new one
    }
    if (std::isnan(p[0])) return ORC_PANIC_NAN;
    val = p[best]; idx = best;
    return ORC_OK;
}

// kernel.rs:245-292 and __pooling_padded kernel.rs:298-349. argmax (2*dy+dx) is this repo's extension.
int pool_2d(const Mat& m, int padding, int pooling, Mat& out, std::vector<uint8_t>* argmax) {
    if (m.r < 2 || m.c < 2) return ORC_PANIC_SHAPE;  // kernel.rs:246-251
    size_t rp = 0, cp = 0;
    if (padding == PAD_SAME) { rp = m.r % 2; cp = m.c % 2; }  // kernel.rs:253-261
    const size_t PH = m.r + rp, PW = m.c + cp;
    // zero-padded copy (bottom/right), kernel.rs:310-319; identical to the unpadded path when rp=cp=0
    auto P = [&](size_t i, size_t j) -> double { return (i < m.r && j < m.c) ? m.at(i, j) : 0.0; };
    const size_t oh = PH / 2, ow = PW / 2;
    if (pooling != POOL_MAX) return ORC_PANIC_NOT_IMPL;  // kernel.rs:283-285,341-343
    out = Mat(oh, ow);
    if (argmax) argmax->assign(oh * ow, 0);
    for (size_t ry = 0; ry < oh; ++ry)
        for (size_t rx = 0; rx < ow; ++rx) {
            double p[4];
            for (int px = 0; px < 2; ++px)
                for (int py = 0; py < 2; ++py) p[py + px * 2] = P(ry * 2 + px, rx * 2 + py);
            double v; int idx;
            int rc = max4_last(p, v, idx);
            if (rc) return rc;
            out.at(ry, rx) = v;
            if (argmax) (*argmax)[rx * oh + ry] = (uint8_t)idx;
        }
    return ORC_OK;
}

// rcn.rs:41-46
const int SEP_OPS[4] = {OP_TOP, OP_LEFT, OP_RIGHT, OP_BOTTOM};

// rcn.rs:317-356
int flatten_feature_set(const int* cfg, size_t n_cfg, const Mat& m, std::vector<double>& out) {
    std::vector<Mat> fs;
    for (size_t li = 0; li < n_cfg; ++li) {
        const int layer = cfg[li];
        if (layer == LAYER_CONV_NONE || layer == LAYER_CONV_SAME) {
            const int p = (layer == LAYER_CONV_SAME) ? PAD_SAME : PAD_NONE;
            if (!fs.empty()) {
                const size_t curr_len = fs.size();  // rcn.rs:325
                for (size_t i = 0; i < curr_len; ++i)
                    for (int oi = 0; oi < 4; ++oi) {
                        Mat r;
                        int rc = convolve_2d_separated(fs[i], SEP_OPS[oi], p, r);
                        if (rc) return rc;
                        if (oi == 3) fs[i] = std::move(r);   // last op overwrites slot i (rcn.rs:331-332)
                        else fs.push_back(std::move(r));     // others are appended (rcn.rs:334)
                    }
            } else {
                for (int oi = 0; oi < 4; ++oi) {             // rcn.rs:339
                    Mat r;
                    int rc = convolve_2d_separated(m, SEP_OPS[oi], p, r);
                    if (rc) return rc;
                    fs.push_back(std::move(r));
                }
            }
        } else {
            const int pooling = (layer == LAYER_POOL_MAX) ? POOL_MAX : POOL_AVERAGE;
            for (auto& f : fs) {                             // rcn.rs:343-345, always Padding::Same
                Mat r;
                int rc = pool_2d(f, PAD_SAME, pooling, r, nullptr);
                if (rc) return rc;
                f = std::move(r);
            }
        }
    }
    out.clear();
    for (auto& f : fs) out.insert(out.end(), f.d.begin(), f.d.end());  // rcn.rs:350-355 column-major concat
    return ORC_OK;
}

// rcn.rs:478-483: 1/(1+powf(E,-x))
inline double sigmoid1(double x) { return 1.0 / (1.0 + std::pow(M_E, -x)); }
// rcn.rs:490-492: sigmoid(v) .* (1 - sigmoid(v)), sigmoid evaluated twice (bitwise identical values)
inline double sigmoid_prime1(double x) { double s = sigmoid1(x); return s * (1.0 - sigmoid1(x)); }

// nalgebra 0.31 gemv (Matrix * Vector): y = W[:,0]*x0; y += W[:,j]*xj for j>=1, multiply and add
// rounded separately. W is rows x cols column-major.
void matvec_colsweep(const double* W, size_t rows, size_t cols, const double* x, double* y) {
    if (cols == 0) { for (size_t i = 0; i < rows; ++i) y[i] = 0.0; return; }
    for (size_t i = 0; i < rows; ++i) y[i] = W[i] * x[0];
    for (size_t j = 1; j < cols; ++j) {
        const double xj = x[j];
        const double* col = W + j * rows;
        for (size_t i = 0; i < rows; ++i) y[i] = col[i] * xj + y[i];
    }
}

struct Net {
    size_t n_layers = 0;
    std::vector<size_t> rows, cols;         // W_l is rows[l] x cols[l]  (rcn.rs:500-503)
    std::vector<size_t> w_off, b_off;       // offsets into the flat [W0|b0|W1|b1|...] buffer
    size_t n_params = 0;
};

Net make_net(const size_t* rows, const size_t* cols, size_t n_layers) {
    Net n; n.n_layers = n_layers;
    size_t off = 0;
    for (size_t l = 0; l < n_layers; ++l) {
        n.rows.push_back(rows[l]); n.cols.push_back(cols[l]);
        n.w_off.push_back(off); off += rows[l] * cols[l];
        n.b_off.push_back(off); off += rows[l];
    }
    n.n_params = off;
    return n;
}

// rcn.rs:105-116
int classify_test(const Net& net, const double* params, const double* x, size_t n_in, double* out) {
    std::vector<double> a(x, x + n_in), z;
    for (size_t l = 0; l < net.n_layers; ++l) {
        if (net.cols[l] != a.size()) return ORC_PANIC_DIM;
        z.assign(net.rows[l], 0.0);
        matvec_colsweep(params + net.w_off[l], net.rows[l], net.cols[l], a.data(), z.data());
        const double* b = params + net.b_off[l];
        for (size_t i = 0; i < net.rows[l]; ++i) z[i] = sigmoid1(z[i] + b[i]);  // sigmoid(&(w*a + b))
        a.swap(z);
    }
    std::copy(a.begin(), a.end(), out);
    return ORC_OK;
}

// rcn.rs:260-314. grads is a flat [dW0|db0|...] buffer that is fully overwritten (the reference
// zero-initialises fresh matrices and then assigns every one of them).
int backprop(const Net& net, const double* params, const double* x, size_t n_in, const double* y,
             double* grads, std::vector<std::vector<double>>* acts_out,
             std::vector<std::vector<double>>* zs_out, std::vector<std::vector<double>>* deltas_out) {
    const size_t L = net.n_layers;
    std::vector<std::vector<double>> activations, zs;
    activations.emplace_back(x, x + n_in);
    for (size_t l = 0; l < L; ++l) {  // rcn.rs:281-291
        if (net.cols[l] != activations.back().size()) return ORC_PANIC_DIM;
        std::vector<double> z(net.rows[l]);
        matvec_colsweep(params + net.w_off[l], net.rows[l], net.cols[l], activations.back().data(), z.data());
        const double* b = params + net.b_off[l];
        for (size_t i = 0; i < z.size(); ++i) z[i] = z[i] + b[i];
        std::vector<double> a(z.size());
        for (size_t i = 0; i < z.size(); ++i) a[i] = sigmoid1(z[i]);
        zs.push_back(std::move(z));
        activations.push_back(std::move(a));
    }
    std::vector<std::vector<double>> deltas(L);
    // rcn.rs:299: (a_L - y) .* sigmoid_prime(z_L)
    std::vector<double> delta(net.rows[L - 1]);
    for (size_t i = 0; i < delta.size(); ++i)
        delta[i] = (activations[L][i] - y[i]) * sigmoid_prime1(zs[L - 1][i]);
    auto emit = [&](size_t l, const std::vector<double>& d) {
        double* db = grads + net.b_off[l];
        double* dW = grads + net.w_off[l];
        const std::vector<double>& a = activations[l];  // input activation of layer l
        for (size_t i = 0; i < d.size(); ++i) db[i] = d[i];                       // rcn.rs:302,309
        for (size_t j = 0; j < a.size(); ++j)                                      // rcn.rs:303,310 outer product
            for (size_t i = 0; i < d.size(); ++i) dW[j * d.size() + i] = d[i] * a[j];
    };
    emit(L - 1, delta);
    deltas[L - 1] = delta;
    for (size_t l = 1; l < L; ++l) {  // rcn.rs:305-311 (feedforward_cfg.len() == L-1)
        const size_t li = L - 1 - l;   // layer whose delta we compute
        const size_t up = li + 1;      // weight_end - l + 1
        // materialised transpose (rcn.rs:308) then column-sweep gemv
        const size_t r = net.rows[up], c = net.cols[up];
        std::vector<double> Wt(r * c);
        const double* W = params + net.w_off[up];
        for (size_t i = 0; i < r; ++i)
            for (size_t j = 0; j < c; ++j) Wt[i * c + j] = W[j * r + i];  // Wt is c x r column-major
        std::vector<double> nd(c);
        matvec_colsweep(Wt.data(), c, r, delta.data(), nd.data());
        for (size_t i = 0; i < c; ++i) nd[i] = nd[i] * sigmoid_prime1(zs[li][i]);
        delta.swap(nd);
        emit(li, delta);
        deltas[li] = delta;
    }
    if (acts_out) *acts_out = activations;
    if (zs_out) *zs_out = zs;
    if (deltas_out) *deltas_out = deltas;
    return ORC_OK;
}

}  // namespace

extern "C" {

int orc_version(void) { return 1; }

// kernel.rs:38-53
void orc_sobel_separated(int op, double* v3, double* h3) { sobel_separated(op, v3, h3); }

// kernel.rs:56-59 full 3x3 constants, column-major 3x3 out.
void orc_sobel_full(int op, double* k9) {
    static const double T[9] = {1, 0, -1, 2, 0, -2, 1, 0, -1};      // rows [1,2,1],[0,0,0],[-1,-2,-1]
    static const double Bm[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
    static const double Lm[9] = {1, 2, 1, 0, 0, 0, -1, -2, -1};     // rows [1,0,-1],[2,0,-2],[1,0,-1]
    static const double R[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};
    const double* s = op == OP_TOP ? T : op == OP_BOTTOM ? Bm : op == OP_LEFT ? Lm : R;
    std::memcpy(k9, s, sizeof(T));
}

// kernel.rs:110-194 (f64). out must hold H*W (Same) or (H-kh+1)*(W-kw+1) (None).
int orc_convolve_2d(const double* m, size_t H, size_t W, const double* k, size_t kh, size_t kw,
                    int padding, double* out, size_t* oh, size_t* ow) {
    std::vector<double> o; size_t a = 0, b = 0;
    int rc = convolve_2d_t<double>(m, H, W, k, kh, kw, padding, o, a, b);
    if (rc) return rc;
    std::copy(o.begin(), o.end(), out);
    if (oh) *oh = a;
    if (ow) *ow = b;
    return ORC_OK;
}

// kernel.rs:110-194 instantiated at i32 (the type the reference test kernel.rs:434-441 uses).
int orc_convolve_2d_i32(const int32_t* m, size_t H, size_t W, const int32_t* k, size_t kh, size_t kw,
                        int padding, int32_t* out, size_t* oh, size_t* ow) {
    std::vector<int32_t> o; size_t a = 0, b = 0;
    int rc = convolve_2d_t<int32_t>(m, H, W, k, kh, kw, padding, o, a, b);
    if (rc) return rc;
    std::copy(o.begin(), o.end(), out);
    if (oh) *oh = a;
    if (ow) *ow = b;
    return ORC_OK;
}

// kernel.rs:196-207
int orc_convolve_2d_separated(const double* m, size_t H, size_t W, int op, int padding, double* out,
                              size_t* oh, size_t* ow) {
    Mat in(H, W); std::copy(m, m + H * W, in.d.begin());
    Mat r;
    int rc = convolve_2d_separated(in, op, padding, r);
    if (rc) return rc;
    std::copy(r.d.begin(), r.d.end(), out);
    if (oh) *oh = r.r;
    if (ow) *ow = r.c;
    return ORC_OK;
}

// kernel.rs:209-216
void orc_relu(const double* m, size_t n, double* out) {
    for (size_t i = 0; i < n; ++i) out[i] = (m[i] >= 0.0) ? m[i] : 0.0;
}

// kernel.rs:245-349; argmax may be NULL.
int orc_pool_2d(const double* m, size_t H, size_t W, int padding, int pooling, double* out,
                uint8_t* argmax, size_t* oh, size_t* ow) {
    Mat in(H, W); std::copy(m, m + H * W, in.d.begin());
    Mat r; std::vector<uint8_t> am;
    int rc = pool_2d(in, padding, pooling, r, argmax ? &am : nullptr);
    if (rc) return rc;
    std::copy(r.d.begin(), r.d.end(), out);
    if (argmax) std::copy(am.begin(), am.end(), argmax);
    if (oh) *oh = r.r;
    if (ow) *ow = r.c;
    return ORC_OK;
}

// Shape walk of rcn.rs:317-356 (no arithmetic): feature length and final map geometry.
int orc_feature_shape(const int* cfg, size_t n_cfg, size_t H, size_t W, size_t* n_maps, size_t* oh,
                      size_t* ow) {
    size_t maps = 0, h = H, w = W;
    for (size_t i = 0; i < n_cfg; ++i) {
        if (cfg[i] == LAYER_CONV_NONE || cfg[i] == LAYER_CONV_SAME) {
            if (h < 3 || w < 3) return ORC_PANIC_SHAPE;
            maps = maps ? maps * 4 : 4;
            if (cfg[i] == LAYER_CONV_NONE) { h -= 2; w -= 2; }
        } else if (maps) {
            if (h < 2 || w < 2) return ORC_PANIC_SHAPE;
            if (cfg[i] != LAYER_POOL_MAX) return ORC_PANIC_NOT_IMPL;
            h = (h + 1) / 2; w = (w + 1) / 2;  // always Padding::Same (rcn.rs:344)
        }
    }
    *n_maps = maps; *oh = h; *ow = w;
    return ORC_OK;
}

// rcn.rs:317-356: one image (H x W column-major f64) -> flattened features.
int orc_flatten_feature_set(const int* cfg, size_t n_cfg, const double* m, size_t H, size_t W,
                            double* out, size_t out_cap, size_t* out_len) {
    Mat in(H, W); std::copy(m, m + H * W, in.d.begin());
    std::vector<double> o;
    int rc = flatten_feature_set(cfg, n_cfg, in, o);
    if (rc) return rc;
    if (out_len) *out_len = o.size();
    if (o.size() > out_cap) return ORC_BAD_ARG;
    std::copy(o.begin(), o.end(), out);
    return ORC_OK;
}

// lib.rs:27-33 + rcn.rs:317-356 over a batch: images are B row-major u8 H x W (the `image` crate's
// buffer order); out is L x B column-major (sample b at out + b*L).
int orc_features_u8(const int* cfg, size_t n_cfg, const uint8_t* images, size_t B, size_t H, size_t W,
                    double* out, size_t L) {
    Mat in(H, W);
    std::vector<double> o;
    for (size_t b = 0; b < B; ++b) {
        const uint8_t* img = images + b * H * W;
        for (size_t r = 0; r < H; ++r)
            for (size_t c = 0; c < W; ++c) in.at(r, c) = (double)img[r * W + c];  // from_row_iterator
        int rc = flatten_feature_set(cfg, n_cfg, in, o);
        if (rc) return rc;
        if (o.size() != L) return ORC_BAD_ARG;
        std::copy(o.begin(), o.end(), out + b * L);
    }
    return ORC_OK;
}

// rcn.rs:230-251: sequential sums over every feature of every sample.
void orc_gen_scales(const double* feats, size_t L, size_t B, double* mean_out, double* sd_out) {
    double mean = 0.0, sd = 0.0;
    const double n = (double)L * (double)B;
    for (size_t i = 0; i < L * B; ++i) mean += feats[i];
    mean /= n;
    for (size_t i = 0; i < L * B; ++i) { double d = feats[i] - mean; sd += d * d; }  // powi(.,2) == d*d
    sd = std::sqrt(sd / n);
    *mean_out = mean; *sd_out = sd;
}

// rcn.rs:407-412 / 86-89
void orc_standardise(double* feats, size_t n, double mean, double sd) {
    for (size_t i = 0; i < n; ++i) {
        double d = (feats[i] - mean) / sd;
        feats[i] = (d >= 0.0) ? d : 0.0;
    }
}

// rcn.rs:425-457 incl. the `4^c / 2^p * l` integer arithmetic (division first). rows/cols need
// n_ff+1 slots.
int orc_layer_shapes(const int* cfg, size_t n_cfg, const size_t* ff, size_t n_ff, size_t classes,
                     size_t l, size_t* rows, size_t* cols) {
    if (n_ff == 0) return ORC_PANIC_OOB;  // feedforward_cfg[0] (rcn.rs:444)
    unsigned c = 0, p = 0;
    for (size_t i = 0; i < n_cfg; ++i) {
        if (cfg[i] == LAYER_CONV_NONE || cfg[i] == LAYER_CONV_SAME) c += 1; else p += 2;
    }
    size_t pc = 1, pp = 1;
    for (unsigned i = 0; i < c; ++i) pc *= 4;
    for (unsigned i = 0; i < p; ++i) pp *= 2;
    size_t a = pc / pp * l;
    size_t b = ff[0];
    for (size_t i = 0; i < n_ff + 1; ++i) {
        rows[i] = b; cols[i] = a;  // get_weight_matrix(a, b): dims (output, input) rcn.rs:500-503
        a = b;
        b = (i + 1 < n_ff) ? ff[i + 1] : classes;
    }
    return ORC_OK;
}

double orc_sigmoid(double x) { return sigmoid1(x); }
double orc_sigmoid_prime(double x) { return sigmoid_prime1(x); }

size_t orc_param_count(const size_t* rows, const size_t* cols, size_t n_layers) {
    return make_net(rows, cols, n_layers).n_params;
}

// rcn.rs:105-116 over a batch: X is n_in x B, out is classes x B (both column-major).
int orc_forward(const size_t* rows, const size_t* cols, size_t n_layers, const double* params,
                const double* X, size_t n_in, size_t B, double* out) {
    Net net = make_net(rows, cols, n_layers);
    const size_t n_out = net.rows[n_layers - 1];
    for (size_t b = 0; b < B; ++b) {
        int rc = classify_test(net, params, X + b * n_in, n_in, out + b * n_out);
        if (rc) return rc;
    }
    return ORC_OK;
}

// rcn.rs:92-97: argmax with max_by(total_cmp) => last maximal element wins.
// (total_cmp orders -0.0 < +0.0 and NaNs by bit pattern; sigmoid outputs are in [0,1] so only the
// plain ordering matters here.)
void orc_argmax_last(const double* acts, size_t n, size_t B, int64_t* labels) {
    for (size_t b = 0; b < B; ++b) {
        const double* a = acts + b * n;
        size_t best = 0;
        for (size_t i = 1; i < n; ++i)
            if (!(a[i] < a[best])) best = i;
        labels[b] = (int64_t)best;
    }
}

// rcn.rs:152-157: correct iff map(v == max) equals the one-hot exactly (ties => wrong).
size_t orc_accuracy(const double* acts, size_t n, size_t B, const int64_t* labels) {
    size_t accept = 0;
    for (size_t b = 0; b < B; ++b) {
        const double* a = acts + b * n;
        double mx = a[0];
        for (size_t i = 1; i < n; ++i) mx = std::max(mx, a[i]);  // DVector::max()
        bool ok = true;
        for (size_t i = 0; i < n; ++i) {
            const double r = (a[i] == mx) ? 1.0 : 0.0;
            const double e = ((int64_t)i == labels[b]) ? 1.0 : 0.0;
            if (r != e) { ok = false; break; }
        }
        accept += ok;
    }
    return accept;
}

// rcn.rs:260-314 for ONE sample; grads = flat [dW0|db0|...]. Optional taps: zs / acts / deltas are
// concatenated per layer (z_0..z_{L-1}; a_1..a_L; delta_0..delta_{L-1}), each sum(rows) long.
int orc_backprop(const size_t* rows, const size_t* cols, size_t n_layers, const double* params,
                 const double* x, size_t n_in, const double* y, double* grads, double* zs, double* acts,
                 double* deltas) {
    Net net = make_net(rows, cols, n_layers);
    std::vector<std::vector<double>> A, Z, D;
    int rc = backprop(net, params, x, n_in, y, grads, &A, &Z, &D);
    if (rc) return rc;
    size_t o = 0;
    for (size_t l = 0; l < n_layers; ++l) {
        for (size_t i = 0; i < net.rows[l]; ++i) {
            if (zs) zs[o + i] = Z[l][i];
            if (acts) acts[o + i] = A[l + 1][i];
            if (deltas) deltas[o + i] = D[l][i];
        }
        o += net.rows[l];
    }
    return ORC_OK;
}

// rcn.rs:176-223. X is n_in x B, Y is classes x B (one-hot columns). params updated in place.
// grad_sum_out (optional) receives the summed gradients before the update.
// n_threads <= 1: samples are summed in index order (deterministic, used for parity checks).
// n_threads  > 1: worker threads pull samples and add under one lock, mirroring the reference's
//                 rayon par_iter + Mutex accumulation (order nondeterministic) -- the timed CPU baseline.
int orc_train_batch(const size_t* rows, const size_t* cols, size_t n_layers, double* params,
                    const double* X, size_t n_in, const double* Y, size_t B, double eta, int n_threads,
                    double* grad_sum_out) {
    Net net = make_net(rows, cols, n_layers);
    const size_t n_out = net.rows[n_layers - 1];
    std::vector<double> acc(net.n_params, 0.0);  // rcn.rs:177-188
    std::atomic<int> err{0};
    if (n_threads <= 1) {
        std::vector<double> g(net.n_params);
        for (size_t b = 0; b < B; ++b) {
            int rc = backprop(net, params, X + b * n_in, n_in, Y + b * n_out, g.data(), nullptr, nullptr, nullptr);
            if (rc) return rc;
            for (size_t i = 0; i < acc.size(); ++i) acc[i] = g[i] + acc[i];  // rcn.rs:195-204 (new + acc)
        }
    } else {
        std::mutex mu;
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            std::vector<double> g(net.n_params);
            for (;;) {
                const size_t b = next.fetch_add(1);
                if (b >= B) break;
                std::fill(g.begin(), g.end(), 0.0);  // rcn.rs:265-274 zero-initialised gradients
                int rc = backprop(net, params, X + b * n_in, n_in, Y + b * n_out, g.data(), nullptr, nullptr, nullptr);
                if (rc) { err = rc; break; }
                std::lock_guard<std::mutex> lk(mu);  // rcn.rs:192-193
                for (size_t i = 0; i < acc.size(); ++i) acc[i] = g[i] + acc[i];
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker);
        for (auto& t : th) t.join();
        if (err) return err;
    }
    if (grad_sum_out) std::copy(acc.begin(), acc.end(), grad_sum_out);
    const double scale = eta / (double)B;  // rcn.rs:214,221: (eta / batch.len() as f64) * w
    for (size_t i = 0; i < acc.size(); ++i) params[i] = params[i] - scale * acc[i];
    return ORC_OK;
}

// Timed CPU baseline leg: features (rcn.rs:317-356) + standardise (rcn.rs:407-412) + train_batch
// (rcn.rs:176-223) on one batch of u8 images, worker threads over samples for both stages.
int orc_train_step_u8(const int* cfg, size_t n_cfg, const size_t* rows, const size_t* cols,
                      size_t n_layers, double* params, const uint8_t* images, const int64_t* labels,
                      size_t B, size_t H, size_t W, double mean, double sd, double eta, int n_threads,
                      double* feats_scratch, double* onehot_scratch) {
    const size_t L = cols[0];
    const size_t classes = rows[n_layers - 1];
    std::atomic<int> err{0};
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= B) break;
            int rc = orc_features_u8(cfg, n_cfg, images + b * H * W, 1, H, W, feats_scratch + b * L, L);
            if (rc) { err = rc; break; }
            orc_standardise(feats_scratch + b * L, L, mean, sd);
            for (size_t i = 0; i < classes; ++i)
                onehot_scratch[b * classes + i] = ((int64_t)i == labels[b]) ? 1.0 : 0.0;  // rcn.rs:466-471
        }
    };
    const int nt = std::max(1, n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
    if (err) return err;
    return orc_train_batch(rows, cols, n_layers, params, feats_scratch, L, onehot_scratch, B, eta,
                           n_threads, nullptr);
}

}  // extern "C"
