// features_device.cuh -- device code of the feature stage (Sobel-separated conv, ReLU, 2x2 max-pool, flatten,
// standardise), shared by the stand-alone feature kernels (features.cu) and the fused training kernel (smallnet.cu).
// Reference: rcn/src/utils/kernel.rs:38-53,110-349; rcn/src/rcn.rs:41-46,317-356,407-412.
#pragma once
#include "features.cuh"

namespace rcn {

// ------------------------------------------------------------------------------------------------
// Device arithmetic.  mac(a, k, c) = c + a*k.  The taps are in {0, +-1, +-2}, so a*k is exact in f64 and a
// fused multiply-add rounds exactly like the reference's separate multiply and add (kernel.rs:164).  Zero
// taps are kept (inf*0 must stay NaN as in the reference); for T = int the compiler folds them away.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mac(double a, double k, double c) { return fma(a, k, c); }
__device__ __forceinline__ int mac(int a, int k, int c) { return a * k + c; }

template <typename T>
__device__ __forceinline__ T relu1(T v) { return (v >= T(0)) ? v : T(0); }  // kernel.rs:214

// Pre-activation responses of the four operators (kernel.rs:38-53) at output pixel (y, x) of
// convolve_2d(3x1, p) . convolve_2d(1x3, p)  (kernel.rs:204-205), f is one h x w column-major map.
// SAME reproduces the reference's padded-copy quirk (kernel.rs:154-158): the result is the
// Sobel response centred at (y-1, x-1); row 0 is zero; the last input column / last intermediate row are
// never read.
template <typename T, bool SAME>
__device__ __forceinline__ void sobel4(const T* __restrict__ f, int h, int w, int y, int x, T& t, T& l, T& r, T& b) {
    const T Z = T(0);
    T ct[3], cb[3], cs[3];  // vertical passes with [1,0,-1], [-1,0,1], [1,2,1] at the three columns
    if (SAME) {
        if (y == 0) { t = l = r = b = Z; return; }
        const int rr = y - 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int j = x + k - 2;
            if (j >= 0 && j <= w - 2) {
                const T* col = f + (size_t)j * h;
                const T x0 = (rr >= 1) ? col[rr - 1] : Z;
                const T x1 = col[rr];
                const T x2 = col[rr + 1];
                ct[k] = mac(x2, T(-1), mac(x1, T(0), mac(x0, T(1), Z)));
                cb[k] = mac(x2, T(1), mac(x1, T(0), mac(x0, T(-1), Z)));
                cs[k] = mac(x2, T(1), mac(x1, T(2), mac(x0, T(1), Z)));
            } else {
                ct[k] = cb[k] = cs[k] = Z;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T* col = f + (size_t)(x + k) * h;
            const T x0 = col[y], x1 = col[y + 1], x2 = col[y + 2];
            ct[k] = mac(x2, T(-1), mac(x1, T(0), mac(x0, T(1), Z)));
            cb[k] = mac(x2, T(1), mac(x1, T(0), mac(x0, T(-1), Z)));
            cs[k] = mac(x2, T(1), mac(x1, T(2), mac(x0, T(1), Z)));
        }
    }
    t = mac(ct[2], T(1), mac(ct[1], T(2), mac(ct[0], T(1), Z)));   // Top:    h = [1,2,1]
    b = mac(cb[2], T(1), mac(cb[1], T(2), mac(cb[0], T(1), Z)));   // Bottom: h = [1,2,1]
    l = mac(cs[2], T(-1), mac(cs[1], T(0), mac(cs[0], T(1), Z)));  // Left:   h = [1,0,-1]
    r = mac(cs[2], T(1), mac(cs[1], T(0), mac(cs[0], T(-1), Z)));  // Right:  h = [-1,0,1]
}

template <typename T>
__device__ __forceinline__ void sobel4_relu(const Stage& st, const T* __restrict__ f, int y, int x, T& t, T& l, T& r, T& b) {
    if (st.same) sobel4<T, true>(f, st.h_in, st.w_in, y, x, t, l, r, b);
    else sobel4<T, false>(f, st.h_in, st.w_in, y, x, t, l, r, b);
    t = relu1(t); l = relu1(l); r = relu1(r); b = relu1(b);
}

// Iterator::max_by keeps the LAST maximal element (kernel.rs:278-281): replace unless strictly smaller.
template <typename T>
__device__ __forceinline__ void max_last(T& best, T v) { if (!(v < best)) best = v; }

// Exact-integer fast path for a Padding::Same conv directly followed by the 2x2 max-pool (the canonical rcn layer
// pair): the four conv pixels of one pooled output share a 4x4 input patch, the vertical passes are shared
// between Top/Bottom ([1,0,-1] and its negation) and Left/Right ([1,2,1]), and ReLU + max collapse to
// max(0, max T), max(0, -min T).  Integer arithmetic is exact, so this equals the literal tap chains of sobel4
// (and therefore the reference) bit for bit.
__device__ __forceinline__ void conv_pool_same_int(const int* __restrict__ f, int h, int w, int y, int x, int& t, int& l,
                                                   int& r, int& b) {
    int p[4][4];  // rows 2y-2 .. 2y+1, cols 2x-2 .. 2x+1
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = 2 * x - 2 + k;
        const bool cok = (c >= 0) && (c <= w - 2);  // intermediate columns outside [0, w-2] are zero
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = 2 * y - 2 + i;
            p[i][k] = (cok && rr >= 0 && rr < h) ? f[c * h + rr] : 0;
        }
    }
    int tmax = 0, tmin = 0, lmax = 0, lmin = 0;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int cy = 2 * y + a;
        const bool rok = (cy >= 1) && (cy < h);  // conv row 0 is zero; rows >= h are the pool's zero padding
        int vT[4], vS[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            vT[k] = p[a][k] - p[a + 2][k];
            vS[k] = p[a][k] + 2 * p[a + 1][k] + p[a + 2][k];
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const bool ok = rok && (2 * x + e < w);
            const int T = ok ? vT[e] + 2 * vT[e + 1] + vT[e + 2] : 0;
            const int L = ok ? vS[e] - vS[e + 2] : 0;
            tmax = max(tmax, T); tmin = min(tmin, T);
            lmax = max(lmax, L); lmin = min(lmin, L);
        }
    }
    t = tmax; b = -tmin; l = lmax; r = -lmin;
}

template <typename T>
__device__ __forceinline__ bool conv_pool_fast(const Stage&, const T*, int, int, T&, T&, T&, T&) { return false; }
template <>
__device__ __forceinline__ bool conv_pool_fast<int>(const Stage& st, const int* f, int y, int x, int& t, int& l, int& r, int& b) {
    if (!st.same) return false;
    conv_pool_same_int(f, st.h_in, st.w_in, y, x, t, l, r, b);
    return true;
}

// Runs one stage for one image.  `in`: n_in maps (column-major each, back to back).  emit(slot, y, x, v).
template <typename T, typename Emit>
__device__ __forceinline__ void run_stage(const Stage& st, const T* __restrict__ in, Emit emit, int tid, int nthreads) {
    const int hw_out = st.h_out * st.w_out;
    const int hw_in = st.h_in * st.w_in;
    const int items = st.n_in * hw_out;
    for (int it = tid; it < items; it += nthreads) {
        const int i = it / hw_out;
        const int rem = it - i * hw_out;
        const int x = rem / st.h_out;
        const int y = rem - x * st.h_out;
        const T* f = in + (size_t)i * hw_in;
        if (st.kind == 2) {
            // pool_2d(Padding::Same, Max): kernel.rs:245-349, zero row/col appended when odd.
            T best = T(0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cy = 2 * y + (q >> 1), cx = 2 * x + (q & 1);
                const T v = (cy < st.h_in && cx < st.w_in) ? f[(size_t)cx * st.h_in + cy] : T(0);
                if (q == 0) best = v; else max_last(best, v);
            }
            emit(i, y, x, best);
            continue;
        }
        T t, l, r, b;
        if (st.kind == 0) {
            sobel4_relu<T>(st, f, y, x, t, l, r, b);
        } else if (!conv_pool_fast<T>(st, f, y, x, t, l, r, b)) {
            t = l = r = b = T(0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cy = 2 * y + (q >> 1), cx = 2 * x + (q & 1);
                T vt = T(0), vl = T(0), vr = T(0), vb = T(0);
                if (cy < st.h_c && cx < st.w_c) sobel4_relu<T>(st, f, cy, cx, vt, vl, vr, vb);
                if (q == 0) { t = vt; l = vl; r = vr; b = vb; }
                else { max_last(t, vt); max_last(l, vl); max_last(r, vr); max_last(b, vb); }
            }
        }
        // Slot order (rcn.rs:325-339): first conv -> [T,L,R,B]; later convs overwrite slot i with Bottom and
        // append Top, Left, Right of map i at n + 3i.
        int sT, sL, sR, sB;
        if (st.first) { sT = 0; sL = 1; sR = 2; sB = 3; }
        else { sB = i; sT = st.n_in + 3 * i; sL = sT + 1; sR = sT + 2; }
        emit(sT, y, x, t); emit(sL, y, x, l); emit(sR, y, x, r); emit(sB, y, x, b);
    }
}

template <typename T>
struct EmitMaps {  // into a map set (shared or global), column-major maps back to back
    T* out; int h, hw;
    __device__ __forceinline__ void operator()(int slot, int y, int x, T v) const { out[(size_t)slot * hw + x * h + y] = v; }
};

template <typename T>
struct EmitFeatures {  // final stage: flatten (rcn.rs:350-355) + optional standardise/clamp (rcn.rs:407-412)
    double* out; int h, hw; Standardise sc;
    __device__ __forceinline__ void operator()(int slot, int y, int x, T v) const {
        double d = (double)v;
        if (sc.mode == 1) {
            d = (d - sc.mean) / sc.sd;                       // IEEE division
            d = (d >= 0.0) ? d : 0.0;
        } else if (sc.mode == 2) {
            // Markstein division: q = RN(a*r), rem = a - q*sd (exact in one fma), RN(q + rem*r) == RN(a/sd).
            // Only enabled after the host has checked bit-equality with IEEE division for EVERY value this
            // stage can produce with this (mean, sd) (standardise_fast_ok), so it is exact, not approximate.
            const double a = d - sc.mean;
            const double q = __dmul_rn(a, sc.rcp);
            const double rem = fma(-q, sc.sd, a);
            d = fma(rem, sc.rcp, q);
            d = (d >= 0.0) ? d : 0.0;
        }
        out[(size_t)slot * hw + x * h + y] = d;
    }
};

template <typename TIN, typename T>
__device__ __forceinline__ void load_image(const TIN* __restrict__ src, T* dst, int H, int W, int tid, int nthreads);

// u8 row-major (image crate) -> column-major T: DMatrix::from_row_iterator (lib.rs:29-33)
template <>
__device__ __forceinline__ void load_image<uint8_t, int>(const uint8_t* __restrict__ src, int* dst, int H, int W, int tid, int nthreads) {
    const int n = H * W;
    for (int i = tid; i < n; i += nthreads) { const int r = i / W, c = i - r * W; dst[c * H + r] = (int)src[i]; }
}
template <>
__device__ __forceinline__ void load_image<uint8_t, double>(const uint8_t* __restrict__ src, double* dst, int H, int W, int tid, int nthreads) {
    const int n = H * W;
    for (int i = tid; i < n; i += nthreads) { const int r = i / W, c = i - r * W; dst[c * H + r] = (double)src[i]; }
}
template <>
__device__ __forceinline__ void load_image<double, double>(const double* __restrict__ src, double* dst, int H, int W, int tid, int nthreads) {
    const int n = H * W;
    for (int i = tid; i < n; i += nthreads) dst[i] = src[i];
}

__device__ __forceinline__ size_t source_image(const BatchIndex& bi, size_t img) {
    if (!bi.cursor) return img;
    const long long pos = *bi.cursor + (long long)img;
    return (size_t)(bi.perm ? bi.perm[pos] : pos);
}
// where that image's pixels live: the dataset itself, or a ring of `window` slots when streaming from the host
__device__ __forceinline__ size_t image_slot(const BatchIndex& bi, size_t src) {
    return bi.window ? (size_t)((unsigned)src % (unsigned)bi.window) : src;   // 32-bit: a streamed epoch has < 2^32 samples
}


}  // namespace rcn
