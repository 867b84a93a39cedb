"""ctypes binding of the CPU oracle ``librcn_oracle.so`` (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package; the product (mercer_research_b200) never does.
Arrays follow the reference's layout: column-major f64 matrices, batches as (n x B) column-major,
i.e. numpy arrays of shape (B, n) C-contiguous ("sample b is contiguous").
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librcn_oracle.so")

PAD_NONE, PAD_SAME = 0, 1
POOL_AVERAGE, POOL_MAX = 0, 1
OP_TOP, OP_BOTTOM, OP_LEFT, OP_RIGHT = 0, 1, 2, 3
LAYER_CONV_NONE, LAYER_CONV_SAME, LAYER_POOL_AVERAGE, LAYER_POOL_MAX = 0, 1, 2, 3

PANIC_NAMES = {1: "shape", 2: "index out of bounds", 3: "Not implemented", 4: "NaN in partial_cmp",
               5: "dimension mismatch", 6: "bad argument"}


class RefPanic(Exception):
    """The reference would panic!() on these inputs."""

    def __init__(self, code):
        super().__init__(f"reference panic: {PANIC_NAMES.get(code, code)}")
        self.code = code


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("rcn_oracle.cpp", "ext_oracle.cpp", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_sp = C.POINTER(C.c_size_t)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_sigmoid.restype = C.c_double
        _lib.orc_sigmoid.argtypes = [C.c_double]
        _lib.orc_sigmoid_prime.restype = C.c_double
        _lib.orc_sigmoid_prime.argtypes = [C.c_double]
        _lib.orc_param_count.restype = C.c_size_t
        _lib.orc_accuracy.restype = C.c_size_t
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _sz(seq):
    return (C.c_size_t * len(seq))(*[int(x) for x in seq])


def _ints(seq):
    return (C.c_int * max(1, len(seq)))(*[int(x) for x in seq])


def _chk(rc):
    if rc:
        raise RefPanic(rc)


def _cm(m):
    """2-D (row, col) numpy array -> column-major flat f64 buffer."""
    return np.asfortranarray(np.asarray(m, dtype=np.float64)).ravel(order="K").copy()


def _from_cm(buf, h, w):
    return np.asarray(buf[: h * w]).reshape((w, h)).T.copy()


def sobel_separated(op):
    v = np.zeros(3)
    h = np.zeros(3)
    lib().orc_sobel_separated(int(op), _d(v), _d(h))
    return v.reshape(3, 1), h.reshape(1, 3)


def sobel_full(op):
    k = np.zeros(9)
    lib().orc_sobel_full(int(op), _d(k))
    return _from_cm(k, 3, 3)


def convolve_2d(m, k, padding):
    m = np.asarray(m)
    k = np.asarray(k)
    H, W = m.shape
    kh, kw = k.shape
    oh, ow = C.c_size_t(), C.c_size_t()
    if np.issubdtype(m.dtype, np.integer):
        mi = np.asfortranarray(m.astype(np.int32)).ravel(order="K").copy()
        ki = np.asfortranarray(k.astype(np.int32)).ravel(order="K").copy()
        out = np.zeros(H * W, dtype=np.int32)
        ip = C.POINTER(C.c_int32)
        _chk(lib().orc_convolve_2d_i32(mi.ctypes.data_as(ip), C.c_size_t(H), C.c_size_t(W), ki.ctypes.data_as(ip),
                                       C.c_size_t(kh), C.c_size_t(kw), int(padding), out.ctypes.data_as(ip),
                                       C.byref(oh), C.byref(ow)))
        return _from_cm(out, oh.value, ow.value)
    out = np.zeros(H * W)
    _chk(lib().orc_convolve_2d(_d(_cm(m)), C.c_size_t(H), C.c_size_t(W), _d(_cm(k)), C.c_size_t(kh),
                               C.c_size_t(kw), int(padding), _d(out), C.byref(oh), C.byref(ow)))
    return _from_cm(out, oh.value, ow.value)


def convolve_2d_separated(m, op, padding):
    m = np.asarray(m, dtype=np.float64)
    H, W = m.shape
    out = np.zeros(H * W)
    oh, ow = C.c_size_t(), C.c_size_t()
    _chk(lib().orc_convolve_2d_separated(_d(_cm(m)), C.c_size_t(H), C.c_size_t(W), int(op), int(padding),
                                         _d(out), C.byref(oh), C.byref(ow)))
    return _from_cm(out, oh.value, ow.value)


def relu(m):
    m = np.ascontiguousarray(m, dtype=np.float64)
    out = np.zeros_like(m)
    lib().orc_relu(_d(m), C.c_size_t(m.size), _d(out))
    return out


def pool_2d(m, padding, pooling, return_argmax=False):
    m = np.asarray(m, dtype=np.float64)
    H, W = m.shape
    out = np.zeros(((H + 1) // 2) * ((W + 1) // 2))
    arg = np.zeros(out.size, dtype=np.uint8)
    oh, ow = C.c_size_t(), C.c_size_t()
    _chk(lib().orc_pool_2d(_d(_cm(m)), C.c_size_t(H), C.c_size_t(W), int(padding), int(pooling), _d(out),
                           arg.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(oh), C.byref(ow)))
    r = _from_cm(out, oh.value, ow.value)
    if return_argmax:
        return r, arg[: oh.value * ow.value].reshape((ow.value, oh.value)).T.copy()
    return r


def feature_shape(cfg, H, W):
    n, oh, ow = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _chk(lib().orc_feature_shape(_ints(cfg), C.c_size_t(len(cfg)), C.c_size_t(H), C.c_size_t(W), C.byref(n),
                                 C.byref(oh), C.byref(ow)))
    return n.value, oh.value, ow.value


def flatten_feature_set(cfg, m):
    m = np.asarray(m, dtype=np.float64)
    H, W = m.shape
    n, oh, ow = feature_shape(cfg, H, W)
    out = np.zeros(max(1, n * oh * ow))
    ln = C.c_size_t()
    _chk(lib().orc_flatten_feature_set(_ints(cfg), C.c_size_t(len(cfg)), _d(_cm(m)), C.c_size_t(H), C.c_size_t(W),
                                       _d(out), C.c_size_t(out.size), C.byref(ln)))
    return out[: ln.value].copy()


def features_u8(cfg, images):
    """images: (B, H, W) uint8 row-major -> (B, L) f64 (un-standardised)."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    B, H, W = images.shape
    n, oh, ow = feature_shape(cfg, H, W)
    L = n * oh * ow
    out = np.zeros((B, L))
    _chk(lib().orc_features_u8(_ints(cfg), C.c_size_t(len(cfg)), images.ctypes.data_as(C.POINTER(C.c_uint8)),
                               C.c_size_t(B), C.c_size_t(H), C.c_size_t(W), _d(out), C.c_size_t(L)))
    return out


def gen_scales(feats):
    feats = np.ascontiguousarray(feats, dtype=np.float64)
    B, L = feats.shape
    mean, sd = C.c_double(), C.c_double()
    lib().orc_gen_scales(_d(feats), C.c_size_t(L), C.c_size_t(B), C.byref(mean), C.byref(sd))
    return mean.value, sd.value


def standardise(feats, mean, sd):
    out = np.ascontiguousarray(feats, dtype=np.float64).copy()
    lib().orc_standardise(_d(out), C.c_size_t(out.size), C.c_double(mean), C.c_double(sd))
    return out


def layer_shapes(cfg, ff, classes, l):
    rows = (C.c_size_t * (len(ff) + 1))()
    cols = (C.c_size_t * (len(ff) + 1))()
    _chk(lib().orc_layer_shapes(_ints(cfg), C.c_size_t(len(cfg)), _sz(ff), C.c_size_t(len(ff)), C.c_size_t(classes),
                                C.c_size_t(l), rows, cols))
    return [(rows[i], cols[i]) for i in range(len(ff) + 1)]


def sigmoid(x):
    return lib().orc_sigmoid(float(x))


def sigmoid_prime(x):
    return lib().orc_sigmoid_prime(float(x))


class Net:
    """Flat-parameter view: params = [W0 (col-major) | b0 | W1 | b1 | ...] (serialization.rs:19-22 order)."""

    def __init__(self, shapes):
        self.shapes = [(int(r), int(c)) for r, c in shapes]
        self.rows = _sz([r for r, _ in self.shapes])
        self.cols = _sz([c for _, c in self.shapes])
        self.n = len(self.shapes)
        self.n_params = sum(r * c + r for r, c in self.shapes)
        self.sum_rows = sum(r for r, _ in self.shapes)

    def pack(self, weights, biases):
        parts = []
        for W, b in zip(weights, biases):
            parts.append(_cm(W))
            parts.append(np.asarray(b, dtype=np.float64).ravel())
        return np.concatenate(parts)

    def unpack(self, flat):
        ws, bs, o = [], [], 0
        for r, c in self.shapes:
            ws.append(_from_cm(flat[o:o + r * c], r, c))
            o += r * c
            bs.append(np.array(flat[o:o + r]))
            o += r
        return ws, bs

    def forward(self, params, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        B, n_in = X.shape
        out = np.zeros((B, self.shapes[-1][0]))
        _chk(lib().orc_forward(self.rows, self.cols, C.c_size_t(self.n), _d(params), _d(X), C.c_size_t(n_in),
                               C.c_size_t(B), _d(out)))
        return out

    def backprop(self, params, x, y):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        g = np.zeros(self.n_params)
        zs, acts, deltas = np.zeros(self.sum_rows), np.zeros(self.sum_rows), np.zeros(self.sum_rows)
        _chk(lib().orc_backprop(self.rows, self.cols, C.c_size_t(self.n), _d(params), _d(x), C.c_size_t(x.size),
                                _d(y), _d(g), _d(zs), _d(acts), _d(deltas)))
        return g, zs, acts, deltas

    def train_batch(self, params, X, Y, eta, n_threads=1):
        """Returns (new_params, grad_sum). X: (B, n_in); Y: (B, classes) one-hot."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        B, n_in = X.shape
        p = np.array(params, dtype=np.float64, copy=True)
        g = np.zeros(self.n_params)
        _chk(lib().orc_train_batch(self.rows, self.cols, C.c_size_t(self.n), _d(p), _d(X), C.c_size_t(n_in), _d(Y),
                                   C.c_size_t(B), C.c_double(eta), int(n_threads), _d(g)))
        return p, g

    def train_step_u8(self, cfg, params, images, labels, mean, sd, eta, n_threads):
        """In-place on params (timed CPU baseline leg)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        labels = np.ascontiguousarray(labels, dtype=np.int64)
        B, H, W = images.shape
        L = self.shapes[0][1]
        if not hasattr(self, "_scratch") or self._scratch[0].shape[0] < B:
            self._scratch = (np.zeros((B, L)), np.zeros((B, self.shapes[-1][0])))
        f, oh = self._scratch
        _chk(lib().orc_train_step_u8(_ints(cfg), C.c_size_t(len(cfg)), self.rows, self.cols, C.c_size_t(self.n),
                                     _d(params), images.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     labels.ctypes.data_as(C.POINTER(C.c_int64)), C.c_size_t(B), C.c_size_t(H),
                                     C.c_size_t(W), C.c_double(mean), C.c_double(sd), C.c_double(eta),
                                     int(n_threads), _d(f), _d(oh)))


def argmax_last(acts):
    acts = np.ascontiguousarray(acts, dtype=np.float64)
    B, n = acts.shape
    out = np.zeros(B, dtype=np.int64)
    lib().orc_argmax_last(_d(acts), C.c_size_t(n), C.c_size_t(B), out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def accuracy(acts, labels):
    acts = np.ascontiguousarray(acts, dtype=np.float64)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    B, n = acts.shape
    return int(lib().orc_accuracy(_d(acts), C.c_size_t(n), C.c_size_t(B), labels.ctypes.data_as(C.POINTER(C.c_int64))))
