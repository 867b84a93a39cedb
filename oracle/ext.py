"""ctypes binding of the EXTENSION oracles in ext_oracle.cpp (TEST INFRASTRUCTURE ONLY; parity unpinned -- the
reference implements none of these ops, see the header of ext_oracle.cpp). Tensors are NHWC float64 numpy arrays;
dense-head matrices are (B, n) C-contiguous == (n x B) column-major."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _chk, _d, lib

ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
_z = C.c_size_t


def _out_hw(H, W, kh, kw, padding):
    return (H, W) if padding == 1 else (H - kh + 1, W - kw + 1)


def _pool_hw(H, W, padding):
    return ((H + 1) // 2, (W + 1) // 2) if padding == 1 else (H // 2, W // 2)


def conv2d_forward(x, w, bias, padding, act=ACT_NONE):
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    B, H, W, Ci = x.shape
    Co, kh, kw, _ = w.shape
    Ho, Wo = _out_hw(H, W, kh, kw, padding)
    y = np.zeros((B, Ho, Wo, Co))
    bp = _d(np.ascontiguousarray(bias, dtype=np.float64)) if bias is not None else None
    _chk(lib().ext_conv2d_forward(_d(x), _z(B), _z(H), _z(W), _z(Ci), _d(w), bp, _z(Co), _z(kh), _z(kw), int(padding),
                                  int(act), _d(y)))
    return y


def activation_backward(y, dy, act):
    y = np.ascontiguousarray(y, dtype=np.float64)
    dy = np.ascontiguousarray(dy, dtype=np.float64)
    dz = np.zeros_like(y)
    lib().ext_activation_backward(_d(y), _d(dy), _z(y.size), int(act), _d(dz))
    return dz


def conv2d_backward_data(dz, w, in_hw, padding):
    dz = np.ascontiguousarray(dz, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    B = dz.shape[0]
    Co, kh, kw, Ci = w.shape
    H, W = in_hw
    dx = np.zeros((B, H, W, Ci))
    _chk(lib().ext_conv2d_backward_data(_d(dz), _z(B), _z(H), _z(W), _z(Ci), _d(w), _z(Co), _z(kh), _z(kw), int(padding),
                                        _d(dx)))
    return dx


def conv2d_backward_weight(x, dz, kh, kw, padding):
    x = np.ascontiguousarray(x, dtype=np.float64)
    dz = np.ascontiguousarray(dz, dtype=np.float64)
    B, H, W, Ci = x.shape
    Co = dz.shape[3]
    dw = np.zeros((Co, kh, kw, Ci))
    db = np.zeros(Co)
    _chk(lib().ext_conv2d_backward_weight(_d(x), _d(dz), _z(B), _z(H), _z(W), _z(Ci), _z(Co), _z(kh), _z(kw),
                                          int(padding), _d(dw), _d(db)))
    return dw, db


def pool2d_forward(x, padding, pooling):
    x = np.ascontiguousarray(x, dtype=np.float64)
    B, H, W, Cc = x.shape
    Ho, Wo = _pool_hw(H, W, padding)
    y = np.zeros((B, Ho, Wo, Cc))
    am = np.zeros((B, Ho, Wo, Cc), dtype=np.uint8)
    _chk(lib().ext_pool2d_forward(_d(x), _z(B), _z(H), _z(W), _z(Cc), int(padding), int(pooling), _d(y),
                                  am.ctypes.data_as(C.POINTER(C.c_uint8))))
    return y, am


def pool2d_backward(dy, argmax, in_hw, padding, pooling):
    dy = np.ascontiguousarray(dy, dtype=np.float64)
    B, _, _, Cc = dy.shape
    H, W = in_hw
    dx = np.zeros((B, H, W, Cc))
    am = np.ascontiguousarray(argmax, dtype=np.uint8) if argmax is not None else np.zeros(1, dtype=np.uint8)
    _chk(lib().ext_pool2d_backward(_d(dy), am.ctypes.data_as(C.POINTER(C.c_uint8)), _z(B), _z(H), _z(W), _z(Cc),
                                   int(padding), int(pooling), _d(dx)))
    return dx


def softmax_xent(z, onehot=None, labels=None):
    """z: (B, n). Returns (probs (B, n), loss (B,), delta (B, n))."""
    z = np.ascontiguousarray(z, dtype=np.float64)
    B, n = z.shape
    p, loss, delta = np.zeros((B, n)), np.zeros(B), np.zeros((B, n))
    oh = _d(np.ascontiguousarray(onehot, dtype=np.float64)) if onehot is not None else None
    lb = np.ascontiguousarray(labels, dtype=np.int64).ctypes.data_as(C.POINTER(C.c_int64)) if labels is not None else None
    _chk(lib().ext_softmax_xent(_d(z), _z(n), _z(B), oh, lb, _d(p), _d(loss), _d(delta)))
    return p, loss, delta
