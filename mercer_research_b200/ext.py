"""EXTENSION ops -- operations BASELINE.json's north_star names that the reference (rcn) does NOT implement
(SURVEY.md section 8a rows x1-x3): learned convolution forward / backward-data / backward-weight, NHWC average / max
pooling with backward, softmax + cross-entropy. "Parity unpinned": checked against oracle/ext_oracle.cpp only.

Tensors with channels are NHWC float64 ``(B, H, W, C)``; weights ``(Co, kh, kw, Ci)``; dense-head matrices ``(B, n)``
C-contiguous (== n x B column-major, sample contiguous). numpy arrays are staged through the device, torch CUDA tensors
are used in place on torch's current stream. Conventions extend the reference's own (cross-correlation, Padding,
2x2 / stride-2 pooling window with the last-maximal-element rule) -- see include/rcn_cuda.h.
"""
from __future__ import annotations

import enum

import numpy as np

from . import _lib
from .kernel import Padding, Pooling, _is_torch


class Activation(enum.IntEnum):
    None_ = 0
    ReLU = 1
    Sigmoid = 2


class _T:
    """Pointer + keep-alive for a contiguous float64 / uint8 / int64 host or device tensor."""

    def __init__(self, x, dtype=np.float64):
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise ValueError("torch tensors must live on a CUDA device (pass numpy arrays for host data)")
            tdt = {np.float64: torch.float64, np.uint8: torch.uint8, np.int64: torch.int64}[dtype]
            self.store = x.to(tdt).contiguous()
            self.ptr = self.store.data_ptr()
            self.torch, self.device = True, x.device
        else:
            self.store = np.ascontiguousarray(x, dtype=dtype)
            self.ptr = self.store.ctypes.data
            self.torch, self.device = False, None
        self.shape = tuple(self.store.shape)


def _ctx(*ts):
    """(device index, stream handle, allocator for outputs) from the first torch tensor, else device 0 / host outputs."""
    for t in ts:
        if t is not None and t.torch:
            import torch
            dev = t.device

            def alloc(shape, dtype=np.float64):
                tdt = {np.float64: torch.float64, np.uint8: torch.uint8}[dtype]
                return torch.empty(shape, dtype=tdt, device=dev)
            return dev.index or 0, torch.cuda.current_stream(dev).cuda_stream, alloc
    return 0, None, lambda shape, dtype=np.float64: np.zeros(shape, dtype=dtype)


def _ptr(o):
    return o.data_ptr() if _is_torch(o) else o.ctypes.data


def conv_out_hw(H, W, kh, kw, padding):
    return (H, W) if int(padding) == Padding.Same else (H - kh + 1, W - kw + 1)


def pool_out_hw(H, W, padding):
    return ((H + 1) // 2, (W + 1) // 2) if int(padding) == Padding.Same else (H // 2, W // 2)


def conv2d_forward(x, w, bias=None, padding=Padding.Same, activation=Activation.None_, device=None):
    X, Wt = _T(x), _T(w)
    Bi = _T(bias) if bias is not None else None
    B, H, W, Ci = X.shape
    Co, kh, kw, Ci2 = Wt.shape
    if Ci2 != Ci:
        raise ValueError(f"weight has {Ci2} input channels, x has {Ci}")
    dev, stream, alloc = _ctx(X, Wt)
    Ho, Wo = conv_out_hw(H, W, kh, kw, padding)
    y = alloc((B, max(Ho, 0), max(Wo, 0), Co))
    _lib.check(_lib.load().rcn_cuda_ext_conv2d_forward(dev if device is None else device, stream, X.ptr, B, H, W, Ci, Wt.ptr,
                                                       Bi.ptr if Bi else None, Co, kh, kw, int(padding), int(activation), _ptr(y)))
    return y


def activation_backward(y, dy, activation):
    Y, D = _T(y), _T(dy)
    dev, stream, alloc = _ctx(Y, D)
    dz = alloc(Y.shape)
    _lib.check(_lib.load().rcn_cuda_ext_activation_backward(dev, stream, Y.ptr, D.ptr, int(np.prod(Y.shape)), int(activation),
                                                            _ptr(dz)))
    return dz


def conv2d_backward_data(dz, w, in_hw, padding=Padding.Same, y_prev=None, activation_prev=Activation.None_):
    D, Wt = _T(dz), _T(w)
    Yp = _T(y_prev) if y_prev is not None else None
    B = D.shape[0]
    Co, kh, kw, Ci = Wt.shape
    H, W = in_hw
    dev, stream, alloc = _ctx(D, Wt)
    dx = alloc((B, H, W, Ci))
    _lib.check(_lib.load().rcn_cuda_ext_conv2d_backward_data(dev, stream, D.ptr, B, H, W, Ci, Wt.ptr, Co, kh, kw, int(padding),
                                                             Yp.ptr if Yp else None, int(activation_prev), _ptr(dx)))
    return dx


def conv2d_backward_weight(x, dz, kh, kw, padding=Padding.Same):
    X, D = _T(x), _T(dz)
    B, H, W, Ci = X.shape
    Co = D.shape[3]
    dev, stream, alloc = _ctx(X, D)
    dw, db = alloc((Co, kh, kw, Ci)), alloc((Co,))
    _lib.check(_lib.load().rcn_cuda_ext_conv2d_backward_weight(dev, stream, X.ptr, D.ptr, B, H, W, Ci, Co, kh, kw, int(padding),
                                                               _ptr(dw), _ptr(db)))
    return dw, db


def pool2d_forward(x, padding=Padding.Same, pooling=Pooling.Max, return_argmax=True):
    X = _T(x)
    B, H, W, Cc = X.shape
    dev, stream, alloc = _ctx(X)
    Ho, Wo = pool_out_hw(H, W, padding)
    y = alloc((B, Ho, Wo, Cc))
    am = alloc((B, Ho, Wo, Cc), np.uint8) if (return_argmax and int(pooling) == Pooling.Max) else None
    _lib.check(_lib.load().rcn_cuda_ext_pool2d_forward(dev, stream, X.ptr, B, H, W, Cc, int(padding), int(pooling), _ptr(y),
                                                       _ptr(am) if am is not None else None))
    return (y, am) if return_argmax else y


def pool2d_backward(dy, argmax, in_hw, padding=Padding.Same, pooling=Pooling.Max):
    D = _T(dy)
    A = _T(argmax, np.uint8) if argmax is not None else None
    B, _, _, Cc = D.shape
    H, W = in_hw
    dev, stream, alloc = _ctx(D)
    dx = alloc((B, H, W, Cc))
    _lib.check(_lib.load().rcn_cuda_ext_pool2d_backward(dev, stream, D.ptr, A.ptr if A else None, B, H, W, Cc, int(padding),
                                                        int(pooling), _ptr(dx)))
    return dx


def softmax_xent(z, onehot=None, labels=None):
    """z: (B, n) logits. Returns (probs (B, n), loss (B,), delta (B, n) = probs - y)."""
    Z = _T(z)
    Oh = _T(onehot) if onehot is not None else None
    Lb = _T(labels, np.int64) if labels is not None else None
    B, n = Z.shape
    dev, stream, alloc = _ctx(Z)
    p, loss, delta = alloc((B, n)), alloc((B,)), alloc((B, n))
    _lib.check(_lib.load().rcn_cuda_ext_softmax_xent(dev, stream, Z.ptr, n, B, Oh.ptr if Oh else None, Lb.ptr if Lb else None,
                                                     _ptr(p), _ptr(loss), _ptr(delta)))
    return p, loss, delta


def gemm_f64(a, b, impl: int = 1, a_kcontig: bool = True, b_kcontig: bool = True):
    """C (M x N) = A (M x K) @ B (K x N) through the library's GEMM building block (include/rcn_cuda.h).
    ``a`` is given as a numpy (M, K) array, ``b`` as (K, N); *_kcontig chooses the memory layout handed to the kernel
    (k-contiguous or row-contiguous), which selects the slicing / loader variant. impl 0 = DMMA, 1 = tcgen05 int8 slices."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    M, K = a.shape
    K2, N = b.shape
    assert K == K2
    abuf = np.ascontiguousarray(a) if a_kcontig else np.ascontiguousarray(a.T)     # [m][k] or [k][m]
    bbuf = np.ascontiguousarray(b.T) if b_kcontig else np.ascontiguousarray(b)     # [n][k] or [k][n]
    lda = K if a_kcontig else M
    ldb = K if b_kcontig else N
    c = np.zeros((N, M))                                                           # column-major M x N
    _lib.check(_lib.load().rcn_cuda_ext_gemm_f64(0, None, abuf.ctypes.data, lda, int(a_kcontig), bbuf.ctypes.data, ldb,
                                                 int(b_kcontig), M, N, K, int(impl), c.ctypes.data))
    return c.T
