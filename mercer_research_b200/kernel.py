"""Host-side mirror of rcn's op layer (rcn/src/utils/kernel.rs): the ``Convolve2D`` and ``Pool2D`` extension
traits, and the ``Padding`` / ``Pooling`` / ``SeparableOperator`` enums -- same names, argument meaning and
error behaviour, executed by librcn_cuda.so on a B200.

Matrices are 2-D arrays indexed ``(row, col)`` like ``nalgebra::DMatrix``; numpy arrays are staged through the
device, torch CUDA tensors are used in place. Where the reference ``panic!``s, ``RcnCudaError`` is raised with the
reference's message.
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib


class SeparableOperator(enum.IntEnum):
    """kernel.rs:16-21"""
    Top = 0
    Bottom = 1
    Left = 2
    Right = 3


class Padding(enum.IntEnum):
    """kernel.rs:25-28"""
    None_ = 0
    Same = 1


class Pooling(enum.IntEnum):
    """kernel.rs:32-35"""
    Average = 0
    Max = 1


def sobel_separated(op: SeparableOperator):
    """kernel.rs:38-53 -> (3x1, 1x3) as numpy arrays."""
    op = SeparableOperator(op)
    if op == SeparableOperator.Top:
        v, h = [1.0, 0.0, -1.0], [1.0, 2.0, 1.0]
    elif op == SeparableOperator.Bottom:
        v, h = [-1.0, 0.0, 1.0], [1.0, 2.0, 1.0]
    elif op == SeparableOperator.Left:
        v, h = [1.0, 2.0, 1.0], [1.0, 0.0, -1.0]
    else:
        v, h = [1.0, 2.0, 1.0], [-1.0, 0.0, 1.0]
    return np.array(v).reshape(3, 1), np.array(h).reshape(1, 3)


# kernel.rs:56-59
TOP_SOBEL = np.array([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]])
BOTTOM_SOBEL = np.array([[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]])
LEFT_SOBEL = np.array([[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]])
RIGHT_SOBEL = np.array([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Mat:
    """Column-major view of a (row, col) matrix for the C ABI; keeps the backing storage alive."""

    def __init__(self, m, dtype=np.float64):
        if _is_torch(m):
            import torch
            if m.dim() != 2:
                raise ValueError("expected a 2-D matrix")
            self.torch = True
            self.H, self.W = int(m.shape[0]), int(m.shape[1])
            self.store = m.to(torch.float64).t().contiguous()  # (W, H) contiguous == H x W column-major
            self.ptr = self.store.data_ptr()
            self.device = self.store.device
        else:
            a = np.asarray(m, dtype=dtype)
            if a.ndim != 2:
                raise ValueError("expected a 2-D matrix")
            self.torch = False
            self.H, self.W = a.shape
            self.store = np.asfortranarray(a)
            self.ptr = self.store.ctypes.data
            self.device = None


def _out(like: _Mat, oh: int, ow: int, dtype=np.float64):
    """Allocates a column-major output; returns (ptr, finish()) where finish() yields the (row, col) matrix."""
    if like.torch:
        import torch
        tdt = torch.float64 if dtype == np.float64 else torch.uint8
        buf = torch.empty((ow, oh), dtype=tdt, device=like.device)
        return buf.data_ptr(), (lambda: buf.t())
    buf = np.zeros((oh, ow), dtype=dtype, order="F")
    return buf.ctypes.data, (lambda: buf)


def _stream(like: _Mat):
    if like.torch:
        import torch
        return torch.cuda.current_stream(like.device).cuda_stream, (like.device.index or 0)
    return None, 0


def convolve_2d(m, kernel, padding: Padding, device: int = 0):
    """``Convolve2D::convolve_2d`` (kernel.rs:110-194): 2-D cross-correlation with ``kernel``."""
    lib = _lib.load()
    a, k = _Mat(m), _Mat(kernel)
    same = Padding(padding) == Padding.Same
    oh, ow = (a.H, a.W) if same else (a.H - k.H + 1, a.W - k.W + 1)
    stream, dev = _stream(a)
    if not a.torch:
        dev = device
    optr, fin = _out(a, max(oh, 0), max(ow, 0))
    _lib.check(lib.rcn_cuda_convolve_2d(dev, stream, a.ptr, a.H, a.W, k.ptr, k.H, k.W, int(padding), optr))
    return fin()


def convolve_2d_separated(m, op: SeparableOperator, padding: Padding, device: int = 0):
    """``Convolve2D::convolve_2d_separated`` (kernel.rs:196-207): 3x1 then 1x3 Sobel pass, then ReLU."""
    lib = _lib.load()
    a = _Mat(m)
    same = Padding(padding) == Padding.Same
    oh, ow = (a.H, a.W) if same else (a.H - 2, a.W - 2)
    stream, dev = _stream(a)
    if not a.torch:
        dev = device
    optr, fin = _out(a, max(oh, 0), max(ow, 0))
    _lib.check(lib.rcn_cuda_convolve_2d_separated(dev, stream, a.ptr, a.H, a.W, int(op), int(padding), optr))
    return fin()


def relu(m, device: int = 0):
    """``Convolve2D::relu`` (kernel.rs:209-216)."""
    lib = _lib.load()
    a = _Mat(m)
    stream, dev = _stream(a)
    if not a.torch:
        dev = device
    optr, fin = _out(a, a.H, a.W)
    _lib.check(lib.rcn_cuda_relu(dev, stream, a.ptr, a.H * a.W, optr))
    return fin()


def pool_2d(m, padding: Padding, pooling: Pooling, return_argmax: bool = False, device: int = 0):
    """``Pool2D::pool_2d`` (kernel.rs:245-349): 2x2 / stride-2 max pooling (``Pooling::Average`` raises "Not
    implemented" like the reference). ``return_argmax`` additionally returns the index ``2*dy+dx`` of the chosen
    element (last maximal element wins) -- an extension the reference does not have."""
    lib = _lib.load()
    a = _Mat(m)
    same = Padding(padding) == Padding.Same
    oh = (a.H + 1) // 2 if same else a.H // 2
    ow = (a.W + 1) // 2 if same else a.W // 2
    stream, dev = _stream(a)
    if not a.torch:
        dev = device
    optr, fin = _out(a, oh, ow)
    aptr, afin = (None, None)
    if return_argmax:
        aptr, afin = _out(a, oh, ow, dtype=np.uint8)
    _lib.check(lib.rcn_cuda_pool_2d(dev, stream, a.ptr, a.H, a.W, int(padding), int(pooling), optr, aptr))
    return (fin(), afin()) if return_argmax else fin()
