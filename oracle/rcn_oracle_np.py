"""Independent numpy / pure-Python restatement of rcn's CPU hot path (TEST INFRASTRUCTURE ONLY).

A second, deliberately naive restatement of the same reference lines as ``rcn_oracle.cpp``; the two
are cross-checked against each other in ``tests/test_oracle.py`` so that a transcription slip in one
is caught by the other. Loops are literal (small cases only). Every function cites the reference
file:line (relative to /root/reference/) it follows.

PARITY STATUS: "parity unpinned" except for the two result-pinning tests the reference owns
(kernel.rs:400-417 and kernel.rs:434-441) -- see the header of rcn_oracle.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PAD_NONE, PAD_SAME = 0, 1
POOL_AVERAGE, POOL_MAX = 0, 1
OP_TOP, OP_BOTTOM, OP_LEFT, OP_RIGHT = 0, 1, 2, 3
LAYER_CONV_NONE, LAYER_CONV_SAME, LAYER_POOL_AVERAGE, LAYER_POOL_MAX = 0, 1, 2, 3
SEP_OPS = (OP_TOP, OP_LEFT, OP_RIGHT, OP_BOTTOM)  # rcn.rs:41-46


class RefPanic(Exception):
    """The reference would panic!() here."""


def sobel_separated(op, dtype=np.float64):
    """kernel.rs:38-53 -> (3x1, 1x3)."""
    one, two, neg, zero = 1, 2, -1, 0
    if op == OP_TOP:
        v, h = [one, zero, neg], [one, two, one]
    elif op == OP_BOTTOM:
        v, h = [neg, zero, one], [one, two, one]
    elif op == OP_LEFT:
        v, h = [one, two, one], [one, zero, neg]
    else:
        v, h = [one, two, one], [neg, zero, one]
    return np.array(v, dtype=dtype).reshape(3, 1), np.array(h, dtype=dtype).reshape(1, 3)


# kernel.rs:56-59
TOP_SOBEL = np.array([[1.0, 2.0, 1.0], [0.0, 0.0, 0.0], [-1.0, -2.0, -1.0]])
BOTTOM_SOBEL = np.array([[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]])
LEFT_SOBEL = np.array([[1.0, 0.0, -1.0], [2.0, 0.0, -2.0], [1.0, 0.0, -1.0]])
RIGHT_SOBEL = np.array([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]])


def convolve_2d(x, k, padding):
    """kernel.rs:110-194, literal loops. x, k are 2-D arrays indexed (row, col)."""
    x = np.asarray(x)
    k = np.asarray(k)
    H, W = x.shape
    kh, kw = k.shape
    if (kh, kw) == (0, 0) or kh > H or kw > W or kh == 0 or kw == 0:  # kernel.rs:123-128
        raise RefPanic("convolve_2d shape")
    if (kh % 2 == 0 or kw % 2 == 0) and padding == PAD_SAME:  # kernel.rs:131-135
        raise RefPanic("even kernel with SAME")
    if padding == PAD_SAME:
        ph, pw = kh // 2, kw // 2
        P = np.zeros((H + 2 * ph, W + 2 * pw), dtype=x.dtype)
        for cy in range(1, H + ph):  # kernel.rs:154-158
            for cx in range(1, W + pw):
                if cy - 1 >= H or cx - 1 >= W:
                    raise RefPanic("index out of bounds")
                P[cy, cx] = x[cy - 1, cx - 1]
        conv = np.zeros((H, W), dtype=x.dtype)
        for cy in range(H):  # kernel.rs:160-168
            for cx in range(W):
                acc = x.dtype.type(0)
                for ky in range(kh):
                    for kx in range(kw):
                        acc = acc + P[cy + ky, cx + kx] * k[ky, kx]
                conv[cy, cx] = acc
        return conv
    oh, ow = H - kh + 1, W - kw + 1  # kernel.rs:171-192
    conv = np.zeros((oh, ow), dtype=x.dtype)
    for cy in range(oh):
        for cx in range(ow):
            acc = x.dtype.type(0)
            for ky in range(kh):
                for kx in range(kw):
                    acc = acc + x[cy + ky, cx + kx] * k[ky, kx]
            conv[cy, cx] = acc
    return conv


def relu(x):
    """kernel.rs:209-216"""
    x = np.asarray(x)
    out = x.copy()
    out[~(x >= 0)] = 0
    return out


def convolve_2d_separated(x, op, padding):
    """kernel.rs:196-207"""
    x = np.asarray(x)
    if x.shape[0] < 3 or x.shape[1] < 3:
        raise RefPanic("convolve_2d_separated shape")
    v, h = sobel_separated(op, x.dtype)
    return relu(convolve_2d(convolve_2d(x, v, padding), h, padding))


def pool_2d(x, padding, pooling, return_argmax=False):
    """kernel.rs:245-349. argmax (2*dy+dx, last maximal element wins) is this repo's extension."""
    x = np.asarray(x)
    H, W = x.shape
    if H < 2 or W < 2:
        raise RefPanic("pool_2d shape")
    rp = cp = 0
    if padding == PAD_SAME:
        rp, cp = H % 2, W % 2
    P = np.zeros((H + rp, W + cp), dtype=x.dtype)
    P[:H, :W] = x
    oh, ow = (H + rp) // 2, (W + cp) // 2
    if pooling != POOL_MAX:
        raise RefPanic("Not implemented")
    res = np.zeros((oh, ow), dtype=x.dtype)
    arg = np.zeros((oh, ow), dtype=np.uint8)
    for ry in range(oh):
        for rx in range(ow):
            pooler = [None] * 4
            for px in range(2):
                for py in range(2):
                    pooler[py + px * 2] = P[ry * 2 + px, rx * 2 + py]
            best = 0
            for i in range(1, 4):  # Iterator::max_by: later element wins unless strictly smaller
                if math.isnan(pooler[i]) or math.isnan(pooler[best]):
                    raise RefPanic("partial_cmp on NaN")
                if not (pooler[i] < pooler[best]):
                    best = i
            res[ry, rx] = pooler[best]
            arg[ry, rx] = best
    return (res, arg) if return_argmax else res


def feature_maps(cfg, m):
    """rcn.rs:317-348 -> list of maps in the reference's slot order."""
    fs = []
    for layer in cfg:
        if layer in (LAYER_CONV_NONE, LAYER_CONV_SAME):
            p = PAD_SAME if layer == LAYER_CONV_SAME else PAD_NONE
            if fs:
                curr_len = len(fs)
                for i in range(curr_len):
                    for oi, op in enumerate(SEP_OPS):
                        r = convolve_2d_separated(fs[i], op, p)
                        if oi == len(SEP_OPS) - 1:
                            fs[i] = r
                        else:
                            fs.append(r)
            else:
                fs.extend(convolve_2d_separated(m, op, p) for op in SEP_OPS)
        else:
            pooling = POOL_MAX if layer == LAYER_POOL_MAX else POOL_AVERAGE
            fs = [pool_2d(f, PAD_SAME, pooling) for f in fs]
    return fs


def flatten_feature_set(cfg, m):
    """rcn.rs:317-356: concat of column-major maps."""
    fs = feature_maps(cfg, np.asarray(m, dtype=np.float64))
    if not fs:
        return np.zeros(0)
    return np.concatenate([f.flatten(order="F") for f in fs])


def gen_scales(feats):
    """rcn.rs:230-251, sequential sums. feats: iterable of 1-D vectors."""
    mean = 0.0
    sd = 0.0
    n = float(len(feats[0])) * float(len(feats))
    for v in feats:
        for r in range(len(v)):
            mean += float(v[r])
    mean /= n
    for v in feats:
        for r in range(len(v)):
            d = float(v[r]) - mean
            sd += d * d
    sd = math.sqrt(sd / n)
    return mean, sd


def standardise(v, mean, sd):
    """rcn.rs:407-412"""
    d = (np.asarray(v, dtype=np.float64) - mean) / sd
    return np.where(d >= 0, d, 0.0)


def layer_shapes(cfg, ff, classes, l):
    """rcn.rs:425-457 -> [(rows, cols)]"""
    c = sum(1 for x in cfg if x in (LAYER_CONV_NONE, LAYER_CONV_SAME))
    p = sum(2 for x in cfg if x in (LAYER_POOL_AVERAGE, LAYER_POOL_MAX))
    a = (4 ** c) // (2 ** p) * l
    b = ff[0]
    shapes = []
    for i in range(len(ff) + 1):
        shapes.append((b, a))
        a, b = b, (ff[i + 1] if i + 1 < len(ff) else classes)
    return shapes


def sigmoid(v):
    """rcn.rs:478-483"""
    return np.array([1.0 / (1.0 + math.pow(math.e, -float(x))) if -float(x) < 709.0 else 0.0 for x in v])


def sigmoid_prime(v):
    """rcn.rs:490-492"""
    s = sigmoid(v)
    return s * (1.0 - s)


def _matvec(W, x):
    """nalgebra 0.31 gemv: column sweep, separate mul and add."""
    y = W[:, 0] * x[0]
    for j in range(1, W.shape[1]):
        y = W[:, j] * x[j] + y
    return y


def classify_test(weights, biases, x):
    """rcn.rs:105-116"""
    a = np.asarray(x, dtype=np.float64)
    for W, b in zip(weights, biases):
        a = sigmoid(_matvec(W, a) + b)
    return a


def backprop(weights, biases, x, y):
    """rcn.rs:260-314 -> (del_b list, del_w list, zs, activations, deltas)"""
    L = len(weights)
    a = np.asarray(x, dtype=np.float64)
    activations = [a]
    zs = []
    for W, b in zip(weights, biases):
        z = _matvec(W, a) + b
        zs.append(z)
        a = sigmoid(z)
        activations.append(a)
    del_b = [None] * L
    del_w = [None] * L
    deltas = [None] * L
    delta = (activations[-1] - y) * sigmoid_prime(zs[-1])
    del_b[-1] = delta
    del_w[-1] = np.outer(delta, activations[-2])
    deltas[-1] = delta
    for l in range(1, L):
        sp = sigmoid_prime(zs[-1 - l])
        delta = _matvec(np.ascontiguousarray(weights[L - l].T), delta) * sp
        del_b[-1 - l] = delta
        del_w[-1 - l] = np.outer(delta, activations[-2 - l])
        deltas[-1 - l] = delta
    return del_b, del_w, zs, activations, deltas


def train_batch(weights, biases, X, Y, eta):
    """rcn.rs:176-223, samples summed in index order. X: list of inputs, Y: list of one-hots."""
    acc_w = [np.zeros_like(W) for W in weights]
    acc_b = [np.zeros_like(b) for b in biases]
    for x, y in zip(X, Y):
        db, dw, *_ = backprop(weights, biases, x, y)
        acc_b = [d + a for d, a in zip(db, acc_b)]
        acc_w = [d + a for d, a in zip(dw, acc_w)]
    scale = eta / float(len(X))
    new_w = [W - scale * g for W, g in zip(weights, acc_w)]
    new_b = [b - scale * g for b, g in zip(biases, acc_b)]
    return new_w, new_b, acc_w, acc_b


def argmax_last(a):
    """rcn.rs:92-97"""
    best = 0
    for i in range(1, len(a)):
        if not (a[i] < a[best]):
            best = i
    return best


def accuracy_hit(a, label):
    """rcn.rs:153-157"""
    mx = max(a)
    res = [1.0 if v == mx else 0.0 for v in a]
    exp = [1.0 if i == label else 0.0 for i in range(len(a))]
    return res == exp
