"""CPU tests of the EXTENSION oracle (oracle/ext_oracle.cpp): the definitions are self-consistent -- every backward
function is the numerical gradient of its forward function -- and pin the conventions (cross-correlation, padding,
pool window order / last-max-wins, divisor 4). Parity unpinned: the reference implements none of these ops."""
import numpy as np
import pytest

import oracle
import oracle.ext as E


def numgrad(f, x, dz, eps=1e-6):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        g[i] = ((f(xp) - f(xm)) * dz).sum() / (2 * eps)
    return g


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("k", [(3, 3), (1, 3), (5, 3)])
def test_conv_backward_is_gradient_of_forward(pad, k):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 6, 5, 2))
    w = rng.standard_normal((3, k[0], k[1], 2))
    b = rng.standard_normal(3)
    y = E.conv2d_forward(x, w, b, pad)
    dz = rng.standard_normal(y.shape)
    dx = E.conv2d_backward_data(dz, w, (6, 5), pad)
    dw, db = E.conv2d_backward_weight(x, dz, k[0], k[1], pad)
    assert np.allclose(dx, numgrad(lambda t: E.conv2d_forward(t, w, b, pad), x, dz), atol=1e-6)
    assert np.allclose(dw, numgrad(lambda t: E.conv2d_forward(x, t, b, pad), w, dz), atol=1e-6)
    assert np.allclose(db, dz.sum((0, 1, 2)))


def test_conv_matches_reference_convolve_2d_on_one_channel():
    """With one input and one output channel the extension IS Convolve2D::convolve_2d (kernel.rs:110-194) for 3x3."""
    rng = np.random.default_rng(2)
    m = rng.integers(-5, 6, size=(7, 9)).astype(np.float64)
    k = rng.integers(-2, 3, size=(3, 3)).astype(np.float64)
    for pad in (0, 1):
        want = oracle.convolve_2d(m, k, pad)
        got = E.conv2d_forward(m[None, :, :, None], k[None, :, :, None], None, pad)[0, :, :, 0]
        assert np.array_equal(got, want)


def test_activation_backward():
    rng = np.random.default_rng(3)
    z = rng.standard_normal(50)
    dy = rng.standard_normal(50)
    s = 1 / (1 + np.exp(-z))
    assert np.allclose(E.activation_backward(s, dy, E.ACT_SIGMOID), dy * s * (1 - s))
    r = np.maximum(z, 0)
    assert np.array_equal(E.activation_backward(r, dy, E.ACT_RELU), dy * (r > 0))
    assert np.array_equal(E.activation_backward(z, dy, E.ACT_NONE), dy)


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("hw", [(6, 4), (7, 5)])
def test_pool_matches_reference_pool_2d_per_channel(pad, hw):
    rng = np.random.default_rng(4)
    x = rng.integers(0, 4, size=(2, hw[0], hw[1], 3)).astype(np.float64)   # many ties
    y, am = E.pool2d_forward(x, pad, oracle.POOL_MAX)
    for b in range(2):
        for c in range(3):
            want, wam = oracle.pool_2d(x[b, :, :, c], pad, oracle.POOL_MAX, return_argmax=True)
            assert np.array_equal(y[b, :, :, c], want) and np.array_equal(am[b, :, :, c], wam)


@pytest.mark.parametrize("pad", [0, 1])
def test_pool_backward(pad):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 5, 7, 2))
    for pooling in (oracle.POOL_MAX, oracle.POOL_AVERAGE):
        y, am = E.pool2d_forward(x, pad, pooling)
        dy = rng.standard_normal(y.shape)
        dx = E.pool2d_backward(dy, am, (5, 7), pad, pooling)
        num = numgrad(lambda t: E.pool2d_forward(t, pad, pooling)[0], x, dy, eps=1e-7)
        assert np.allclose(dx, num, atol=1e-6)
    avg, _ = E.pool2d_forward(np.ones((1, 3, 3, 1)), 1, oracle.POOL_AVERAGE)
    assert np.array_equal(avg[0, :, :, 0], [[1.0, 0.5], [0.5, 0.25]])       # zero padding counts: divisor is always 4


def test_softmax_xent():
    rng = np.random.default_rng(6)
    z = rng.standard_normal((6, 10)) * 5
    lab = np.arange(6) % 10
    p, loss, d = E.softmax_xent(z, labels=lab)
    e = np.exp(z - z.max(1, keepdims=True))
    sm = e / e.sum(1, keepdims=True)
    assert np.allclose(p, sm, rtol=1e-14) and np.allclose(loss, -np.log(sm[np.arange(6), lab]), rtol=1e-13)
    assert np.allclose(d, sm - np.eye(10)[lab], atol=1e-15)
    p2, loss2, d2 = E.softmax_xent(z, onehot=np.eye(10)[lab])
    assert np.array_equal(p, p2) and np.array_equal(loss, loss2) and np.array_equal(d, d2)
    big = np.array([[1000.0, 0.0, -1000.0]])
    pb, lb, _ = E.softmax_xent(big, labels=[2])
    assert np.isfinite(pb).all() and np.isclose(lb[0], 2000.0)
