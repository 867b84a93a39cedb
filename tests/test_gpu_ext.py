"""GPU parity tests of the EXTENSION ops through the C ABI against oracle/ext_oracle.cpp.
Parity unpinned (the reference implements none of them): f64 values within 1e-9 relative, pool values / argmax bytes
bit-exact."""
import numpy as np
import pytest

import oracle as O
import oracle.ext as E
from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def X(built_library):
    from mercer_research_b200 import ext
    return ext


CONV_CASES = [
    # B, H, W, Ci, Co, kh, kw, pad, act          path
    (2, 9, 7, 1, 5, 3, 3, 1, 1),                 # direct small-channel kernel (first layer, 1 channel)
    (3, 12, 10, 3, 16, 3, 3, 1, 2),              # direct, CIFAR-like 3 channels, sigmoid
    (2, 8, 8, 3, 7, 5, 3, 0, 0),                 # direct, valid padding, rectangular kernel
    (2, 10, 9, 16, 24, 3, 3, 1, 1),              # implicit GEMM (K = 144)
    (1, 7, 6, 20, 40, 3, 3, 0, 0),               # implicit GEMM, K not a multiple of the k-tile, valid padding
    (2, 16, 16, 64, 64, 3, 3, 1, 1),             # implicit GEMM, wide-stack shape (reduced spatial size)
    (4, 6, 5, 8, 136, 1, 1, 1, 2),               # 1x1 kernel, M > 128
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_forward_backward(X, case):
    B, H, W, Ci, Co, kh, kw, pad, act = case
    rng = np.random.default_rng(hash(case) % (2 ** 32))
    x = rng.standard_normal((B, H, W, Ci))
    w = rng.standard_normal((Co, kh, kw, Ci)) / np.sqrt(kh * kw * Ci)
    b = rng.standard_normal(Co)
    y = X.conv2d_forward(x, w, b, pad, act)
    want = E.conv2d_forward(x, w, b, pad, act)
    assert_close(y, want, what="conv forward")
    assert_close(X.conv2d_forward(x, w, None, pad, 0), E.conv2d_forward(x, w, None, pad, 0), what="conv forward, no bias")
    dy = rng.standard_normal(want.shape)
    dz = X.activation_backward(want, dy, act)
    wdz = E.activation_backward(want, dy, act)
    assert_close(dz, wdz, what="activation backward")
    dx = X.conv2d_backward_data(wdz, w, (H, W), pad)
    assert_close(dx, E.conv2d_backward_data(wdz, w, (H, W), pad), what="conv backward-data")
    dw, db = X.conv2d_backward_weight(x, wdz, kh, kw, pad)
    wdw, wdb = E.conv2d_backward_weight(x, wdz, kh, kw, pad)
    assert_close(dw, wdw, what="conv backward-weight")
    assert_close(db, wdb, what="conv bias gradient")


def test_conv2d_backward_data_fused_previous_activation(X):
    rng = np.random.default_rng(7)
    B, H, W, Ci, Co = 2, 8, 8, 16, 16
    yprev = 1 / (1 + np.exp(-rng.standard_normal((B, H, W, Ci))))
    w = rng.standard_normal((Co, 3, 3, Ci))
    dz = rng.standard_normal((B, H, W, Co))
    got = X.conv2d_backward_data(dz, w, (H, W), 1, y_prev=yprev, activation_prev=2)
    want = E.activation_backward(yprev, E.conv2d_backward_data(dz, w, (H, W), 1), 2)
    assert_close(got, want, what="fused backward-data")


def test_conv2d_igemm_equals_direct_kernel(X, monkeypatch):
    """The same small-channel problem through the implicit-GEMM path and the direct kernel (two code paths, one result)."""
    import subprocess, sys, os, json
    code = ("import numpy as np, json, sys; sys.path.insert(0, %r); from mercer_research_b200 import ext;"
            "rng = np.random.default_rng(3); x = rng.standard_normal((2, 9, 9, 3)); w = rng.standard_normal((8, 3, 3, 3));"
            "print(json.dumps(ext.conv2d_forward(x, w, None, 1, 0).tolist()))") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        env = dict(os.environ, RCN_CUDA_CONV_DIRECT=flag)
        outs.append(np.array(json.loads(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                                                       check=True).stdout)))
    assert_close(outs[0], outs[1], what="direct vs implicit GEMM")


def test_conv2d_large_batch_weight_gradient_is_deterministic(X):
    """Split over pixels with a fixed-order combine: two runs are bit-identical; value matches the oracle."""
    rng = np.random.default_rng(8)
    x = rng.standard_normal((16, 16, 16, 32))
    dz = rng.standard_normal((16, 16, 16, 32))
    dw1, db1 = X.conv2d_backward_weight(x, dz, 3, 3, 1)
    dw2, db2 = X.conv2d_backward_weight(x, dz, 3, 3, 1)
    assert np.array_equal(dw1, dw2) and np.array_equal(db1, db2)
    wdw, wdb = E.conv2d_backward_weight(x, dz, 3, 3, 1)
    assert_close(dw1, wdw, what="dw")
    assert_close(db1, wdb, what="db")


def test_conv2d_on_device_tensors(X):
    import torch
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, 8, 8, 16)); w = rng.standard_normal((16, 3, 3, 16)); b = rng.standard_normal(16)
    y = X.conv2d_forward(torch.tensor(x, device="cuda"), torch.tensor(w, device="cuda"), torch.tensor(b, device="cuda"), 1, 1)
    assert y.is_cuda
    torch.cuda.synchronize()
    assert_close(y.cpu().numpy(), E.conv2d_forward(x, w, b, 1, 1), what="device conv")


def test_conv2d_contract_violations(X):
    from mercer_research_b200 import RcnCudaError
    x = np.zeros((1, 4, 4, 2))
    for w, pad in [(np.zeros((2, 2, 2, 2)), 1), (np.zeros((2, 5, 5, 2)), 0)]:
        with pytest.raises(RcnCudaError) as e:
            X.conv2d_forward(x, w, None, pad, 0)
        assert e.value.status == 2


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("shape", [(2, 6, 4, 3), (3, 7, 5, 8), (1, 2, 2, 1), (2, 9, 16, 64)])
def test_pool2d_forward_backward(X, shape, pad):
    rng = np.random.default_rng(10)
    x = rng.integers(-3, 4, size=shape).astype(np.float64)                 # many ties, incl. -0.0 / +0.0 candidates
    x[x == 0] = rng.choice([0.0, -0.0], size=int((x == 0).sum()))
    for pooling in (O.POOL_MAX, O.POOL_AVERAGE):
        y, am = X.pool2d_forward(x, pad, pooling)
        wy, wam = E.pool2d_forward(x, pad, pooling)
        assert np.array_equal(np.asarray(y).view(np.uint64), wy.view(np.uint64)), "pooled values must be bit-exact"
        if pooling == O.POOL_MAX:
            assert np.array_equal(am, wam), "argmax indices must be bit-exact (last maximal element wins)"
        dy = rng.standard_normal(wy.shape)
        dx = X.pool2d_backward(dy, wam if pooling == O.POOL_MAX else None, shape[1:3], pad, pooling)
        assert np.array_equal(dx, E.pool2d_backward(dy, wam, shape[1:3], pad, pooling))


def test_pool2d_nan_and_shape_errors(X):
    from mercer_research_b200 import RcnCudaError
    x = np.zeros((1, 4, 4, 2)); x[0, 1, 1, 0] = np.nan
    with pytest.raises(RcnCudaError) as e:
        X.pool2d_forward(x, 1, O.POOL_MAX)
    assert e.value.status == 7
    with pytest.raises(RcnCudaError) as e:
        X.pool2d_forward(np.zeros((1, 1, 4, 2)), 1, O.POOL_MAX)
    assert e.value.status == 2


@pytest.mark.parametrize("n,B", [(10, 64), (3, 5), (100, 33), (1000, 4)])
def test_softmax_xent(X, n, B):
    rng = np.random.default_rng(11)
    z = rng.standard_normal((B, n)) * 4
    lab = rng.integers(0, n, size=B)
    p, loss, d = X.softmax_xent(z, labels=lab)
    wp, wl, wd = E.softmax_xent(z, labels=lab)
    assert_close(p, wp, what="softmax")
    assert_close(loss, wl, what="cross-entropy")
    assert_close(d, wd, what="softmax-xent delta")
    p2, l2, d2 = X.softmax_xent(z, onehot=np.eye(n)[lab])
    assert np.array_equal(p, p2) and np.array_equal(loss, l2) and np.array_equal(d, d2)
    assert np.array_equal(np.argmax(p, 1), np.argmax(wp, 1))


def test_small_cnn_training_step_matches_oracle(X):
    """conv(3->8, relu) -> maxpool -> conv(8->16, relu) -> maxpool -> dense softmax head: one full forward/backward
    through the extension ops (the CIFAR-shaped 3-channel variant of BASELINE config 3, reduced) against the oracle."""
    rng = np.random.default_rng(12)
    B, H, W = 8, 12, 12
    x = rng.standard_normal((B, H, W, 3))
    w1 = rng.standard_normal((8, 3, 3, 3)) * 0.3; b1 = rng.standard_normal(8) * 0.1
    w2 = rng.standard_normal((16, 3, 3, 8)) * 0.2; b2 = rng.standard_normal(16) * 0.1
    wd = rng.standard_normal((10, 16 * 3 * 3)) * 0.1
    lab = rng.integers(0, 10, size=B)

    def run(M):
        a1 = M.conv2d_forward(x, w1, b1, 1, 1)
        p1, am1 = M.pool2d_forward(a1, 1, O.POOL_MAX)
        a2 = M.conv2d_forward(p1, w2, b2, 1, 1)
        p2, am2 = M.pool2d_forward(a2, 1, O.POOL_MAX)
        flat = np.asarray(p2).reshape(B, -1)
        z = flat @ wd.T
        _, loss, delta = M.softmax_xent(z, labels=lab)
        dflat = np.asarray(delta) @ wd
        dp2 = dflat.reshape(np.asarray(p2).shape)
        da2 = M.pool2d_backward(dp2, am2, np.asarray(a2).shape[1:3], 1, O.POOL_MAX)
        dz2 = M.activation_backward(a2, da2, 1)
        dw2, db2 = M.conv2d_backward_weight(p1, dz2, 3, 3, 1)
        dp1 = M.conv2d_backward_data(dz2, w2, np.asarray(p1).shape[1:3], 1)
        da1 = M.pool2d_backward(dp1, am1, (H, W), 1, O.POOL_MAX)
        dz1 = M.activation_backward(a1, da1, 1)
        dw1, db1 = M.conv2d_backward_weight(x, dz1, 3, 3, 1)
        return [np.asarray(t) for t in (loss, dw2, db2, dw1, db1)]

    got, want = run(X), run(E)
    for g, w_, name in zip(got, want, ["loss", "dw2", "db2", "dw1", "db1"]):
        assert_close(g, w_, what=name)
