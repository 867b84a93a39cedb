"""CPU tests of the oracle: it must reproduce (a) every result-pinning test the reference owns, (b) the derived
known-answer vectors of SURVEY.md Appendix B, (c) an independent numpy restatement, (d) the committed golden files."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import rcn_oracle_np as N

GOLD = os.path.join(os.path.dirname(__file__), "golden")
X65 = np.array([[8, 6, 5, 2, 3], [0, 0, 0, 1, 8], [6, 9, 5, 6, 9], [7, 6, 5, 5, 9], [2, 8, 6, 0, 3], [8, 5, 0, 7, 7]], dtype=float)

KAT = {  # op -> (separated SAME conv, pooled, argmax)
    O.OP_TOP: ([[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [2, 1, 0, 0, 0], [0, 0, 0, 0, 0], [4, 9, 5, 5, 11], [0, 0, 6, 9, 1]],
               [[0, 0, 0], [2, 0, 0], [9, 9, 11]], [[3, 3, 3], [0, 3, 3], [1, 3, 0]]),
    O.OP_LEFT: ([[0, 0, 0, 0, 0], [0, 0, 6, 7, 10], [0, 0, 4, 5, 10], [0, 0, 4, 6, 15], [0, 0, 1, 13, 21], [0, 0, 2, 15, 17]],
                [[0, 7, 10], [0, 6, 15], [0, 15, 21]], [[3, 3, 2], [3, 3, 2], [3, 3, 0]]),
    O.OP_RIGHT: ([[0, 0, 0, 0, 0], [16, 12, 0, 0, 0], [14, 15, 0, 0, 0], [19, 24, 0, 0, 0], [22, 29, 0, 0, 0], [19, 27, 0, 0, 0]],
                 [[16, 0, 0], [24, 0, 0], [29, 0, 0]], [[2, 3, 3], [3, 3, 3], [1, 3, 3]]),
    O.OP_BOTTOM: ([[0, 0, 0, 0, 0], [0, 0, 0, 1, 2], [0, 0, 4, 7, 8], [7, 20, 24, 20, 13], [0, 0, 0, 0, 0], [1, 1, 0, 0, 0]],
                  [[0, 1, 2], [20, 24, 13], [1, 0, 0]], [[3, 3, 2], [3, 2, 2], [3, 3, 3]]),
}


# ---- the reference's own result-pinning tests -------------------------------------------------------------------
def test_ref_verify_separated_sobels():
    """kernel.rs:400-417: 3x1 * 1x3 equals the four full 3x3 Sobel constants."""
    for op in (O.OP_TOP, O.OP_BOTTOM, O.OP_LEFT, O.OP_RIGHT):
        v, h = O.sobel_separated(op)
        assert np.array_equal(v @ h, O.sobel_full(op))
        vn, hn = N.sobel_separated(op)
        assert np.array_equal(v, vn) and np.array_equal(h, hn)
    assert np.array_equal(O.sobel_full(O.OP_TOP), N.TOP_SOBEL)
    assert np.array_equal(O.sobel_full(O.OP_BOTTOM), N.BOTTOM_SOBEL)
    assert np.array_equal(O.sobel_full(O.OP_LEFT), N.LEFT_SOBEL)
    assert np.array_equal(O.sobel_full(O.OP_RIGHT), N.RIGHT_SOBEL)


def test_ref_convolve_2d_padding_same():
    """kernel.rs:434-441: SAME conv of the 30x30 i32 ramp with the 3x3 identity returns the input."""
    m = np.arange(900, dtype=np.int32).reshape(30, 30)  # from_row_iterator
    k = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=np.int32)
    assert np.array_equal(O.convolve_2d(m, k, O.PAD_SAME), m)
    assert np.array_equal(N.convolve_2d(m, k, N.PAD_SAME), m)
    assert np.array_equal(O.convolve_2d(m.astype(float), k.astype(float), O.PAD_SAME), m.astype(float))


def test_ref_validate_padding_calc():
    """kernel.rs:419-432 (shape arithmetic only)."""
    m_size, k_size = (28, 28), (3, 3)
    reg = (m_size[0] - k_size[0] + 1, m_size[1] - k_size[1] + 1)
    pad = (m_size[0] - reg[0], m_size[1] - reg[1])
    assert pad[0] < k_size[0] and pad[1] < k_size[1]
    assert O.convolve_2d(np.zeros(m_size), np.zeros(k_size), O.PAD_NONE).shape == reg


def test_ref_weight_init_shape():
    """rcn.rs:530-538: a (100 -> 32) weight matrix has 32*100 elements; shapes follow rcn.rs:425-457."""
    shapes = O.layer_shapes([], [32], 10, 100)
    assert shapes[0] == (32, 100) and shapes[0][0] * shapes[0][1] == 3200
    # CLI default topology CPCP + [30] on 28x28 (main.rs:51-62): 4^2/2^4 * 784 = 784
    assert O.layer_shapes([1, 3, 1, 3], [30], 10, 784) == [(30, 784), (10, 30)]
    # the integer-division quirk (rcn.rs:443): one conv, two pools -> 4/16 = 0 -> zero-width first layer
    assert O.layer_shapes([1, 3, 3], [30], 10, 196)[0] == (30, 0)
    assert O.layer_shapes([1, 3, 3], [30], 10, 196) == N.layer_shapes([1, 3, 3], [30], 10, 196)
    # two convs, one pool: 16/4*l
    assert O.layer_shapes([1, 1, 3], [7, 5], 3, 10) == [(7, 40), (5, 7), (3, 5)] == N.layer_shapes([1, 1, 3], [7, 5], 3, 10)


# ---- Appendix B known answers ----------------------------------------------------------------------------------
@pytest.mark.parametrize("op", list(KAT))
def test_kat_separated_and_pool(op):
    conv, pooled, arg = KAT[op]
    c = O.convolve_2d_separated(X65, op, O.PAD_SAME)
    assert np.array_equal(c, np.array(conv, dtype=float))
    p, a = O.pool_2d(c, O.PAD_SAME, O.POOL_MAX, return_argmax=True)
    assert np.array_equal(p, np.array(pooled, dtype=float))
    assert np.array_equal(a, np.array(arg, dtype=np.uint8))
    assert np.array_equal(N.convolve_2d_separated(X65, op, N.PAD_SAME), c)
    pn, an = N.pool_2d(c, N.PAD_SAME, N.POOL_MAX, return_argmax=True)
    assert np.array_equal(pn, p) and np.array_equal(an, a)


def test_kat_unquirked_contrast():
    """Appendix B: the plain 3x3 SAME Top Sobel differs from the separated result (shifted by (+1,+1))."""
    full = O.convolve_2d(X65, O.sobel_full(O.OP_TOP), O.PAD_SAME)
    assert full[1].tolist() == [1, -4, -7, -14, -16]
    assert not np.array_equal(np.maximum(full, 0), O.convolve_2d_separated(X65, O.OP_TOP, O.PAD_SAME))
    # with Padding::None the separated form equals the full valid convolution (kernel.rs:171-192)
    for op in range(4):
        a = O.convolve_2d_separated(X65, op, O.PAD_NONE)
        b = np.maximum(O.convolve_2d(X65, O.sobel_full(op), O.PAD_NONE), 0)
        assert np.array_equal(a, b)


def test_kat_mlp_step():
    """Appendix B: 3-2-2 MLP, one sample, eta = 3."""
    W1 = np.array([[0.1, -0.2, 0.3], [0.4, 0.5, -0.6]]); b1 = np.array([0.01, -0.02])
    W2 = np.array([[0.7, -0.8], [-0.9, 1.0]]); b2 = np.array([0.03, 0.04])
    x = np.array([1.0, 0.5, -1.5]); y = np.array([0.0, 1.0])
    net = O.Net([(2, 3), (2, 2)])
    p = net.pack([W1, W2], [b1, b2])
    g, zs, acts, deltas = net.backprop(p, x, y)
    np.testing.assert_allclose(zs, [-0.43999999999999995, 1.5299999999999998, -0.3533863728935629, 0.5094394418856164], rtol=1e-14)
    np.testing.assert_allclose(acts, [0.3917409692534856, 0.8220063142137535, 0.41256147606104177, 0.6246750573701205], rtol=1e-14)
    np.testing.assert_allclose(deltas, [0.035548466979736765, -0.024578376854694016, 0.099986132119507, -0.08799723356765517], rtol=1e-13)
    newp, _ = net.train_batch(p, x[None], y[None], 3.0)
    ws, bs = net.unpack(newp)
    np.testing.assert_allclose(ws[1], [[0.5824940070747917, -1.046567695808136], [-0.7965836352917443, 1.217002844877865]], rtol=1e-13)
    np.testing.assert_allclose(bs[1], [-0.26995839635852104, 0.3039917007029655], rtol=1e-13)
    np.testing.assert_allclose(ws[0], [[-0.006645400939210289, -0.25332270046960514, 0.45996810140881544],
                                       [0.47373513056408206, 0.536867565282041, -0.710602695846123]], rtol=1e-12)
    np.testing.assert_allclose(bs[0], [-0.0966454009392103, 0.05373513056408204], rtol=1e-13)
    # numpy restatement agrees bit for bit
    nw, nb, *_ = N.train_batch([W1, W2], [b1, b2], [x], [y], 3.0)
    assert np.array_equal(nw[0], ws[0]) and np.array_equal(nw[1], ws[1])
    assert np.array_equal(nb[0], bs[0]) and np.array_equal(nb[1], bs[1])


# ---- C++ oracle vs independent numpy restatement ---------------------------------------------------------------
@pytest.mark.parametrize("shape", [(6, 5), (28, 28), (7, 9), (3, 3), (4, 3), (11, 16)])
@pytest.mark.parametrize("pad", [O.PAD_NONE, O.PAD_SAME])
def test_cross_separated(shape, pad):
    rng = np.random.default_rng(hash(shape) % 1000 + pad)
    for x in (rng.integers(0, 256, shape).astype(float), rng.standard_normal(shape) * 100):
        for op in range(4):
            assert np.array_equal(O.convolve_2d_separated(x, op, pad), N.convolve_2d_separated(x, op, pad))


@pytest.mark.parametrize("kshape", [(3, 3), (1, 1), (3, 1), (1, 3), (2, 2), (2, 3), (5, 5), (1, 5)])
@pytest.mark.parametrize("pad", [O.PAD_NONE, O.PAD_SAME])
def test_cross_convolve_2d(kshape, pad):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((9, 8))
    k = rng.standard_normal(kshape)
    try:
        want = N.convolve_2d(x, k, pad)
    except N.RefPanic:
        with pytest.raises(O.RefPanic):
            O.convolve_2d(x, k, pad)
        return
    assert np.array_equal(O.convolve_2d(x, k, pad), want)


@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (6, 5), (7, 7), (28, 28), (5, 8)])
@pytest.mark.parametrize("pad", [O.PAD_NONE, O.PAD_SAME])
def test_cross_pool(shape, pad):
    rng = np.random.default_rng(11)
    x = rng.integers(-3, 4, shape).astype(float)  # many ties
    x[x == 0] = np.where(rng.random(np.count_nonzero(x == 0)) < 0.5, 0.0, -0.0)  # +0 / -0 ties
    p, a = O.pool_2d(x, pad, O.POOL_MAX, return_argmax=True)
    pn, an = N.pool_2d(x, pad, N.POOL_MAX, return_argmax=True)
    assert np.array_equal(p, pn) and np.array_equal(np.signbit(p), np.signbit(pn)) and np.array_equal(a, an)


def test_panics():
    with pytest.raises(O.RefPanic):
        O.pool_2d(np.zeros((1, 5)), O.PAD_SAME, O.POOL_MAX)               # kernel.rs:246
    with pytest.raises(O.RefPanic):
        O.pool_2d(np.zeros((4, 4)), O.PAD_SAME, O.POOL_AVERAGE)           # kernel.rs:284 "Not implemented"
    with pytest.raises(O.RefPanic):
        O.convolve_2d_separated(np.zeros((2, 5)), 0, O.PAD_SAME)          # kernel.rs:200
    with pytest.raises(O.RefPanic):
        O.convolve_2d(np.zeros((4, 4)), np.zeros((2, 2)), O.PAD_SAME)     # kernel.rs:133
    with pytest.raises(O.RefPanic):
        O.convolve_2d(np.zeros((8, 8)), np.zeros((5, 5)), O.PAD_SAME)     # kernel.rs:156 out of bounds
    with pytest.raises(O.RefPanic):
        O.convolve_2d(np.zeros((2, 2)), np.zeros((3, 3)), O.PAD_NONE)     # kernel.rs:127
    x = np.zeros((4, 4)); x[1, 1] = np.nan
    with pytest.raises(O.RefPanic):
        O.pool_2d(x, O.PAD_SAME, O.POOL_MAX)                              # kernel.rs:280 partial_cmp unwrap


@pytest.mark.parametrize("cfg", [[1, 3], [1, 3, 1, 3], [0, 3, 1], [1, 1], [0, 0, 3], [3, 1, 3, 3], [1, 3, 1, 3, 1, 3], []])
def test_cross_flatten(cfg):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (28, 28)).astype(float)
    a, b = O.flatten_feature_set(cfg, img), N.flatten_feature_set(cfg, img)
    assert np.array_equal(a, b)
    n, h, w = O.feature_shape(cfg, 28, 28)
    assert a.size == n * h * w


def test_feature_order_cpcp():
    """slot order after the second conv is [B(f0..f3), T(f0),L(f0),R(f0), T(f1), ...]."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (28, 28)).astype(float)
    f = [O.pool_2d(O.convolve_2d_separated(img, op, 1), 1, 1) for op in (O.OP_TOP, O.OP_LEFT, O.OP_RIGHT, O.OP_BOTTOM)]
    maps = [O.convolve_2d_separated(fi, O.OP_BOTTOM, 1) for fi in f]
    for fi in f:
        maps += [O.convolve_2d_separated(fi, op, 1) for op in (O.OP_TOP, O.OP_LEFT, O.OP_RIGHT)]
    want = np.concatenate([O.pool_2d(m, 1, 1).flatten(order="F") for m in maps])
    assert np.array_equal(O.flatten_feature_set([1, 3, 1, 3], img), want)


def test_cross_backprop_and_train():
    rng = np.random.default_rng(9)
    shapes = [(7, 12), (5, 7), (3, 5)]
    net = O.Net(shapes)
    ws = [rng.standard_normal(s) for s in shapes]
    bs = [rng.standard_normal(s[0]) for s in shapes]
    p = net.pack(ws, bs)
    X = np.maximum(rng.standard_normal((6, 12)), 0)
    labels = np.arange(6) % 3
    Y = np.eye(3)[labels]
    newp, g = net.train_batch(p, X, Y, 3.0)
    nw, nb, gw, gb = N.train_batch(ws, bs, list(X), list(Y), 3.0)
    w2, b2 = net.unpack(newp)
    for a, b in zip(w2 + b2, nw + nb):
        assert np.array_equal(a, b)
    gw2, gb2 = net.unpack(g)
    for a, b in zip(gw2 + gb2, gw + gb):
        assert np.array_equal(a, b)
    # threaded reduction (rayon + mutex order) only changes summation order
    newp_t, _ = net.train_batch(p, X, Y, 3.0, n_threads=4)
    np.testing.assert_allclose(newp_t, newp, rtol=1e-12, atol=1e-14)
    acts = net.forward(p, X)
    for i in range(6):
        assert np.array_equal(acts[i], N.classify_test(ws, bs, X[i]))
    assert O.argmax_last(acts).tolist() == [N.argmax_last(a) for a in acts]
    assert O.accuracy(acts, labels) == sum(N.accuracy_hit(a, l) for a, l in zip(acts, labels))


def test_argmax_and_accuracy_ties():
    acts = np.array([[0.5, 0.9, 0.9], [1.0, 0.0, 0.0], [0.2, 0.2, 0.2]])
    assert O.argmax_last(acts).tolist() == [2, 0, 2]                    # last max wins (rcn.rs:92-97)
    assert O.accuracy(acts, np.array([2, 0, 2])) == 1                   # ties are never correct (rcn.rs:153-157)


def test_gen_scales_and_standardise():
    rng = np.random.default_rng(2)
    f = np.maximum(rng.standard_normal((5, 40)) * 50, 0)
    mean, sd = O.gen_scales(f)
    mn, sn = N.gen_scales(list(f))
    assert mean == mn and sd == sn
    assert np.array_equal(O.standardise(f, mean, sd), N.standardise(f, mean, sd))


# ---- committed golden files -------------------------------------------------------------------------------------
def test_golden_features():
    g = np.load(os.path.join(GOLD, "features_mnist8.npz"))
    for name, cfg in [("cp", [1, 3]), ("cpcp", [1, 3, 1, 3]), ("c_none_p", [0, 3]), ("cc", [1, 1])]:
        assert np.array_equal(O.features_u8(cfg, g["images"]), g["feat_" + name])
        one = N.flatten_feature_set(cfg, g["images"][3].astype(float))
        assert np.array_equal(one, g["feat_" + name][3])
    mean, sd = O.gen_scales(g["feat_cp"])
    assert [mean, sd] == g["scale_cp"].tolist()
    assert np.array_equal(O.standardise(g["feat_cp"], mean, sd), g["std_cp"])


def test_golden_dense():
    g = np.load(os.path.join(GOLD, "dense_mnist8.npz"))
    net = O.Net([(30, 784), (10, 30)])
    Y = np.eye(10)[g["labels"]]
    newp, grads = net.train_batch(g["params"], g["X"], Y, 3.0)
    assert np.array_equal(newp, g["new_params"]) and np.array_equal(grads, g["grads"])
    assert np.array_equal(net.forward(g["params"], g["X"]), g["acts"])
    ws, bs = net.unpack(g["params"])
    db, dw, zs, acts, deltas = N.backprop(ws, bs, g["X"][0], Y[0])
    assert np.array_equal(np.concatenate(zs), g["sample0_zs"])
    assert np.array_equal(np.concatenate(deltas), g["sample0_deltas"])


# ---- randomised cross-checks of the two restatements (hypothesis): shapes and configurations nobody hand-picked ---------
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

_SLOW_OK = settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))


@_SLOW_OK
@given(h=st.integers(3, 17), w=st.integers(3, 17), cfg=st.lists(st.sampled_from([0, 1, 3]), max_size=4), seed=st.integers(0, 2 ** 31))
def test_cross_flatten_random(h, w, cfg, seed):
    """C++ oracle == numpy twin, bit for bit, on random image sizes and convpool stacks (valid / SAME convolutions, max
    pools in any order); a stack that shrinks a map below what the next layer accepts must panic in both."""
    img = np.random.default_rng(seed).integers(0, 256, (h, w)).astype(float)
    try:
        want = N.flatten_feature_set(cfg, img)
    except N.RefPanic:
        with pytest.raises(O.RefPanic):
            O.flatten_feature_set(cfg, img)
        return
    got = O.flatten_feature_set(cfg, img)
    assert np.array_equal(got, want)
    n, fh, fw = O.feature_shape(cfg, h, w)
    assert got.size == n * fh * fw


@_SLOW_OK
@given(sizes=st.lists(st.integers(1, 9), min_size=2, max_size=5), batch=st.integers(1, 7), seed=st.integers(0, 2 ** 31),
       eta=st.sampled_from([0.1, 0.5, 3.0]))
def test_cross_train_batch_random(sizes, batch, seed, eta):
    """backprop + minibatch sum + SGD step: C++ oracle == numpy twin, bit for bit, on random layer stacks (incl. width-1
    layers and single-sample batches) with unscaled N(0, 1) parameters like the reference's initialisation."""
    rng = np.random.default_rng(seed)
    shapes = [(sizes[i + 1], sizes[i]) for i in range(len(sizes) - 1)]
    net = O.Net(shapes)
    ws = [rng.standard_normal(s) for s in shapes]
    bs = [rng.standard_normal(s[0]) for s in shapes]
    X = np.maximum(rng.standard_normal((batch, sizes[0])), 0)
    labels = rng.integers(0, sizes[-1], batch)
    Y = np.eye(sizes[-1])[labels]
    newp, g = net.train_batch(net.pack(ws, bs), X, Y, eta)
    nw, nb, gw, gb = N.train_batch(ws, bs, list(X), list(Y), eta)
    w2, b2 = net.unpack(newp)
    gw2, gb2 = net.unpack(g)
    for a, b in zip(w2 + b2 + gw2 + gb2, nw + nb + gw + gb):
        assert np.array_equal(a, b)
    acts = net.forward(net.pack(ws, bs), X)
    assert O.argmax_last(acts).tolist() == [N.argmax_last(a) for a in acts]
    assert O.accuracy(acts, labels) == sum(N.accuracy_hit(a, l) for a, l in zip(acts, labels))


@_SLOW_OK
@given(h=st.integers(2, 12), w=st.integers(2, 12), pad=st.sampled_from([0, 1]), seed=st.integers(0, 2 ** 31), ties=st.booleans())
def test_cross_pool_random(h, w, pad, seed, ties):
    """max pool values and the last-max-wins argmax on random maps, with many ties when asked for (small value range)."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 3 if ties else 256, (h, w)).astype(float)
    a, ai = O.pool_2d(x, pad, O.POOL_MAX, return_argmax=True)
    b, bi = N.pool_2d(x, pad, O.POOL_MAX, return_argmax=True)
    assert np.array_equal(a, b) and np.array_equal(ai, bi)
