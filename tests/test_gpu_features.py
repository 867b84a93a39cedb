"""GPU parity tests of the feature stage through the C ABI: bit-exact against the oracle (integer/byte work)."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def api(built_library):
    import mercer_research_b200 as m
    return m


def layers(api, codes):
    out = []
    for c in codes:
        out.append(api.RCNLayer.Convolve2D(api.Padding(c)) if c < 2 else api.RCNLayer.Pool2D(api.Pooling(c - 2)))
    return out


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


# ---- op level -----------------------------------------------------------------------------------------------------
def test_ref_identity_same(api):
    """kernel.rs:434-441 through the CUDA path."""
    m = np.arange(900, dtype=np.float64).reshape(30, 30)
    k = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 0]], dtype=np.float64)
    assert bits_equal(api.convolve_2d(m, k, api.Padding.Same), m)


@pytest.mark.parametrize("kshape", [(3, 3), (1, 1), (3, 1), (1, 3), (2, 2), (2, 3), (1, 5), (4, 1)])
@pytest.mark.parametrize("pad", [0, 1])
def test_convolve_2d_generic(api, kshape, pad):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((19, 23))
    k = rng.standard_normal(kshape)
    try:
        want = O.convolve_2d(x, k, pad)
    except O.RefPanic:
        with pytest.raises(api.RcnCudaError):
            api.convolve_2d(x, k, api.Padding(pad))
        return
    assert bits_equal(api.convolve_2d(x, k, api.Padding(pad)), want)


@pytest.mark.parametrize("shape", [(3, 3), (6, 5), (28, 28), (7, 9), (64, 64), (33, 130)])
@pytest.mark.parametrize("pad", [0, 1])
def test_convolve_2d_separated(api, shape, pad):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    for x in (rng.integers(0, 256, shape).astype(float), rng.standard_normal(shape) * 1e3):
        for op in range(4):
            got = api.convolve_2d_separated(x, api.SeparableOperator(op), api.Padding(pad))
            assert bits_equal(got, O.convolve_2d_separated(x, op, pad)), (shape, pad, op)


def test_kat_appendix_b(api):
    X = np.array([[8, 6, 5, 2, 3], [0, 0, 0, 1, 8], [6, 9, 5, 6, 9], [7, 6, 5, 5, 9], [2, 8, 6, 0, 3], [8, 5, 0, 7, 7]], dtype=float)
    top = api.convolve_2d_separated(X, api.SeparableOperator.Top, api.Padding.Same)
    assert top.tolist() == [[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [2, 1, 0, 0, 0], [0, 0, 0, 0, 0], [4, 9, 5, 5, 11], [0, 0, 6, 9, 1]]
    p, a = api.pool_2d(top, api.Padding.Same, api.Pooling.Max, return_argmax=True)
    assert p.tolist() == [[0, 0, 0], [2, 0, 0], [9, 9, 11]]
    assert a.tolist() == [[3, 3, 3], [0, 3, 3], [1, 3, 0]]


def test_relu(api):
    x = np.array([[1.5, -2.0, 0.0], [-0.0, np.inf, -np.inf], [np.nan, 1e-300, -1e-300]])
    assert bits_equal(api.relu(x), O.relu(x))   # NaN -> 0, -0.0 stays -0.0 (kernel.rs:214)


@pytest.mark.parametrize("shape", [(2, 2), (3, 3), (6, 5), (7, 7), (28, 28), (5, 8), (129, 65)])
@pytest.mark.parametrize("pad", [0, 1])
def test_pool_2d_argmax_bit_exact(api, shape, pad):
    rng = np.random.default_rng(11)
    x = rng.integers(-3, 4, shape).astype(float)
    z = x == 0
    x[z] = np.where(rng.random(np.count_nonzero(z)) < 0.5, 0.0, -0.0)
    p, a = api.pool_2d(x, api.Padding(pad), api.Pooling.Max, return_argmax=True)
    pw, aw = O.pool_2d(x, pad, O.POOL_MAX, return_argmax=True)
    assert bits_equal(p, pw) and np.array_equal(a, aw)


def test_pool_nan_is_an_error(api):
    x = np.zeros((4, 4)); x[2, 1] = np.nan
    with pytest.raises(api.RcnCudaError) as e:
        api.pool_2d(x, api.Padding.Same, api.Pooling.Max)
    assert e.value.status == 7


def test_ops_on_torch_device_tensors(api):
    import torch
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (28, 28)).astype(float)
    xt = torch.from_numpy(x).cuda()
    got = api.convolve_2d_separated(xt, api.SeparableOperator.Left, api.Padding.Same)
    assert got.is_cuda and bits_equal(got.cpu().numpy(), O.convolve_2d_separated(x, O.OP_LEFT, 1))
    p = api.pool_2d(got, api.Padding.Same, api.Pooling.Max)
    assert bits_equal(p.cpu().numpy(), O.pool_2d(O.convolve_2d_separated(x, O.OP_LEFT, 1), 1, 1))


# ---- model level ----------------------------------------------------------------------------------------------------
CFGS = [[1, 3], [1, 3, 1, 3], [0, 3], [1, 1], [0, 0, 3], [3, 1, 3, 3], [1, 3, 1, 3, 1, 3], [0, 3, 1], [1]]


@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("hw", [(28, 28), (32, 32), (9, 14)])
def test_features_u8_bit_exact(api, cfg, hw):
    rng = np.random.default_rng(len(cfg) * 100 + hw[0])
    imgs = rng.integers(0, 256, size=(37, hw[0], hw[1]), dtype=np.uint8)
    try:
        want = O.features_u8(cfg, imgs)
    except O.RefPanic:
        model = api.RCN(10, layers(api, cfg), [30])
        with pytest.raises(api.RcnCudaError):
            model.flatten_feature_set(imgs)
        return
    model = api.RCN(10, layers(api, cfg), [30])
    assert model.feature_len(*hw) == want.shape[1]
    assert bits_equal(model.flatten_feature_set(imgs), want)
    # f64 column-major images (DMatrix layout) take the f64 kernel and must agree bit for bit as well
    assert bits_equal(model.flatten_feature_set(imgs.astype(np.float64)), want)


@pytest.mark.parametrize("cfg", [[1, 3], [1, 3, 1, 3], [1, 3, 1, 3, 1, 3]])
@pytest.mark.parametrize("hw", [(64, 64), (12, 20), (20, 12), (28, 28), (8, 4), (4, 4), (33, 32)])
def test_features_staged_kernel_bit_exact(api, cfg, hw):
    """The bulk-async staged kernel (u8, conv(Same)+pool stacks, W % 4 == 0): more images than resident CTAs so every
    CTA walks several images through its double-buffered staging ring; with and without the standardise epilogue
    (IEEE and host-verified fast division), contiguous and 1-byte-offset (non-bulk) sources."""
    import torch
    rng = np.random.default_rng(hw[0] * 1000 + hw[1] + len(cfg))
    B = 2500 if hw[0] * hw[1] <= 1024 else 1300
    imgs = rng.integers(0, 256, size=(B,) + hw, dtype=np.uint8)
    imgs[0] = 255; imgs[1] = 0; imgs[2, ::2] = 255; imgs[2, 1::2] = 0
    try:
        want = O.features_u8(cfg, imgs)
    except O.RefPanic:
        model = api.RCN(10, layers(api, cfg), [30])
        with pytest.raises(api.RcnCudaError):
            model.flatten_feature_set(imgs)
        return
    model = api.RCN(10, layers(api, cfg), [30])
    assert bits_equal(model.flatten_feature_set(imgs), want)
    mean, sd = O.gen_scales(want)
    if sd > 0:
        model.scale_set = (mean, sd)
        assert bits_equal(model.flatten_feature_set(imgs, standardise=True), O.standardise(want, mean, sd))
        model.scale_set = (mean * 1.0000001, sd * 0.9999999)     # a pair the fast-division check may reject
        assert bits_equal(model.flatten_feature_set(imgs, standardise=True),
                          O.standardise(want, mean * 1.0000001, sd * 0.9999999))
    # misaligned device source: the kernel falls back from bulk copies to plain loads into the same staging ring
    flat = torch.zeros(imgs.size + 1, dtype=torch.uint8, device="cuda")
    flat[1:] = torch.from_numpy(imgs.reshape(-1)).cuda()
    shifted = flat[1:].view(B, *hw)
    assert shifted.data_ptr() % 16 != 0
    assert bits_equal(model.flatten_feature_set(shifted).cpu().numpy(), want)


def test_features_f64_arbitrary_values(api):
    """Non-integer pixels: the f64 path keeps the reference's accumulation order, so it is still bit-exact."""
    rng = np.random.default_rng(17)
    imgs = rng.standard_normal((5, 28, 28)) * 37.0
    for cfg in ([1, 3], [1, 3, 1, 3], [0, 3, 1]):
        want = np.stack([O.flatten_feature_set(cfg, im) for im in imgs])
        model = api.RCN(10, layers(api, cfg), [30])
        assert bits_equal(model.flatten_feature_set(imgs), want)


def test_features_standardise_epilogue_and_scales(api):
    g = np.load(os.path.join(GOLD, "features_mnist8.npz"))
    model = api.RCN(10, layers(api, [1, 3]), [30])
    raw = model.flatten_feature_set(g["images"])
    assert bits_equal(raw, g["feat_cp"])
    mean, sd = model.gen_scales(raw)                    # parallel sum: agrees to rounding, not bitwise
    assert abs(mean - g["scale_cp"][0]) <= 1e-12 * abs(g["scale_cp"][0])
    assert abs(sd - g["scale_cp"][1]) <= 1e-12 * abs(g["scale_cp"][1])
    assert model.scale_set == (mean, sd)                # stored like rcn.rs:249-250
    model.scale_set = tuple(g["scale_cp"])              # (mean, sd) as inputs => elementwise IEEE ops => bit-exact
    assert bits_equal(model.flatten_feature_set(g["images"], standardise=True), g["std_cp"])
    assert bits_equal(model.standardise(raw), g["std_cp"])
    for name, cfg in [("cpcp", [1, 3, 1, 3]), ("c_none_p", [0, 3]), ("cc", [1, 1])]:
        m2 = api.RCN(10, layers(api, cfg), [30])
        assert bits_equal(m2.flatten_feature_set(g["images"]), g["feat_" + name])


def test_features_large_maps_take_the_layered_path(api):
    """64x64 with three unpooled convs: 64 maps of 64x64 per image do not fit shared memory."""
    rng = np.random.default_rng(23)
    imgs = rng.integers(0, 256, size=(3, 64, 64), dtype=np.uint8)
    cfg = [1, 1, 1]
    model = api.RCN(10, layers(api, cfg), [30])
    assert bits_equal(model.flatten_feature_set(imgs), O.features_u8(cfg, imgs))
    assert bits_equal(model.flatten_feature_set(imgs.astype(np.float64)), O.features_u8(cfg, imgs))


def test_features_device_buffers_and_empty(api):
    import torch
    rng = np.random.default_rng(29)
    imgs = rng.integers(0, 256, size=(1024, 28, 28), dtype=np.uint8)
    model = api.RCN(10, layers(api, [1, 3]), [30])
    got = model.flatten_feature_set(torch.from_numpy(imgs).cuda())
    assert got.is_cuda and bits_equal(got.cpu().numpy(), O.features_u8([1, 3], imgs))
    # no conv layer => empty feature vector (rcn.rs:323,339); zero-sized batch => nothing to do
    empty = api.RCN(10, layers(api, [3]), [30])
    assert empty.flatten_feature_set(imgs[:4]).shape == (4, 0)
    assert model.flatten_feature_set(imgs[:0]).shape == (0, 784)


def test_features_full_size_property(api):
    """At bench size (B = 8192, larger than the oracle should chew in a test) check a size-independent property:
    features of a batch are the per-image features (no cross-image mixing), via a checksum of spot images."""
    rng = np.random.default_rng(31)
    imgs = rng.integers(0, 256, size=(8192, 28, 28), dtype=np.uint8)
    model = api.RCN(10, layers(api, [1, 3, 1, 3]), [30])
    got = model.flatten_feature_set(imgs)
    idx = [0, 1, 4095, 4096, 8190, 8191]
    assert bits_equal(got[idx], O.features_u8([1, 3, 1, 3], imgs[idx]))
    assert (got >= 0).all() and np.array_equal(got, np.round(got))   # ReLU'd integers
