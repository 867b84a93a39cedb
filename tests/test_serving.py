"""Serving path (backend/src/main.rs:10-74): request batcher + HTTP routes.
The host logic runs on CPU against a stand-in model; the GPU test serves a real model and checks every label the
endpoint returns against the oracle (bit-exact, rcn.rs:82-98)."""
import base64
import io
import json
import os
import random
import threading
import urllib.error
import urllib.request

import numpy as np
import pytest

from mercer_research_b200 import serving

CP = [1, 3]


class FakeModel:
    """classify_images stand-in: label = (sum of pixels) mod 10; records the batch sizes it was called with."""

    def __init__(self, gate=None):
        self.calls = []
        self.gate = gate

    def classify_images(self, images):
        if self.gate is not None:
            self.gate.wait(5)
        a = np.asarray(images)
        assert a.ndim == 3 and a.dtype == np.uint8
        self.calls.append(a.shape)
        return a.reshape(a.shape[0], -1).sum(axis=1).astype(np.int64) % 10


def want(img):
    return int(img.astype(np.int64).sum() % 10)


def test_batcher_routes_results_and_coalesces():
    gate = threading.Event()
    model = FakeModel(gate)
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, size=(28, 28), dtype=np.uint8) for _ in range(64)]
    with serving.ClassifyBatcher(model, max_batch=16, max_delay_s=0.05) as b:
        first = b.submit(imgs[0])          # the worker blocks inside the model on this one ...
        futs = [b.submit(im) for im in imgs[1:]]   # ... so these 63 pile up and must go out in chunks of <= 16
        gate.set()
        got = [first.result(5)] + [f.result(5) for f in futs]
    assert got == [want(im) for im in imgs]
    assert sum(s[0] for s in model.calls) == 64 and max(s[0] for s in model.calls) <= 16
    assert len(model.calls) <= 1 + 4 + 1   # first image alone (or with early arrivals), then full 16-image batches
    assert b.images == 64 and b.batches == len(model.calls)


def test_batcher_mixed_shapes_and_validation():
    model = FakeModel()
    rng = np.random.default_rng(2)
    a = [rng.integers(0, 256, size=(28, 28), dtype=np.uint8) for _ in range(5)]
    c = [rng.integers(0, 256, size=(32, 32), dtype=np.uint8) for _ in range(3)]
    with serving.ClassifyBatcher(model, max_batch=64, max_delay_s=0.05) as b:
        futs = [b.submit(x) for x in (a[0], c[0], a[1], c[1], a[2], a[3], c[2], a[4])]
        got = [f.result(5) for f in futs]
        with pytest.raises(ValueError):
            b.submit(np.zeros((28, 28)))              # not uint8
        with pytest.raises(ValueError):
            b.submit(np.zeros((2, 28, 28), np.uint8))  # not one image
    assert got == [want(x) for x in (a[0], c[0], a[1], c[1], a[2], a[3], c[2], a[4])]
    assert all(s[1:] in ((28, 28), (32, 32)) for s in model.calls)
    with pytest.raises(RuntimeError):
        b.submit(a[0])                                # closed


def test_batcher_delivers_model_errors_and_keeps_serving():
    class Flaky(FakeModel):
        def classify_images(self, images):
            if np.asarray(images).shape[1] == 1:
                raise RuntimeError("matrices must be at least 2x2 to be pooled")   # kernel.rs:247 as a status
            return super().classify_images(images)

    with serving.ClassifyBatcher(Flaky(), max_batch=8, max_delay_s=0.01) as b:
        bad = b.submit(np.zeros((1, 5), np.uint8))
        with pytest.raises(RuntimeError, match="at least 2x2"):
            bad.result(5)
        ok = np.full((4, 4), 3, np.uint8)
        assert b.classify(ok, 5) == want(ok)


@pytest.fixture()
def image_tree(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(3)
    pixels = {}
    for c in range(3):
        d = tmp_path / "images" / str(c)
        d.mkdir(parents=True)
        for k in range(4):
            a = rng.integers(0, 256, size=(28, 28), dtype=np.uint8)
            p = d / "{}.png".format(k)
            Image.fromarray(a, mode="L").save(p)
            pixels[str(p)] = a
    return str(tmp_path / "images"), pixels


def _get(url):
    with urllib.request.urlopen(url, timeout=10) as r:
        return r.status, dict(r.headers), r.read()


def test_http_routes(image_tree):
    from PIL import Image
    root, pixels = image_tree
    paths = serving.list_images(root, random.Random(7))
    assert sorted(paths) == sorted(pixels)
    state = serving.RCNState(FakeModel(), paths, max_batch=8, max_delay_s=0.001, rng=random.Random(11))
    srv = serving.make_server(state, port=0)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    base = "http://127.0.0.1:{}".format(srv.server_address[1])
    try:
        code, hdr, body = _get(base + "/health")
        assert code == 200 and body == b"Healthy!"                       # backend/src/main.rs:44-47
        assert hdr["Access-Control-Allow-Origin"] == "*"
        valid = {want(a) for a in pixels.values()}
        for _ in range(12):
            code, hdr, body = _get(base + "/")
            assert code == 200 and hdr["Content-Type"] == "application/json"
            res = json.loads(body)
            assert set(res) == {"output", "img"}                         # RCNResult (main.rs:15-19)
            img = np.asarray(Image.open(io.BytesIO(base64.standard_b64decode(res["img"]))).convert("L"))
            assert any(np.array_equal(img, a) for a in pixels.values())  # the PNG sent back is the classified file
            assert res["output"] == want(img) and res["output"] in valid
        with pytest.raises(urllib.error.HTTPError) as e:
            _get(base + "/nope")
        assert e.value.code == 404
        # concurrent requests are answered correctly and share device calls
        out = []
        lock = threading.Lock()

        def hit():
            r = json.loads(_get(base + "/")[2])
            im = np.asarray(Image.open(io.BytesIO(base64.standard_b64decode(r["img"]))).convert("L"))
            with lock:
                out.append(r["output"] == want(im))

        ts = [threading.Thread(target=hit) for _ in range(16)]
        [x.start() for x in ts]
        [x.join(20) for x in ts]
        assert len(out) == 16 and all(out)
    finally:
        srv.shutdown()
        srv.server_close()
        state.close()


def test_state_needs_images():
    with pytest.raises(ValueError):
        serving.RCNState(FakeModel(), [])


@pytest.mark.gpu
def test_serving_real_model_labels_exact(built_library, image_tree, tmp_path):
    """rcn.bin -> RCN.load -> batcher: every served label equals the oracle's classify (rcn.rs:82-98)."""
    import oracle as O
    import mercer_research_b200 as m
    root, pixels = image_tree
    model = m.RCN(10, [m.RCNLayer.Convolve2D(m.Padding.Same), m.RCNLayer.Pool2D(m.Pooling.Max)], [30])
    model.load_weights_and_bias(model.feature_len(28, 28))
    net = O.Net(model.layer_shapes)
    params = np.random.default_rng(0xC0FFEE).standard_normal(net.n_params) * 0.05
    model.set_params(params)
    paths = sorted(pixels)
    imgs = np.stack([pixels[p] for p in paths])
    raw = O.features_u8(CP, imgs)
    mean, sd = O.gen_scales(raw)
    model.scale_set = (mean, sd)
    labels = O.argmax_last(net.forward(params, O.standardise(raw, mean, sd)))
    ckpt = str(tmp_path / "rcn.bin")
    model.save(ckpt)
    served = m.RCN.load(ckpt)                                            # what the backend does (main.rs:54,68)
    with serving.ClassifyBatcher(served, max_batch=8, max_delay_s=0.01) as b:
        futs = [b.submit(pixels[p]) for p in paths]
        got = np.array([f.result(60) for f in futs])
    assert np.array_equal(got, labels)
    assert b.batches < len(paths)                                        # requests were coalesced
    state = serving.RCNState(served, paths, rng=random.Random(5))
    try:
        for _ in range(6):
            r = state.get_rcn_result()
            from PIL import Image
            im = np.asarray(Image.open(io.BytesIO(base64.standard_b64decode(r["img"]))).convert("L"))
            k = [i for i, p in enumerate(paths) if np.array_equal(pixels[p], im)][0]
            assert r["output"] == int(labels[k])
    finally:
        state.close()
