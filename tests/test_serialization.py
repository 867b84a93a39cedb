"""CPU tests of the bincode-compatible checkpoint layout (rcn/src/utils/serialization.rs,
rcn/src/rcn.rs:13-25, main.rs:47-50,77). The reference holds no fixture for rcn.bin ("parity unpinned"): the expected
bytes below are assembled by hand from bincode 1.3's documented default encoding."""
import struct

import numpy as np
import pytest

from mercer_research_b200 import serialization as S


def u64(*v):
    return struct.pack(f"<{len(v)}Q", *v)


def test_known_bytes():
    w0 = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])          # 2 x 3; column-major data = 1,4,2,5,3,6
    w1 = np.array([[7.0, 8.0]])
    state = dict(classes=1, convpool_cfg=[1, 3, 0, 2], feedforward_cfg=[2], weights=[w0, w1],
                 biases=[np.array([0.5, -0.5]), np.array([9.0])], scale_set=(40.25, 60.5), training_path="tr", testing_path="tést")
    want = b"".join([
        u64(1),
        u64(4), struct.pack("<IIIIIIII", 0, 1, 1, 1, 0, 0, 1, 0),   # Convolve2D(Same), Pool2D(Max), Convolve2D(None), Pool2D(Average)
        u64(1), u64(2),
        u64(2),
        u64(2, 3), u64(6), struct.pack("<6d", 1, 4, 2, 5, 3, 6),
        u64(1, 2), u64(2), struct.pack("<2d", 7, 8),
        u64(2),
        u64(2), struct.pack("<2d", 0.5, -0.5),
        u64(1), struct.pack("<d", 9.0),
        struct.pack("<dd", 40.25, 60.5),
        u64(2), b"tr",
        u64(5), "tést".encode("utf-8"),
    ])
    got = S.encode_state(state)
    assert got == want
    back = S.decode_state(got)
    assert back["classes"] == 1 and back["convpool_cfg"] == [1, 3, 0, 2] and back["feedforward_cfg"] == [2]
    assert np.array_equal(back["weights"][0], w0) and np.array_equal(back["weights"][1], w1)
    assert np.array_equal(back["biases"][0], [0.5, -0.5]) and back["scale_set"] == (40.25, 60.5)
    assert back["training_path"] == "tr" and back["testing_path"] == "tést"


def test_untrained_model_and_round_trip_random():
    empty = dict(classes=10, convpool_cfg=[1, 3], feedforward_cfg=[30], weights=[], biases=[], scale_set=(1.0, 1.0),
                 training_path="", testing_path="")                 # RCN::new before train(): empty vectors (rcn.rs:65-74)
    b = S.encode_state(empty)
    assert len(b) == 8 + (8 + 2 * 8) + (8 + 8) + 8 + 8 + 16 + 8 + 8
    assert S.decode_state(b)["weights"] == []
    rng = np.random.default_rng(0)
    st = dict(classes=10, convpool_cfg=[1, 3, 1, 3], feedforward_cfg=[30, 20], weights=[rng.standard_normal((30, 784)),
              rng.standard_normal((20, 30)), rng.standard_normal((10, 20))], biases=[rng.standard_normal(30), rng.standard_normal(20),
              rng.standard_normal(10)], scale_set=(3.5, 7.25), training_path="images/mnist_png/training", testing_path="images/mnist_png/testing")
    back = S.decode_state(S.encode_state(st))
    for a, b_ in zip(st["weights"] + st["biases"], back["weights"] + back["biases"]):
        assert np.array_equal(a, b_)
    assert S.encode_state(back) == S.encode_state(st)


def test_malformed_checkpoints_are_rejected():
    good = S.encode_state(dict(classes=2, convpool_cfg=[1], feedforward_cfg=[3], weights=[np.zeros((3, 4)), np.zeros((2, 3))],
                               biases=[np.zeros(3), np.zeros(2)], scale_set=(1.0, 1.0), training_path="a", testing_path="b"))
    with pytest.raises(ValueError):
        S.decode_state(good[:-1])
    bad_variant = bytearray(good)
    bad_variant[16:20] = struct.pack("<I", 7)
    with pytest.raises(ValueError):
        S.decode_state(bytes(bad_variant))
    with pytest.raises(ValueError):
        S.encode_state(dict(classes=2, convpool_cfg=[9], feedforward_cfg=[], weights=[], biases=[], scale_set=(1, 1)))


from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None, derandomize=True, database=None, suppress_health_check=list(HealthCheck))
@given(classes=st.integers(1, 50), cfg=st.lists(st.integers(0, 3), max_size=6), widths=st.lists(st.integers(1, 12), min_size=1, max_size=4),
       trained=st.booleans(), seed=st.integers(0, 2 ** 31), tr=st.text(max_size=20), te=st.text(max_size=20))
def test_round_trip_random_states(classes, cfg, widths, trained, seed, tr, te):
    """encode -> decode -> encode is the identity on random models (any layer stack, any widths, untrained or trained,
    non-ASCII paths): the byte layout has no state-dependent ambiguity."""
    rng = np.random.default_rng(seed)
    sizes = [int(rng.integers(1, 40))] + widths + [classes]
    ws = [rng.standard_normal((sizes[i + 1], sizes[i])) for i in range(len(sizes) - 1)] if trained else []
    bs = [rng.standard_normal(w.shape[0]) for w in ws]
    state = dict(classes=classes, convpool_cfg=cfg, feedforward_cfg=widths, weights=ws, biases=bs,
                 scale_set=(float(rng.standard_normal()), float(abs(rng.standard_normal()) + 0.1)), training_path=tr, testing_path=te)
    blob = S.encode_state(state)
    back = S.decode_state(blob)
    assert S.encode_state(back) == blob
    assert back["classes"] == classes and back["convpool_cfg"] == cfg and back["feedforward_cfg"] == widths
    assert back["training_path"] == tr and back["testing_path"] == te and back["scale_set"] == state["scale_set"]
    assert len(back["weights"]) == len(ws) and all(np.array_equal(a, b) for a, b in zip(back["weights"], ws))
    for cut in sorted({0, 1, len(blob) // 2, len(blob) - 1}):
        if cut < len(blob):
            with pytest.raises(ValueError):
                S.decode_state(blob[:cut])                      # a truncated file never decodes silently
