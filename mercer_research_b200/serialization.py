"""bincode-compatible checkpoint of an ``RCN``.

The reference persists its model with ``bincode::serialize(&model)`` into ``rcn.bin`` and reloads it in the CLI and the
backend (rcn/src/main.rs:47-50,77; backend/src/main.rs:54,68). The byte layout follows from the serde derives on
``struct RCN`` (rcn/src/rcn.rs:13-25), ``enum RCNLayer`` (rcn.rs:35-38), ``Padding`` / ``Pooling`` (kernel.rs:23-35)
and the hand-written impls for ``Weights`` / ``Bias`` (rcn/src/utils/serialization.rs:11-24, 96-108) under bincode 1.3's
default options (crate not vendored: little-endian, fixed-width integers, ``usize`` as u64, sequence and string
lengths as u64, enum variant index as u32, struct fields and tuple members concatenated in declaration order):

    classes            u64
    convpool_cfg       u64 n, then n x (u32 RCNLayer variant {0 Convolve2D, 1 Pool2D}, u32 inner variant
                                        {Padding: 0 None, 1 Same | Pooling: 0 Average, 1 Max})
    feedforward_cfg    u64 n, then n x u64
    layer_weights      u64 n, then n x (dims: u64 rows, u64 cols; data: u64 len, len x f64 COLUMN-MAJOR)
    layer_bias         u64 n, then n x (u64 len, len x f64)
    scale_set          f64 mean, f64 sd
    training_path      u64 len, utf-8 bytes
    testing_path       u64 len, utf-8 bytes

A model trained on the B200 can therefore be served by the unmodified reference backend and vice versa. The reference
holds no fixture for this format ("parity unpinned"); tests/test_serialization.py pins the layout byte by byte.
"""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np


def encode_state(state: Dict) -> bytes:
    """state: classes, convpool_cfg (RCN_LAYER_* codes), feedforward_cfg, weights (list of (rows, cols) arrays),
    biases (list of vectors), scale_set (mean, sd), training_path, testing_path."""
    out: List[bytes] = [struct.pack("<Q", int(state["classes"]))]
    cfg = [int(c) for c in state["convpool_cfg"]]
    out.append(struct.pack("<Q", len(cfg)))
    for code in cfg:
        if not 0 <= code <= 3:
            raise ValueError(f"unknown RCNLayer code {code}")
        out.append(struct.pack("<II", 0 if code < 2 else 1, code & 1 if code < 2 else code - 2))
    ff = [int(x) for x in state["feedforward_cfg"]]
    out.append(struct.pack("<Q", len(ff)) + struct.pack(f"<{len(ff)}Q", *ff))
    ws = [np.asarray(w, dtype=np.float64) for w in state["weights"]]
    out.append(struct.pack("<Q", len(ws)))
    for w in ws:
        if w.ndim != 2:
            raise ValueError("weights must be 2-D")
        out.append(struct.pack("<QQQ", w.shape[0], w.shape[1], w.size))
        out.append(np.asfortranarray(w).astype("<f8").tobytes(order="F"))     # self.0.iter(): column-major
    bs = [np.asarray(b, dtype=np.float64).ravel() for b in state["biases"]]
    out.append(struct.pack("<Q", len(bs)))
    for b in bs:
        out.append(struct.pack("<Q", b.size) + b.astype("<f8").tobytes())
    out.append(struct.pack("<dd", float(state["scale_set"][0]), float(state["scale_set"][1])))
    for key in ("training_path", "testing_path"):
        raw = str(state.get(key, "")).encode("utf-8")
        out.append(struct.pack("<Q", len(raw)) + raw)
    return b"".join(out)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = memoryview(data), 0

    def take(self, fmt: str):
        n = struct.calcsize(fmt)
        if self.o + n > len(self.d):
            raise ValueError("unexpected end of checkpoint (io error: UnexpectedEof in bincode terms)")
        v = struct.unpack_from(fmt, self.d, self.o)
        self.o += n
        return v

    def f64s(self, n: int) -> np.ndarray:
        if self.o + 8 * n > len(self.d):
            raise ValueError("unexpected end of checkpoint")
        a = np.frombuffer(self.d, dtype="<f8", count=n, offset=self.o).copy()
        self.o += 8 * n
        return a

    def string(self) -> str:
        (n,) = self.take("<Q")
        if self.o + n > len(self.d):
            raise ValueError("unexpected end of checkpoint")
        s = bytes(self.d[self.o:self.o + n]).decode("utf-8")
        self.o += n
        return s


def decode_state(data: bytes) -> Dict:
    r = _Reader(data)
    (classes,) = r.take("<Q")
    (n,) = r.take("<Q")
    cfg = []
    for _ in range(n):
        outer, inner = r.take("<II")
        if outer > 1 or inner > 1:
            raise ValueError(f"invalid RCNLayer variant ({outer}, {inner})")
        cfg.append(inner if outer == 0 else 2 + inner)
    (n,) = r.take("<Q")
    ff = list(r.take(f"<{n}Q")) if n else []
    (n,) = r.take("<Q")
    weights = []
    for _ in range(n):
        rows, cols, ln = r.take("<QQQ")
        if ln != rows * cols:       # DMatrix::from_vec panics on a length mismatch
            raise ValueError(f"Weights data length {ln} does not match dims ({rows}, {cols})")
        weights.append(r.f64s(ln).reshape((cols, rows)).T.copy())
    (n,) = r.take("<Q")
    biases = []
    for _ in range(n):
        (ln,) = r.take("<Q")
        biases.append(r.f64s(ln))
    mean, sd = r.take("<dd")
    training_path, testing_path = r.string(), r.string()
    return dict(classes=classes, convpool_cfg=cfg, feedforward_cfg=ff, weights=weights, biases=biases, scale_set=(mean, sd),
                training_path=training_path, testing_path=testing_path)


# ---- model-level helpers (need a GPU: they read / write the device-resident parameters) ----------------------------------
def model_state(model) -> Dict:
    shapes = model.layer_shapes
    return dict(classes=model.classes, convpool_cfg=[l.code if hasattr(l, "code") else int(l) for l in model.convpool_cfg],
                feedforward_cfg=model.feedforward_cfg, weights=[model.get_weights(i) for i in range(len(shapes))],
                biases=[model.get_bias(i) for i in range(len(shapes))], scale_set=model.scale_set,
                training_path=model.training_path, testing_path=model.testing_path)


def dumps(model) -> bytes:
    """``bincode::serialize(&model)`` (main.rs:77)."""
    return encode_state(model_state(model))


def loads(data: bytes, device: int = 0):
    """``bincode::deserialize(&data)`` (main.rs:50, backend/src/main.rs:68) onto a B200."""
    from .rcn import RCN
    st = decode_state(data)
    model = RCN(st["classes"], st["convpool_cfg"], st["feedforward_cfg"], st["training_path"], st["testing_path"], device=device)
    if st["weights"]:
        model.load_weights_and_bias_shapes([w.shape for w in st["weights"]])
        for i, (w, b) in enumerate(zip(st["weights"], st["biases"])):
            model.set_weights(i, w)
            model.set_bias(i, b)
    model.scale_set = st["scale_set"]
    return model


def save(model, path: str = "./rcn.bin"):
    with open(path, "wb") as f:
        f.write(dumps(model))


def load(path: str = "./rcn.bin", device: int = 0):
    with open(path, "rb") as f:
        return loads(f.read(), device=device)
