"""Serving path: the reference's actix backend (backend/src/main.rs:10-74) on the GPU model.

The reference answers ``GET /`` by classifying ONE random image per request on a CPU model held per worker
(backend/src/main.rs:22-42, 64-71) and ``GET /health`` with ``"Healthy!"`` (:44-47). Here every request thread hands
its decoded image to a :class:`ClassifyBatcher`; one worker thread coalesces whatever is waiting into a single
``RCN.classify_images`` call (feature kernel + standardise + forward + argmax on the device, rcn.rs:82-98), so
concurrent requests share one launch instead of one model replica each. The batcher only needs an object with
``classify_images(uint8 (B, H, W)) -> (B,) labels``; it never computes anything itself (no CPU fallback).

Same wire format as the reference: ``{"output": <class index>, "img": <base64 PNG>}`` (``RCNResult``,
backend/src/main.rs:15-19), permissive CORS (:66), default bind 127.0.0.1:8080 (:72).
"""
from __future__ import annotations

import base64
import io
import json
import os
import random
import threading
import time
from concurrent.futures import Future
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .data import grayscale_u8


class ClassifyBatcher:
    """Coalesces concurrent ``classify`` requests into batched device calls.

    ``submit(image)`` returns a ``Future`` resolving to the class index. The worker drains the queue whenever it is
    idle: the first waiting request opens a batch, which closes after ``max_delay_s`` or at ``max_batch`` images;
    images of different sizes go to the device in separate calls (one call takes one H x W). An exception raised by
    the model is delivered to every future of the failing call and the worker keeps serving."""

    def __init__(self, model, max_batch: int = 1024, max_delay_s: float = 0.002):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.model = model
        self.max_batch = int(max_batch)
        self.max_delay_s = float(max_delay_s)
        self.batches = 0                      # device calls made
        self.images = 0                       # images classified
        self._pending: List[Tuple[np.ndarray, Future]] = []
        self._cv = threading.Condition()
        self._closed = False
        self._worker = threading.Thread(target=self._run, name="rcn-classify-batcher", daemon=True)
        self._worker.start()

    def submit(self, image) -> Future:
        a = np.asarray(image)
        if a.ndim != 2 or a.dtype != np.uint8:
            raise ValueError("image must be a decoded grayscale (H, W) uint8 array")
        fut: Future = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("batcher is closed")
            self._pending.append((a, fut))
            self._cv.notify()
        return fut

    def classify(self, image, timeout: Optional[float] = None) -> int:
        return int(self.submit(image).result(timeout))

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _take(self) -> List[Tuple[np.ndarray, Future]]:
        with self._cv:
            while not self._pending and not self._closed:
                self._cv.wait()
            if not self._pending:
                return []
            deadline = time.monotonic() + self.max_delay_s
            while len(self._pending) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            taken, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
            return taken

    def _run(self):
        while True:
            taken = self._take()
            if not taken:
                return                            # closed and drained
            by_shape = {}
            for a, fut in taken:
                by_shape.setdefault(a.shape, []).append((a, fut))
            for group in by_shape.values():
                live = [(a, f) for a, f in group if f.set_running_or_notify_cancel()]
                if not live:
                    continue
                try:
                    labels = np.asarray(self.model.classify_images(np.stack([a for a, _ in live])))
                    if labels.shape != (len(live),):
                        raise RuntimeError("classify_images returned shape {} for {} images".format(labels.shape, len(live)))
                except BaseException as e:        # delivered to the callers; the worker keeps serving
                    for _, f in live:
                        f.set_exception(e)
                    continue
                self.batches += 1
                self.images += len(live)
                for (_, f), lab in zip(live, labels):
                    f.set_result(int(lab))


def list_images(root: str, rng: Optional[random.Random] = None) -> List[str]:
    """backend/src/main.rs:55-62: every file of every sub-directory of ``root``, shuffled once."""
    paths: List[str] = []
    for sub in sorted(os.listdir(root)):
        d = os.path.join(root, sub)
        paths.extend(os.path.join(d, f) for f in sorted(os.listdir(d)))   # (a non-directory entry is an error there too)
    (rng or random).shuffle(paths)
    return paths


class RCNState:
    """``RCNState`` (backend/src/main.rs:10-13): the model and the image pool, plus the request batcher."""

    def __init__(self, model, image_paths: Sequence[str], max_batch: int = 1024, max_delay_s: float = 0.002,
                 rng: Optional[random.Random] = None):
        if not image_paths:
            raise ValueError("no images to serve")   # gen_range(0..0) panics in the reference (main.rs:28)
        self.model = model
        self.image_paths = list(image_paths)
        self.batcher = ClassifyBatcher(model, max_batch, max_delay_s)
        self.rng = rng or random.Random()
        self._rng_lock = threading.Lock()

    def get_rcn_result(self) -> dict:
        """``get_rcn_result`` (backend/src/main.rs:22-42): pick a random file, classify it, return class + PNG bytes."""
        from PIL import Image
        with self._rng_lock:
            path = self.image_paths[self.rng.randrange(len(self.image_paths))]
        with Image.open(path) as im:
            im.load()
            pixels = grayscale_u8(im)                               # .grayscale() of rcn.rs:83 (image-crate luma weights)
            buf = io.BytesIO()
            im.save(buf, format="PNG")                              # img.write_to(.., Png) of main.rs:33-35
        output = self.batcher.classify(pixels)
        return {"output": output, "img": base64.standard_b64encode(buf.getvalue()).decode("ascii")}

    def close(self):
        self.batcher.close()


def _handler(state: RCNState):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _send(self, code: int, body: bytes, ctype: str):
            self.send_response(code)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(body)))
            self.send_header("Access-Control-Allow-Origin", "*")   # Cors::permissive() (main.rs:66)
            self.end_headers()
            self.wfile.write(body)

        def do_OPTIONS(self):
            self.send_response(200)
            self.send_header("Access-Control-Allow-Origin", "*")
            self.send_header("Access-Control-Allow-Methods", "GET, OPTIONS")
            self.send_header("Access-Control-Allow-Headers", "*")
            self.send_header("Content-Length", "0")
            self.end_headers()

        def do_GET(self):
            route = self.path.split("?", 1)[0]
            if route == "/health":
                self._send(200, b"Healthy!", "text/plain; charset=utf-8")
            elif route == "/":
                try:
                    body = json.dumps(state.get_rcn_result()).encode()
                except Exception as e:                            # `?` on classify -> 500 with the error text (main.rs:30)
                    self._send(500, str(e).encode(), "text/plain; charset=utf-8")
                    return
                self._send(200, body, "application/json")
            else:
                self._send(404, b"", "text/plain; charset=utf-8")

        def log_message(self, fmt, *args):                        # the reference logs at debug level only (main.rs:37)
            if os.environ.get("RCN_SERVE_LOG"):
                super().log_message(fmt, *args)

    return Handler


def make_server(state: RCNState, host: str = "127.0.0.1", port: int = 8080) -> ThreadingHTTPServer:
    """The reference binds 127.0.0.1:8080 (backend/src/main.rs:72); ``port=0`` picks a free one (tests)."""
    srv = ThreadingHTTPServer((host, port), _handler(state))
    srv.daemon_threads = True
    return srv


def main(argv=None):
    """``python -m mercer_research_b200.serving [--model ../rcn/rcn.bin] [--images images]`` (backend/src/main.rs:49-74)."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="../rcn/rcn.bin")
    ap.add_argument("--images", default="images")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8080)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--max-batch", type=int, default=1024)
    ap.add_argument("--max-delay-ms", type=float, default=2.0)
    args = ap.parse_args(argv)
    from .rcn import RCN
    model = RCN.load(args.model, device=args.device)              # bincode::deserialize(rcn.bin) (main.rs:54,68)
    state = RCNState(model, list_images(args.images), args.max_batch, args.max_delay_ms * 1e-3)
    srv = make_server(state, args.host, args.port)
    try:
        srv.serve_forever()
    finally:
        srv.server_close()
        state.close()


if __name__ == "__main__":
    main()
