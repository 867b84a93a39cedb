"""CPU (gloo, world_size 2) tests of the data-parallel host logic: shard -> local gradient sums -> ONE all-reduce ->
identical update on every replica == the single-process train_batch of the global minibatch (rcn.rs:176-223).
The local gradient producer here is the oracle (tests may use it); on a GPU box it is the CUDA model."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = [1, 3]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class CheckerBackend:
    """Same interface as mercer_research_b200.RCN for the trainer, computed by the CPU oracle."""

    def __init__(self, shapes, params, mean, sd, classes):
        import oracle as O
        self.O = O
        self.net = O.Net(shapes)
        self.params = np.array(params, copy=True)
        self.mean, self.sd, self.classes = mean, sd, classes
        self.n_params = self.net.n_params
        self.g = None
        self.device = 0

    def bind_gradient_buffer(self, t):
        self.g = t

    def accumulate_gradients_images(self, images, labels):
        import torch
        O = self.O
        X = O.standardise(O.features_u8(CFG, np.asarray(images)), self.mean, self.sd)
        _, g = self.net.train_batch(self.params, X, np.eye(self.classes)[np.asarray(labels)], 0.0)
        self.g.copy_(torch.from_numpy(g))

    def apply_gradients(self, eta, batch):
        self.params = self.params - (eta / float(batch)) * self.g.numpy()

    def last_batch_stats(self):
        return 0.0, 0


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle as O
    from mercer_research_b200.trainer import DataParallelTrainer, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0x5EED)
    B = 16
    images = rng.integers(0, 256, size=(B, 28, 28), dtype=np.uint8)
    labels = (np.arange(B) % 10).astype(np.int64)
    shapes = [(30, 784), (10, 30)]
    params = np.random.default_rng(0xC0FFEE).standard_normal(O.Net(shapes).n_params) * 0.05
    raw = O.features_u8(CFG, images)
    mean, sd = O.gen_scales(raw)
    backend = CheckerBackend(shapes, params, mean, sd, 10)
    trainer = DataParallelTrainer(backend, eta=3.0, device=torch.device("cpu"))
    assert trainer.world == world and trainer.rank == rank
    for _ in range(3):
        trainer.step_global_images(images, labels)
    lo, hi = shard_bounds(B, rank, world)
    np.save(os.path.join(out_dir, f"params_{rank}.npy"), backend.params)
    np.save(os.path.join(out_dir, f"shard_{rank}.npy"), np.array([lo, hi]))
    dist.destroy_process_group()


def test_shard_bounds():
    from mercer_research_b200.trainer import shard_bounds
    assert [shard_bounds(1024, r, 4) for r in range(4)] == [(0, 256), (256, 512), (512, 768), (768, 1024)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)


def test_dp_two_ranks_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    import oracle as O
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0 = np.load(tmp_path / "params_0.npy")
    p1 = np.load(tmp_path / "params_1.npy")
    assert np.array_equal(p0, p1), "replicas diverged"
    assert np.load(tmp_path / "shard_0.npy").tolist() == [0, 8] and np.load(tmp_path / "shard_1.npy").tolist() == [8, 16]
    # single-process reference: three train_batch steps on the global batch
    rng = np.random.default_rng(0x5EED)
    images = rng.integers(0, 256, size=(16, 28, 28), dtype=np.uint8)
    labels = (np.arange(16) % 10).astype(np.int64)
    net = O.Net([(30, 784), (10, 30)])
    params = np.random.default_rng(0xC0FFEE).standard_normal(net.n_params) * 0.05
    raw = O.features_u8(CFG, images)
    mean, sd = O.gen_scales(raw)
    X = O.standardise(raw, mean, sd)
    for _ in range(3):
        params, _ = net.train_batch(params, X, np.eye(10)[labels], 3.0)
    err = np.max(np.abs(p0 - params)) / np.max(np.abs(params))
    assert err < 1e-12, err   # only the summation order differs


def test_captured_step_is_recaptured_when_its_baked_in_state_moved(monkeypatch):
    """Host logic of DataParallelTrainer._replay (no GPU): a captured step holds device pointers and scale_set by value,
    so it is re-captured when the library's allocation generation or the model's scale_set differ from capture time, and
    only then; epoch_steps() = whole replays + an eager remainder."""
    from mercer_research_b200 import _lib
    from mercer_research_b200.trainer import DataParallelTrainer

    class Graph:
        def __init__(self, log):
            self.log = log

        def replay(self):
            self.log.append("replay")

    class Model:
        scale_set = (1.0, 1.0)

    gen = {"v": 7}
    monkeypatch.setattr(_lib, "allocation_generation", lambda: gen["v"])
    log = []
    tr = DataParallelTrainer.__new__(DataParallelTrainer)
    tr.model, tr.world, tr.p2p = Model(), 1, False

    def capture():
        log.append("capture")
        tr.graph = Graph(log)
        tr._graph_generation = _lib.allocation_generation()
        tr._graph_scale = tr.model.scale_set

    monkeypatch.setattr(tr, "_capture_graph", capture)
    monkeypatch.setattr(tr, "_epoch_step_eager", lambda: log.append("eager"))
    tr.steps_per_graph = 4
    capture()
    del log[:]
    tr.epoch_steps(8)
    assert log == ["replay", "replay"]
    gen["v"] = 8                                   # a library buffer moved
    del log[:]
    tr.epoch_step()
    tr.epoch_step()
    assert log == ["capture", "replay", "replay"]
    tr.model.scale_set = (2.0, 3.0)                # scale_set changed
    del log[:]
    tr.epoch_steps(10)                             # 2 replays + 2 eager steps
    assert log == ["capture", "replay", "replay", "eager", "eager"]
    tr.graph = None
    del log[:]
    tr.epoch_steps(3)
    assert log == ["eager"] * 3
