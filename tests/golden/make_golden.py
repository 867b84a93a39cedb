"""Generates tests/golden/*.npz from the CPU oracle (oracle/rcn_oracle.cpp).

The reference (Rust) cannot be built in this environment, so these are NOT outputs of the reference itself:
they are outputs of the line-by-line restatement, cross-checked against the independent numpy restatement and
the hand-derived KATs of SURVEY.md Appendix B. Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(0x5EED)
    # --- feature stage: 8 MNIST-shaped u8 images through [C(Same),P] and the CLI default CPCP ---------------
    imgs = rng.integers(0, 256, size=(8, 28, 28), dtype=np.uint8)
    out = {"images": imgs}
    for name, cfg in [("cp", [1, 3]), ("cpcp", [1, 3, 1, 3]), ("c_none_p", [0, 3]), ("cc", [1, 1])]:
        out["feat_" + name] = O.features_u8(cfg, imgs)
    f = out["feat_cp"]
    mean, sd = O.gen_scales(f)
    out["scale_cp"] = np.array([mean, sd])
    out["std_cp"] = O.standardise(f, mean, sd)
    np.savez_compressed(os.path.join(HERE, "features_mnist8.npz"), **out)

    # --- dense stage: 784-30-10 (config M), 8 samples, one train_batch step at eta = 3 -----------------------
    prng = np.random.default_rng(0xC0FFEE)
    net = O.Net([(30, 784), (10, 30)])
    params = prng.standard_normal(net.n_params)
    X = out["std_cp"]
    labels = np.arange(8) % 10
    Y = np.eye(10)[labels]
    new_params, grads = net.train_batch(params, X, Y, 3.0, n_threads=1)
    acts = net.forward(params, X)
    g0, zs, a, d = net.backprop(params, X[0], Y[0])
    np.savez_compressed(os.path.join(HERE, "dense_mnist8.npz"), params=params, X=X, labels=labels, new_params=new_params,
                        grads=grads, acts=acts, sample0_grads=g0, sample0_zs=zs, sample0_acts=a, sample0_deltas=d,
                        pred=O.argmax_last(acts))
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
