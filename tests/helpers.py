"""Shared test helpers: golden loading and oracle parameter reconstruction."""
import os

import numpy as np

from oracle import st_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DFT_ROWS = np.array([0, 1, 2, 3, 100, 255, 256, 511, 512, 513, 514, 700, 1022, 1023])


def perturbation(shape, seed, scale):
    # must match tests/golden/make_goldens.py:perturbation
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def load_case(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = O.model_dims(float(g["meta/scale_factor"]), float(g["meta/shrink_factor"]), int(g["meta/num_knobs"]))
    assert d.C == int(g["meta/in_chunk_size"]) and d.L == int(g["meta/out_chunk_size"])
    return g, d


def initial_params(g, d):
    """Rebuild the 40 initial tensors: AE weights from the golden, DFT tensors from the oracle's own
    init (checked against the golden's row subset by test_oracle_vs_golden), plus the case's perturbation."""
    P = {}
    Wr, Wi, Sr, Si = O.dft_init(d.N, d.H)
    for i, (k, w) in enumerate(zip(O.DFT_KEYS, (Wr, Wi, Sr, Si))):
        w = w.reshape(d.N, 1, d.N).copy()
        if "meta/perturb_seed" in g:
            w = w + perturbation(w.shape, int(g["meta/perturb_seed"]) + i, float(g["meta/perturb_scale"]))
        P[k] = w
    for name, shape in O.param_order(d)[4:]:
        P[name] = g["init/" + name].astype(np.float32)
        assert P[name].shape == tuple(shape)
    return P


def dft_summary(a, N):
    a = np.asarray(a).reshape(N, -1)
    return a[DFT_ROWS], np.array([np.abs(a.astype(np.float64)).sum(), a.astype(np.float64).sum()])


def assert_params_close_after_adam(got, ref, name, atol=3e-6, hard=2.5e-5, frac=2e-3):
    """Parameters after a few Adam steps.  Adam moves every weight by ~lr per step whatever the gradient's size,
    so a weight whose gradient is at fp32-noise level can take a step of the opposite sign in two correct
    implementations: allow a fraction `frac` of elements to differ by up to `hard` (~ steps * 2 * lr)."""
    diff = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64))
    assert diff.max() <= hard, (name, diff.max())
    assert (diff > atol).mean() <= frac, (name, (diff > atol).mean(), diff.max())
