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


def _describe(value):
    """JSON-able description of one checkpoint value: type, and shape/dtype or the value itself when small."""
    import torch
    if isinstance(value, torch.Tensor):
        return {"type": "Tensor", "shape": list(value.shape), "dtype": str(value.dtype)}
    if isinstance(value, np.ndarray):
        return {"type": "ndarray", "shape": list(value.shape), "dtype": str(value.dtype), "value": value.tolist()}
    if isinstance(value, (list, tuple)):
        return {"type": type(value).__name__,
                "value": [v if isinstance(v, (int, float, str, bool, type(None))) else str(v) for v in value]}
    if isinstance(value, (int, float, str, bool, type(None))):
        return {"type": type(value).__name__, "value": value}
    return {"type": type(value).__name__}


def checkpoint_manifest(ckpt):
    """Structure of a `modelcheckpoint.tar` dict (misc.py:21-35): must match tests/golden/make_checkpoint_manifest.py's use."""
    opt = ckpt["optimizer"]
    group = opt["param_groups"][0]
    first = opt["state"][group["params"][0]]
    return {
        "top_level": {k: _describe(v) if k not in ("state_dict", "optimizer") else {"type": type(v).__name__} for k, v in ckpt.items()},
        "state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in ckpt["state_dict"].items()],
        "optimizer": {
            "keys": sorted(opt.keys()),
            "n_param_groups": len(opt["param_groups"]),
            "param_group_keys": sorted(group.keys()),
            "param_group_values": {k: _describe(v) for k, v in group.items() if k != "params"},
            "params": list(group["params"]),
            "n_state": len(opt["state"]),
            "state_entry": {k: _describe(v) for k, v in first.items()},
            "state_shapes": [list(opt["state"][i]["exp_avg"].shape) for i in group["params"]],
        },
    }
