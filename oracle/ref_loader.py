"""Test infrastructure: import the UNMODIFIED reference package from where it lies (read-only tree), with the four
environment shims of SURVEY.md section 8(c), and run its own train step on the CPU.  Nothing of the reference is copied.

Used by tests/golden/make_goldens.py-style scripts, scripts/measure_port_vs_reference.py and bench.py's `--impl reference`
arm when the reference tree is present (it is NOT present on the GPU box; the arm then times the numpy port instead)."""
import os
import sys
import time
from unittest import mock

import numpy as np

DEFAULT_PATHS = (os.environ.get("SIGNALTRAIN_REFERENCE", ""), "/root/reference")


def find_reference():
    for p in DEFAULT_PATHS:
        if p and os.path.isdir(os.path.join(p, "signaltrain")):
            return p
    return None


def load_reference(path=None):
    """Returns the reference's `signaltrain` package (imported with shims, no edits to its files), or None."""
    path = path or find_reference()
    if path is None:
        return None
    import scipy.signal
    import scipy.signal.windows
    import torch
    scipy.signal.hamming = scipy.signal.windows.hamming        # removed from scipy.signal (cls_fe_dft.py:38,148)
    scipy.signal.cosine = scipy.signal.windows.cosine          # cls_fe_dct_bases.py:10
    torch.has_cudnn = False                                    # build flag that gates unconditional .cuda() (cls_fe_dft.py:43-45)
    for name in ("librosa", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        sys.modules.setdefault(name, mock.MagicMock())
    if path not in sys.path:
        sys.path.insert(0, path)
    import signaltrain as st
    return st


def reference_step_rate(st, x, y, knobs, steps, warmup, threads, scale=1, shrink=4):
    """The reference's own loop body (train.py:104-151) on the CPU: model.forward, calc_loss, zero_grad, backward,
    clip_grad_norm_, Adam.step, lr poke.  x (B, C), y (B, L), knobs (B, K) numpy float32.  Returns (frames/s, ms/step)."""
    import torch
    torch.set_num_threads(threads)
    torch.manual_seed(218)
    B, C = x.shape
    model = st.nn_proc.st_model(scale_factor=scale, shrink_factor=shrink, num_knobs=knobs.shape[1], sr=44100)
    lr_sched, _ = st.learningrate.get_1cycle_schedule(lr_max=1e-4, n_data_points=200000, epochs=1000, batch_size=200)
    opt = torch.optim.Adam(model.parameters(), lr=lr_sched[0], weight_decay=0)
    xt, yt, kt = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(knobs)
    sbf = None

    def step(i):
        nonlocal sbf
        y_hat, mag, mag_hat = model.forward(xt, kt)
        if sbf is None:
            F = mag_hat.size()[-1]
            sbf = torch.exp(7.0 / F * torch.arange(0., F)).expand_as(mag_hat)
        loss = st.loss_functions.calc_loss(y_hat.float(), yt.float(), mag_hat.float(), batch_size=B, scale_by_freq=sbf)
        opt.zero_grad()
        loss.backward()
        model.clip_grad_norm_()
        opt.step()
        opt.param_groups[0]['lr'] = float(lr_sched[min(i, len(lr_sched) - 1)])
        return float(loss.item())

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    return B * C * steps / dt, 1e3 * dt / steps
