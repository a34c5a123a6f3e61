"""1-cycle learning-rate look-up table (host side): mirror of `signaltrain/learningrate.py:14-52`.
lr rises on a half cosine from lr_max/15 to lr_max over the first 30 % of the iterations, then anneals on a
half cosine to lr_max/1500.  The momentum table is produced for interface parity; Adam ignores it
(train.py:151)."""
import numpy as np


def _half_cosine(a, b, n):
    """n points from a to b along (1 - cos)/2."""
    return a + (b - a) * (1.0 - np.cos(np.linspace(0, np.pi, n))) / 2.0


def get_1cycle_schedule(lr_max=1e-3, n_data_points=8000, epochs=200, batch_size=40):
    n_iter = n_data_points * epochs // batch_size
    n_up = int(n_iter * 0.3)
    n_down = n_iter - n_up
    lr_lo = lr_max / 15.0
    lrs = np.concatenate((_half_cosine(lr_lo, lr_max, n_up), _half_cosine(lr_max, lr_lo / 100.0, n_down)))
    moms = np.concatenate((_half_cosine(0.95, 0.85, n_up), _half_cosine(0.85, 0.95, n_down)))
    return lrs, moms
