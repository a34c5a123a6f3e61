"""CPU: the synthetic window generator bench.py, smoke runs and run_train.py feed the path with when the reference's data layer is
not installed (signaltrain_b200/data.py -- a stand-in with the reference's shapes and knob statistics, not a port of audio.py):
shapes, dtypes, determinism by seed, knob range, the target really is a function of the knobs, recycling for validation sets."""
import numpy as np
import pytest

from signaltrain_b200 import data


@pytest.mark.parametrize("name,K", [("comp_4c", 4), ("comp_2k", 2), ("denoise", 1)])
def test_make_pool_shapes_and_determinism(name, K):
    effect = data.EFFECTS[name]()
    assert len(effect.knob_names) == K and np.asarray(effect.knob_ranges).shape == (K, 2)
    x, y, k = data.make_pool(16, 8192, 2048, effect, 44100, seed=5)
    assert x.shape == (16, 8192) and y.shape == (16, 2048) and k.shape == (16, K)
    assert x.dtype == y.dtype == k.dtype == np.float32
    assert x.flags.c_contiguous and y.flags.c_contiguous and k.flags.c_contiguous        # the C ABI takes dense rows
    assert np.isfinite(x).all() and np.isfinite(y).all()
    assert k.min() >= -0.5 and k.max() <= 0.5                                           # Beta(0.8, 0.8) - 0.5, datasets.py:325
    assert 0.05 < np.abs(x).max() <= 1.5
    x2, y2, k2 = data.make_pool(16, 8192, 2048, effect, 44100, seed=5)
    assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(k, k2)
    x3, _, _ = data.make_pool(16, 8192, 2048, effect, 44100, seed=6)
    assert not np.array_equal(x, x3)


def test_compressor_target_depends_on_the_knobs():
    fx = data.Compressor_4c()
    rng = np.random.RandomState(0)
    x = 0.8 * np.sin(2 * np.pi * 440.0 * np.arange(8192) / 44100.0)[None, :].repeat(2, axis=0)
    knobs = np.array([[-0.5, 0.5, 0.0, 0.0], [0.5, -0.5, 0.0, 0.0]])        # (threshold -30 dB, ratio 5) vs (0 dB, ratio 1)
    y, x_out = fx.apply(x, knobs, 44100)
    assert np.array_equal(x_out, x)
    assert np.abs(y[0, 4096:]).max() < 0.5 * np.abs(y[1, 4096:]).max()       # the hard setting really compresses
    np.testing.assert_allclose(y[1], x[1], atol=1e-6)                        # ratio 1 above a 0 dB threshold: identity
    wc = fx.knobs_wc(knobs[0])
    np.testing.assert_allclose(wc, [-30.0, 5.0, 0.0205, 0.0205])
    del rng


def test_window_batches_regenerate_or_recycle():
    fx = data.Denoise()
    fresh = data.SynthWindowBatches(4096, fx, datapoints=12, batch_size=4, y_size=1024, seed=3)
    assert len(fresh) == 3
    e0 = [tuple(t.clone() for t in b) for b in fresh]
    e1 = [tuple(t.clone() for t in b) for b in fresh]
    assert len(e0) == len(e1) == 3 and e0[0][0].shape == (4, 4096) and e0[0][1].shape == (4, 1024) and e0[0][2].shape == (4, 1)
    assert not all(np.array_equal(a[0].numpy(), b[0].numpy()) for a, b in zip(e0, e1))       # a new epoch, new windows
    val = data.SynthWindowBatches(4096, fx, datapoints=8, batch_size=4, y_size=1024, recycle=True, augment=False, seed=3)
    v0 = [b[0].numpy().copy() for b in val]
    v1 = [b[0].numpy().copy() for b in val]
    assert all(np.array_equal(a, b) for a, b in zip(v0, v1))                                 # validation sets are recycled
