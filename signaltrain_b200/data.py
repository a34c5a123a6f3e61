"""Synthetic (x, y, knobs) windows of the reference's shapes for benchmarks, smoke runs and run_train.py when
the reference's own data layer (signaltrain/datasets.py + audio.py, out of scope here: SURVEY.md section 2.1
rows 9-10) is not installed.  Not a port of those files: a small vectorised generator that produces whole
batches at once -- test tones / plucks / gated noise through a feed-forward compressor (4 knobs), a 2-knob
variant, or additive-noise "denoise" pairs (1 knob)."""
import numpy as np
from scipy.signal import lfilter


class Effect:
    """Duck-type of the reference's audio.Effect (audio.py:449-480): name, knob_names, knob_ranges, info()."""
    name = "effect"
    knob_names = ["knob"]
    knob_ranges = np.array([[0.0, 1.0]])

    def info(self):
        print(f"Effect: {self.name}.  Knobs:")
        for n, r in zip(self.knob_names, self.knob_ranges):
            print(f"    {n}: {r[0]} to {r[1]}")

    def knobs_wc(self, knobs_nn):
        r = np.asarray(self.knob_ranges, dtype=np.float64)
        return r[:, 0] + (np.asarray(knobs_nn) + 0.5) * (r[:, 1] - r[:, 0])


def _compress(x, thresh_db, ratio, t_attack, t_release, sr):
    """Batch feed-forward compressor: static curve in dB, then one-pole smoothing of the gain with a time constant
    between attack and release (an LTI stand-in for the reference's branchy smoother, audio.py:380-426)."""
    x_db = 20.0 * np.log10(np.abs(x) + 1e-8)
    x_db = np.maximum(x_db, -96.0)
    over = np.maximum(x_db - thresh_db[:, None], 0.0)
    gain_db = over / ratio[:, None] - over
    tau = 0.5 * (t_attack + t_release)
    alpha = np.exp(-np.log(9.0) / (sr * tau))
    out = np.empty_like(gain_db)
    for i in range(x.shape[0]):                       # per-window filter coefficient
        out[i] = lfilter([1.0 - alpha[i]], [1.0, -alpha[i]], gain_db[i])
    return x * np.power(10.0, out / 20.0)


class Compressor_4c(Effect):
    name = "Compressor_4c"
    knob_names = ["threshold", "ratio", "attackTime", "releaseTime"]
    knob_ranges = np.array([[-30, 0], [1, 5], [1e-3, 4e-2], [1e-3, 4e-2]], dtype=np.float64)

    def apply(self, x, knobs_nn, sr):
        k = np.stack([self.knobs_wc(kk) for kk in knobs_nn])
        return _compress(x, k[:, 0], k[:, 1], k[:, 2], k[:, 3], sr), x


class Compressor_2knob(Effect):
    """LA2A-style: threshold and ratio free, attack/release pinned (BASELINE.json configs[3])."""
    name = "Compressor_2knob"
    knob_names = ["threshold", "ratio"]
    knob_ranges = np.array([[-30, 0], [1, 5]], dtype=np.float64)

    def apply(self, x, knobs_nn, sr):
        k = np.stack([self.knobs_wc(kk) for kk in knobs_nn])
        n = x.shape[0]
        return _compress(x, k[:, 0], k[:, 1], np.full(n, 0.01), np.full(n, 0.02), sr), x


class Denoise(Effect):
    """Input = clean + uniform noise of knob-controlled strength, target = clean (cf. audio.py:558-571)."""
    name = "Denoise"
    knob_names = ["strength"]
    knob_ranges = np.array([[0.0, 0.1]], dtype=np.float64)

    def apply(self, x, knobs_nn, sr):
        s = np.stack([self.knobs_wc(kk) for kk in knobs_nn])[:, 0]
        noisy = x + (2 * np.random.rand(*x.shape) - 1) * s[:, None]
        return x, noisy


EFFECTS = {"comp_4c": Compressor_4c, "comp_2k": Compressor_2knob, "denoise": Denoise}


def synth_inputs(n, chunk, sr=44100, rng=None):
    """n test signals of `chunk` samples: sines, noisy sines, decaying plucks, gated ('box') tones and noise."""
    rng = rng or np.random
    t = np.arange(chunk, dtype=np.float64) / sr
    kind = rng.randint(0, 5, size=n)
    freq = np.exp(rng.uniform(np.log(40.0), np.log(6000.0), size=(n, 1)))
    amp = rng.uniform(0.05, 0.9, size=(n, 1))
    phase = rng.uniform(0, 2 * np.pi, size=(n, 1))
    x = amp * np.sin(2 * np.pi * freq * t[None, :] + phase)
    t0 = rng.uniform(0.0, 0.6, size=(n, 1)) * t[-1]
    decay = np.exp(-np.maximum(t[None, :] - t0, 0.0) * rng.uniform(20, 200, size=(n, 1))) * (t[None, :] >= t0)
    width = rng.uniform(0.1, 0.5, size=(n, 1)) * t[-1]
    box = ((t[None, :] >= t0) & (t[None, :] < t0 + width)).astype(np.float64)
    noise = rng.uniform(-1, 1, size=(n, chunk))
    out = np.where((kind == 0)[:, None], x, 0.0)
    out = out + np.where((kind == 1)[:, None], x + 0.1 * amp * noise, 0.0)
    out = out + np.where((kind == 2)[:, None], x * decay, 0.0)
    out = out + np.where((kind == 3)[:, None], x * box, 0.0)
    out = out + np.where((kind == 4)[:, None], amp * noise * box, 0.0)
    return out * rng.choice([-1.0, 1.0], size=(n, 1))


def make_pool(n, chunk, y_size, effect, sr=44100, seed=218, augment=True):
    """Returns float32 arrays x (n, chunk), y (n, y_size), knobs (n, K) with knobs ~ Beta(0.8, 0.8) - 0.5."""
    rng = np.random.RandomState(seed)
    x = synth_inputs(n, chunk, sr, rng)
    knobs = rng.beta(0.8, 0.8, size=(n, len(effect.knob_names))) - 0.5
    state = np.random.get_state()
    np.random.seed(seed + 1)
    y, x = effect.apply(x, knobs, sr)
    np.random.set_state(state)
    if augment:                                           # random polarity flip of the pair
        flip = rng.choice([-1.0, 1.0], size=(n, 1))
        x, y = x * flip, y * flip
    return (np.ascontiguousarray(x, dtype=np.float32), np.ascontiguousarray(y[:, -y_size:], dtype=np.float32),
            np.ascontiguousarray(knobs, dtype=np.float32))


class SynthWindowBatches:
    """Iterable of (x, y, knobs) CPU tensors, `datapoints // batch_size` batches per epoch, regenerated on the fly
    (or recycled for a validation set)."""

    def __init__(self, chunk_size, effect, sr=44100, datapoints=8000, batch_size=200, y_size=None, augment=True,
                 recycle=False, seed=218):
        self.chunk, self.effect, self.sr = chunk_size, effect, sr
        self.datapoints, self.batch_size = datapoints, batch_size
        self.y_size = chunk_size if y_size is None else y_size
        self.augment, self.recycle, self.seed = augment, recycle, seed
        self._epoch = 0
        self._cache = None

    def __len__(self):
        return max(1, self.datapoints // self.batch_size)

    def __iter__(self):
        import torch
        if self.recycle and self._cache is not None:
            yield from self._cache
            return
        out = []
        for i in range(len(self)):
            x, y, k = make_pool(self.batch_size, self.chunk, self.y_size, self.effect, self.sr,
                                seed=self.seed + 7919 * self._epoch + i, augment=self.augment)
            item = (torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(k))
            if self.recycle:
                out.append(item)
            yield item
        self._epoch += 1
        if self.recycle:
            self._cache = out
