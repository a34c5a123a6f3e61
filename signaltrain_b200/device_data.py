"""Data step in front of the train step, on the device.

The reference generates its windows on CPU DataLoader workers (~1 k windows/s per core): `audio.compressor_4controls`
(audio.py:380-426) for the comp_4c target and `AudioFileDataSet.get_single_chunk` (datasets.py:225-253) + `do_augment`
(:21-30) for the crop / polarity flip.  Here a corpus lives in HBM, the host only draws the random numbers (same calls,
same order as the reference), and the cropping, the flip and -- with rerun_effect -- the compressor run as CUDA kernels
(st_crop_windows, st_compressor_4c)."""
import ctypes

import numpy as np
import torch

from .engine import Engine, Geometry, _ptr

_engines = {}


def _engine(device):
    eng = _engines.get(device.index)
    if eng is None:
        eng = Engine(Geometry(1, 4, 1), device)
        _engines[device.index] = eng
    return eng


def _need_cuda(t, what, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"signaltrain_b200: {what} must be a contiguous CUDA {dtype} tensor (no CPU fallback)")


def compressor_4controls(x, knobs_wc, sr=44100.0):
    """Batched audio.compressor_4controls: x (B, n) float32 CUDA, knobs_wc (B, 4) world coordinates [threshold dB, ratio,
    attackTime s, releaseTime s] (any float dtype / host array) -> y (B, n) float32 CUDA."""
    _need_cuda(x, "x")
    B, n = x.shape
    k = torch.as_tensor(np.asarray(knobs_wc.cpu() if isinstance(knobs_wc, torch.Tensor) else knobs_wc, dtype=np.float64)).reshape(B, 4)
    k = k.to(x.device).contiguous()
    y = torch.empty_like(x)
    eng = _engine(x.device)
    eng._ok(eng.lib.st_compressor_4c(eng.h, _ptr(x), ctypes.c_void_p(k.data_ptr()), B, n, float(sr), _ptr(y), eng._stream()),
            "st_compressor_4c")
    return y


def crop_windows(corpus_x, corpus_y, offsets, signs, chunk, y_size):
    """x[b] = s_b corpus_x[o_b : o_b + chunk],  y[b] = s_b corpus_y[o_b + chunk - y_size : o_b + chunk]."""
    _need_cuda(corpus_x, "corpus_x")
    _need_cuda(corpus_y, "corpus_y")
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    B = off.shape[0]
    sg = None
    if signs is not None:
        sg = torch.as_tensor(np.asarray(signs, dtype=np.float32)).to(corpus_x.device).contiguous()
    x = torch.empty((B, chunk), device=corpus_x.device, dtype=torch.float32)
    y = torch.empty((B, y_size), device=corpus_x.device, dtype=torch.float32)
    eng = _engine(corpus_x.device)
    eng._ok(eng.lib.st_crop_windows(eng.h, _ptr(corpus_x), _ptr(corpus_y), int(corpus_x.numel()), off.ctypes.data_as(ctypes.c_void_p),
                                    _ptr(sg) if sg is not None else None, B, int(chunk), int(y_size), _ptr(x), _ptr(y), eng._stream()),
            "st_crop_windows")
    torch.cuda.current_stream(corpus_x.device).synchronize()      # `off` is a pageable host array: keep it alive until copied
    return x, y


class DeviceAudioFileBatches:
    """Batches of (x, y, knobs_nn) CUDA tensors from preloaded input / target pairs held on the device: the on-device
    counterpart of AudioFileDataSet(preload=True) + DataLoader (datasets.py:64-259, train.py:235-248).

    files_x / files_y: lists of equal-length 1-D float arrays (one per file pair), knobs_wc: (nfiles, K) world coordinates.
    Per window the host draws, in the reference's order (datasets.py:230, :237, :22): the file index, the start sample and
    the polarity flip.  rerun_effect=True recomputes the target from the cropped input with the comp_4c compressor on the
    device (datasets.py:241-242) instead of cropping the stored target."""

    def __init__(self, files_x, files_y, knobs_wc, knob_ranges, chunk_size, y_size, batch_size, datapoints, device="cuda:0",
                 augment=True, rerun_effect=False, sr=44100.0):
        dev = torch.device(device)
        self.lens = [len(f) for f in files_x]
        for n in self.lens:
            assert n > chunk_size, f"Error: len(in_audio)={n}, must be > self.chunk_size={chunk_size}"        # datasets.py:236
        self.starts = np.concatenate([[0], np.cumsum(self.lens)[:-1]]).astype(np.int64)
        self.cx = torch.from_numpy(np.concatenate([np.asarray(f, np.float32) for f in files_x])).to(dev)
        self.cy = torch.from_numpy(np.concatenate([np.asarray(f, np.float32) for f in files_y])).to(dev)
        self.knobs_wc = np.asarray(knobs_wc, np.float64)
        self.kr = np.asarray(knob_ranges, np.float64)
        self.chunk, self.y_size, self.batch, self.datapoints = chunk_size, y_size, batch_size, datapoints
        self.augment, self.rerun, self.sr = augment, rerun_effect, sr

    def __len__(self):
        return max(1, self.datapoints // self.batch)

    def __iter__(self):
        for _ in range(len(self)):
            idx = np.empty(self.batch, np.int64)
            off = np.empty(self.batch, np.int64)
            sign = np.ones(self.batch, np.float32)
            for b in range(self.batch):
                i = np.random.randint(0, high=len(self.lens))
                ibgn = np.random.randint(0, self.lens[i] - self.chunk)
                idx[b], off[b] = i, self.starts[i] + ibgn
                if self.augment and np.random.choice([True, False]):
                    sign[b] = -1.0
            k_wc = self.knobs_wc[idx]
            x, y = crop_windows(self.cx, self.cy, off, None if self.rerun else sign, self.chunk, self.y_size)
            if self.rerun:                                    # effect on the un-flipped chunk, then the flip of both
                y = compressor_4controls(x, k_wc, self.sr)[:, -self.y_size:]
                s = torch.from_numpy(sign).to(x.device)[:, None]
                x, y = x * s, (y * s).contiguous()
            knobs_nn = (k_wc - self.kr[:, 0]) / (self.kr[:, 1] - self.kr[:, 0]) - 0.5                     # datasets.py:246-247
            yield x, y, torch.from_numpy(knobs_nn.astype(np.float32)).to(x.device)


# ---- file datasets: what AudioFileDataSet(preload=True) reads from disk (datasets.py:102-160), handed to the device batches ----

def parse_knob_string(filename, ext=".wav"):
    """Knob settings from a target filename (datasets.py:177-185): double underscores precede every setting, e.g.
    'target_9400_Compressor_4c__-10.95__3.428__0.005043__0.01308.wav' -> [-10.95, 3.428, 0.005043, 0.01308]."""
    import os
    parts = os.path.basename(filename).replace(ext, "").split("__")[1:]
    return np.array([float(p) for p in parts], dtype=np.float32)


def read_wav(filename, sr=44100):
    """audio.read_audio_file's scipy branch (audio.py:207-233): first channel of a multi-channel file, int16 -> float / 32767.
    Resampling (the reference falls back to librosa, audio.py:234-242) is outside this path: a sample-rate mismatch raises."""
    from scipy.io import wavfile
    read_sr, signal = wavfile.read(filename)
    if signal.ndim > 1:
        signal = signal[:, 0]
    if signal.dtype == np.int16:
        signal = np.array(signal / 32767.0, dtype=np.float32)
    if int(read_sr) != int(sr):
        raise RuntimeError(f"{filename}: sample rate {read_sr} Hz, expected {sr} Hz (resample the corpus first; the reference "
                           "would call librosa here)")
    return np.asarray(signal, dtype=np.float32)


def mu_compand(y, mu=32):
    """audio.mu_compand (audio.py:339-340)."""
    return (np.sign(y) * np.log(1 + mu * np.abs(y)) / np.log(1 + mu)).astype(np.float32)


def load_file_pairs(path, sr=44100, is_inverse=False, compand=False, max_files=100000):
    """The preload of AudioFileDataSet (datasets.py:104-160): sorted input_* / target_* pairs under `path`, knobs from the target
    names, unequal lengths aligned to their ends (align_end, :146-152), input and target swapped for inverse effects (:154-155),
    optional mu-law companding (:198-200).  Returns (files_x, files_y, knobs_wc[nfiles, K])."""
    import glob
    inputs = sorted(glob.glob(path + "/" + "input_*"))
    targets = sorted(glob.glob(path + "/" + "target_*"))
    print("AudioFileDataSet: Found", len(inputs), "input files and", len(targets), " target files in path", path)
    assert len(inputs) == len(targets)
    if not inputs:
        raise RuntimeError(f"no input_* / target_* file pairs under {path}")
    n = min(max_files, len(inputs))
    xs, ys, knobs = [], [], []
    for i in range(n):
        x, y = read_wav(inputs[i], sr), read_wav(targets[i], sr)
        knobs.append(parse_knob_string(targets[i]))
        if compand:
            x, y = mu_compand(x), mu_compand(y)
        if len(x) != len(y):
            m = min(len(x), len(y))
            x, y = x[-m:], y[-m:]
        if is_inverse:
            x, y = y, x
        xs.append(x)
        ys.append(y)
    return xs, ys, np.stack(knobs).astype(np.float32)


def file_batches(path, effect, chunk_size, y_size, batch_size, datapoints, device, sr=44100, rerun=False, augment=True,
                 compand=False):
    """DeviceAudioFileBatches over the file pairs under `path` -- the on-device counterpart of
    DataLoader(AudioFileDataSet(chunk_size, effect, path=..., preload=True, rerun=..., augment=..., compand=...)) (train.py:240-248)."""
    if rerun and len(getattr(effect, "knob_names", [])) != 4:
        raise NotImplementedError("target_type != 'stream' re-runs the effect on every chunk (datasets.py:241-242); on the device "
                                  "that exists for the 4-knob compressor only (st_compressor_4c)")
    xs, ys, knobs_wc = load_file_pairs(path, sr=sr, is_inverse=bool(getattr(effect, "is_inverse", False)), compand=compand)
    return DeviceAudioFileBatches(xs, ys, knobs_wc, np.asarray(effect.knob_ranges, np.float64), chunk_size, y_size, batch_size,
                                  datapoints, device=device, augment=augment, rerun_effect=rerun, sr=float(sr))
