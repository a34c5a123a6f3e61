"""Data step in front of the path: the comp_4c compressor and the window cropper.
CPU: the numpy oracle against a golden minted by the reference's numba-compiled audio.compressor_4controls.
GPU: st_compressor_4c / st_crop_windows against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import st_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    return np.load(os.path.join(HERE, "golden", "compressor4c_n4096_b4.npz"))


def test_oracle_compressor_matches_reference():
    g = _golden()
    k = g["knobs_wc"]
    y = O.compressor_4controls(g["x"], k[:, 0], k[:, 1], k[:, 2], k[:, 3], 44100.0)
    assert y.dtype == np.float64 and g["y"].dtype == np.float64                 # what the reference hands to train.py:120
    np.testing.assert_allclose(y, g["y"], atol=1e-12, rtol=0)
    # edge cases of the reference: silence sits on the -96 dB floor (no gain change), ratio 1 is the identity
    z = O.compressor_4controls(np.zeros((1, 64), np.float32), -20.0, 4.0, 0.01, 0.01)
    assert np.all(z == 0)
    x = g["x"][:1]
    np.testing.assert_allclose(O.compressor_4controls(x, -30.0, 1.0, 0.01, 0.01), x.astype(np.float64), atol=1e-7)


def test_oracle_crop_windows():
    cx, cy = np.arange(100, dtype=np.float32), -np.arange(100, dtype=np.float32)
    x, y = O.crop_windows(cx, cy, [0, 90, 37], [1, -1, 1], 10, 4)
    np.testing.assert_array_equal(x[1], -np.arange(90, 100))
    np.testing.assert_array_equal(y[2], -np.arange(43, 47))


@pytest.mark.gpu
def test_cuda_compressor_matches_oracle_and_reference():
    from signaltrain_b200.device_data import compressor_4controls
    g = _golden()
    y = compressor_4controls(torch.from_numpy(g["x"]).cuda(), g["knobs_wc"], 44100.0).cpu().numpy()
    ref = g["y"].astype(np.float32)                                           # train.py:120: y.float()
    assert np.abs(y - ref).max() < 2e-7                                       # double log10 / pow: last-ulp differences of float32
    # full-size batch against the oracle (size-independent check: every window is independent of the batch it is in)
    rng = np.random.RandomState(1)
    x = (0.5 * rng.standard_normal((200, 8192))).astype(np.float32)
    k = np.stack([rng.uniform(-30, 0, 200), rng.uniform(1, 5, 200), rng.uniform(1e-3, 4e-2, 200), rng.uniform(1e-3, 4e-2, 200)], 1)
    yb = compressor_4controls(torch.from_numpy(x).cuda(), k).cpu().numpy()
    sub = [0, 57, 199]
    ref = O.compressor_4controls(x[sub], k[sub, 0], k[sub, 1], k[sub, 2], k[sub, 3]).astype(np.float32)
    assert np.abs(yb[sub] - ref).max() < 5e-7
    np.testing.assert_array_equal(yb[5:9], compressor_4controls(torch.from_numpy(x[5:9]).cuda(), k[5:9]).cpu().numpy())


@pytest.mark.gpu
def test_cuda_crop_and_device_batches():
    import signaltrain_b200 as st
    from signaltrain_b200.device_data import DeviceAudioFileBatches, crop_windows
    rng = np.random.RandomState(2)
    cx, cy = rng.standard_normal(50000).astype(np.float32), rng.standard_normal(50000).astype(np.float32)
    off, sg = np.array([0, 41808, 12345, 777]), np.array([1, -1, -1, 1], np.float32)
    x, y = crop_windows(torch.from_numpy(cx).cuda(), torch.from_numpy(cy).cuda(), off, sg, 8192, 2048)
    rx, ry = O.crop_windows(cx, cy, off, sg, 8192, 2048)
    np.testing.assert_array_equal(x.cpu().numpy(), rx)
    np.testing.assert_array_equal(y.cpu().numpy(), ry)
    with pytest.raises(RuntimeError):
        crop_windows(torch.from_numpy(cx).cuda(), torch.from_numpy(cy).cuda(), np.array([49000]), None, 8192, 2048)
    # the iterable: same host RNG calls as the reference's get_single_chunk -> reproducible against a numpy replay
    files_x = [rng.standard_normal(20000).astype(np.float32) * 0.3 for _ in range(3)]
    kwc = np.array([[-20, 3, 0.01, 0.02], [-10, 2, 0.005, 0.01], [-25, 4.5, 0.03, 0.03]])
    kr = st.data.Compressor_4c.knob_ranges
    files_y = [O.compressor_4controls(f[None], *kwc[i]).astype(np.float32)[0] for i, f in enumerate(files_x)]
    np.random.seed(7)
    ds = DeviceAudioFileBatches(files_x, files_y, kwc, kr, 8192, 2048, batch_size=4, datapoints=8, rerun_effect=True)
    got = [(a.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()) for a, b, c in ds]
    np.random.seed(7)
    for gx, gy, gk in got:
        for b in range(4):
            i = np.random.randint(0, high=3)
            ibgn = np.random.randint(0, 20000 - 8192)
            s = -1.0 if np.random.choice([True, False]) else 1.0
            xi = files_x[i][ibgn:ibgn + 8192]
            yi = O.compressor_4controls(xi[None], *kwc[i])[0, -2048:]
            np.testing.assert_array_equal(gx[b], s * xi)
            assert np.abs(gy[b] - (s * yi).astype(np.float32)).max() < 5e-7
            np.testing.assert_allclose(gk[b], (kwc[i] - kr[:, 0]) / (kr[:, 1] - kr[:, 0]) - 0.5, atol=1e-7)


def _write_pairs(root, n_files, n_samples, rng, extra_target=0):
    """input_*/target_* wav pairs with the reference's naming (datasets.py:177-185); returns (xs, ys, knobs)."""
    from scipy.io import wavfile
    os.makedirs(root, exist_ok=True)
    xs, ys, ks = [], [], []
    for i in range(n_files):
        x = (0.3 * rng.standard_normal(n_samples)).astype(np.float32)
        y = np.tanh(np.concatenate([np.zeros(extra_target, np.float32), x])).astype(np.float32)
        k = [round(float(rng.uniform(-30, 0)), 3), round(float(rng.uniform(1, 5)), 3), 0.01, 0.02]
        wavfile.write(os.path.join(root, f"input_{i}_.wav"), 44100, x)
        wavfile.write(os.path.join(root, f"target_{i}_Compressor_4c__{k[0]}__{k[1]}__{k[2]}__{k[3]}.wav"), 44100, y)
        xs.append(x); ys.append(y); ks.append(k)
    return xs, ys, np.array(ks, np.float32)


def test_file_pair_loader_follows_the_reference_conventions(tmp_path):
    """load_file_pairs = the preload of AudioFileDataSet (datasets.py:104-160, audio.py:207-233): sorted pairs, knobs from the
    target name, int16 scaling, first channel of stereo, unequal lengths aligned to their ends, inverse effects swapped,
    mu-law companding, and a loud error instead of a silent resample."""
    from scipy.io import wavfile
    from signaltrain_b200 import device_data as dd
    rng = np.random.RandomState(0)
    root = str(tmp_path / "Train")
    xs, ys, ks = _write_pairs(root, 3, 3000, rng, extra_target=7)
    fx, fy, kw = dd.load_file_pairs(root)
    assert len(fx) == 3 and kw.shape == (3, 4)
    np.testing.assert_array_equal(kw, ks)
    for a, b, x, y in zip(fx, fy, xs, ys):
        assert len(a) == len(b) == 3000                       # the 7 extra leading target samples are cut: ends aligned
        np.testing.assert_array_equal(a, x)
        np.testing.assert_array_equal(b, y[-3000:])
    ix, iy, _ = dd.load_file_pairs(root, is_inverse=True)
    np.testing.assert_array_equal(ix[0], fy[0]); np.testing.assert_array_equal(iy[0], fx[0])
    cx, _, _ = dd.load_file_pairs(root, compand=True)
    np.testing.assert_allclose(cx[1], np.sign(xs[1]) * np.log(1 + 32 * np.abs(xs[1])) / np.log(33), rtol=1e-6, atol=1e-7)
    assert list(dd.parse_knob_string("target_9400_Compressor_4c__-10.95__3.428__0.005043__0.01308.wav")) == \
        [np.float32(-10.95), np.float32(3.428), np.float32(0.005043), np.float32(0.01308)]
    # int16 stereo file: first channel, / 32767
    st = (rng.randint(-3000, 3000, size=(500, 2))).astype(np.int16)
    wavfile.write(str(tmp_path / "s.wav"), 44100, st)
    np.testing.assert_array_equal(dd.read_wav(str(tmp_path / "s.wav")), np.array(st[:, 0] / 32767.0, dtype=np.float32))
    wavfile.write(str(tmp_path / "r.wav"), 22050, xs[0])
    with pytest.raises(RuntimeError, match="sample rate"):
        dd.read_wav(str(tmp_path / "r.wav"))
    with pytest.raises(RuntimeError, match="no input_"):
        dd.load_file_pairs(str(tmp_path / "nothing"))


def test_file_batches_refuses_a_chunk_rerun_it_cannot_do_on_the_device(tmp_path):
    """target_type != "stream" re-runs the effect per chunk (datasets.py:241-242); only the 4-knob compressor exists as a device
    kernel, so any other effect is refused up front instead of silently training on stale targets."""
    from signaltrain_b200 import data, device_data as dd
    with pytest.raises(NotImplementedError, match="st_compressor_4c"):
        dd.file_batches(str(tmp_path), data.Denoise(), 8192, 2048, 4, 40, "cuda:0", rerun=True)
