"""Long-signal inference: host-side mirror of the reference's `utils/predict_long.py:30-79` (`predict_long`) and
`signaltrain/audio.py:23-49` (`sliding_window`).

The reference cuts the signal into overlapping windows on the CPU and uploads 200 windows at a time; here the waveform
goes to the device once, the windows are a strided view of it, and every batch runs the forward-only CUDA path
(st_forward with training off: no activation records are written).  Windows are independent, so the result does not
depend on how they are batched; the output bookkeeping (zero padding of the tail, `num_extra` trimming, float64 result
of `np.append`) follows the reference.
"""
import numpy as np
import torch


def sliding_window(x, size, overlap=0):
    """Stack a 1-D array into windows of `size` samples advancing by `size - overlap`, zero-padding the end so the windows
    cover it evenly (audio.py:23-49).  Returns (nwin, size)."""
    x = np.asarray(x)
    step = size - overlap
    remainder = (x.shape[-1] - size) % step
    if remainder != 0:
        x = np.pad(x, (0, step - remainder), mode="constant")
    nwin = (x.shape[-1] - size) // step + 1
    return np.lib.stride_tricks.as_strided(x, shape=(nwin, size), strides=(step * x.strides[-1], x.strides[-1]), writeable=False)


def predict_long(signal, knobs_nn, model, chunk_size, out_chunk_size, sr=44100, effect=None, device="cuda:0", compand=False,
                 batch_size=200):
    """Run `model` over a long 1-D `signal` with fixed knob settings `knobs_nn` (normalised, length K).  Returns the predicted
    signal as a float64 numpy array, like the reference (utils/predict_long.py:30-79)."""
    if compand:
        raise RuntimeError("signaltrain_b200: mu-law companding (audio.mu_compand) is outside the accelerated path")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("signaltrain_b200: predict_long needs a CUDA device (no CPU fallback)")
    signal = np.asarray(signal)
    overlap = chunk_size - out_chunk_size
    step = chunk_size - overlap
    # the same zero padding sliding_window() applies, then ONE upload; the windows are a strided view on the device
    remainder = (signal.shape[-1] - chunk_size) % step
    padded = np.pad(signal, (0, step - remainder), mode="constant") if remainder != 0 else signal
    nwin = (padded.shape[-1] - chunk_size) // step + 1
    wave = torch.from_numpy(np.ascontiguousarray(padded, dtype=np.float32)).to(dev)
    windows = wave.as_strided((nwin, chunk_size), (step, 1))
    knobs_row = torch.from_numpy(np.asarray(knobs_nn, dtype=np.float32).reshape(1, -1)).to(dev)
    was_training = model.training
    model.eval()
    outs = []
    with torch.no_grad():
        for b0 in range(0, nwin, batch_size):
            xb = windows[b0:b0 + batch_size].contiguous()
            y_hat, _, _ = model.forward(xb, knobs_row.expand(xb.shape[0], -1).contiguous())
            outs.append(y_hat.reshape(-1))
    y_pred = torch.cat(outs).cpu().numpy().astype(np.float64)          # np.append onto np.empty(0) yields float64 (:45, :70)
    model.train(was_training)
    unique = chunk_size + (nwin - 1) * (chunk_size - overlap)          # :73
    num_extra = unique - signal.size
    return y_pred[0:-num_extra] if num_extra > 0 else y_pred
