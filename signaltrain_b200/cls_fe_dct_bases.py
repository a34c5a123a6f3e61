"""Cosine-modulated (MDCT-style) front-end variant: host-side mirror of the reference's `signaltrain/cls_fe_dct_bases.py`
(`core_modulation` :57-97, `Analysis` :100-135, `Synthesis` :138-179, `tied_transform` :36-54).

The reference ships these classes but imports them nowhere; they are mirrored here as stand-alone transforms with the
reference's constructor arguments, attribute names and `state_dict` keys (`conv_analysis.weight`, `conv_analysis.bias`,
`conv_synthesis.weight`).  The torch conv modules are parameter containers only (built in the reference's order, so a
seeded construction yields the same random bias); the arithmetic runs in the CUDA library through the same tensor-core
contraction + overlap-add machinery as the DFT front-end (st_dct_analysis / st_dct_synthesis).  Forward only.
"""
import numpy as np
import torch
import torch.nn as nn

from .engine import Engine, Geometry


def core_modulation(freq_subbands, window_size):
    """Cosine-modulated analysis / synthesis matrix (freq_subbands, window_size):
    w[n] cos(pi / M (k + 1/2)(n + 1/2 + M/2)) sqrt(2 / M), w the sine ("cosine") window  (cls_fe_dct_bases.py:57-97)."""
    n = np.arange(window_size)
    w = np.sin(np.pi * (n + 0.5) / window_size)                       # scipy.signal.cosine
    kvec = np.arange(0, freq_subbands) + 0.5
    nvec = n + 0.5 + freq_subbands / 2
    cos_an = w * np.cos(np.pi / freq_subbands * kvec[np.newaxis].T * nvec) * np.sqrt(2. / freq_subbands)
    return cos_an.astype(np.float32, copy=False)


_engines = {}


def _engine(device):
    """Any handle on the device will do: the DCT entry points take their geometry per call."""
    eng = _engines.get(device.index)
    if eng is None:
        eng = Engine(Geometry(1, 4, 1), device)
        _engines[device.index] = eng
    return eng


def _cuda_float(t, what):
    if isinstance(t, np.ndarray):                                     # the reference's Analysis.forward takes numpy (:130)
        t = torch.from_numpy(t).cuda()
    if not t.is_cuda:
        raise RuntimeError(f"signaltrain_b200: {what} needs a CUDA tensor (no CPU fallback)")
    return t.float().contiguous()


class Analysis(nn.Module):
    """Conv1d(1 -> ft_size, kernel w_size, stride hop_size, padding ft_size, bias=True), output transposed to
    (B, frames, ft_size)."""

    def __init__(self, ft_size=1024, w_size=2048, hop_size=1024, shrink=False):
        super().__init__()
        self.batch_size = None
        self.time_domain_samples = None
        self.sz = ft_size
        self.wsz = w_size
        self.hop = hop_size
        self.conv_analysis = nn.Conv1d(1, self.sz, self.wsz, padding=self.sz, stride=self.hop, bias=True)
        self.initialize()

    def initialize(self):
        with torch.no_grad():
            self.conv_analysis.weight.copy_(torch.from_numpy(core_modulation(self.sz, self.wsz)[:, None, :]))

    def forward(self, wave_form):
        x = _cuda_float(wave_form, "Analysis.forward")
        x = x.reshape(x.shape[0], -1)
        with torch.no_grad():
            return _engine(x.device).dct_analysis(x, self.conv_analysis.weight.detach().contiguous(),
                                                  self.conv_analysis.bias.detach().contiguous(), self.sz, self.wsz, self.hop)


class Synthesis(nn.Module):
    """ConvTranspose1d(ft_size -> 1, kernel w_size, stride hop_size, no bias), ft_size samples trimmed on both sides."""

    def __init__(self, ft_size=1024, w_size=2048, hop_size=1024):
        super().__init__()
        self.batch_size = None
        self.time_domain_samples = None
        self.sz = ft_size
        self.wsz = w_size
        self.hop = hop_size
        self.half_N = int(self.sz / 2 + 1)
        self.conv_synthesis = nn.ConvTranspose1d(self.sz, 1, self.wsz, padding=0, stride=self.hop, bias=False)
        self.h_tanh = torch.nn.Hardtanh()
        self.tanh = torch.nn.Tanh()
        self.initialize()

    def initialize(self):
        with torch.no_grad():
            self.conv_synthesis.weight.copy_(torch.from_numpy(core_modulation(self.sz, self.wsz)[:, None, :]))

    def forward(self, x_ft):
        x = _cuda_float(x_ft, "Synthesis.forward")
        with torch.no_grad():
            return _engine(x.device).dct_synthesis(x, self.conv_synthesis.weight.detach().contiguous(), self.sz, self.wsz, self.hop)


def tied_transform(analysis, x_ft, hop):
    """Reconstruction through the ANALYSIS weights (transposed convolution, padding = ft_size; cls_fe_dct_bases.py:36-54)."""
    x = _cuda_float(x_ft, "tied_transform")
    w = analysis.conv_analysis.weight.detach().contiguous()
    with torch.no_grad():
        return _engine(x.device).dct_synthesis(x, w, w.shape[0], w.shape[2], hop)
