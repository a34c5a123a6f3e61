"""Trainable windowed-DFT front-end: host-side mirror of the reference's `signaltrain/cls_fe_dft.py`
classes `Analysis` (:12-58) and `Synthesis` (:61-163).

The modules own the parameters under the reference's names (`conv_analysis_real.weight`, ...,
shape (N,1,N)) so `state_dict()` is wire-compatible; the arithmetic runs in the
CUDA library (analysis / synthesis contractions with the frame gather and overlap-add fused, see
csrc/).  The torch conv modules below are parameter containers only and are never called.
"""
import numpy as np
import torch
import torch.nn as nn


def hamming_window(n):
    """Symmetric Hamming window (what the reference obtains from scipy.signal.hamming, cls_fe_dft.py:38)."""
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))


def lsee_synthesis_window(wsz, hop):
    """Griffin-Lim LSEE-MSTFT synthesis window: w / sum_k w^2(n - k*hop)   (reference: Synthesis.GLA,
    cls_fe_dft.py:133-163)."""
    w = hamming_window(wsz)
    env = np.zeros(wsz)
    for k in range(-(wsz // hop), wsz // hop + 1):
        lo, hi = max(0, hop * k), min(wsz, wsz + hop * k)
        if hi > lo:
            env[lo:hi] += w[lo - hop * k: hi - hop * k] ** 2
    return w / env


def ortho_dft_rows(n):
    """Real and imaginary parts of the unitary DFT matrix exp(-2 pi i k m / n)/sqrt(n).

    Taken from numpy's FFT of the identity, as the reference does (cls_fe_dft.py:37), rather than from cos/sin:
    rows k = 0 and k = n/2 of the imaginary part are pure rounding residue (~1e-17), and the SIGN of that residue
    decides whether atan2 returns +pi or -pi for those two bins (nn_proc.py:310).  Bit-identical residue is what
    makes a model built here train exactly like one built by the reference from the same seed."""
    f = np.fft.fft(np.eye(n), norm="ortho")
    return f.real, f.imag


class _FrontEnd(nn.Module):
    def __init__(self, ft_size, hop_size):
        super().__init__()
        self.sz = ft_size
        self.hop = hop_size
        self.half_N = int(self.sz / 2 + 1)
        self.batch_size = None
        self.time_domain_samples = None

    def _load(self, conv_real, conv_imag, window):
        re, im = ortho_dft_rows(self.sz)
        with torch.no_grad():
            conv_real.weight.copy_(torch.from_numpy((re * window).astype(np.float32)[:, None, :]))
            conv_imag.weight.copy_(torch.from_numpy((im * window).astype(np.float32)[:, None, :]))


class Analysis(_FrontEnd):
    """Conv1d(1 -> N, kernel N, stride hop, padding N, no bias) x2, bins [:N/2+1] kept."""

    def __init__(self, ft_size=1024, hop_size=384):
        super().__init__(ft_size, hop_size)
        self.conv_analysis_real = nn.Conv1d(1, self.sz, self.sz, padding=self.sz, stride=self.hop, bias=False)
        self.conv_analysis_imag = nn.Conv1d(1, self.sz, self.sz, padding=self.sz, stride=self.hop, bias=False)
        self.initialize()

    def initialize(self):
        self._load(self.conv_analysis_real, self.conv_analysis_imag, hamming_window(self.sz))

    def forward(self, wave_form):
        """(B, C) -> an_real, an_imag each (B, T, N/2+1).  Stand-alone use goes through the same CUDA
        analysis contraction as the model (signaltrain_b200.frontend_ops)."""
        from . import frontend_ops
        return frontend_ops.analysis_forward(self, wave_form)


class Synthesis(_FrontEnd):
    """Hermitian mirror + ConvTranspose1d(N -> 1, kernel N, stride hop) x2 summed, N samples trimmed each side."""

    def __init__(self, ft_size=1024, hop_size=384):
        super().__init__(ft_size, hop_size)
        self.conv_synthesis_real = nn.ConvTranspose1d(self.sz, 1, self.sz, padding=0, stride=self.hop, bias=False)
        self.conv_synthesis_imag = nn.ConvTranspose1d(self.sz, 1, self.sz, padding=0, stride=self.hop, bias=False)
        self.initialize()

    def initialize(self):
        self._load(self.conv_synthesis_real, self.conv_synthesis_imag, Synthesis.GLA(self.sz, self.hop, self.sz))

    def forward(self, real, imag):
        from . import frontend_ops
        return frontend_ops.synthesis_forward(self, real, imag)

    @staticmethod
    def GLA(wsz, hop, N=4096):
        return lsee_synthesis_window(wsz, hop)
