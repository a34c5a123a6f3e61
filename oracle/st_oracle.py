"""CPU oracle for the SignalTrain train-step hot path.   *** TEST INFRASTRUCTURE, NOT PRODUCT ***

A plain-numpy restatement of the reference's algorithm for the path named in BASELINE.json
(trainable-STFT front-end -> magnitude/phase autoencoders -> synthesis -> log-cosh/L1 loss ->
backward -> L1 grad clip -> Adam).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under signaltrain_b200/
does.  Every function cites the reference file:line it restates (paths relative to the
reference tree, commit 7d93cb4).

Parity status: PINNED.  The reference ships no tests or golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of the unmodified reference executed in the
build container by tests/golden/make_goldens.py (forward activations, loss, all 40 gradients,
clipped gradients, parameters after 1 and 3 Adam steps; four cases) and checked in
tests/test_oracle_vs_golden.py.

The reference computes in float32 through PyTorch (conv1d / addmm / autograd); this oracle takes a
`dtype` argument: float64 gives the mathematically exact answer the CUDA path is compared against
(tolerances in the tests), float32 is used for the timed CPU baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

AE_LAYERS = ("fnn_enc", "fnn_enc2", "fnn_enc3", "fnn_enc4", "fnn_addknobs",
             "fnn_dec4", "fnn_dec3", "fnn_dec2", "fnn_dec")
DFT_KEYS = (
    "mpaec.dft_analysis.conv_analysis_real.weight",
    "mpaec.dft_analysis.conv_analysis_imag.weight",
    "mpaec.dft_synthesis.conv_synthesis_real.weight",
    "mpaec.dft_synthesis.conv_synthesis_imag.weight",
)


# --------------------------------------------------------------------------------------------
# geometry                                                                  nn_proc.py:348-385
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Dims:
    C: int        # input chunk (samples)
    N: int        # DFT size / taps
    H: int        # hop
    F: int        # kept bins = N/2+1
    T: int        # analysis frames
    OT: int       # output frames
    L: int        # output samples
    K: int        # knobs
    R: int = 64   # AE rank                                           nn_proc.py:279


def model_dims(scale_factor=1, shrink_factor=4, num_knobs=3, scale_scheme="lean") -> Dims:
    """nn_proc.py:357-384."""
    C = int(8192 * scale_factor)
    out_chunk = int(C / shrink_factor)
    N, H = 1024, 384
    if scale_scheme != "lean":
        N, H = int(N * scale_factor), int(H * scale_factor)
    T = int(np.ceil(C / float(H)) + np.ceil(N / float(H)))
    OT = int(np.ceil(out_chunk / float(H)) + np.ceil(N / float(H)))
    L = (OT - 1) * H - N
    return Dims(C=C, N=N, H=H, F=N // 2 + 1, T=T, OT=OT, L=L, K=num_knobs)


def ae_layer_shapes(d: Dims):
    """(out, in) of the nine Linear layers, in layer_list order.  nn_proc.py:45-60."""
    R, r2, r4 = d.R, d.R // 2, d.R // 4
    return [(R, d.T), (r2, R), (r4, r2), (r4, r4), (r4, r4 + d.K), (r4, r4), (r2, r4), (R, r2), (d.OT, R)]


def param_order(d: Dims):
    """state_dict key order and shapes (40 tensors).  SURVEY.md section 8(b)."""
    out = [(DFT_KEYS[0], (d.N, 1, d.N)), (DFT_KEYS[1], (d.N, 1, d.N)),
           (DFT_KEYS[2], (d.N, 1, d.N)), (DFT_KEYS[3], (d.N, 1, d.N))]
    for ae in ("aenc", "phs_aenc"):
        for name, (o, i) in zip(AE_LAYERS, ae_layer_shapes(d)):
            out.append((f"mpaec.{ae}.{name}.weight", (o, i)))
            out.append((f"mpaec.{ae}.{name}.bias", (o,)))
    return out


# --------------------------------------------------------------------------------------------
# front-end initialisation                                       cls_fe_dft.py:36-48, 87-100
# --------------------------------------------------------------------------------------------
def hamming_sym(n: int) -> np.ndarray:
    """scipy.signal.hamming(n) (symmetric).  Used at cls_fe_dft.py:38,148."""
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))


def gla_window(wsz: int, hop: int) -> np.ndarray:
    """Griffin-Lim LSEE synthesis window.  cls_fe_dft.py:133-163."""
    w = hamming_sym(wsz)
    w2 = w ** 2
    env = np.zeros(wsz)
    red = wsz // hop
    for k in range(-red, red + 1):
        lo, hi = max(0, hop * k), min(wsz, wsz + hop * k)     # env index range hit by shift k
        if hi > lo:
            env[lo:hi] += w2[lo - hop * k: hi - hop * k]
    return w / env


def dft_init(N: int, H: int):
    """Initial analysis (Wr, Wi) and synthesis (Sr, Si) matrices, each (N, N) float32.
    cls_fe_dft.py:36-41 (ortho DFT x Hamming) and :87-92 (ortho DFT x GLA window)."""
    f = np.fft.fft(np.eye(N), norm="ortho")
    wa = hamming_sym(N)
    ws = gla_window(N, H)
    return ((f.real * wa).astype(np.float32), (f.imag * wa).astype(np.float32),
            (f.real * ws).astype(np.float32), (f.imag * ws).astype(np.float32))


def xavier_normal(rng: np.random.RandomState, shape):
    """torch.nn.init.xavier_normal_ statistics (nn_proc.py:71-75); own RNG stream, not torch's."""
    o, i = shape
    return (rng.standard_normal(shape) * math.sqrt(2.0 / (i + o))).astype(np.float32)


def init_params(d: Dims, seed=218):
    p = {}
    Wr, Wi, Sr, Si = dft_init(d.N, d.H)
    for k, w in zip(DFT_KEYS, (Wr, Wi, Sr, Si)):
        p[k] = w.reshape(d.N, 1, d.N).copy()
    rng = np.random.RandomState(seed)
    for name, shape in param_order(d)[4:]:
        p[name] = xavier_normal(rng, shape) if name.endswith("weight") else np.zeros(shape, np.float32)
    return p


# --------------------------------------------------------------------------------------------
# forward                                                                   nn_proc.py:305-340
# --------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------
# reduced-precision emulation (checker for st_set_precision(ST_PRECISION_TF32))
# --------------------------------------------------------------------------------------------
_OPERAND_ROUNDING = None


class operand_rounding:
    """Context manager: inside it every contraction of the path (the five DFT contractions and the autoencoder layers,
    forward and backward) sees its two operands rounded to TF32 (10 explicit mantissa bits, round-to-nearest ties away,
    as `cvt.rna.tf32.f32`), products and sums exact in the working dtype.  The reference's counterpart is apex mixed
    precision (train.py:133-136); there is no reference run of it here (apex absent), so this is a model of the CUDA
    path's arithmetic used to bound its error, not a pin."""

    def __init__(self, mode="tf32"):
        assert mode in (None, "tf32")
        self.mode = mode

    def __enter__(self):
        global _OPERAND_ROUNDING
        self.prev, _OPERAND_ROUNDING = _OPERAND_ROUNDING, self.mode
        return self

    def __exit__(self, *exc):
        global _OPERAND_ROUNDING
        _OPERAND_ROUNDING = self.prev
        return False


def round_tf32(a):
    """fp32 -> tf32 (rna): add half an ulp of the 13 dropped bits to the magnitude, clear them."""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    bits = a32.view(np.uint32)
    out = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return out.astype(np.asarray(a).dtype if np.asarray(a).dtype in (np.float32, np.float64) else np.float32)


def _q(a):
    return a if _OPERAND_ROUNDING is None else round_tf32(a)


def elu(z):
    """nn.ELU(alpha=1), nn_proc.py:63."""
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))


def _frames(xpad, T, N, H):
    B = xpad.shape[0]
    s0, s1 = xpad.strides
    return np.lib.stride_tricks.as_strided(xpad, shape=(B, T, N), strides=(s0, H * s1, s1), writeable=False)


def analysis_forward(d: Dims, Wr, Wi, x):
    """cls_fe_dft.py:50-58: Conv1d(1->N, k=N, stride=H, padding=N, no bias), keep bins [:F]."""
    B = x.shape[0]
    xpad = np.zeros((B, d.C + 2 * d.N), dtype=x.dtype)
    xpad[:, d.N:d.N + d.C] = x
    Tc = (d.C + d.N) // d.H + 1
    assert Tc == d.T, f"conv output frames {Tc} != expected_time_frames {d.T}"
    fr = _frames(xpad, d.T, d.N, d.H)
    re = _q(fr) @ _q(Wr[:d.F]).T
    im = _q(fr) @ _q(Wi[:d.F]).T
    return re, im, fr


def ae_forward(d: Dims, P, prefix, v_btf, knobs, mode):
    """AsymAutoEncoder.forward, nn_proc.py:77-126.  v_btf (B,T,F).  Returns (out (B,OT,F), cache)."""
    v = np.transpose(v_btf, (0, 2, 1))                       # (B,F,T)            :79
    acts, pre = [v], []
    h = v
    for li, name in enumerate(AE_LAYERS):
        W, b = P[f"{prefix}.{name}.weight"], P[f"{prefix}.{name}.bias"]
        if li == 4:                                           # knob concat        :95-96
            kr = np.broadcast_to(knobs[:, None, :], (h.shape[0], h.shape[1], knobs.shape[1]))
            h = np.concatenate([h, kr], axis=2)
            acts[-1] = h
        z = _q(h) @ _q(W).T + b
        pre.append(z)
        h = elu(z)
        acts.append(h)
    tail = v[:, :, -d.OT:]
    if mode == "sf":                                          # skip-filter        :114-115
        out = h * tail
    elif mode == "res":
        raise NotImplementedError("'res' is not on the live path (nn_proc.py:315-316)")
    else:
        out = h                                               #                    :117
    return np.transpose(out, (0, 2, 1)), dict(acts=acts, pre=pre, tail=tail, mode=mode)


def synthesis_forward(d: Dims, Sr, Si, an_re, an_im):
    """cls_fe_dft.py:102-115: Hermitian mirror to N channels, two ConvTranspose1d summed, trim N."""
    Rf = np.concatenate([an_re, an_re[:, :, -2:0:-1]], axis=2)        # (B,OT,N)   :109
    If = np.concatenate([an_im, -an_im[:, :, -2:0:-1]], axis=2)       #            :110
    fo = _q(Rf) @ _q(Sr) + _q(If) @ _q(Si)                            # (B,OT,N) per-frame output
    B = an_re.shape[0]
    wave = np.zeros((B, (d.OT - 1) * d.H + d.N), dtype=fo.dtype)
    for t in range(d.OT):
        wave[:, t * d.H: t * d.H + d.N] += fo[:, t]
    return wave[:, d.N:-d.N], Rf, If                                  #            :113


def forward(d: Dims, P, x, knobs, dtype=np.float64, keep=True):
    """AsymMPAEC.forward, nn_proc.py:305-340.  Returns dict with outputs + everything backward needs."""
    P = {k: v.astype(dtype) for k, v in P.items()}
    x = x.astype(dtype)
    knobs = knobs.astype(dtype)
    Wr, Wi, Sr, Si = (P[k].reshape(d.N, d.N) for k in DFT_KEYS)
    xh = x / 2                                                        #            :307
    re, im, fr = analysis_forward(d, Wr, Wi, xh)
    mag = np.sqrt(re * re + im * im)                                  #            :309
    phs = np.arctan2(im, re + dtype(1e-7))                            #            :310
    mag_hat, mc = ae_forward(d, P, "mpaec.aenc", mag, knobs, "sf")    #            :315
    phs_o, pc = ae_forward(d, P, "mpaec.phs_aenc", phs, knobs, "")    #            :316
    phs_hat = phs_o + phs[:, -d.OT:, :]                               #            :322
    an_re = mag_hat * np.cos(phs_hat)                                 #            :325
    an_im = mag_hat * np.sin(phs_hat)                                 #            :326
    xs, Rf, If = synthesis_forward(d, Sr, Si, an_re, an_im)           #            :329
    y_half = xs + xh[:, -d.L:]                                        #            :332
    r = dict(y_hat=2 * y_half, mag=mag, mag_hat=mag_hat)              #            :340
    if keep:
        r.update(P=P, x=x, knobs=knobs, re=re, im=im, fr=fr, phs=phs, phs_hat=phs_hat, an_re=an_re,
                 an_im=an_im, x_fwdsyn=xs, y_half=y_half, Rf=Rf, If=If, mc=mc, pc=pc)
    return r


# --------------------------------------------------------------------------------------------
# loss                                                         loss_functions.py:9-10,22-43
# --------------------------------------------------------------------------------------------
# ---- data step in front of the path: the comp_4c target generator and the window cropper ------------
def compressor_4controls(x, thresh=-24.0, ratio=2.0, attackTime=0.01, releaseTime=0.01, sr=44100.0):
    """audio.py:380-426, batched over the leading axis, with the dtypes the reference's numba-compiled function produces for a
    float32 x: level detection in float64 (x_uni + 1e-8 promotes, :401-402), the static gain change and its smoothed copy
    stored in float32 arrays (:396, :409) but updated with float64 arithmetic (:414-418), 10^(lin_A/20) * x in float64
    (:420-422).  x (B, n) float32; scalars or (B,) arrays for the knobs.  Returns float64 (B, n)."""
    x = np.asarray(x, np.float32)
    B, N = x.shape
    col = lambda v: np.broadcast_to(np.asarray(v, np.float64), (B,)).reshape(B, 1)
    thresh, ratio = col(thresh), col(ratio)
    aA = np.exp(-np.log(9) / (sr * col(attackTime)))[:, 0]
    aR = np.exp(-np.log(9) / (sr * col(releaseTime)))[:, 0]
    xd = x.astype(np.float64)
    x_dB = np.maximum(20 * np.log10(np.abs(xd) + 1e-8), -96)
    g = np.where(x_dB > thresh, thresh + (x_dB - thresh) / ratio - x_dB, 0.0).astype(np.float32)
    lin = np.zeros((B, N), np.float32)
    for n in range(1, N):
        prev, gn = lin[:, n - 1], g[:, n]
        a = np.where(gn < prev, aA, aR)
        lin[:, n] = ((1 - a) * gn.astype(np.float64) + a * prev.astype(np.float64)).astype(np.float32)
    return np.power(10.0, lin.astype(np.float64) / 20) * xd


def crop_windows(corpus_x, corpus_y, offsets, signs, chunk, y_size):
    """datasets.py:236-241 (random chunk of a preloaded pair, last y_size samples of the target) and :21-30 (do_augment's
    polarity flip), for given start offsets and signs."""
    x = np.stack([corpus_x[o:o + chunk] for o in offsets]) * np.asarray(signs)[:, None]
    y = np.stack([corpus_y[o + chunk - y_size:o + chunk] for o in offsets]) * np.asarray(signs)[:, None]
    return x.astype(np.float32), y.astype(np.float32)


# ---- DCT / MDCT front-end variant (signaltrain/cls_fe_dct_bases.py; shipped by the reference, wired to nothing) ----------
def dct_core_modulation(freq_subbands, window_size):
    """cls_fe_dct_bases.py:57-97 ('scott' method :85-90): w[n] cos(pi/M (k+1/2)(n+1/2+M/2)) sqrt(2/M), w = scipy.signal.cosine
    (= sin(pi (n + 1/2) / window_size)), cast to float32."""
    n = np.arange(window_size)
    w = np.sin(np.pi * (n + 0.5) / window_size)
    kvec = np.arange(0, freq_subbands) + 0.5
    nvec = n + 0.5 + freq_subbands / 2
    return (w * np.cos(np.pi / freq_subbands * kvec[np.newaxis].T * nvec) * np.sqrt(2. / freq_subbands)).astype(np.float32)


def dct_analysis_forward(x, W, bias, hop, dtype=np.float64):
    """Analysis.forward, cls_fe_dct_bases.py:128-135: Conv1d(1 -> sz, kernel wsz, stride hop, padding sz, bias) (:116-117),
    transposed to (B, frames, sz).  W (sz, wsz), bias (sz)."""
    W = np.asarray(W, dtype).reshape(W.shape[0], -1)
    sz, wsz = W.shape
    x = np.asarray(x, dtype)
    B, C = x.shape
    xp = np.zeros((B, C + 2 * sz), dtype)
    xp[:, sz:sz + C] = x
    nf = (C + 2 * sz - wsz) // hop + 1
    frames = np.stack([xp[:, t * hop:t * hop + wsz] for t in range(nf)], axis=1)          # (B, nf, wsz)
    return frames @ W.T + np.asarray(bias, dtype)


def dct_synthesis_forward(x_ft, W, hop, dtype=np.float64):
    """Synthesis.forward, cls_fe_dct_bases.py:173-179: ConvTranspose1d(sz -> 1, kernel wsz, stride hop) (:156-157), sz samples
    trimmed on both sides.  Returns (B, 1, C) like the reference.  tied_transform (:36-54) is this with the analysis weights."""
    W = np.asarray(W, dtype).reshape(W.shape[0], -1)
    sz, wsz = W.shape
    x_ft = np.asarray(x_ft, dtype)
    B, nf, _ = x_ft.shape
    out = np.zeros((B, (nf - 1) * hop + wsz), dtype)
    fo = x_ft @ W                                                                         # (B, nf, wsz)
    for t in range(nf):
        out[:, t * hop:t * hop + wsz] += fo[:, t]
    return out[:, None, sz:out.shape[1] - sz]


def scale_by_freq(F, dtype=np.float32):
    """train.py:115-117: exp(7/F * arange(F)) (computed in float32 by the reference)."""
    return np.exp(np.float32(7.0 / F) * np.arange(F, dtype=np.float32)).astype(dtype)


def logcosh(y_hat, y):
    """loss_functions.py:9-10."""
    return np.mean(np.log(np.cosh(y - y_hat)))


def mae(x, x_hat):
    """loss_functions.py:22-23."""
    return np.mean(np.abs(x - x_hat))


def calc_loss(y_hat, y, mag_hat, sbf=None, l1_lambda=2e-5):
    """loss_functions.py:26-43, live branches :34 (sbf None) and :36."""
    if sbf is None:
        return logcosh(y_hat, y) + l1_lambda * np.abs(mag_hat).mean()
    return logcosh(y_hat, y) + l1_lambda / 10 * np.abs(mag_hat * sbf).mean()


# --------------------------------------------------------------------------------------------
# backward (what autograd does for train.py:138, written out)
# --------------------------------------------------------------------------------------------
def _ae_backward(d: Dims, P, prefix, cache, g_out_bof, grads):
    """Reverse of ae_forward.  g_out_bof (B,OT,F) = dL/d(out).  Returns dL/dv as (B,T,F)."""
    g = np.transpose(g_out_bof, (0, 2, 1))                    # (B,F,OT)
    acts, pre, tail = cache["acts"], cache["pre"], cache["tail"]
    h_last = acts[-1]
    g_tail = None
    if cache["mode"] == "sf":
        g_tail = g * h_last
        g = g * tail
    for li in range(8, -1, -1):
        name = AE_LAYERS[li]
        z = pre[li]
        gz = g * np.where(z > 0, 1.0, np.exp(np.minimum(z, 0)))          # ELU'
        hin = acts[li]
        W = P[f"{prefix}.{name}.weight"]
        grads[f"{prefix}.{name}.weight"] = np.einsum("bfo,bfi->oi", _q(gz), _q(hin), optimize=True)
        grads[f"{prefix}.{name}.bias"] = gz.sum(axis=(0, 1))
        g = _q(gz) @ _q(W)
        if li == 4:
            g = g[:, :, : d.R // 4]                            # knobs carry no gradient
    if g_tail is not None:
        g = g.copy()
        g[:, :, -d.OT:] += g_tail
    return np.transpose(g, (0, 2, 1))


def backward(d: Dims, fw, g_y_hat, g_mag_hat, g_mag=None):
    """Gradients of all 40 parameters given dL/d(y_hat), dL/d(mag_hat) [, dL/d(mag)]."""
    P = fw["P"]
    Wr, Wi, Sr, Si = (P[k].reshape(d.N, d.N) for k in DFT_KEYS)
    B = g_y_hat.shape[0]
    grads = {}
    # y_hat = 2*(x_fwdsyn + x/2): into the trimmed overlap-added wave
    g_wave = np.zeros((B, (d.OT - 1) * d.H + d.N), dtype=g_y_hat.dtype)
    g_wave[:, d.N:-d.N] = 2 * g_y_hat
    g_fo = _frames(g_wave, d.OT, d.N, d.H)                   # (B,OT,N) gather = adjoint of overlap-add
    Rf, If = fw["Rf"], fw["If"]
    grads[DFT_KEYS[2]] = np.einsum("btk,btn->kn", _q(Rf), _q(g_fo), optimize=True).reshape(d.N, 1, d.N)
    grads[DFT_KEYS[3]] = np.einsum("btk,btn->kn", _q(If), _q(g_fo), optimize=True).reshape(d.N, 1, d.N)
    gRf = _q(g_fo) @ _q(Sr).T
    gIf = _q(g_fo) @ _q(Si).T
    F = d.F
    g_re_o = gRf[:, :, :F].copy()
    g_im_o = gIf[:, :, :F].copy()
    g_re_o[:, :, 1:F - 1] += gRf[:, :, :F - 1:-1]            # adjoint of the mirror (cat + flip)
    g_im_o[:, :, 1:F - 1] -= gIf[:, :, :F - 1:-1]
    mag_hat, phs_hat = fw["mag_hat"], fw["phs_hat"]
    c, s = np.cos(phs_hat), np.sin(phs_hat)
    gm = g_re_o * c + g_im_o * s + g_mag_hat
    gp = mag_hat * (g_im_o * c - g_re_o * s)
    g_mag_in = _ae_backward(d, P, "mpaec.aenc", fw["mc"], gm, grads)
    g_phs_in = _ae_backward(d, P, "mpaec.phs_aenc", fw["pc"], gp, grads)
    g_phs_in = g_phs_in.copy()
    g_phs_in[:, -d.OT:, :] += gp                             # phase residual, nn_proc.py:322
    if g_mag is not None:
        g_mag_in = g_mag_in + g_mag
    re, im, mag = fw["re"], fw["im"], fw["mag"]
    inv = np.where(mag > 0, 1.0 / np.where(mag > 0, mag, 1.0), 0.0)      # norm subgradient 0 at 0
    u = re + re.dtype.type(1e-7)
    den = u * u + im * im
    den = np.where(den > 0, den, 1.0)
    g_re = g_mag_in * re * inv - g_phs_in * im / den
    g_im = g_mag_in * im * inv + g_phs_in * u / den
    fr = fw["fr"]
    gWr = np.zeros((d.N, d.N), dtype=g_re.dtype)
    gWi = np.zeros((d.N, d.N), dtype=g_re.dtype)
    gWr[:F] = np.einsum("btk,btn->kn", _q(g_re), _q(fr), optimize=True)  # rows >= F: sliced off, zero
    gWi[:F] = np.einsum("btk,btn->kn", _q(g_im), _q(fr), optimize=True)
    grads[DFT_KEYS[0]] = gWr.reshape(d.N, 1, d.N)
    grads[DFT_KEYS[1]] = gWi.reshape(d.N, 1, d.N)
    return grads


def loss_and_grads(d: Dims, P, x, y, knobs, sbf, dtype=np.float64, l1_lambda=2e-5):
    """forward + calc_loss + backward; returns (loss, grads, fw)."""
    fw = forward(d, P, x, knobs, dtype=dtype)
    y = y.astype(dtype)
    y_hat, mag_hat = fw["y_hat"], fw["mag_hat"]
    sbf = None if sbf is None else sbf.astype(dtype)
    loss = calc_loss(y_hat, y, mag_hat, sbf, l1_lambda)
    g_y = -np.tanh(y - y_hat) / y_hat.size
    if sbf is None:
        g_m = l1_lambda * np.sign(mag_hat) / mag_hat.size
    else:
        g_m = (l1_lambda / 10) * np.sign(mag_hat * sbf) * sbf / mag_hat.size
    grads = backward(d, fw, g_y, g_m)
    return loss, grads, fw


# --------------------------------------------------------------------------------------------
# clip + Adam + schedule
# --------------------------------------------------------------------------------------------
def clip_grad_norm_(grads, max_norm=1.0):
    """nn_proc.py:299-302 -> torch.nn.utils.clip_grad_norm_(4 DFT tensors, max_norm=1, norm_type=1).
    Returns the total L1 norm; scales the four DFT gradients in place."""
    total = sum(float(np.abs(grads[k]).sum()) for k in DFT_KEYS)
    coef = min(1.0, max_norm / (total + 1e-6))
    for k in DFT_KEYS:
        grads[k] = grads[k] * grads[k].dtype.type(coef)
    return total


def adam_step(P, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (train.py:228,147), weight_decay 0, amsgrad off.  state: dict with 't', 'm', 'v'."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    m, v = state.setdefault("m", {}), state.setdefault("v", {})
    for k, p in P.items():
        g = grads[k].astype(p.dtype)
        mk = m.get(k, np.zeros_like(p))
        vk = v.get(k, np.zeros_like(p))
        mk = mk + (g - mk) * p.dtype.type(1 - beta1)
        vk = vk * p.dtype.type(beta2) + g * g * p.dtype.type(1 - beta2)
        m[k], v[k] = mk, vk
        denom = np.sqrt(vk) / p.dtype.type(math.sqrt(bc2)) + p.dtype.type(eps)
        P[k] = p - p.dtype.type(lr / bc1) * (mk / denom)
    return P


def get_1cycle_schedule(lr_max=1e-3, n_data_points=8000, epochs=200, batch_size=40):
    """learningrate.py:14-52 (lr LUT and the momentum LUT Adam ignores)."""
    lr_start = lr_max / 15.0
    lr_end = lr_start / 1e2
    n_iter = n_data_points * epochs // batch_size
    a1 = int(n_iter * 0.3)
    a2 = n_iter - a1
    up = (lr_max - lr_start) * (1 - np.cos(np.linspace(0, np.pi, a1))) / 2 + lr_start
    down = (lr_max - lr_end) * (1 + np.cos(np.linspace(0, np.pi, a2))) / 2 + lr_end
    mom_up = 0.9 + 0.05 * np.cos(np.linspace(0, np.pi, a1))
    mom_down = 0.9 - 0.05 * np.cos(np.linspace(0, np.pi, a2))
    return np.concatenate((up, down)), np.concatenate((mom_up, mom_down))


class Trainer:
    """The loop body of train.py:104-151 (lr lag included: step i runs Adam with the lr installed
    at the end of step i-1, initial lr = lr_sched[0])."""

    def __init__(self, d: Dims, P, lr_sched, dtype=np.float32):
        self.d, self.dtype = d, dtype
        self.P = {k: v.astype(dtype) for k, v in P.items()}
        self.lr_sched = lr_sched
        self.lr = float(lr_sched[0])
        self.iter = 0
        self.state = {}
        self.sbf = scale_by_freq(d.F)
        self.last_total_norm = None

    def step(self, x, y, knobs):
        loss, grads, fw = loss_and_grads(self.d, self.P, x, y.astype(np.float32), knobs, self.sbf, dtype=self.dtype)
        self.last_total_norm = clip_grad_norm_(grads)
        self.P = adam_step(self.P, grads, self.state, self.lr)
        self.lr = float(self.lr_sched[min(self.iter, len(self.lr_sched) - 1)])
        self.iter += 1
        return float(loss), grads, fw
