#!/usr/bin/env python3
"""Mint golden vectors by EXECUTING the unmodified reference (read-only at /root/reference).

Run here (CPU container) only:   python tests/golden/make_goldens.py
Writes tests/golden/<case>.npz.  Nothing from the reference tree is copied: the reference is
imported from where it lies, with the four environment shims of SURVEY.md §8(c) (scipy.signal.hamming
alias, torch.has_cudnn=False, stub librosa, stub matplotlib), and executed through its own public
API, following the call order of signaltrain/train.py:104-151 (forward, calc_loss, zero_grad,
backward, clip_grad_norm_, Adam.step, lr poke).

Every array that would be too large to commit in full (the four 1024x1024 DFT tensors) is stored
as (a) a fixed set of rows, (b) its L1 norm and plain sum in float64.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

REF = os.environ.get("SIGNALTRAIN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

# ---- environment shims (no edits to reference files) ------------------------------------
scipy.signal.hamming = scipy.signal.windows.hamming
scipy.signal.cosine = scipy.signal.windows.cosine
torch.has_cudnn = False
for name in ("librosa", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
    sys.modules.setdefault(name, mock.MagicMock())
sys.path.insert(0, REF)
import signaltrain as st  # noqa: E402  (the reference package)

DFT_ROWS = np.array([0, 1, 2, 3, 100, 255, 256, 511, 512, 513, 514, 700, 1022, 1023])
DFT_KEYS = [
    "mpaec.dft_analysis.conv_analysis_real.weight",
    "mpaec.dft_analysis.conv_analysis_imag.weight",
    "mpaec.dft_synthesis.conv_synthesis_real.weight",
    "mpaec.dft_synthesis.conv_synthesis_imag.weight",
]
ACT_NAMES = (["x_real", "x_imag", "mag", "phs"]
             + ["m_act%d" % i for i in range(10)] + ["p_act%d" % i for i in range(10)]
             + ["mag_hat", "phs_hat", "an_real", "an_imag", "x_fwdsyn", "y_hat_half"])


def perturbation(shape, seed, scale):
    """Deterministic pseudo-random perturbation reproducible from numpy alone (legacy RandomState)."""
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def summarize_dft(t, nrows):
    a = t.detach().cpu().numpy().reshape(t.shape[0], -1)
    rows = DFT_ROWS[DFT_ROWS < nrows]
    return a[rows].copy(), np.array([np.abs(a.astype(np.float64)).sum(), a.astype(np.float64).sum()])


def pack_state(prefix, model, out, grads=False):
    for k, p in model.named_parameters():
        t = p.grad if grads else p.data
        if k in DFT_KEYS:
            rows, sums = summarize_dft(t, t.shape[0])
            out[f"{prefix}/{k}/rows"] = rows
            out[f"{prefix}/{k}/sums"] = sums
        else:
            out[f"{prefix}/{k}"] = t.detach().cpu().numpy().copy()


def make_case(name, scale, shrink, effect, B, nsteps=3, perturb=None, data_seed=218):
    np.random.seed(data_seed)
    torch.manual_seed(218)
    K = len(effect.knob_names)
    model = st.nn_proc.st_model(scale_factor=scale, shrink_factor=shrink, num_knobs=K, sr=44100)
    if perturb is not None:   # general (non-DFT) front-end matrices
        with torch.no_grad():
            for i, k in enumerate(DFT_KEYS):
                p = dict(model.named_parameters())[k]
                p.add_(torch.from_numpy(perturbation(tuple(p.shape), perturb["seed"] + i, perturb["scale"])))
    C, L = model.in_chunk_size, model.out_chunk_size
    ds = st.datasets.SynthAudioDataSet(C, effect, sr=44100, datapoints=64, y_size=L, augment=True)
    out = {}
    out["meta/scale_factor"] = np.array(scale, dtype=np.float64)
    out["meta/shrink_factor"] = np.array(shrink, dtype=np.float64)
    out["meta/num_knobs"] = np.array(K)
    out["meta/batch"] = np.array(B)
    out["meta/in_chunk_size"] = np.array(C)
    out["meta/out_chunk_size"] = np.array(L)
    if perturb is not None:
        out["meta/perturb_seed"] = np.array(perturb["seed"])
        out["meta/perturb_scale"] = np.array(perturb["scale"], dtype=np.float64)
    # initial AE weights in full; DFT tensors summarised (they are a deterministic function of (N, hop))
    pack_state("init", model, out)

    lr_sched, mom_sched = st.learningrate.get_1cycle_schedule(lr_max=1e-4, n_data_points=200000,
                                                              epochs=1000, batch_size=200)
    out["meta/lr_sched_head"] = np.asarray(lr_sched[:8], dtype=np.float64)
    out["meta/lr_sched_len"] = np.array(len(lr_sched))
    out["meta/lr_sched_probe_idx"] = np.array([0, 1, 1000, 299999, 300000, 300001, 650000, len(lr_sched) - 1])
    out["meta/lr_sched_probe"] = np.asarray(lr_sched[out["meta/lr_sched_probe_idx"]], dtype=np.float64)
    optimizer = torch.optim.Adam(list(model.parameters()), lr=lr_sched[0], weight_decay=0)

    scale_by_freq = None
    for step in range(nsteps):
        xs, ys, ks = zip(*[ds[i] for i in range(B)])
        x = torch.from_numpy(np.stack(xs).astype(np.float32))
        y = torch.from_numpy(np.stack(ys))                      # float64, as the DataLoader delivers it
        knobs = torch.from_numpy(np.stack(ks).astype(np.float32))
        out[f"step{step}/x"] = x.numpy().copy()
        out[f"step{step}/y"] = y.numpy().astype(np.float64)
        out[f"step{step}/knobs"] = knobs.numpy().copy()
        lr = lr_sched[min(step, len(lr_sched) - 1)]

        if step == 0:
            y_hat, mag, mag_hat, acts = model.forward(x, knobs, return_acts=True)
            assert len(acts) == len(ACT_NAMES)
            for n, a in zip(ACT_NAMES, acts):
                a = a.detach().numpy()
                if n.startswith(("m_act", "p_act")):      # (B, F, width): keep b=0, every 16th bin
                    a = a[0, ::16]
                out[f"step0/acts/{n}"] = np.ascontiguousarray(a)
        else:
            y_hat, mag, mag_hat = model.forward(x, knobs)
        if scale_by_freq is None:
            expfac = 7. / mag_hat.size()[-1]
            scale_by_freq = torch.exp(expfac * torch.arange(0., mag_hat.size()[-1])).expand_as(mag_hat).float()
        loss = st.loss_functions.calc_loss(y_hat.float(), y.float(), mag_hat.float(), scale_by_freq=scale_by_freq)
        out[f"step{step}/y_hat"] = y_hat.detach().numpy().copy()
        out[f"step{step}/loss"] = np.array(loss.item(), dtype=np.float64)
        out[f"step{step}/logcosh"] = np.array(st.loss_functions.logcosh(y_hat.float(), y.float()).item())
        out[f"step{step}/mae"] = np.array(st.loss_functions.mae(y_hat.float(), y.float()).item())
        if step == 0:
            out["step0/mag"] = mag.detach().numpy().copy()
            out["step0/mag_hat"] = mag_hat.detach().numpy().copy()
        optimizer.zero_grad()
        loss.backward()
        if step == 0:
            pack_state("step0/grad", model, out, grads=True)
        model.clip_grad_norm_()
        if step == 0:
            pack_state("step0/grad_clipped", model, out, grads=True)
        optimizer.step()
        optimizer.param_groups[0]['lr'] = lr
        optimizer.param_groups[0]['momentum'] = mom_sched[min(step, len(mom_sched) - 1)]
        if step in (0, nsteps - 1):
            pack_state(f"step{step}/params_after", model, out)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


class TwoKnobComp(st.audio.Compressor_4c):
    """LA2A-style 2-knob case of BASELINE.json configs[3]: threshold+ratio free, attack/release pinned.
    Built here (golden minting only) from the reference's own compressor_4controls via Compressor_4c."""
    def __init__(self):
        super().__init__()
        self.name = "comp_2knob"
        self.knob_names = self.knob_names[:2]
        self._full_ranges = np.array(self.knob_ranges)
        self.knob_ranges = self._full_ranges[:2]

    def go(self, x, knobs_nn, **kw):
        saved = self.knob_ranges
        self.knob_ranges = self._full_ranges
        try:
            return super().go(x, np.concatenate([knobs_nn, np.zeros(2)]), **kw)
        finally:
            self.knob_ranges = saved


def make_dct_case(name, ft_size, w_size, hop, chunk, B, seed=218):
    """cls_fe_dct_bases.Analysis / Synthesis / tied_transform.  The reference imports this module nowhere and its
    Analysis.forward begins with an unconditional numpy -> .cuda() (:130), so the golden applies the REST of that forward --
    the reference's own conv_analysis layer, viewed and transposed exactly as :131-134 do -- to a CPU tensor; Synthesis.forward
    and tied_transform run unmodified.  Weights: the reference's own initialize() (bias = Conv1d's default random init under the
    seed) plus a deterministic perturbation, so the transforms are general matrices rather than the analytic cosine basis."""
    import importlib
    dct = importlib.import_module("signaltrain.cls_fe_dct_bases")
    torch.manual_seed(seed)
    ana = dct.Analysis(ft_size=ft_size, w_size=w_size, hop_size=hop)
    syn = dct.Synthesis(ft_size=ft_size, w_size=w_size, hop_size=hop)
    out = {"ft_size": ft_size, "w_size": w_size, "hop": hop, "chunk": chunk}
    out["core_rows"] = np.array([0, 1, 7, ft_size // 2, ft_size - 1])
    out["core_modulation_rows"] = dct.core_modulation(ft_size, w_size)[out["core_rows"]]
    out["bias"] = ana.conv_analysis.bias.detach().numpy().copy()
    with torch.no_grad():
        ana.conv_analysis.weight.add_(torch.from_numpy(perturbation(tuple(ana.conv_analysis.weight.shape), 777, 1e-3)))
        syn.conv_synthesis.weight.add_(torch.from_numpy(perturbation(tuple(syn.conv_synthesis.weight.shape), 778, 1e-3)))
    rng = np.random.RandomState(seed)
    t = np.arange(chunk) / 44100.0
    x = (0.5 * np.sin(2 * np.pi * rng.uniform(60, 4000, (B, 1)) * t) + 0.1 * rng.standard_normal((B, chunk))).astype(np.float32)
    out["x"] = x
    out["pert_seeds"] = np.array([777, 778])
    with torch.no_grad():
        wave = torch.from_numpy(x)
        x_ft = torch.transpose(ana.conv_analysis(wave.view(B, 1, chunk)), 2, 1)          # cls_fe_dct_bases.py:131-134
        out["x_ft"] = x_ft.numpy().copy()
        out["wave"] = syn.forward(x_ft).numpy().copy()                                    # :173-179, unmodified
        out["tied"] = dct.tied_transform(ana, x_ft, hop).numpy().copy()                   # :36-54, unmodified
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def make_compressor_case(name, B=4, n=4096, seed=218):
    """audio.compressor_4controls (audio.py:380-426, numba-compiled) on float32 windows, one knob setting per window."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / 44100.0
    x = (rng.uniform(0.2, 0.9, (B, 1)) * np.sin(2 * np.pi * rng.uniform(80, 3000, (B, 1)) * t) * np.exp(-rng.uniform(0, 8, (B, 1)) * t)
         + 0.03 * rng.standard_normal((B, n))).astype(np.float32)
    x[1, 100:400] = 0.0                                          # silence: the -96 dB floor and the 1e-8 offset
    eff = st.audio.Compressor_4c()
    knobs_nn = rng.beta(0.8, 0.8, size=(B, 4)) - 0.5
    kr = eff.knob_ranges
    knobs_wc = kr[:, 0] + (knobs_nn + 0.5) * (kr[:, 1] - kr[:, 0])
    y = np.stack([st.audio.compressor_4controls(x[b].copy(), thresh=knobs_wc[b, 0], ratio=knobs_wc[b, 1], attackTime=knobs_wc[b, 2],
                                                releaseTime=knobs_wc[b, 3], sr=44100.0) for b in range(B)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, knobs_wc=knobs_wc, y=y)
    print(name, x.shape, y.dtype, float(np.abs(y).max()))


def main():
    make_compressor_case("compressor4c_n4096_b4")
    make_dct_case("dct_ft256_w512_h256_c4096_b3", 256, 512, 256, 4096, 3)
    make_dct_case("dct_ft1024_w2048_h1024_c8192_b2", 1024, 2048, 1024, 8192, 2)
    make_case("comp4c_c8192_k4_b3", 1, 4, st.audio.Compressor_4c(), B=3)
    make_case("comp2k_c16384_k2_b2", 2, 4, TwoKnobComp(), B=2)
    make_case("denoise_c8192_k1_b2", 1, 4, st.audio.Denoise(), B=2)
    make_case("general_c8192_k4_b2", 1, 4, st.audio.Compressor_4c(), B=2,
              perturb={"seed": 4242, "scale": 2e-3})


if __name__ == "__main__":
    main()
