"""Thin host wrapper over one st_handle: torch tensors in, kernels launched on torch's current stream.
PyTorch is used for device memory and streams only; all arithmetic of the path runs in the CUDA library."""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import NUM_ACTS, NUM_PARAMS, ActTable, PtrTable, StAdam, StConfig


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check(t, name, shape=None, device=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"signaltrain_b200: {name} must be a CUDA tensor (no CPU fallback for this path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"signaltrain_b200: {name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"signaltrain_b200: {name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"signaltrain_b200: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    if device is not None and t.device != device:
        raise RuntimeError(f"signaltrain_b200: {name} is on {t.device}, engine is on {device}")


class Geometry:
    """nn_proc.py:357-384 (st_model.__init__ arithmetic)."""

    def __init__(self, scale_factor=1, shrink_factor=4, num_knobs=3, scale_scheme="lean"):
        self.C = int(8192 * scale_factor)
        out_chunk = int(self.C / shrink_factor)
        self.N, self.H = 1024, 384
        if scale_scheme != "lean":
            self.N, self.H = int(self.N * scale_factor), int(self.H * scale_factor)
        self.T = int(np.ceil(self.C / float(self.H)) + np.ceil(self.N / float(self.H)))
        self.OT = int(np.ceil(out_chunk / float(self.H)) + np.ceil(self.N / float(self.H)))
        self.L = (self.OT - 1) * self.H - self.N
        self.F = self.N // 2 + 1
        self.K = num_knobs
        self.R = 64
        self.intended_out_chunk = out_chunk

    def ae_shapes(self):
        R, r2, r4 = self.R, self.R // 2, self.R // 4
        return [(R, self.T), (r2, R), (r4, r2), (r4, r4), (r4, r4 + self.K), (r4, r4), (r2, r4), (R, r2), (self.OT, R)]

    def act_shapes(self, B):
        """Logical shapes of the reference's 30 layer_acts (nn_proc.py:311-335)."""
        T, OT, F, K = self.T, self.OT, self.F, self.K
        ae = [(B, F, 64), (B, F, 32), (B, F, 16), (B, F, 16), (B, F, 16 + K), (B, F, 16), (B, F, 16), (B, F, 32), (B, F, 64),
              (B, F, OT)]
        return [(B, T, F)] * 4 + ae + ae + [(B, OT, F)] * 4 + [(B, self.L)] * 2


class Engine:
    _registry = []          # weak references to live engines (loss_functions looks engines up by shape)

    @classmethod
    def any_on(cls, device):
        """Some live engine on `device` (for the geometry-free reductions), or None."""
        for ref in list(cls._registry):
            e = ref()
            if e is None:
                cls._registry.remove(ref)
            elif e.device == device:
                return e
        return None

    @classmethod
    def find(cls, device, L, OT=None, F=None):
        for ref in list(cls._registry):
            e = ref()
            if e is None:
                cls._registry.remove(ref)
            elif e.device == device and e.g.L == L and (OT is None or (e.g.OT, e.g.F) == (OT, F)):
                return e
        return None

    def __init__(self, geom: Geometry, device):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("signaltrain_b200: no CUDA device visible; this path has no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"signaltrain_b200: engine needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.g = geom
        cfg = StConfig(geom.C, geom.N, geom.H, geom.T, geom.OT, geom.K, geom.R)
        h = ctypes.c_void_p()
        if self.lib.st_create(ctypes.byref(cfg), self.device.index, ctypes.byref(h)) != 0:
            raise RuntimeError("signaltrain_b200: st_create failed: " + self.lib.st_last_error(None).decode())
        self.h = h
        assert self.lib.st_out_samples(h) == geom.L and self.lib.st_bins(h) == geom.F
        self.param_names = [self.lib.st_param_name(h, i).decode() for i in range(NUM_PARAMS)]
        self.param_numel = [self.lib.st_param_numel(h, i) for i in range(NUM_PARAMS)]
        self._table_cache = {}
        Engine._registry.append(weakref.ref(self))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.st_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- helpers ---------------------------------------------------------------------------
    def _ok(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"signaltrain_b200: {what} failed: " + self.lib.st_last_error(self.h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def table(self, tensors, what):
        """Host array of 40 device pointers; cached per tuple of addresses."""
        if len(tensors) != NUM_PARAMS:
            raise RuntimeError(f"signaltrain_b200: {what}: expected {NUM_PARAMS} tensors, got {len(tensors)}")
        key = tuple(t.data_ptr() for t in tensors)
        tab = self._table_cache.get(key)
        if tab is None:
            for i, t in enumerate(tensors):
                _check(t, f"{what}[{i}] ({self.param_names[i]})", device=self.device)
                if t.numel() != self.param_numel[i]:
                    raise RuntimeError(f"signaltrain_b200: {what}[{i}] ({self.param_names[i]}) has {t.numel()} elements, "
                                       f"expected {self.param_numel[i]}")
            tab = PtrTable(*key)
            if len(self._table_cache) > 64:
                self._table_cache.clear()
            self._table_cache[key] = tab
        return tab

    # ---- entry points ----------------------------------------------------------------------
    def init_frontend(self, params4):
        tab = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in params4])
        for t in params4:
            _check(t, "front-end weight", (self.g.N, 1, self.g.N), self.device)
        self._ok(self.lib.st_init_frontend(self.h, tab, self._stream()), "st_init_frontend")

    def analysis(self, x, w_real, w_imag):
        g = self.g
        B = x.shape[0]
        _check(x, "wave_form", (B, g.C), self.device)
        _check(w_real, "conv_analysis_real.weight", (g.N, 1, g.N), self.device)
        _check(w_imag, "conv_analysis_imag.weight", (g.N, 1, g.N), self.device)
        re = torch.empty((B, g.T, g.F), device=self.device, dtype=torch.float32)
        im = torch.empty_like(re)
        self._ok(self.lib.st_analysis(self.h, _ptr(x), _ptr(w_real), _ptr(w_imag), B, _ptr(re), _ptr(im), self._stream()), "st_analysis")
        return re, im

    def synthesis(self, real, imag, w_real, w_imag):
        g = self.g
        B = real.shape[0]
        _check(real, "real", (B, g.OT, g.F), self.device)
        _check(imag, "imag", (B, g.OT, g.F), self.device)
        _check(w_real, "conv_synthesis_real.weight", (g.N, 1, g.N), self.device)
        _check(w_imag, "conv_synthesis_imag.weight", (g.N, 1, g.N), self.device)
        wave = torch.empty((B, g.L), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_synthesis(self.h, _ptr(real), _ptr(imag), _ptr(w_real), _ptr(w_imag), B, _ptr(wave), self._stream()),
                 "st_synthesis")
        return wave

    def dct_analysis(self, x, weight, bias, ft_size, w_size, hop):
        """cls_fe_dct_bases.Analysis.forward: x (B, C) -> (B, frames, ft_size)."""
        B, C = x.shape
        _check(x, "wave_form", (B, C), self.device)
        _check(weight, "conv_analysis.weight", (ft_size, 1, w_size), self.device)
        _check(bias, "conv_analysis.bias", (ft_size,), self.device)
        nf = (C + 2 * ft_size - w_size) // hop + 1
        out = torch.empty((B, nf, ft_size), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_dct_analysis(self.h, _ptr(x), _ptr(weight), _ptr(bias), B, C, ft_size, w_size, hop, _ptr(out),
                                          self._stream()), "st_dct_analysis")
        return out

    def dct_synthesis(self, x_ft, weight, ft_size, w_size, hop):
        """cls_fe_dct_bases.Synthesis.forward: x_ft (B, frames, ft_size) -> (B, 1, (frames - 1) hop + w_size - 2 ft_size)."""
        B, nf, _ = x_ft.shape
        _check(x_ft, "x_ft", (B, nf, ft_size), self.device)
        _check(weight, "conv_synthesis.weight", (ft_size, 1, w_size), self.device)
        C = (nf - 1) * hop + w_size - 2 * ft_size
        wave = torch.empty((B, 1, max(C, 1)), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_dct_synthesis(self.h, _ptr(x_ft), _ptr(weight), B, nf, ft_size, w_size, hop, _ptr(wave), self._stream()),
                 "st_dct_synthesis")
        return wave

    PRECISIONS = {"fp32": 0, "tf32": 1}

    def set_precision(self, mode):
        """"fp32" (default): 3xTF32 GEMMs + exact-fp32 autoencoder chains, waveforms within 1e-5 of the reference.
        "tf32": one tensor-core TF32 multiply per product (fp32 storage / accumulation / loss / optimiser) -- the
        mixed-precision class the reference reaches through apex (train.py:133-136,169,184)."""
        if mode not in self.PRECISIONS:
            raise ValueError(f"signaltrain_b200: precision must be one of {sorted(self.PRECISIONS)}, got {mode!r}")
        self._ok(self.lib.st_set_precision(self.h, self.PRECISIONS[mode]), "st_set_precision")

    @property
    def precision(self):
        return {v: k for k, v in self.PRECISIONS.items()}[self.lib.st_get_precision(self.h)]

    def fallback_count(self):
        """Calls served by a SIMT fallback kernel since the handle was created (0 on the benchmarked path)."""
        return int(self.lib.st_debug_fallbacks(self.h))

    def set_training(self, on):
        self._ok(self.lib.st_set_training(self.h, int(bool(on))), "st_set_training")

    def forward(self, x, knobs, params, return_acts=False):
        g = self.g
        if x.dim() != 2 or x.shape[1] != g.C:
            raise RuntimeError(f"signaltrain_b200: x must be (B, {g.C}), got {tuple(x.shape)}")
        B = x.shape[0]
        _check(x, "x", (B, g.C), self.device)
        _check(knobs, "knobs", (B, g.K), self.device)
        y_hat = torch.empty((B, g.L), device=self.device, dtype=torch.float32)
        mag = torch.empty((B, g.T, g.F), device=self.device, dtype=torch.float32)
        mag_hat = torch.empty((B, g.OT, g.F), device=self.device, dtype=torch.float32)
        acts, acts_tab = None, None
        if return_acts:
            acts = [torch.zeros(s, device=self.device, dtype=torch.float32) for s in g.act_shapes(B)]
            acts_tab = ActTable(*[a.data_ptr() for a in acts])
        self._ok(self.lib.st_forward(self.h, _ptr(x), _ptr(knobs), B, self.table(params, "params"), _ptr(y_hat), _ptr(mag),
                                     _ptr(mag_hat), acts_tab, self._stream()), "st_forward")
        return y_hat, mag, mag_hat, acts

    def loss(self, y_hat, y, mag_hat, sbf, l1_coef, want_grads=True):
        B = y_hat.shape[0]
        g = self.g
        _check(y_hat, "y_hat", (B, g.L), self.device)
        _check(y, "y", (B, g.L), self.device)
        _check(mag_hat, "mag_hat", (B, g.OT, g.F), self.device)
        if sbf is not None:
            _check(sbf, "scale_by_freq", (g.F,), self.device)
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        g_y = torch.empty_like(y_hat) if want_grads else None
        g_m = torch.empty_like(mag_hat) if want_grads else None
        self._ok(self.lib.st_loss(self.h, _ptr(y_hat), _ptr(y), _ptr(mag_hat), _ptr(sbf), float(l1_coef), B, _ptr(loss),
                                  _ptr(g_y), _ptr(g_m), self._stream()), "st_loss")
        return loss, g_y, g_m

    def loss_shaped(self, y_hat, y, mag_hat, sbf, l1_coef, want_grads=True):
        """calc_loss for tensors of any (B, n_wave) / (B, n_frames, n_bins) shape (utils/lr_finder.py:38 passes the input
        magnitude as mag_hat): only the sizes are read off the tensors, the handle lends scratch and device."""
        B, n_wave = y_hat.shape
        _check(y_hat, "y_hat", (B, n_wave), self.device)
        _check(y, "y", (B, n_wave), self.device)
        if mag_hat.dim() != 3 or mag_hat.shape[0] != B:
            raise RuntimeError(f"signaltrain_b200: mag_hat must be (B, frames, bins), got {tuple(mag_hat.shape)}")
        _check(mag_hat, "mag_hat", device=self.device)
        n_frames, n_bins = int(mag_hat.shape[1]), int(mag_hat.shape[2])
        if sbf is not None:
            _check(sbf, "scale_by_freq", (n_bins,), self.device)
        loss = torch.empty((), device=self.device, dtype=torch.float32)
        g_y = torch.empty_like(y_hat) if want_grads else None
        g_m = torch.empty_like(mag_hat) if want_grads else None
        self._ok(self.lib.st_loss_shaped(self.h, _ptr(y_hat), _ptr(y), _ptr(mag_hat), _ptr(sbf), float(l1_coef), B, int(n_wave),
                                         n_frames, n_bins, _ptr(loss), _ptr(g_y), _ptr(g_m), self._stream()), "st_loss_shaped")
        return loss, g_y, g_m

    def mae(self, a, b):
        _check(a, "a", device=self.device)
        _check(b, "b", a.shape, self.device)
        out = torch.empty((), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_mae(self.h, _ptr(a), _ptr(b), a.numel(), _ptr(out), self._stream()), "st_mae")
        return out

    def backward(self, g_y_hat, g_mag, g_mag_hat, params, grads, part=None):
        """part: None = whole backward; "begin" = until the synthesis gradients (grads[2], grads[3]) are final; "finish" = rest."""
        B = g_y_hat.shape[0]
        g = self.g
        _check(g_y_hat, "g_y_hat", (B, g.L), self.device)
        if g_mag is not None:
            _check(g_mag, "g_mag", (B, g.T, g.F), self.device)
        if g_mag_hat is not None:
            _check(g_mag_hat, "g_mag_hat", (B, g.OT, g.F), self.device)
        fn = {None: self.lib.st_backward, "begin": self.lib.st_backward_begin, "finish": self.lib.st_backward_finish}[part]
        self._ok(fn(self.h, _ptr(g_y_hat), _ptr(g_mag), _ptr(g_mag_hat), B, self.table(params, "params"),
                    self.table(grads, "grads"), self._stream()), "st_backward" + ("_" + part if part else ""))

    def clip_grad_norm(self, grads4, max_norm=1.0):
        for t in grads4:
            _check(t, "DFT gradient", device=self.device)
        tab = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in grads4])
        total = torch.empty((), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_clip_grad_norm(self.h, tab, float(max_norm), _ptr(total), self._stream()), "st_clip_grad_norm")
        return total

    @staticmethod
    def adam_hp(lr, step, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, max_norm=0.0):
        return StAdam(float(lr), float(betas[0]), float(betas[1]), float(eps), int(step), float(grad_scale), float(max_norm))

    def adam_step(self, params, grads, exp_avg, exp_avg_sq, hp):
        self._ok(self.lib.st_adam_step(self.h, self.table(params, "params"), self.table(grads, "grads"),
                                       self.table(exp_avg, "exp_avg"), self.table(exp_avg_sq, "exp_avg_sq"),
                                       ctypes.byref(hp), self._stream()), "st_adam_step")

    def train_step(self, x, y, knobs, params, grads, exp_avg, exp_avg_sq, sbf, l1_coef, hp, loss_out=None):
        g = self.g
        B = x.shape[0]
        _check(x, "x", (B, g.C), self.device)
        _check(y, "y", (B, g.L), self.device)
        _check(knobs, "knobs", (B, g.K), self.device)
        if sbf is not None:
            _check(sbf, "scale_by_freq", (g.F,), self.device)
        loss = loss_out if loss_out is not None else torch.empty((), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_train_step(self.h, _ptr(x), _ptr(y), _ptr(knobs), B, self.table(params, "params"),
                                        self.table(grads, "grads"), self.table(exp_avg, "exp_avg"),
                                        self.table(exp_avg_sq, "exp_avg_sq"), _ptr(sbf), float(l1_coef), ctypes.byref(hp),
                                        _ptr(loss), self._stream()), "st_train_step")
        return loss

    def grad_step(self, x, y, knobs, params, grads, sbf, l1_coef, loss_out=None):
        """forward + loss + backward in one call (st_grad_step): gradients into `grads`, no update -- the data-parallel step."""
        g = self.g
        B = x.shape[0]
        _check(x, "x", (B, g.C), self.device)
        _check(y, "y", (B, g.L), self.device)
        _check(knobs, "knobs", (B, g.K), self.device)
        if sbf is not None:
            _check(sbf, "scale_by_freq", (g.F,), self.device)
        loss = loss_out if loss_out is not None else torch.empty((), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_grad_step(self.h, _ptr(x), _ptr(y), _ptr(knobs), B, self.table(params, "params"),
                                       self.table(grads, "grads"), _ptr(sbf), float(l1_coef), _ptr(loss), self._stream()),
                 "st_grad_step")
        return loss

    def packed_grad_floats(self):
        return int(self.lib.st_packed_grad_floats(self.h))

    def pack_grads(self, grads, packed):
        """The data-parallel exchange payload (8.5 of 16.8 MB: live analysis rows, Hermitian half of the synthesis pair, the
        autoencoders) gathered into one contiguous buffer -- one collective reduces it."""
        _check(packed, "packed", (self.packed_grad_floats(),), self.device)
        self._ok(self.lib.st_pack_grads(self.h, self.table(grads, "grads"), _ptr(packed), self._stream()), "st_pack_grads")

    def unpack_grads(self, packed, grads):
        _check(packed, "packed", (self.packed_grad_floats(),), self.device)
        self._ok(self.lib.st_unpack_grads(self.h, _ptr(packed), self.table(grads, "grads"), self._stream()), "st_unpack_grads")

    def grad_step_packed(self, x, y, knobs, params, packed, sbf, l1_coef, loss_out=None):
        """st_grad_step whose gradients leave directly as the packed data-parallel payload (st_grad_step_packed)."""
        g = self.g
        B = x.shape[0]
        _check(x, "x", (B, g.C), self.device)
        _check(y, "y", (B, g.L), self.device)
        _check(knobs, "knobs", (B, g.K), self.device)
        _check(packed, "packed", (self.packed_grad_floats(),), self.device)
        if sbf is not None:
            _check(sbf, "scale_by_freq", (g.F,), self.device)
        loss = loss_out if loss_out is not None else torch.empty((), device=self.device, dtype=torch.float32)
        self._ok(self.lib.st_grad_step_packed(self.h, _ptr(x), _ptr(y), _ptr(knobs), B, self.table(params, "params"), _ptr(packed),
                                              _ptr(sbf), float(l1_coef), _ptr(loss), self._stream()), "st_grad_step_packed")
        return loss

    def unpack_clip(self, packed, grads, grad_scale, max_norm, total_norm=None):
        """Scatter the reduced payload back into the 40 gradient tensors and leave the L1 clip coefficient of the four DFT
        tensors (times grad_scale) in the engine for adam_step_clipped (st_unpack_clip)."""
        _check(packed, "packed", (self.packed_grad_floats(),), self.device)
        self._ok(self.lib.st_unpack_clip(self.h, _ptr(packed), self.table(grads, "grads"), float(grad_scale), float(max_norm),
                                         _ptr(total_norm), self._stream()), "st_unpack_clip")

    def adam_step_clipped(self, params, grads, exp_avg, exp_avg_sq, hp):
        self._ok(self.lib.st_adam_step_clipped(self.h, self.table(params, "params"), self.table(grads, "grads"),
                                               self.table(exp_avg, "exp_avg"), self.table(exp_avg_sq, "exp_avg_sq"),
                                               ctypes.byref(hp), self._stream()), "st_adam_step_clipped")

    def launch_count(self):
        return int(self.lib.st_launch_count(self.h))

    def graph_replays(self):
        """Train steps served by replaying a captured CUDA graph (st_train_step on a capturable stream)."""
        return int(self.lib.st_debug_graph_replays(self.h))

    def profile(self, enable):
        self._ok(self.lib.st_profile(self.h, int(bool(enable))), "st_profile")

    def profile_read(self):
        """{stage: (total_ms, calls)} since the last read (synchronises the device)."""
        n = self.lib.st_profile_stage_count()
        ms = (ctypes.c_float * n)()
        calls = (ctypes.c_long * n)()
        self._ok(self.lib.st_profile_read(self.h, ms, calls), "st_profile_read")
        return {self.lib.st_profile_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    def debug_gemm(self, use_tc, a_mn, b_mn, a_hi, a_lo, a_ld, b_hi, b_lo, b_ld, M, N, K, splits=1):
        """Test hook (st_debug_gemm): returns the (planes, M, N) tensor of split-K planes."""
        planes = max(1, splits)
        C = torch.full((planes, M, N), float("nan"), device=self.device, dtype=torch.float32)
        r = self.lib.st_debug_gemm(self.h, int(use_tc), int(a_mn), int(b_mn), _ptr(a_hi), _ptr(a_lo), a_ld, _ptr(b_hi), _ptr(b_lo),
                                   b_ld, _ptr(C), N, M, N, K, splits, self._stream())
        if r < 0:
            raise RuntimeError("signaltrain_b200: st_debug_gemm: shape not covered or launch failed")
        return C[:r]

    def debug_read(self, name):
        n = self.lib.st_debug_numel(self.h, name.encode())
        if n < 0:
            raise RuntimeError(f"signaltrain_b200: unknown debug buffer {name}")
        out = np.empty(n, dtype=np.float32)
        self._ok(self.lib.st_debug_read(self.h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n), "st_debug_read")
        return out
