"""Host-side mirror of the reference model classes (`signaltrain/nn_proc.py`): `AsymAutoEncoder` (:28-126),
`AsymMPAEC` (:264-340) and `st_model` (:344-393).  Same constructor arguments, attribute names, parameter
names/shapes and call results; `forward` and its gradient run in the CUDA library through one
autograd.Function.  There is no PyTorch/CPU fallback: inputs must be float32 CUDA tensors."""
import torch
import torch.nn as nn

from .cls_fe_dft import Analysis, Synthesis
from .engine import Engine, Geometry

_LAYER_NAMES = ("fnn_enc", "fnn_enc2", "fnn_enc3", "fnn_enc4", "fnn_addknobs", "fnn_dec4", "fnn_dec3", "fnn_dec2", "fnn_dec")


class AsymAutoEncoder(nn.Module):
    """Nine Linear layers + ELU with the knobs concatenated before `fnn_addknobs`.  Parameter container:
    the layers are evaluated by the fused CUDA autoencoder kernels, never by torch."""

    def __init__(self, T=25, R=64, K=3, OT=None, use_bias=True, use_dropout=False):
        super().__init__()
        if not use_bias or use_dropout:
            raise NotImplementedError("signaltrain_b200 implements the live configuration only: use_bias=True, "
                                      "use_dropout=False (nn_proc.py:29 defaults)")
        self._T, self._R, self._K = T, R, K
        self._OT = T if OT is None else OT
        self.use_bias, self.use_dropout = use_bias, use_dropout
        widths = [(T, R), (R, R // 2), (R // 2, R // 4), (R // 4, R // 4), (R // 4 + K, R // 4), (R // 4, R // 4),
                  (R // 4, R // 2), (R // 2, R), (R, self._OT)]
        for name, (i, o) in zip(_LAYER_NAMES, widths):
            setattr(self, name, nn.Linear(i, o, bias=True))
        self.layer_list = [getattr(self, n) for n in _LAYER_NAMES]
        self.relu = nn.ELU()       # the reference calls its ELU "relu" (nn_proc.py:63)
        self.initialize()

    def initialize(self):
        for layer in self.layer_list:
            torch.nn.init.xavier_normal_(layer.weight)
            layer.bias.data.zero_()

    def forward(self, x_input, knobs, skip_connections='res', return_acts=False):
        raise NotImplementedError("AsymAutoEncoder is evaluated inside AsymMPAEC.forward by the fused CUDA kernels "
                                  "(magnitude: 'sf', phase: ''); it has no stand-alone forward in signaltrain_b200")


class _MPAECFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mpaec, return_acts, need_grad, x, knobs, *params):
        eng = mpaec._engine_for(x)
        eng.set_training(need_grad)        # saves the autoencoder activations for backward only when one can follow
        y_hat, mag, mag_hat, acts = eng.forward(x, knobs, [p.detach() for p in params], return_acts)
        ctx.mpaec, ctx.eng = mpaec, eng
        ctx.ticket = mpaec._new_ticket()
        ctx.params = params
        outs = (y_hat, mag, mag_hat) + (tuple(acts) if return_acts else ())
        if return_acts:
            ctx.mark_non_differentiable(*acts)
        return outs

    @staticmethod
    def backward(ctx, g_y, g_mag, g_mag_hat, *unused):
        mpaec, eng = ctx.mpaec, ctx.eng
        if ctx.ticket != mpaec._ticket:
            raise RuntimeError("signaltrain_b200: backward() must follow the forward() of the same batch (the "
                               "activations live in the engine workspace and a later forward overwrote them)")
        if g_y is None:
            g_y = torch.zeros((ctx.params[0].shape[0] * 0 + mpaec._last_B, eng.g.L), device=eng.device)
        grads = [torch.empty_like(p) for p in ctx.params]
        eng.backward(g_y.contiguous(), None if g_mag is None else g_mag.contiguous(),
                     None if g_mag_hat is None else g_mag_hat.contiguous(), [p.detach() for p in ctx.params], grads)
        return (None, None, None, None, None) + tuple(grads)


class AsymMPAEC(nn.Module):
    """Asymmetric magnitude/phase autoencoder with knobs: analysis -> (mag, phase) -> two autoencoders ->
    polar-to-rectangular -> synthesis -> input residual."""

    def __init__(self, expected_time_frames, ft_size=1024, hop_size=384, decomposition_rank=64, n_knobs=4, output_tf=None):
        super().__init__()
        self.output_tf = expected_time_frames if output_tf is None else output_tf
        self.expected_time_frames = expected_time_frames
        self.dft_analysis = Analysis(ft_size=ft_size, hop_size=hop_size)
        self.dft_synthesis = Synthesis(ft_size=ft_size, hop_size=hop_size)
        self.aenc = AsymAutoEncoder(T=expected_time_frames, R=decomposition_rank, K=n_knobs, OT=self.output_tf)
        self.phs_aenc = AsymAutoEncoder(T=expected_time_frames, R=decomposition_rank, K=n_knobs, OT=self.output_tf)
        self._ft, self._hop, self._R, self._K = ft_size, hop_size, decomposition_rank, n_knobs
        self._engines = {}
        self._ticket = 0
        self._last_B = 0
        self.precision = "fp32"

    def set_precision(self, mode):
        """"fp32" | "tf32" (Engine.set_precision); applies to existing and future engines of this model."""
        if mode not in Engine.PRECISIONS:
            raise ValueError(f"signaltrain_b200: precision must be one of {sorted(Engine.PRECISIONS)}, got {mode!r}")
        self.precision = mode
        for eng in self._engines.values():
            eng.set_precision(mode)

    # ---- engine plumbing ---------------------------------------------------------------------
    def _geometry(self, chunk):
        g = Geometry.__new__(Geometry)
        g.C, g.N, g.H = int(chunk), self._ft, self._hop
        g.T, g.OT = self.expected_time_frames, self.output_tf
        g.L = (g.OT - 1) * g.H - g.N
        g.F, g.K, g.R = g.N // 2 + 1, self._K, self._R
        g.intended_out_chunk = g.L
        return g

    def _engine_for(self, x):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("signaltrain_b200: model input must be a CUDA tensor; this path has no CPU fallback")
        key = (x.device.index, int(x.shape[1]))
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self._geometry(x.shape[1]), x.device)
            if self.precision != "fp32":
                eng.set_precision(self.precision)
            self._engines[key] = eng
        return eng

    def _new_ticket(self):
        self._ticket += 1
        return self._ticket

    def ordered_parameters(self):
        """The 40 tensors in state_dict order (the C ABI's pointer-table order)."""
        ps = [self.dft_analysis.conv_analysis_real.weight, self.dft_analysis.conv_analysis_imag.weight,
              self.dft_synthesis.conv_synthesis_real.weight, self.dft_synthesis.conv_synthesis_imag.weight]
        for ae in (self.aenc, self.phs_aenc):
            for layer in ae.layer_list:
                ps += [layer.weight, layer.bias]
        return ps

    def reinitialize(self):
        self.aenc.initialize()
        self.phs_aenc.initialize()

    def clip_grad_norm_(self):
        """L1-norm clip (max 1) over the four front-end tensors only (reference nn_proc.py:299-302)."""
        ps = self.ordered_parameters()[:4]
        if any(p.grad is None for p in ps):
            raise RuntimeError("clip_grad_norm_: front-end gradients are missing (call backward() first)")
        eng = self._engine_for_device(ps[0].device)
        return eng.clip_grad_norm([p.grad for p in ps], 1.0)

    def _engine_for_device(self, device):
        for (idx, _), eng in self._engines.items():
            if idx == device.index:
                return eng
        raise RuntimeError("signaltrain_b200: no engine exists yet on %s (run a forward pass first)" % (device,))

    def forward(self, x_cuda, knobs_cuda, return_acts=False):
        self._last_B = x_cuda.shape[0]
        params = self.ordered_parameters()
        if x_cuda.dtype != torch.float32:
            raise RuntimeError(f"signaltrain_b200: float32 only (got {x_cuda.dtype}); .double()/.half() models are not supported")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        outs = _MPAECFunction.apply(self, bool(return_acts), need_grad, x_cuda.contiguous(), knobs_cuda.contiguous(), *params)
        if return_acts:
            return outs[0], outs[1], outs[2], list(outs[3:])
        return outs[0], outs[1], outs[2]


class st_model(nn.Module):
    """Wrapper with the reference's sizing rules (nn_proc.py:348-385)."""

    def __init__(self, scale_factor=1, shrink_factor=4, num_knobs=3, sr=44100, scale_scheme='lean'):
        super().__init__()
        g = Geometry(scale_factor, shrink_factor, num_knobs, scale_scheme)
        self.scale_factor, self.shrink_factor = scale_factor, shrink_factor
        self.in_chunk_size, self.out_chunk_size = g.C, g.L
        self.num_knobs = num_knobs
        print(f"st_model: in_chunk_size = {g.C}, intended out chunk = {g.intended_out_chunk}, sample rate = {sr}")
        if g.L != g.intended_out_chunk:
            print(f"st_model: out_chunk_size set to y_size = {g.L} (frames: in {g.T}, out {g.OT}, ft {g.N}, hop {g.H})")
        self.mpaec = AsymMPAEC(g.T, ft_size=g.N, hop_size=g.H, n_knobs=num_knobs, output_tf=g.OT)

    def clip_grad_norm_(self):
        return self.mpaec.clip_grad_norm_()

    def set_precision(self, mode):
        self.mpaec.set_precision(mode)
        return self

    def forward(self, x_cuda, knobs_cuda, return_acts=False):
        return self.mpaec.forward(x_cuda, knobs_cuda, return_acts=return_acts)

    def ordered_parameters(self):
        return self.mpaec.ordered_parameters()
