"""Stand-alone forward of the front-end classes (`Analysis.forward` cls_fe_dft.py:50-58, `Synthesis.forward`
:102-115) through the same CUDA contractions the model uses (st_analysis / st_synthesis).  Forward only: inside the
model the front-end trains through `AsymMPAEC.forward`'s fused backward."""
import torch

from .engine import Engine, Geometry

_engines = {}


def _engine(kind, n, hop, frames_or_chunk, device):
    key = (kind, n, hop, int(frames_or_chunk), device.index)
    eng = _engines.get(key)
    if eng is None:
        g = Geometry.__new__(Geometry)
        g.N, g.H, g.F, g.K, g.R = n, hop, n // 2 + 1, 1, 64
        if kind == "analysis":
            g.C = int(frames_or_chunk)
            g.T = (g.C + n) // hop + 1
            g.OT = g.T
        else:
            g.OT = int(frames_or_chunk)
            g.T = g.OT
            g.C = (g.OT - 1) * hop - n            # a chunk whose conv yields exactly OT frames
        g.L = (g.OT - 1) * hop - n
        g.intended_out_chunk = g.L
        if g.L <= 0 or g.C <= 0:
            raise RuntimeError(f"signaltrain_b200: {kind} of this size is too short for ft_size={n}, hop={hop}")
        eng = Engine(g, device)
        _engines[key] = eng
    return eng


def analysis_forward(mod, wave_form):
    if not wave_form.is_cuda:
        raise RuntimeError("signaltrain_b200: Analysis.forward needs a CUDA tensor (no CPU fallback)")
    x = wave_form.reshape(wave_form.shape[0], -1).float().contiguous()
    eng = _engine("analysis", mod.sz, mod.hop, x.shape[1], x.device)
    with torch.no_grad():
        return eng.analysis(x, mod.conv_analysis_real.weight.detach(), mod.conv_analysis_imag.weight.detach())


def synthesis_forward(mod, real, imag):
    if not real.is_cuda:
        raise RuntimeError("signaltrain_b200: Synthesis.forward needs CUDA tensors (no CPU fallback)")
    eng = _engine("synthesis", mod.sz, mod.hop, real.shape[1], real.device)
    with torch.no_grad():
        return eng.synthesis(real.float().contiguous(), imag.float().contiguous(),
                             mod.conv_synthesis_real.weight.detach(), mod.conv_synthesis_imag.weight.detach())
